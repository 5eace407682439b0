// Drop-in module with noble-bls12-381's export surface (index.ts:22, 94, 287, 466, 715-821) whose hot functions run on the
// B200 engine through the native addon (js/addon.cc -> include/bls381_b200.h).
//
// SOURCE ONLY in this repository: the build image has no node / tsc (INTEGRATION.md).  tests/test_js_shim.py checks what can
// be checked without them: every `native.*` name used here is exported by addon.cc, every function the reference exports is
// exported here, brackets balance.  The value classes (Fp, Fr, Fp2, Fp12, PointG1, PointG2), `utils` and CURVE are
// re-exported from the reference package, so existing user code keeps compiling and `instanceof` keeps working; only the
// functions below are replaced.  Semantics (order of checks, error messages, throw-versus-false) follow the reference line by
// line; the citations are to its index.ts.
import * as ref from '@noble/bls12-381';
// eslint-disable-next-line @typescript-eslint/no-var-requires
const native = require('./build/Release/bls381_b200.node');

export { Fp, Fr, Fp2, Fp12, CURVE, PointG1, PointG2, utils } from '@noble/bls12-381';
type Hex = Uint8Array | string;
type PrivateKey = Hex | bigint | number;
type G1Hex = Hex | ref.PointG1;
type G2Hex = Hex | ref.PointG2;

// per-item status codes of include/bls381_b200.h
const ST_OK = 0, ST_INFINITY = 1, ST_NOT_ON_CURVE = 2, ST_NOT_IN_SUBGROUP = 3, ST_BAD_ENCODING = 4, ST_NO_SQRT = 5;
const R_ORDER = ref.CURVE.r;

function concat(...arrays: Uint8Array[]): Uint8Array {
  const out = new Uint8Array(arrays.reduce((n, a) => n + a.length, 0));
  let pos = 0;
  for (const a of arrays) { out.set(a, pos); pos += a.length; }
  return out;
}
function hexToBytes(hex: string): Uint8Array { // math.ts:169-180
  if (hex.length % 2) throw new Error('hexToBytes: received invalid unpadded hex');
  const out = new Uint8Array(hex.length / 2);
  for (let i = 0; i < out.length; i++) {
    const byte = Number.parseInt(hex.slice(2 * i, 2 * i + 2), 16);
    if (Number.isNaN(byte) || byte < 0) throw new Error('Invalid byte sequence');
    out[i] = byte;
  }
  return out;
}
// ensureBytes (index.ts:159-163): ALWAYS a copy, the caller's buffer is never touched
const toBytes = (h: Hex): Uint8Array => (typeof h === 'string' ? hexToBytes(h) : Uint8Array.from(h));
const numberTo32 = (n: bigint): Uint8Array => hexToBytes(n.toString(16).padStart(64, '0'));
const g1Wire = (P: ref.PointG1): Uint8Array => { const [x, y] = P.toAffine(); return concat(x.toBytes(), y.toBytes()); };
const g2Wire = (Q: ref.PointG2): Uint8Array => { const [x, y] = Q.toAffine(); return concat(x.toBytes(), y.toBytes()); }; // c0 || c1
// htfDefaults.DST is mutable global state (index.ts:60-64, 138-144): read at call time, charCode per char (index.ts:166-172)
const dstBytes = (): Uint8Array => Uint8Array.from(ref.utils.getDSTLabel(), (c) => c.charCodeAt(0) & 0xff);
const isOne = (f12: Uint8Array): boolean => f12.every((b, i) => b === (i === 47 ? 1 : 0)); // Fp12.ONE, math.ts:707
function packMessages(msgs: Uint8Array[]): { packed: Uint8Array; off: BigUint64Array } {
  const off = new BigUint64Array(msgs.length + 1);
  msgs.forEach((m, i) => { off[i + 1] = off[i] + BigInt(m.length); });
  return { packed: concat(...msgs), off };
}

function raise(st: number, g: 'G1' | 'G2'): void {
  if (st === ST_NOT_ON_CURVE) throw new Error(`Invalid ${g} point: not on curve ${g === 'G1' ? 'Fp' : 'Fp2'}`);   // index.ts:385, 635
  if (st === ST_NOT_IN_SUBGROUP) throw new Error(`Invalid ${g} point: must be of prime-order subgroup`);            // index.ts:386, 636
  if (st === ST_BAD_ENCODING) throw new Error(`Invalid compressed ${g} point`);                                     // index.ts:312
  if (st === ST_NO_SQRT) throw new Error('Failed to find a square root');                                           // index.ts:518
}

function normalizePrivKey(key: PrivateKey): bigint { // index.ts:269-279
  let int: bigint;
  if (key instanceof Uint8Array && key.length === 32) int = BigInt('0x' + Array.from(key, (b) => b.toString(16).padStart(2, '0')).join(''));
  else if (typeof key === 'string' && key.length === 64) int = BigInt(`0x${key}`);
  else if (typeof key === 'number' && key > 0 && Number.isSafeInteger(key)) int = BigInt(key);
  else if (typeof key === 'bigint' && key > 0n) int = key;
  else throw new TypeError('Expected valid private key');
  int = ((int % R_ORDER) + R_ORDER) % R_ORDER;
  if (!(0n < int && int < R_ORDER)) throw new Error('Private key must be 0 < key < CURVE.r');
  return int;
}

// ---- pairing (index.ts:715-722) -------------------------------------------------------------------------------------
export function pairing(P: ref.PointG1, Q: ref.PointG2, withFinalExponent = true): ref.Fp12 {
  if (P.isZero() || Q.isZero()) throw new Error('No pairings at point of Infinity');               // :716
  const { out, status } = native.pairingBatch(g1Wire(P), g2Wire(Q), withFinalExponent, true);       // :717-720 on the device
  if (status[0] !== ST_OK) {
    // P.assertValidity() runs before Q.assertValidity(); the device reports the first failing check in that order
    const pBad = native.g1Validate(g1Wire(P))[0];
    raise(pBad !== ST_OK ? pBad : status[0], pBad !== ST_OK ? 'G1' : 'G2');
  }
  return ref.Fp12.fromBytes(out);
}

/** n pairings in one device batch: the B200-native entry point (the reference has no batched form).
 *  `checked = false` skips the per-point assertValidity (points that were validated when they were decoded). */
export function pairingBatch(Ps: ref.PointG1[], Qs: ref.PointG2[], withFinalExponent = true, checked = true): ref.Fp12[] {
  if (Ps.length !== Qs.length) throw new Error('pairingBatch: point count mismatch');
  if (!Ps.length) return [];
  if (Ps.some((p) => p.isZero()) || Qs.some((q) => q.isZero())) throw new Error('No pairings at point of Infinity');
  const g1 = concat(...Ps.map(g1Wire)), g2 = concat(...Qs.map(g2Wire));
  const { out, status } = native.pairingBatch(g1, g2, withFinalExponent, checked);
  for (let i = 0; i < Ps.length; i++) {
    if (status[i] === ST_OK) continue;
    const pBad = native.g1Validate(g1.subarray(96 * i, 96 * i + 96))[0];
    raise(pBad !== ST_OK ? pBad : status[i], pBad !== ST_OK ? 'G1' : 'G2');
  }
  return Ps.map((_, i) => ref.Fp12.fromBytes(out.subarray(576 * i, 576 * i + 576)));
}

// ---- keys and signatures ----------------------------------------------------------------------------------------------
export function getPublicKey(privateKey: PrivateKey): Uint8Array { // index.ts:738-740
  return native.getPublicKeyBatch(numberTo32(normalizePrivKey(privateKey)));
}

/** n public keys in one device batch. */
export function getPublicKeys(privateKeys: PrivateKey[]): Uint8Array[] {
  if (!privateKeys.length) return [];
  const out: Uint8Array = native.getPublicKeyBatch(concat(...privateKeys.map((k) => numberTo32(normalizePrivKey(k)))));
  return privateKeys.map((_, i) => out.slice(48 * i, 48 * i + 48));
}

export async function sign(message: Hex, privateKey: PrivateKey): Promise<Uint8Array>;
export async function sign(message: ref.PointG2, privateKey: PrivateKey): Promise<ref.PointG2>;
export async function sign(message: G2Hex, privateKey: PrivateKey): Promise<Uint8Array | ref.PointG2> { // index.ts:746-752
  if (message instanceof ref.PointG2) {
    message.assertValidity();                                                                      // :748
    const sk = numberTo32(normalizePrivKey(privateKey));
    if (message.isZero()) return message;                                                          // ZERO.multiply(k) == ZERO
    const { points, flags } = native.g2ScalarMul(g2Wire(message), sk);                             // :749, constant-time ladder on the device
    if (flags[0] & 2) return ref.PointG2.ZERO;
    return new ref.PointG2(ref.Fp2.fromBytes(points.subarray(0, 96)), ref.Fp2.fromBytes(points.subarray(96, 192)), ref.Fp2.ONE);
  }
  const sk = numberTo32(normalizePrivKey(privateKey));
  const { packed, off } = packMessages([toBytes(message)]);
  return native.signBatch(sk, packed, off, dstBytes());                                            // hashToCurve + multiply + toSignature
}

/** n signatures in one device batch (BASELINE config 4). */
export async function signBatch(messages: Hex[], privateKeys: PrivateKey[]): Promise<Uint8Array[]> {
  if (messages.length !== privateKeys.length) throw new Error('signBatch: key count should equal msg count');
  if (!messages.length) return [];
  const { packed, off } = packMessages(messages.map(toBytes));
  const out: Uint8Array = await native.signBatch(concat(...privateKeys.map((k) => numberTo32(normalizePrivKey(k)))), packed, off, dstBytes());
  return messages.map((_, i) => out.slice(96 * i, 96 * i + 96));
}

// ---- verification -----------------------------------------------------------------------------------------------------
export async function verify(signature: G2Hex, message: G2Hex, publicKey: G1Hex): Promise<boolean> { // index.ts:756-767
  const P = publicKey instanceof ref.PointG1 ? publicKey : ref.PointG1.fromHex(publicKey);
  const Hm = message instanceof ref.PointG2 ? message : await ref.PointG2.hashToCurve(message);
  const S = signature instanceof ref.PointG2 ? signature : ref.PointG2.fromSignature(signature);
  // pairing(P.negate(), Hm, false) then pairing(G, S, false): each throws on infinity / invalid points (:716-718)
  for (const [a, b] of [[P, Hm], [ref.PointG1.BASE, S]] as const) {
    if (a.isZero() || b.isZero()) throw new Error('No pairings at point of Infinity');
    a.assertValidity(); b.assertValidity();
  }
  const out: Uint8Array = native.millerProduct(concat(g1Wire(P.negate()), g1Wire(ref.PointG1.BASE)), concat(g2Wire(Hm), g2Wire(S)), true);
  return isOne(out);                                                                                // :765-766
}

export async function verifyBatch(signature: G2Hex, messages: G2Hex[], publicKeys: G1Hex[]): Promise<boolean> { // index.ts:792-821
  if (!messages.length) throw new Error('Expected non-empty messages array');                      // :797
  if (publicKeys.length !== messages.length) throw new Error('Pubkey count should equal msg count'); // :798
  // fused device path: a compressed 96-byte signature, byte messages, compressed 48-byte keys (hex strings never dedup in
  // the reference: `new Set` compares PointG2 objects by identity, :804)
  let fused = !(signature instanceof ref.PointG2) && messages.every((m) => !(m instanceof ref.PointG2)) &&
    publicKeys.every((p) => !(p instanceof ref.PointG1));
  let sigBytes: Uint8Array | undefined, pkBytes: Uint8Array[] | undefined;
  if (fused) {
    sigBytes = toBytes(signature as Hex);
    pkBytes = (publicKeys as Hex[]).map(toBytes);
    fused = sigBytes.length === 96 && pkBytes.every((p) => p.length === 48);   // 192-byte signatures / 96-byte keys: Point path
  }
  if (fused) {
    const n = messages.length;
    const { packed, off } = packMessages((messages as Hex[]).map(toBytes));
    const { verdict, status } = await native.verifyBatch(sigBytes, packed, off, concat(...(pkBytes as Uint8Array[])), dstBytes());
    if (verdict < 0) {
      // decoding errors reject: normP2(signature) (:799) runs before publicKeys.map(normP1) (:801)
      for (const i of [n, ...Array(n).keys()]) if (status[i] !== ST_OK && status[i] !== ST_INFINITY) raise(status[i], i === n ? 'G2' : 'G1');
    }
    return verdict === 1;
  }
  // Point inputs: identity grouping stays in JS (:804-809)
  const sig = signature instanceof ref.PointG2 ? signature : ref.PointG2.fromSignature(signature);
  const nMessages = await Promise.all(messages.map((m) => (m instanceof ref.PointG2 ? m : ref.PointG2.hashToCurve(m))));
  const nPublicKeys = publicKeys.map((p) => (p instanceof ref.PointG1 ? p : ref.PointG1.fromHex(p)));
  try {
    const g1: Uint8Array[] = [], g2: Uint8Array[] = [];
    for (const message of new Set(nMessages)) {
      const gpk = nMessages.reduce((acc, m, i) => (m === message ? acc.add(nPublicKeys[i]) : acc), ref.PointG1.ZERO);
      if (gpk.isZero() || message.isZero()) throw new Error('No pairings at point of Infinity');
      gpk.assertValidity(); message.assertValidity();
      g1.push(g1Wire(gpk)); g2.push(g2Wire(message));
    }
    if (sig.isZero()) throw new Error('No pairings at point of Infinity');
    sig.assertValidity();
    g1.push(g1Wire(ref.PointG1.BASE.negate())); g2.push(g2Wire(sig));                              // :814
    return isOne(native.millerProduct(concat(...g1), concat(...g2), true));                        // :815-817
  } catch { return false; }                                                                        // :818-820
}

// ---- aggregation (index.ts:773-788) -------------------------------------------------------------------------------------
export function aggregatePublicKeys(publicKeys: Hex[]): Uint8Array;
export function aggregatePublicKeys(publicKeys: ref.PointG1[]): ref.PointG1;
export function aggregatePublicKeys(publicKeys: G1Hex[]): Uint8Array | ref.PointG1 {
  if (!publicKeys.length) throw new Error('Expected non-empty array');                             // :774
  const bytes = publicKeys.map((p) => (p instanceof ref.PointG1 ? undefined : toBytes(p)));
  if (bytes.every((b) => b !== undefined && b.length === 48)) {                                    // compressed keys: one device call
    const { point, status } = native.aggregateG1(concat(...(bytes as Uint8Array[])));
    for (let i = 0; i < status.length; i++) if (status[i] !== ST_OK && status[i] !== ST_INFINITY) raise(status[i], 'G1');
    return point;
  }
  const agg = publicKeys.map((p) => (p instanceof ref.PointG1 ? p : ref.PointG1.fromHex(p))).reduce((sum, p) => sum.add(p), ref.PointG1.ZERO);
  if (publicKeys[0] instanceof ref.PointG1) { agg.assertValidity(); return agg; }                  // :776
  return agg.toRawBytes(true);
}

export function aggregateSignatures(signatures: Hex[]): Uint8Array;
export function aggregateSignatures(signatures: ref.PointG2[]): ref.PointG2;
export function aggregateSignatures(signatures: G2Hex[]): Uint8Array | ref.PointG2 {
  if (!signatures.length) throw new Error('Expected non-empty array');                             // :784
  const bytes = signatures.map((s) => (s instanceof ref.PointG2 ? undefined : toBytes(s)));
  if (bytes.every((b) => b !== undefined && b.length === 96)) {                                    // compressed signatures: one device call
    const { point, status } = native.aggregateG2(concat(...(bytes as Uint8Array[])));
    for (let i = 0; i < status.length; i++) if (status[i] !== ST_OK && status[i] !== ST_INFINITY) raise(status[i], 'G2');
    return point;
  }
  const agg = signatures.map((s) => (s instanceof ref.PointG2 ? s : ref.PointG2.fromSignature(s))).reduce((sum, s) => sum.add(s), ref.PointG2.ZERO);
  if (signatures[0] instanceof ref.PointG2) { agg.assertValidity(); return agg; }                  // :786
  return agg.toSignature();
}

/** GPUs the addon shards verifyBatch over (BLS381_B200_DEVICES bit mask; bls381_verify_batch_multi). */
export const deviceCount = (): number => native.deviceCount();

// ---- batch forms of the (de)serialisers and of hashToCurve (SURVEY 8f): one device call for n inputs -------------------
const g1FromWire = (w: Uint8Array): ref.PointG1 =>
  w.every((b) => b === 0) ? ref.PointG1.ZERO : new ref.PointG1(ref.Fp.fromBytes(w.subarray(0, 48)), ref.Fp.fromBytes(w.subarray(48, 96)), ref.Fp.ONE);
const g2FromWire = (w: Uint8Array): ref.PointG2 =>
  w.every((b) => b === 0) ? ref.PointG2.ZERO : new ref.PointG2(ref.Fp2.fromBytes(w.subarray(0, 96)), ref.Fp2.fromBytes(w.subarray(96, 192)), ref.Fp2.ONE);

/** PointG1.fromHex (index.ts:298-327) + assertValidity for n encodings of ONE length (48 compressed or 96 uncompressed). */
export function pointsG1FromHex(hexes: Hex[]): ref.PointG1[] {
  const bytes = hexes.map(toBytes);
  const len = bytes.length ? bytes[0].length : 0;
  if ((len !== 48 && len !== 96) || bytes.some((b) => b.length !== len)) throw new Error('Invalid point G1, expected 48/96 bytes');   // :326
  const { points, status } = len === 48 ? native.g1Decompress(concat(...bytes)) : native.g1FromUncompressed(concat(...bytes));
  return bytes.map((_, i) => { raise(status[i], 'G1'); return status[i] === ST_INFINITY ? ref.PointG1.ZERO : g1FromWire(points.subarray(96 * i, 96 * i + 96)); });
}
/** PointG2.fromSignature / fromHex (index.ts:500-580) + assertValidity for n encodings of ONE length (96 or 192). */
export function pointsG2FromHex(hexes: Hex[]): ref.PointG2[] {
  const bytes = hexes.map(toBytes);
  const len = bytes.length ? bytes[0].length : 0;
  if ((len !== 96 && len !== 192) || bytes.some((b) => b.length !== len)) throw new Error('Invalid compressed signature length, must be 96 or 192');   // :504
  const { points, status } = len === 96 ? native.g2Decompress(concat(...bytes)) : native.g2FromUncompressed(concat(...bytes));
  return bytes.map((_, i) => { raise(status[i], 'G2'); return status[i] === ST_INFINITY ? ref.PointG2.ZERO : g2FromWire(points.subarray(192 * i, 192 * i + 192)); });
}
/** PointG1#toRawBytes(isCompressed) (index.ts:355-376) for n points. */
export function pointsG1ToRawBytes(points: ref.PointG1[], isCompressed = false): Uint8Array[] {
  const w = isCompressed ? 48 : 96;
  const out: Uint8Array = native.g1Encode(concat(...points.map((P) => (P.isZero() ? new Uint8Array(96) : g1Wire(P)))), isCompressed);
  return points.map((_, i) => out.slice(w * i, w * i + w));
}
/** PointG2#toSignature (compressed, index.ts:586-606) / toRawBytes (index.ts:608-631) for n points. */
export function pointsG2ToRawBytes(points: ref.PointG2[], isCompressed = false): Uint8Array[] {
  const w = isCompressed ? 96 : 192;
  const out: Uint8Array = native.g2Encode(concat(...points.map((Q) => (Q.isZero() ? new Uint8Array(192) : g2Wire(Q)))), isCompressed);
  return points.map((_, i) => out.slice(w * i, w * i + w));
}
/** PointG2.hashToCurve (index.ts:481-490) for n messages with the current DST. */
export function hashToCurveG2Batch(messages: Hex[]): ref.PointG2[] {
  const { packed, off } = packMessages(messages.map(toBytes));
  const out: Uint8Array = native.hashToG2(packed, off, dstBytes());
  return messages.map((_, i) => g2FromWire(out.subarray(192 * i, 192 * i + 192)));
}
/** PointG1.hashToCurve (index.ts:331-339) for n messages with the current DST. */
export function hashToCurveG1Batch(messages: Hex[]): ref.PointG1[] {
  const { packed, off } = packMessages(messages.map(toBytes));
  const out: Uint8Array = native.hashToG1(packed, off, dstBytes());
  return messages.map((_, i) => g1FromWire(out.subarray(96 * i, 96 * i + 96)));
}
