// Drop-in module with noble-bls12-381's export surface (index.ts:22, 715-821) whose hot functions call the native
// addon (js/addon.cc -> include/bls381_b200.h).  SOURCE ONLY (no node/tsc in the build image, see INTEGRATION.md).
// Everything not on the hot path (Fp/Fp2/Fp12/Point classes as value objects, utils, getPublicKey) is re-exported
// from the reference package so that existing user code keeps compiling.
import * as ref from '@noble/bls12-381';
// eslint-disable-next-line @typescript-eslint/no-var-requires
const native = require('./build/Release/bls381_b200.node');

export { Fp, Fr, Fp2, Fp12, CURVE, PointG1, PointG2, utils, getPublicKey } from '@noble/bls12-381';
type Hex = Uint8Array | string;
type G1Hex = Hex | ref.PointG1;
type G2Hex = Hex | ref.PointG2;

const ST_OK = 0, ST_INFINITY = 1, ST_NOT_ON_CURVE = 2, ST_NOT_IN_SUBGROUP = 3, ST_BAD_ENCODING = 4, ST_NO_SQRT = 5;
const concat = ref.utils.concatBytes ?? ((...a: Uint8Array[]) => Uint8Array.from(a.flatMap((x) => [...x])));
const toBytes = (h: Hex) => (typeof h === 'string' ? ref.utils.hexToBytes(h) : Uint8Array.from(h)); // ensureBytes index.ts:159
const g1Wire = (P: ref.PointG1) => { const [x, y] = P.toAffine(); return concat(x.toBytes(), y.toBytes()); };
const g2Wire = (Q: ref.PointG2) => { const [x, y] = Q.toAffine(); return concat(x.toBytes(), y.toBytes()); }; // c0||c1
const dstBytes = () => ref.utils.stringToBytes(ref.utils.getDSTLabel()); // htfDefaults.DST is mutable: read per call

function raise(st: number, g: 'G1' | 'G2'): void {
  if (st === ST_NOT_ON_CURVE) throw new Error(`Invalid ${g} point: not on curve ${g === 'G1' ? 'Fp' : 'Fp2'}`);
  if (st === ST_NOT_IN_SUBGROUP) throw new Error(`Invalid ${g} point: must be of prime-order subgroup`);
  if (st === ST_BAD_ENCODING) throw new Error(`Invalid compressed ${g} point`);
  if (st === ST_NO_SQRT) throw new Error('Failed to find a square root');
}

export function pairing(P: ref.PointG1, Q: ref.PointG2, withFinalExponent = true): ref.Fp12 {
  if (P.isZero() || Q.isZero()) throw new Error('No pairings at point of Infinity'); // index.ts:716
  P.assertValidity(); Q.assertValidity();                                            // index.ts:717-718
  return ref.Fp12.fromBytes(native.pairingBatch(g1Wire(P), g2Wire(Q), withFinalExponent));
}

/** n pairings in one device batch: the B200-native entry point (noble has no batched form). */
export function pairingBatch(Ps: ref.PointG1[], Qs: ref.PointG2[], withFinalExponent = true): ref.Fp12[] {
  const out: Uint8Array = native.pairingBatch(concat(...Ps.map(g1Wire)), concat(...Qs.map(g2Wire)), withFinalExponent);
  return Ps.map((_, i) => ref.Fp12.fromBytes(out.subarray(576 * i, 576 * i + 576)));
}

export async function verifyBatch(signature: G2Hex, messages: G2Hex[], publicKeys: G1Hex[]): Promise<boolean> {
  if (!messages.length) throw new Error('Expected non-empty messages array');                 // index.ts:797
  if (publicKeys.length !== messages.length) throw new Error('Pubkey count should equal msg count');
  const allBytes = !(signature instanceof ref.PointG2) && messages.every((m) => !(m instanceof ref.PointG2)) &&
    publicKeys.every((p) => !(p instanceof ref.PointG1));
  if (allBytes) {
    const msgs = (messages as Hex[]).map(toBytes);
    const off = new BigUint64Array(msgs.length + 1);
    msgs.forEach((m, i) => (off[i + 1] = off[i] + BigInt(m.length)));
    const { verdict, status } = await native.verifyBatch(toBytes(signature as Hex), concat(...msgs), off,
      concat(...(publicKeys as Hex[]).map(toBytes)), dstBytes());
    if (verdict < 0) {          // decoding errors reject (they are thrown before the try block, index.ts:799-801)
      const n = msgs.length;
      for (const i of [n, ...Array(n).keys()]) if (status[i] !== ST_OK && status[i] !== ST_INFINITY) raise(status[i], i === n ? 'G2' : 'G1');
    }
    return verdict === 1;
  }
  // Point inputs: identity grouping stays in JS (`new Set` dedups by object identity, index.ts:804-809)
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
    g1.push(g1Wire(ref.PointG1.BASE.negate())); g2.push(g2Wire(sig));
    const out: Uint8Array = native.millerProduct(concat(...g1), concat(...g2), true);          // index.ts:815-816
    return out.every((b, i) => b === (i === 47 ? 1 : 0));                                      // == Fp12.ONE
  } catch { return false; }                                                                    // index.ts:818-820
}

export async function verify(signature: G2Hex, message: G2Hex, publicKey: G1Hex): Promise<boolean> { // index.ts:756-767
  const P = publicKey instanceof ref.PointG1 ? publicKey : ref.PointG1.fromHex(publicKey);
  const Hm = message instanceof ref.PointG2 ? message : await ref.PointG2.hashToCurve(message);
  const S = signature instanceof ref.PointG2 ? signature : ref.PointG2.fromSignature(signature);
  for (const [a, b] of [[P, Hm], [ref.PointG1.BASE, S]] as const) {
    if (a.isZero() || b.isZero()) throw new Error('No pairings at point of Infinity');
    a.assertValidity(); b.assertValidity();
  }
  const out: Uint8Array = native.millerProduct(concat(g1Wire(P.negate()), g1Wire(ref.PointG1.BASE)), concat(g2Wire(Hm), g2Wire(S)), true);
  return out.every((b, i) => b === (i === 47 ? 1 : 0));
}
// sign / aggregatePublicKeys / aggregateSignatures follow noble_bls12_381_b200/api.py line for line
// (native.signBatch, native.aggregateG1, native.aggregateG2).
