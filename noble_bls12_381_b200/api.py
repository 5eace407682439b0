"""Host-side mirror of noble-bls12-381's export surface (index.ts:22, 715-821) over the device engine.

Same names, argument meaning and error behaviour as the reference so that the parity tests read like the
reference's own tests (test/index.test.ts, test/pairing.test.ts):

    pairing(P, Q, withFinalExponent=True) -> Fp12          index.ts:715-722
    getPublicKey(privateKey) -> bytes(48)                  index.ts:738-740   (host-side: out of kernel scope)
    sign(message, privateKey) -> bytes(96) | PointG2       index.ts:746-752
    verify(signature, message, publicKey) -> bool          index.ts:756-767
    aggregatePublicKeys / aggregateSignatures              index.ts:773-788
    verifyBatch(signature, messages, publicKeys) -> bool   index.ts:792-821
    PointG1 / PointG2 / Fp12 value classes, utils.getDSTLabel / setDSTLabel

`Hex` = bytes | bytearray | hex str.  Points are thin value objects holding AFFINE wire bytes (or infinity);
all group / field arithmetic runs on the device through the C ABI (include/bls381_b200.h).  The reference's
async functions are plain synchronous functions here.  There is no CPU fallback.

Known deviations from the reference (none on the verify / verifyBatch / sign / aggregate* path):
  * points are always affine: `aggregate*` on Point inputs and `sign(PointG2, ...)` return the normalised point, equal under
    `.equals()` / `toHex()` but not field-for-field identical to the reference's un-normalised projective result;
  * `PointG1.add` / `PointG2.add` go through the device's aggregate path, which decodes with the subgroup check: adding a point
    that is on the curve but outside the prime-order subgroup raises here, the reference adds it silently (math.ts:993-1025).
"""
from __future__ import annotations

from . import _lib

P = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
R_ORDER = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
ST_OK, ST_INFINITY, ST_NOT_ON_CURVE, ST_NOT_IN_SUBGROUP, ST_BAD_ENCODING, ST_NO_SQRT = range(6)

_htf = {"DST": "BLS_SIG_BLS12381G2_XMD:SHA-256_SSWU_RO_NUL_"}  # htfDefaults.DST index.ts:64


class utils:  # index.ts:94-145 (only the parts that touch the path)
    @staticmethod
    def getDSTLabel() -> str:
        return _htf["DST"]

    @staticmethod
    def setDSTLabel(new_label: str) -> None:
        if not isinstance(new_label, str) or len(new_label) > 2048 or len(new_label) == 0:
            raise TypeError("Invalid DST")
        _htf["DST"] = new_label


def _dst() -> bytes:
    return bytes(ord(c) & 0xFF for c in _htf["DST"])  # stringToBytes index.ts:166-172: charCodeAt per char into a Uint8Array (mod 256)


def _eng():
    return _lib.engine()


def _ensure_bytes(h) -> bytes:  # ensureBytes index.ts:159-163
    if isinstance(h, (bytes, bytearray, memoryview)):
        return bytes(h)
    if isinstance(h, str):
        if len(h) % 2:
            raise ValueError("hexToBytes: received invalid unpadded hex")
        return bytes.fromhex(h)
    raise TypeError("Expected hex string or bytes")


def _raise_status(st: int, group: str):
    if st == ST_NOT_ON_CURVE:
        raise ValueError(f"Invalid {group} point: not on curve {'Fp' if group == 'G1' else 'Fp2'}")
    if st == ST_NOT_IN_SUBGROUP:
        raise ValueError(f"Invalid {group} point: must be of prime-order subgroup")
    if st == ST_BAD_ENCODING:
        raise ValueError(f"Invalid compressed {group} point")
    if st == ST_NO_SQRT:
        raise ValueError("Failed to find a square root")


class Fp12:
    """576-byte value in the reference's Fp12.toBytes order (math.ts:882-884)."""

    def __init__(self, raw: bytes):
        assert len(raw) == 576
        self.raw = bytes(raw)

    ONE: "Fp12"

    @staticmethod
    def fromBytes(b) -> "Fp12":
        b = _ensure_bytes(b)
        if len(b) != 576:
            raise ValueError(f"fromBytes wrong length={len(b)}")
        return Fp12(b)

    def toBytes(self) -> bytes:
        return self.raw

    def coefficients(self):
        return [int.from_bytes(self.raw[48 * i : 48 * i + 48], "big") for i in range(12)]

    def equals(self, o: "Fp12") -> bool:
        return self.raw == o.raw

    __eq__ = lambda self, o: isinstance(o, Fp12) and self.raw == o.raw
    __hash__ = lambda self: hash(self.raw)

    def multiply(self, o: "Fp12") -> "Fp12":
        return Fp12(_eng().fp12_product(self.raw + o.raw, 2, False))

    def finalExponentiate(self) -> "Fp12":
        return Fp12(_eng().final_exp_batch(self.raw, 1))


Fp12.ONE = Fp12((1).to_bytes(48, "big") + bytes(576 - 48))


class PointG1:
    """Affine G1 point (or infinity).  fromHex / toHex follow index.ts:298-381."""

    def __init__(self, x: int | None, y: int | None, infinity: bool = False):
        self.x, self.y, self.inf = x, y, infinity

    BASE: "PointG1"
    ZERO: "PointG1"

    def isZero(self):
        return self.inf

    def wire(self) -> bytes:
        return self.x.to_bytes(48, "big") + self.y.to_bytes(48, "big")

    @staticmethod
    def fromHex(h) -> "PointG1":
        b = _ensure_bytes(h)
        if len(b) == 48:
            out, st = _eng().g1_decompress_batch(b, 1)
            if st[0] == ST_INFINITY:
                return PointG1.ZERO
            _raise_status(st[0], "G1")
            return PointG1(int.from_bytes(out[:48], "big"), int.from_bytes(out[48:], "big"))
        if len(b) == 96:  # index.ts:315-325 on the device: infinity flag, coordinates reduced by `new Fp`, assertValidity
            out, st = _eng().g1_from_uncompressed_batch(b, 1)
            if st[0] == ST_INFINITY:
                return PointG1.ZERO
            _raise_status(st[0], "G1")
            return PointG1(int.from_bytes(out[:48], "big"), int.from_bytes(out[48:], "big"))
        raise ValueError("Invalid point G1, expected 48/96 bytes")

    def assertValidity(self) -> "PointG1":
        if self.inf:
            return self
        _raise_status(_eng().g1_validate_batch(self.wire(), 1)[0], "G1")
        return self

    def negate(self) -> "PointG1":
        return self if self.inf else PointG1(self.x, (-self.y) % P)

    def equals(self, o: "PointG1") -> bool:
        return (self.inf and o.inf) or (not self.inf and not o.inf and self.x == o.x and self.y == o.y)

    def toRawBytes(self, isCompressed=False) -> bytes:
        self.assertValidity()
        if isCompressed:
            if self.inf:
                return bytes([0xC0]) + bytes(47)
            v = self.x + ((self.y * 2) // P) * (1 << 381) + (1 << 383)
            return v.to_bytes(48, "big")
        if self.inf:
            return bytes([0x40]) + bytes(95)
        return self.wire()

    def toHex(self, isCompressed=False) -> str:
        return self.toRawBytes(isCompressed).hex()

    # group law on the device (math.ts:974-1078); results are canonical affine points
    def add(self, o: "PointG1") -> "PointG1":
        if self.inf:
            return o
        if o.inf:
            return self
        out, _ = _eng().aggregate_g1(self.toRawBytes(True) + o.toRawBytes(True), 2)
        return PointG1.fromHex(out)

    def subtract(self, o: "PointG1") -> "PointG1":
        return self.add(o.negate())

    def double(self) -> "PointG1":
        return self.add(self)

    def multiply(self, scalar: int) -> "PointG1":
        if not isinstance(scalar, int) or scalar <= 0 or scalar > R_ORDER:
            raise ValueError(f"Point#multiply: invalid scalar, expected positive integer < CURVE.r. Got: {scalar}")
        if self.inf:
            return self
        out, fl = _eng().g1_scalar_mul_batch(self.wire(), scalar.to_bytes(32, "big"), 1)
        return PointG1.ZERO if fl[0] & 2 else PointG1(int.from_bytes(out[:48], "big"), int.from_bytes(out[48:], "big"))

    multiplyUnsafe = multiply
    multiplyPrecomputed = multiply

    @staticmethod
    def fromPrivateKey(pk) -> "PointG1":  # index.ts:351-353
        """BASE * normalizePrivKey(pk) on the device (constant-time fixed-window ladder, vmprog/curves.py mul_secret): the
        reference's precomputed wNAF table (math.ts:1086-1167) yields the same point; no secret-dependent host arithmetic."""
        k = normalizePrivKey(pk)
        out, fl = _eng().g1_scalar_mul_batch(PointG1.BASE.wire(), k.to_bytes(32, "big"), 1)
        return PointG1.ZERO if fl[0] & 2 else PointG1(int.from_bytes(out[:48], "big"), int.from_bytes(out[48:], "big"))


PointG1.BASE = PointG1(
    0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB,
    0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1,
)
PointG1.ZERO = PointG1(None, None, True)


class PointG2:
    """Affine G2 point (or infinity); x = (c0, c1), y = (c0, c1).  index.ts:466-712."""

    def __init__(self, x, y, infinity: bool = False):
        self.x, self.y, self.inf = x, y, infinity

    BASE: "PointG2"
    ZERO: "PointG2"

    def isZero(self):
        return self.inf

    def wire(self) -> bytes:  # C-ABI order x.c0 x.c1 y.c0 y.c1
        return b"".join(v.to_bytes(48, "big") for v in (self.x[0], self.x[1], self.y[0], self.y[1]))

    @staticmethod
    def _from_wire(w: bytes) -> "PointG2":
        v = [int.from_bytes(w[48 * i : 48 * i + 48], "big") for i in range(4)]
        return PointG2((v[0], v[1]), (v[2], v[3]))

    @staticmethod
    def fromSignature(h) -> "PointG2":  # index.ts:500-530
        b = _ensure_bytes(h)
        if len(b) == 192:
            # index.ts:503-514 with half = 96: z1 / z2 are the two 96-byte halves read as ONE number each; the flags are bits
            # 382 / 381 of z1, x = (z2 mod p, (z1 mod 2^381) mod p).  Re-encode as the 96-byte form and decode that.
            z1, z2 = int.from_bytes(b[:96], "big"), int.from_bytes(b[96:], "big")
            if (z1 >> 382) & 1:
                return PointG2.ZERO
            hi = (z1 % (1 << 381)) % P + (((z1 >> 381) & 1) << 381) + (1 << 383)
            b = hi.to_bytes(48, "big") + (z2 % P).to_bytes(48, "big")
        if len(b) != 96:
            raise ValueError("Invalid compressed signature length, must be 96 or 192")
        out, st = _eng().g2_decompress_batch(b, 1)
        if st[0] == ST_INFINITY:
            return PointG2.ZERO
        _raise_status(st[0], "G2")
        return PointG2._from_wire(out)

    @staticmethod
    def fromHex(h) -> "PointG2":  # index.ts:532-580 (uncompressed form; compressed goes through fromSignature rules)
        b = _ensure_bytes(h)
        m = b[0] & 0xE0
        if m in (0x20, 0x60, 0xE0):
            raise ValueError(f"Invalid encoding flag: {m}")
        if len(b) == 192 and not (m & 0x80):  # index.ts:563-579 on the device
            out, st = _eng().g2_from_uncompressed_batch(b, 1)
            if st[0] == ST_INFINITY:
                return PointG2.ZERO
            _raise_status(st[0], "G2")
            return PointG2._from_wire(out)
        if len(b) == 96 and (m & 0x80):
            # index.ts:542-562: an infinity flag requires an all-zero body, the square root picks the sign bit, and there is
            # NO assertValidity (a decodable point outside the subgroup is returned)
            if m & 0x40:
                if any(b[1:]) or (b[0] & 0x1F):
                    raise ValueError("Invalid compressed G2 point")
                return PointG2.ZERO
            out, st = _eng().g2_decompress_batch(b, 1)
            if st[0] == ST_NO_SQRT:
                raise ValueError("Invalid compressed G2 point")
            return PointG2._from_wire(out)
        raise ValueError("Invalid point G2, expected 96/192 bytes")

    @staticmethod
    def hashToCurve(msg, options=None) -> "PointG2":  # index.ts:481-490
        dst = bytes(ord(c) for c in (options or {}).get("DST", _htf["DST"]))
        return PointG2._from_wire(_eng().hash_to_g2_batch([_ensure_bytes(msg)], dst))

    def assertValidity(self) -> "PointG2":
        if self.inf:
            return self
        _raise_status(_eng().g2_validate_batch(self.wire(), 1)[0], "G2")
        return self

    def negate(self) -> "PointG2":
        return self if self.inf else PointG2(self.x, ((-self.y[0]) % P, (-self.y[1]) % P))

    def equals(self, o: "PointG2") -> bool:
        return (self.inf and o.inf) or (not self.inf and not o.inf and self.x == o.x and self.y == o.y)

    def multiply(self, scalar: int) -> "PointG2":  # math.ts:1061-1078
        if not isinstance(scalar, int) or scalar <= 0 or scalar > R_ORDER:
            raise ValueError(f"Point#multiply: invalid scalar, expected positive integer < CURVE.r. Got: {scalar}")
        if self.inf:
            return self
        out, fl = _eng().g2_scalar_mul_batch(self.wire(), scalar.to_bytes(32, "big"), 1)
        return PointG2.ZERO if fl[0] & 2 else PointG2._from_wire(out)

    def add(self, o: "PointG2") -> "PointG2":
        if self.inf:
            return o
        if o.inf:
            return self
        out, _ = _eng().aggregate_g2(self.toSignature() + o.toSignature(), 2)
        return PointG2.fromSignature(out)

    def subtract(self, o: "PointG2") -> "PointG2":
        return self.add(o.negate())

    def double(self) -> "PointG2":
        return self.add(self)

    multiplyUnsafe = multiply

    def toSignature(self) -> bytes:  # index.ts:586-598
        if self.inf:
            return bytes([0xC0]) + bytes(95)
        y0, y1 = self.y
        aflag = ((y1 if y1 > 0 else y0) * 2) // P
        z1 = self.x[1] + aflag * (1 << 381) + (1 << 383)
        return z1.to_bytes(48, "big") + self.x[0].to_bytes(48, "big")

    def toRawBytes(self, isCompressed=False) -> bytes:  # index.ts:604-631
        self.assertValidity()
        if isCompressed:
            if self.inf:
                return bytes([0xC0]) + bytes(95)
            y0, y1 = self.y
            flag = (y0 * 2) // P if y1 == 0 else (1 if (y1 * 2) // P else 0)
            return (self.x[1] + flag * (1 << 381) + (1 << 383)).to_bytes(48, "big") + self.x[0].to_bytes(48, "big")
        if self.inf:
            return bytes([0x40]) + bytes(191)
        return b"".join(v.to_bytes(48, "big") for v in (self.x[1], self.x[0], self.y[1], self.y[0]))

    def toHex(self, isCompressed=False) -> str:
        return self.toRawBytes(isCompressed).hex()


PointG2.BASE = PointG2(
    (0x024AA2B2F08F0A91260805272DC51051C6E47AD4FA403B02B4510B647AE3D1770BAC0326A805BBEFD48056C8C121BDB8,
     0x13E02B6052719F607DACD3A088274F65596BD0D09920B61AB5DA61BBDC7F5049334CF11213945D57E5AC7D055D042B7E),
    (0x0CE5D527727D6E118CC9CDC6DA2E351AADFD9BAA8CBDD3A76D429A695160D12C923AC9CC3BACA289E193548608B82801,
     0x0606C4A02EA734CC32ACD2B02BC28B99CB3E287E85A763AF267492AB572E99AB3F370D275CEC1DA1AAA9075FF05F79BE),
)
PointG2.ZERO = PointG2(None, None, True)


# ------------------------------------------------------------------------------------------ BLS API
def normalizePrivKey(key) -> int:  # index.ts:269-279
    if isinstance(key, (bytes, bytearray)) and len(key) == 32:
        n = int.from_bytes(key, "big")
    elif isinstance(key, str) and len(key) == 64:
        n = int(key, 16)
    elif isinstance(key, int) and not isinstance(key, bool) and key > 0:
        n = key
    else:
        raise TypeError("Expected valid private key")
    n %= R_ORDER
    if not (0 < n < R_ORDER):
        raise ValueError("Private key must be 0 < key < CURVE.r")
    return n


def pairing(Pt: PointG1, Q: PointG2, withFinalExponent: bool = True) -> Fp12:  # index.ts:715-722
    if Pt.isZero() or Q.isZero():
        raise ValueError("No pairings at point of Infinity")
    Pt.assertValidity()
    Q.assertValidity()
    return Fp12(_eng().pairing_batch(Pt.wire(), Q.wire(), 1, withFinalExponent))


def pairingBatch(Ps, Qs, withFinalExponent: bool = True):
    """n pairings in one device batch (the B200-native entry point; noble has no batched form)."""
    if len(Ps) != len(Qs):
        raise ValueError("length mismatch")
    for a, c in zip(Ps, Qs):
        if a.isZero() or c.isZero():
            raise ValueError("No pairings at point of Infinity")
    n = len(Ps)
    g1 = b"".join(p.wire() for p in Ps)
    g2 = b"".join(q.wire() for q in Qs)
    for st in _eng().g1_validate_batch(g1, n):
        _raise_status(st, "G1")
    for st in _eng().g2_validate_batch(g2, n):
        _raise_status(st, "G2")
    out = _eng().pairing_batch(g1, g2, n, withFinalExponent)
    return [Fp12(out[576 * i : 576 * i + 576]) for i in range(n)]


def _normP1(p) -> PointG1:  # index.ts:726-728
    return p if isinstance(p, PointG1) else PointG1.fromHex(p)


def _normP2(p) -> PointG2:  # index.ts:729-731
    return p if isinstance(p, PointG2) else PointG2.fromSignature(p)


def _normP2Hash(p) -> PointG2:  # index.ts:732-734
    return p if isinstance(p, PointG2) else PointG2.hashToCurve(p)


def getPublicKey(privateKey) -> bytes:  # index.ts:738-740
    return PointG1.fromPrivateKey(privateKey).toRawBytes(True)


def sign(message, privateKey):  # index.ts:746-752
    sk = normalizePrivKey(privateKey)
    if isinstance(message, PointG2):
        message.assertValidity()
        return message.multiply(sk)
    return _eng().sign_batch(sk.to_bytes(32, "big"), [_ensure_bytes(message)], _dst())


def signBatch(messages, privateKeys):
    """n signatures in one device batch (config 4 of BASELINE.json)."""
    sks = b"".join(normalizePrivKey(k).to_bytes(32, "big") for k in privateKeys)
    out = _eng().sign_batch(sks, [_ensure_bytes(m) for m in messages], _dst())
    return [out[96 * i : 96 * i + 96] for i in range(len(messages))]


def verify(signature, message, publicKey) -> bool:  # index.ts:756-767
    Pk = _normP1(publicKey)
    Hm = _normP2Hash(message)
    S = _normP2(signature)
    # pairing(P.negate(), Hm, false), pairing(G, S, false): both throw on infinity / invalid points
    for a, c in ((Pk.negate(), Hm), (PointG1.BASE, S)):
        if a.isZero() or c.isZero():
            raise ValueError("No pairings at point of Infinity")
        a.assertValidity()
        c.assertValidity()
    out = _eng().miller_product(Pk.negate().wire() + PointG1.BASE.wire(), Hm.wire() + S.wire(), 2, True)
    return out == Fp12.ONE.raw


def aggregatePublicKeys(publicKeys):  # index.ts:773-778
    if not len(publicKeys):
        raise ValueError("Expected non-empty array")
    if all(not isinstance(p, PointG1) for p in publicKeys):
        raw = [_ensure_bytes(p) for p in publicKeys]
        if all(len(r) == 48 for r in raw):
            out, st = _eng().aggregate_g1(b"".join(raw), len(raw))
            for code in st:
                if code not in (ST_OK, ST_INFINITY):
                    _raise_status(code, "G1")
            return out
    pts = [_normP1(p) for p in publicKeys]
    out, st = _eng().aggregate_g1(b"".join(p.toRawBytes(True) for p in pts), len(pts))
    if isinstance(publicKeys[0], PointG1):
        return PointG1.fromHex(out).assertValidity()
    return out


def aggregateSignatures(signatures):  # index.ts:783-788
    if not len(signatures):
        raise ValueError("Expected non-empty array")
    if all(not isinstance(s, PointG2) for s in signatures):
        raw = [_ensure_bytes(s) for s in signatures]
        if all(len(r) == 96 for r in raw):
            out, st = _eng().aggregate_g2(b"".join(raw), len(raw))
            for code in st:
                if code not in (ST_OK, ST_INFINITY):
                    _raise_status(code, "G2")
            return out
    pts = [_normP2(s) for s in signatures]
    out, st = _eng().aggregate_g2(b"".join(p.toSignature() for p in pts), len(pts))
    if isinstance(signatures[0], PointG2):
        return PointG2.fromSignature(out).assertValidity()
    return out


def verifyBatch(signature, messages, publicKeys) -> bool:  # index.ts:792-821
    if not len(messages):
        raise ValueError("Expected non-empty messages array")
    if len(publicKeys) != len(messages):
        raise ValueError("Pubkey count should equal msg count")
    all_bytes = not isinstance(signature, PointG2) and not any(isinstance(m, PointG2) for m in messages) and not any(
        isinstance(p, PointG1) for p in publicKeys)
    if all_bytes:
        pks = [_ensure_bytes(p) for p in publicKeys]
        sig = _ensure_bytes(signature)
        if len(sig) == 96 and all(len(p) == 48 for p in pks):
            # fused device pipeline (decompress + hash + Miller product + final exponentiation)
            v, st = _eng().verify_batch(sig, [_ensure_bytes(m) for m in messages], b"".join(pks), _dst())
            if v < 0:  # the reference throws before its try block (index.ts:799-801): signature first, then keys
                order = [len(pks)] + list(range(len(pks)))
                for i in order:
                    if st[i] not in (ST_OK, ST_INFINITY):
                        _raise_status(st[i], "G2" if i == len(pks) else "G1")
            return v == 1
    sig = _normP2(signature)
    nMessages = [_normP2Hash(m) for m in messages]
    nPublicKeys = [_normP1(p) for p in publicKeys]
    try:
        # `new Set(nMessages)` groups by OBJECT IDENTITY (index.ts:804-809)
        groups = []
        for m in nMessages:
            if not any(m is g for g in groups):
                groups.append(m)
        g1, g2 = [], []
        for m in groups:
            keys = [nPublicKeys[i] for i, mm in enumerate(nMessages) if mm is m]
            if len(keys) == 1:
                gpk = keys[0]
            else:
                out, _ = _eng().aggregate_g1(b"".join(k.toRawBytes(True) for k in keys), len(keys))
                gpk = PointG1.fromHex(out)
            if gpk.isZero() or m.isZero():
                raise ValueError("No pairings at point of Infinity")
            gpk.assertValidity()
            m.assertValidity()
            g1.append(gpk.wire())
            g2.append(m.wire())
        if sig.isZero():
            raise ValueError("No pairings at point of Infinity")
        sig.assertValidity()
        g1.append(PointG1.BASE.negate().wire())
        g2.append(sig.wire())
        out = _eng().miller_product(b"".join(g1), b"".join(g2), len(g1), True)
        return out == Fp12.ONE.raw
    except ValueError:
        return False


# snake_case aliases
get_public_key, aggregate_public_keys, aggregate_signatures, verify_batch = getPublicKey, aggregatePublicKeys, aggregateSignatures, verifyBatch
