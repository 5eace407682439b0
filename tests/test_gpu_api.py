"""GPU tests of the host-side mirror of noble's API (noble_bls12_381_b200/api.py); they follow the reference's own
tests: test/index.test.ts (sign KATs, verify, aggregate, verifyBatch truth tables, error cases) and
test/pairing.test.ts (bilinearity etc.)."""
import os
import random

import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def bls():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from noble_bls12_381_b200 import api
    api._eng()
    return api


@pytest.fixture(scope="module")
def vectors():
    return [l.split(":") for l in open(os.path.join(GOLDEN, "sign_g2_vectors.txt")).read().strip().split("\n")]


def test_sign_all_559_kats(bls, vectors):
    """index.test.ts:287-293 -- every `priv:msg:sig` line, signed in ONE device batch."""
    sigs = bls.signBatch([bytes.fromhex(v[1]) for v in vectors], [v[0].rjust(64, "0") for v in vectors])
    assert len(sigs) == 559
    for s, v in zip(sigs, vectors):
        assert s.hex() == v[2]
    # single-call form
    assert bls.sign(vectors[7][1], vectors[7][0].rjust(64, "0")).hex() == vectors[7][2]


def test_sign_scalar_digits_edge_cases(bls, vectors):
    """The sign program takes the scalar as four base-|x| digits (device kernel base_z_digits_kernel): scalars whose
    digits are all-zero / all-maximal, and unreduced 32-byte scalars k + r, 2^256 - 1 (the C ABI reduces them mod r:
    [k]H(m) = [k mod r]H(m) on G2) must give the signature of the reduced key."""
    from noble_bls12_381_b200 import _lib
    eng = _lib.engine()
    dst = b"BLS_SIG_BLS12381G2_XMD:SHA-256_SSWU_RO_NUL_"
    r = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
    z = 0xD201000000010000
    keys = [1, z - 1, z, z * z, z**3 + 1, (z - 1) * (1 + z + z * z), r - 1, int(vectors[3][0], 16) % r]
    msgs = [bytes.fromhex(vectors[3][1])] * len(keys)
    base = eng.sign_batch(b"".join(k.to_bytes(32, "big") for k in keys), msgs, dst)
    from oracle import noble_oracle as O
    for i, k in enumerate(keys):  # against the CPU restatement of index.ts:746-752
        assert base[96 * i: 96 * i + 96] == O.sign(msgs[i], k), i
    assert base[96 * 7: 96 * 8].hex() == vectors[3][2]
    unreduced = [(k + r) for k in keys if k + r < 1 << 256] + [(1 << 256) - 1]
    m2 = [msgs[0]] * len(unreduced)
    got = eng.sign_batch(b"".join(k.to_bytes(32, "big") for k in unreduced), m2, dst)
    want = eng.sign_batch(b"".join((k % r).to_bytes(32, "big") for k in unreduced), m2, dst)
    assert got == want


def test_get_public_key_and_verify(bls, vectors):
    """index.test.ts:308-336"""
    for i in range(4):
        priv, msg, sig = vectors[i][0].rjust(64, "0"), vectors[i][1], vectors[i][2]
        pub = bls.getPublicKey(priv)
        assert bls.verify(sig, msg, pub) is True
        assert bls.verify(sig, vectors[i + 1][1] or "00", pub) is False
        assert bls.verify(sig, msg, bls.getPublicKey(vectors[i + 1][0].rjust(64, "0"))) is False


def test_aggregate_and_verify_batch(bls, vectors):
    """index.test.ts:337-426"""
    rng = random.Random(1)
    privs = [rng.randrange(1, bls.R_ORDER) for _ in range(5)]
    msgs = [bytes(rng.randrange(256) for _ in range(32)).hex() for _ in range(5)]
    pubs = [bls.getPublicKey(p) for p in privs]
    sigs = bls.signBatch(msgs, privs)
    agg = bls.aggregateSignatures(sigs)
    assert bls.verifyBatch(agg, msgs, pubs) is True
    wrong = list(msgs); wrong[2] = msgs[3]
    assert bls.verifyBatch(agg, wrong, pubs) is False
    wpubs = list(pubs); wpubs[0] = bls.getPublicKey(privs[0] + 1)
    assert bls.verifyBatch(agg, msgs, wpubs) is False
    # same message signed by many: aggregate signature verifies against the aggregate public key
    m = msgs[0]
    sigs1 = bls.signBatch([m] * 5, privs)
    assert bls.verify(bls.aggregateSignatures(sigs1), m, bls.aggregatePublicKeys(pubs)) is True
    assert bls.verify(bls.aggregateSignatures(sigs1), msgs[1], bls.aggregatePublicKeys(pubs)) is False
    # Point inputs: messages given as PointG2 are used as already hashed and grouped by identity (index.ts:804)
    H = bls.PointG2.hashToCurve(m)
    assert bls.verifyBatch(bls.aggregateSignatures(sigs1), [H] * 5, [bls.PointG1.fromHex(p) for p in pubs]) is True
    # argument errors reject, data errors inside the try block give false
    with pytest.raises(ValueError, match="non-empty"):
        bls.verifyBatch(agg, [], [])
    with pytest.raises(ValueError, match="count"):
        bls.verifyBatch(agg, msgs, pubs[:-1])
    inf = bytes([0xC0]) + bytes(47)
    assert bls.verifyBatch(agg, msgs, [inf] + pubs[1:]) is False
    with pytest.raises(ValueError, match="Expected non-empty array"):
        bls.aggregatePublicKeys([])


def test_invalid_points_raise_reference_messages(bls):
    """index.test.ts:262-285, point.test.ts on-curve vectors"""
    with pytest.raises(ValueError, match="Invalid point G1, expected 48/96 bytes"):
        bls.PointG1.fromHex(b"\x00" * 47)
    with pytest.raises(ValueError, match="Invalid compressed signature length"):
        bls.PointG2.fromSignature(b"\x00" * 95)
    bad = bls.PointG1(bls.PointG1.BASE.x, (bls.PointG1.BASE.y + 1) % bls.P)
    with pytest.raises(ValueError, match="not on curve Fp"):
        bad.assertValidity()
    with pytest.raises(ValueError, match="No pairings at point of Infinity"):
        bls.pairing(bls.PointG1.ZERO, bls.PointG2.BASE)
    valid = bls.getPublicKey(5)
    x = 5
    while pow((x**3 + 4) % bls.P, (bls.P - 1) // 2, bls.P) == 1:
        x += 1
    with pytest.raises(ValueError, match="Invalid compressed G1 point"):
        bls.aggregatePublicKeys([valid, (x + (1 << 383)).to_bytes(48, "big")])
    # serialisation round trips incl. ZERO (index.test.ts:37-261)
    assert bls.PointG1.fromHex(bls.PointG1.BASE.toHex(True)).equals(bls.PointG1.BASE)
    assert bls.PointG1.fromHex(bls.PointG1.BASE.toHex(False)).equals(bls.PointG1.BASE)
    assert bls.PointG2.fromHex(bls.PointG2.BASE.toHex(False)).equals(bls.PointG2.BASE)
    assert bls.PointG2.fromSignature(bls.PointG2.BASE.toSignature()).equals(bls.PointG2.BASE)
    assert bls.PointG1.fromHex(bls.PointG1.ZERO.toHex(True)).isZero()
    assert bls.PointG2.fromSignature(bls.PointG2.ZERO.toSignature()).isZero()


def test_pairing_algebra(bls):
    """pairing.test.ts:8-45"""
    G1, G2 = bls.PointG1.BASE, bls.PointG2.BASE
    p1 = bls.pairing(G1, G2)
    assert p1.multiply(bls.pairing(G1.negate(), G2)) == bls.Fp12.ONE
    assert bls.pairing(G1.negate(), G2) == bls.pairing(G1, G2.negate())
    g1x2 = bls.PointG1.fromPrivateKey(2)
    g2x2 = G2.multiply(2)
    assert p1.multiply(p1) == bls.pairing(g1x2, G2) == bls.pairing(G1, g2x2)
    assert p1 != bls.pairing(g1x2, G2)
    a = bls.pairing(bls.PointG1.fromPrivateKey(37), G2.multiply(27))
    assert a == bls.pairing(bls.PointG1.fromPrivateKey(999), G2)
    # pairing(P, Q, false).finalExponentiate() == pairing(P, Q)
    assert bls.pairing(G1, G2, False).finalExponentiate() == p1
    # batch form
    outs = bls.pairingBatch([G1, g1x2], [G2, G2])
    assert outs[0] == p1 and outs[1] == p1.multiply(p1)


def test_dst_label_is_forwarded(bls, vectors):
    old = bls.utils.getDSTLabel()
    try:
        bls.utils.setDSTLabel("BLS_SIG_BLS12381G2_XMD:SHA-256_SSWU_RO_POP_")
        assert bls.sign(vectors[3][1], vectors[3][0].rjust(64, "0")).hex() != vectors[3][2]
    finally:
        bls.utils.setDSTLabel(old)
    assert bls.sign(vectors[3][1], vectors[3][0].rjust(64, "0")).hex() == vectors[3][2]


def test_point_group_law_and_zkcrypto_multiples(bls):
    """deterministic.test.ts:49-113: i*G by repeated addition and by all multipliers equals the zkcrypto encodings;
    point.test.ts doubling identities."""
    g1c = open(os.path.join(GOLDEN, "zkcrypto_g1_compressed.dat"), "rb").read()
    g2c = open(os.path.join(GOLDEN, "zkcrypto_g2_compressed.dat"), "rb").read()
    p1, p2 = bls.PointG1.ZERO, bls.PointG2.ZERO
    for i in range(6):
        assert p1.toRawBytes(True) == g1c[48 * i : 48 * i + 48]
        assert p2.toRawBytes(True) == g2c[96 * i : 96 * i + 96]
        p1, p2 = p1.add(bls.PointG1.BASE), p2.add(bls.PointG2.BASE)
    for i in (1, 2, 3, 17, 999):
        assert bls.PointG1.BASE.multiply(i).toRawBytes(True) == g1c[48 * i : 48 * i + 48]
        assert bls.PointG1.BASE.multiplyUnsafe(i).toRawBytes(True) == g1c[48 * i : 48 * i + 48]
        assert bls.PointG2.BASE.multiply(i).toRawBytes(True) == g2c[96 * i : 96 * i + 96]
    G, H = bls.PointG1.BASE, bls.PointG2.BASE
    assert G.double().equals(G.add(G)) and G.double().equals(G.multiply(2))
    assert G.add(G.negate()).isZero() and H.subtract(H).isZero()
    assert G.multiply(bls.R_ORDER).isZero() and H.multiply(bls.R_ORDER).isZero()
    assert G.multiply(5).subtract(G.multiply(2)).equals(G.multiply(3))


def test_full_size_properties(bls):
    """Size-independent properties at BASELINE.json's sizes (config 3: 262 144 signatures; config 4 slice):
    sign -> aggregate -> verifyBatch round trip must be true, one flipped message byte must make it false."""
    import hashlib
    from noble_bls12_381_b200 import synth
    n = 262144
    eng = bls._eng()
    dst = bls._dst()
    g1, _ = synth.multiples_wire(n)  # public keys (i+1)*G1 for secret keys i+1
    pks = []
    for i in range(n):
        x = int.from_bytes(g1[96 * i : 96 * i + 48], "big")
        y = int.from_bytes(g1[96 * i + 48 : 96 * i + 96], "big")
        pks.append((x + ((y * 2) // bls.P) * (1 << 381) + (1 << 383)).to_bytes(48, "big"))
    msgs = [hashlib.sha256(i.to_bytes(8, "big")).digest() for i in range(n)]
    sks = b"".join((i + 1).to_bytes(32, "big") for i in range(n))
    sigs = eng.sign_batch(sks, msgs, dst)
    agg, st = eng.aggregate_g2(sigs, n)
    assert all(s == 0 for s in st)
    v, st = eng.verify_batch(agg, msgs, b"".join(pks), dst)
    assert v == 1 and all(s == 0 for s in st)
    msgs[n // 3] = bytes([msgs[n // 3][0] ^ 0x80]) + msgs[n // 3][1:]
    v, _ = eng.verify_batch(agg, msgs, b"".join(pks), dst)
    assert v == 0
    # spot-check a few of the 262 144 signatures against the oracle's sign()
    from oracle import noble_oracle as O
    for i in (0, 1, 31, 32, 99999, n - 1):
        m = hashlib.sha256(i.to_bytes(8, "big")).digest()
        assert sigs[96 * i : 96 * i + 96] == O.sign(m, i + 1)


def test_decoders_follow_the_reference_on_edge_encodings(bls):
    """PointG2.fromHex(96 B compressed) has its own rules (index.ts:542-562: infinity needs a zero body, no
    assertValidity), PointG2.fromSignature parses 192-byte input (index.ts:503-514), DST characters are taken mod 256
    (index.ts:166-172), fromPrivateKey runs on the device."""
    from oracle import noble_oracle as O
    g2c = open(os.path.join(GOLDEN, "zkcrypto_g2_compressed.dat"), "rb").read()
    g2u = open(os.path.join(GOLDEN, "zkcrypto_g2_uncompressed.dat"), "rb").read()
    for i in (1, 2, 77, 999):
        pt = bls.PointG2.fromHex(g2c[96 * i: 96 * i + 96])
        ref = O.g2_from_hex(g2c[96 * i: 96 * i + 96])
        (x0, x1), (y0, y1) = O.pt_to_affine(O.G2, ref)
        assert (pt.x, pt.y) == ((x0, x1), (y0, y1))
        assert bls.PointG2.fromHex(g2u[192 * i: 192 * i + 192]).equals(pt)
    assert bls.PointG2.fromHex(g2c[:96]).isZero() and bls.PointG2.fromHex(g2u[:192]).isZero()
    with pytest.raises(ValueError, match="Invalid compressed G2 point"):
        bls.PointG2.fromHex(bytes([0xC0]) + bytes(94) + b"\x01")      # infinity flag with a non-zero body
    with pytest.raises(ValueError, match="Invalid encoding flag"):
        bls.PointG2.fromHex(bytes([0x20]) + g2u[193: 192 * 2])
    # a decodable point outside the subgroup: fromHex returns it (no assertValidity), fromSignature rejects it
    xx = (1, 1)
    while True:
        xx = (xx[0] + 1, xx[1])
        yy = O.fp2_sqrt(O.fp2_add(O.fp2_pow(xx, 3), O.B2))
        if yy is not None and not O.g2_is_torsion_free((xx, yy, O.FP2_ONE)):
            break
    enc = (xx[1] + (1 << 383)).to_bytes(48, "big") + xx[0].to_bytes(48, "big")
    assert bls.PointG2.fromHex(enc).x == xx
    with pytest.raises(ValueError, match="prime-order subgroup"):
        bls.PointG2.fromSignature(enc)
    # 192-byte input of fromSignature: the two 96-byte halves are read as one number each
    sig = g2c[96 * 5: 96 * 6]
    wide = bytes(48) + sig[:48] + bytes(48) + sig[48:]
    assert bls.PointG2.fromSignature(wide).equals(bls.PointG2.fromSignature(sig))
    ref = O.g2_from_signature(wide)
    assert bls.PointG2.fromSignature(wide).x == O.pt_to_affine(O.G2, ref)[0]
    # DST characters above 255 wrap (stringToBytes writes charCodes into a Uint8Array)
    old = bls.utils.getDSTLabel()
    try:
        bls.utils.setDSTLabel("ABŁC")        # U+0141 -> 0x41
        a = bls.sign(b"msg", 7)
        bls.utils.setDSTLabel("ABAC")
        assert bls.sign(b"msg", 7) == a
    finally:
        bls.utils.setDSTLabel(old)
    g1c = open(os.path.join(GOLDEN, "zkcrypto_g1_compressed.dat"), "rb").read()
    assert bls.PointG1.fromPrivateKey(321).toRawBytes(True) == g1c[48 * 321: 48 * 322]
    assert bls.getPublicKey((321).to_bytes(32, "big")) == g1c[48 * 321: 48 * 322]
