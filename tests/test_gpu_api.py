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
