"""GPU parity tests of the verifyBatch ingest path through the C ABI: G1/G2 decompression + validity,
hash-to-curve (device SHA-256 xmd + SWU + isogeny + cofactor clearing) and verifyBatch truth table."""
import json
import os

import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def eng():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import noble_bls12_381_b200 as bls
    return bls.engine()


@pytest.fixture(scope="module")
def O():
    from oracle import noble_oracle
    return noble_oracle


def test_g1_decompress_zkcrypto_and_invalid(eng, O):
    from tests.test_vm_ingest_emu import _g1_cases, g1_expected
    g1c = open(os.path.join(GOLDEN, "zkcrypto_g1_compressed.dat"), "rb").read()
    g1u = open(os.path.join(GOLDEN, "zkcrypto_g1_uncompressed.dat"), "rb").read()
    out, st = eng.g1_decompress_batch(g1c, 1000)  # deterministic.test.ts:50-66: all 1000 encodings of i*G1
    assert st[0] == 1 and all(s == 0 for s in st[1:])
    assert out[96:] == g1u[96:]
    items = _g1_cases()
    out, st = eng.g1_decompress_batch(b"".join(items), len(items))
    for i, it in enumerate(items):
        exp_st, exp = g1_expected(it)
        assert st[i] == exp_st, i
        if exp is not None:
            assert out[96 * i : 96 * i + 96] == exp, i


def test_g2_decompress_zkcrypto_and_invalid(eng, O):
    from tests.test_vm_ingest_emu import _g2_cases, g2_expected
    g2c = open(os.path.join(GOLDEN, "zkcrypto_g2_compressed.dat"), "rb").read()
    g2u = open(os.path.join(GOLDEN, "zkcrypto_g2_uncompressed.dat"), "rb").read()
    out, st = eng.g2_decompress_batch(g2c, 1000)
    assert st[0] == 1 and all(s == 0 for s in st[1:])
    for i in range(1, 1000):  # zkcrypto order is x.c1 x.c0 y.c1 y.c0 ; C ABI order is c0 first
        u = g2u[192 * i : 192 * i + 192]
        assert out[192 * i : 192 * i + 192] == u[48:96] + u[0:48] + u[144:192] + u[96:144], i
    items = _g2_cases()
    out, st = eng.g2_decompress_batch(b"".join(items), len(items))
    for i, it in enumerate(items):
        exp_st, exp = g2_expected(it)
        assert st[i] == exp_st, i
        if exp is not None:
            assert out[192 * i : 192 * i + 192] == exp, i


def test_hash_to_g2_rfc_and_kilic_vectors(eng):
    """hashToCurve.test.ts:530-697 (G2 RO suites), incl. the 0-, 133- and 517-byte messages."""
    d = json.load(open(os.path.join(GOLDEN, "hash_to_curve.json")))
    for key in ("g2_kilic_ro", "g2_rfc_ro"):
        dst = d[key]["dst"].encode("latin1")
        msgs = [v["msg"].encode("latin1") for v in d[key]["vectors"]]
        out = eng.hash_to_g2_batch(msgs, dst)
        for i, v in enumerate(d[key]["vectors"]):
            o = out[192 * i : 192 * i + 192]
            assert (o[48:96] + o[0:48] + o[144:192] + o[96:144]).hex() == v["expected"], (key, i)


def test_hash_to_g2_random_vs_oracle_and_long_dst(eng, O):
    import random
    rng = random.Random(5)
    msgs = [bytes(rng.randrange(256) for _ in range(rng.choice([0, 1, 31, 32, 55, 56, 64, 200]))) for _ in range(40)]
    for dst in (O.DEFAULT_DST, b"Q" * 300):
        out = eng.hash_to_g2_batch(msgs, dst)
        for i in (0, 1, 7, 33, 39):
            (x0, x1), (y0, y1) = O.pt_to_affine(O.G2, O.g2_hash_to_curve(msgs[i], dst))
            assert out[192 * i : 192 * i + 192] == b"".join(v.to_bytes(48, "big") for v in (x0, x1, y0, y1)), i


def _signed_batch(O, n):
    vec = [l.split(":") for l in open(os.path.join(GOLDEN, "sign_g2_vectors.txt")).read().strip().split("\n")]
    vec = [v for v in vec if v[1]][:n]  # distinct non-empty messages
    privs = [v[0].rjust(64, "0") for v in vec]
    msgs = [bytes.fromhex(v[1]) for v in vec]
    sigs = [bytes.fromhex(v[2]) for v in vec]
    pks = [O.get_public_key(p) for p in privs]
    return msgs, sigs, pks


def test_verify_batch_truth_table(eng, O):
    """index.test.ts:337-398 on the reference's own sign KATs (signatures come from the fixture file)."""
    n = 40
    msgs, sigs, pks = _signed_batch(O, n)
    agg = O.aggregate_signatures(sigs)
    dst = O.DEFAULT_DST
    v, st = eng.verify_batch(agg, msgs, b"".join(pks), dst)
    assert v == 1 and all(s == 0 for s in st)
    # wrong message / swapped keys / wrong signature -> false
    bad = list(msgs); bad[3] = bad[3] + b"\x01"
    assert eng.verify_batch(agg, bad, b"".join(pks), dst)[0] == 0
    sw = list(pks); sw[0], sw[1] = sw[1], sw[0]
    assert eng.verify_batch(agg, msgs, b"".join(sw), dst)[0] == 0
    assert eng.verify_batch(sigs[0], msgs, b"".join(pks), dst)[0] == 0
    # n = 1 equals verify (index.test.ts:308-336)
    assert eng.verify_batch(sigs[5], [msgs[5]], pks[5], dst)[0] == 1
    assert eng.verify_batch(sigs[5], [msgs[6]], pks[5], dst)[0] == 0
    # infinity public key: pairing() throws inside try -> false (index.ts:716, 818-820)
    inf = bytes([0xC0]) + bytes(47)
    v, st = eng.verify_batch(agg, msgs, inf + b"".join(pks[1:]), dst)
    assert v == 0 and st[0] == 1
    # undecodable public key: the reference throws outside the try block -> verdict -1 with the status code
    x = 5
    while O.fp_sqrt((x**3 + 4) % O.P) is not None:
        x += 1
    v, st = eng.verify_batch(agg, msgs, (x + (1 << 383)).to_bytes(48, "big") + b"".join(pks[1:]), dst)
    assert v == -1 and st[0] == 4
    # the oracle agrees on the positive case and on one negative case
    assert O.verify_batch(agg, msgs[:6], pks[:6]) is False
    assert O.verify_batch(O.aggregate_signatures(sigs[:6]), msgs[:6], pks[:6]) is True
    assert eng.verify_batch(O.aggregate_signatures(sigs[:6]), msgs[:6], b"".join(pks[:6]), dst)[0] == 1


def test_sharded_partials_equal_fused_verify_batch(eng, O):
    """Config-5 building blocks on one GPU: shard partials + fp12_product == fused verifyBatch, for 1/2/4/8 shards."""
    from noble_bls12_381_b200 import dist as bdist
    n = 24
    msgs, sigs, pks = _signed_batch(O, n)
    agg = O.aggregate_signatures(sigs)
    dst = O.DEFAULT_DST
    be = bdist.EngineBackend(eng)
    ref = None
    for world in (1, 2, 4, 8):
        parts = []
        for r in range(world):
            lo, hi = bdist.shard_range(n, r, world)
            p, st = be.partial(agg if r == 0 else None, msgs[lo:hi], b"".join(pks[lo:hi]), dst) if (hi > lo or r == 0) else (bdist.FP12_ONE, [])
            assert all(s == 0 for s in st)
            parts.append(p)
        res = be.combine(parts, True)
        assert res == bdist.FP12_ONE
        raw = be.combine(parts, False)
        ref = ref or raw
        assert raw == ref  # identical un-exponentiated product regardless of the shard count
    v, _ = bdist.verify_batch_sharded(be, agg, msgs, pks, dst)
    assert v == 1
