"""GPU parity tests (run on the B200 box): every call goes through the C ABI (ctypes binding) and is
compared bit-for-bit with the reference's golden vectors and with the CPU oracle."""
import json
import os
import random

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


def _miller(O, p, q):
    return O.miller_loop(O.calc_pairing_precomputes(*q), p)


def _random_pairs(O, n, seed):
    rng = random.Random(seed)
    pts = []
    for _ in range(n):
        a, c = rng.randrange(1, O.R_ORDER), rng.randrange(1, O.R_ORDER)
        pts.append((O.pt_to_affine(O.G1, O.pt_multiply_unsafe(O.G1, O.G1_BASE, a)),
                    O.pt_to_affine(O.G2, O.pt_multiply_unsafe(O.G2, O.G2_BASE, c))))
    g1 = b"".join(p[0].to_bytes(48, "big") + p[1].to_bytes(48, "big") for p, _ in pts)
    g2 = b"".join(b"".join(v.to_bytes(48, "big") for v in (q[0][0], q[0][1], q[1][0], q[1][1])) for _, q in pts)
    return pts, g1, g2


def test_kilic_1000_golden_pairings(eng):
    """deterministic.test.ts:34-46: e(i*G1, i*G2), i = 1..1000, all 576,000 bytes identical."""
    from noble_bls12_381_b200 import synth
    g1, g2 = synth.multiples_wire(1000)
    out = eng.pairing_batch(g1, g2, 1000, True)
    gold = open(os.path.join(GOLDEN, "pairing_kilic_1000.bin"), "rb").read()
    assert out == gold
    assert eng.launch_count() > 0


def test_single_pairing_kat(eng):
    """pairing.test.ts:46-64 (config 1: e(G1_BASE, G2_BASE)), batch of one."""
    from noble_bls12_381_b200 import synth
    g1, g2 = synth.multiples_wire(1)
    out = eng.pairing_batch(g1, g2, 1, True)
    k = json.load(open(os.path.join(GOLDEN, "pairing_kats.json")))
    assert [int.from_bytes(out[48 * i : 48 * i + 48], "big") for i in range(12)] == [int(x, 16) for x in k["e_g1_g2"]]


def test_final_exponentiate_kat(eng, O):
    """pairing.test.ts:65-96."""
    k = json.load(open(os.path.join(GOLDEN, "pairing_kats.json")))
    fin = O.fp12_to_bytes(O.fp12_from_twelve([int(x, 16) for x in k["final_exp_in"]]))
    out = eng.final_exp_batch(fin * 33, 33)  # ragged: two batches
    want = b"".join(int(x, 16).to_bytes(48, "big") for x in k["final_exp_out"])
    assert out == want * 33


def test_random_pairs_vs_oracle_with_and_without_final_exp(eng, O):
    n = 70  # three CTAs batches, ragged tail
    pts, g1, g2 = _random_pairs(O, n, 11)
    full = eng.pairing_batch(g1, g2, n, True)
    mil = eng.pairing_batch(g1, g2, n, False)
    for i, (p, q) in enumerate(pts):
        f = _miller(O, p, q)
        assert mil[576 * i : 576 * i + 576] == O.fp12_to_bytes(f), i  # pairing(P, Q, false): exact line formulas
        assert full[576 * i : 576 * i + 576] == O.fp12_to_bytes(O.fp12_final_exponentiate(f)), i


def test_final_exp_random_vs_oracle(eng, O):
    rng = random.Random(3)
    fs = [O.fp12_from_twelve([rng.randrange(O.P) for _ in range(12)]) for _ in range(40)]
    out = eng.final_exp_batch(b"".join(O.fp12_to_bytes(f) for f in fs), len(fs))
    for i, f in enumerate(fs):
        assert out[576 * i : 576 * i + 576] == O.fp12_to_bytes(O.fp12_final_exponentiate(f)), i


@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 100, 1100])
def test_miller_product_tree(eng, O, n):
    """prod_i millerLoop(P_i, Q_i) (verifyBatch's reduction, index.ts:815) for ragged sizes; sharding and
    combination order must not change a single byte."""
    from noble_bls12_381_b200 import synth
    g1, g2 = synth.multiples_wire(n)
    got = eng.miller_product(g1, g2, n, False)
    per_item = eng.pairing_batch(g1, g2, n, False)
    acc = O.FP12_ONE
    for i in range(n):
        acc = O.fp12_mul(acc, O.fp12_from_bytes(per_item[576 * i : 576 * i + 576]))
    assert got == O.fp12_to_bytes(acc)
    if n <= 33:
        fe = eng.miller_product(g1, g2, n, True)
        assert fe == O.fp12_to_bytes(O.fp12_final_exponentiate(acc))


def test_pairs_per_lane_is_result_neutral(eng):
    """The k-items-per-lane Miller products (shared Fp12 squarings, k = 2, 3, 4) and the one-item-per-lane program give
    the same raw (unexponentiated) product bytes for sizes with every remainder and a multi-level product tree."""
    from noble_bls12_381_b200 import synth
    g1, g2 = synth.multiples_wire(3083)
    try:
        for n in (2, 3, 4, 5, 64, 97, 3083):
            outs = []
            for k in (1, 2, 3, 4):
                eng.set_option("pairs_per_lane", k)
                outs.append(eng.miller_product(g1[: 96 * n], g2[: 192 * n], n, False))
            assert outs[0] == outs[1] == outs[2] == outs[3], n
    finally:
        eng.set_option("pairs_per_lane", 3)


def test_bilinearity_at_scale(eng, O):
    """Size-independent property at a large batch: e(a_i*G1, G2) * e(-G1, a_i*G2) == 1 for every i, checked
    through the multi-Miller product + one final exponentiation == ONE, plus e(P,Q)^r == 1 spot checks."""
    from noble_bls12_381_b200 import synth
    n = 4096
    g1, g2 = synth.multiples_wire(n)
    # pair (i*G1, G2) and (-G1, i*G2): product of all 2n Miller loops must exponentiate to ONE
    g1_base = g1[:96]
    g2_base = g2[:192]
    neg_g1 = g1_base[:48] + ((O.P - int.from_bytes(g1_base[48:], "big")) % O.P).to_bytes(48, "big")
    a1 = g1 + neg_g1 * n
    a2 = g2_base * n + g2
    out = eng.miller_product(a1, a2, 2 * n, True)
    assert out == O.fp12_to_bytes(O.FP12_ONE)
    # negative control: perturb one item -> not ONE
    b2 = g2_base * (n - 1) + g2[192:384] + g2
    assert eng.miller_product(a1, b2, 2 * n, True) != O.fp12_to_bytes(O.FP12_ONE)


def test_device_pointer_entry_point_matches_host(eng):
    import torch
    from noble_bls12_381_b200 import synth
    n = 257
    g1, g2 = synth.multiples_wire(n)
    d1 = torch.frombuffer(bytearray(g1), dtype=torch.uint8).cuda()
    d2 = torch.frombuffer(bytearray(g2), dtype=torch.uint8).cuda()
    do = torch.zeros(576 * n, dtype=torch.uint8, device="cuda")
    s = torch.cuda.current_stream()
    eng.pairing_batch_dev(d1.data_ptr(), d2.data_ptr(), n, True, do.data_ptr(), s.cuda_stream)
    s.synchronize()
    gold = open(os.path.join(GOLDEN, "pairing_kilic_1000.bin"), "rb").read()
    assert bytes(do.cpu().numpy().tobytes()) == gold[: 576 * n]


def test_engine_options_do_not_change_results(eng):
    """More batches than resident CTAs (148 SMs x 2 CTAs x 32 items), ragged tail: the round-robin and the dynamically
    claimed batch assignment, TMA-staged and direct input reads must all give the same bytes; the first 1000 are the
    reference's fixtures (pairing.test.ts:98-116)."""
    from noble_bls12_381_b200 import synth
    n = 296 * 32 * 2 + 777
    g1, g2 = synth.multiples_wire(n)
    gold = open(os.path.join(GOLDEN, "pairing_kilic_1000.bin"), "rb").read()
    outs = []
    try:
        for dyn, no_tma in ((1, 0), (0, 0), (1, 1)):
            eng.set_option("dynamic_batches", dyn)
            eng.set_option("no_tma", no_tma)
            outs.append(eng.pairing_batch(g1, g2, n, True))
    finally:
        eng.set_option("dynamic_batches", 1)
        eng.set_option("no_tma", 0)
    assert outs[0][: 576 * 1000] == gold
    assert outs[0] == outs[1] == outs[2]
    with pytest.raises(Exception):
        eng.set_option("no_such_option", 1)
    assert eng.last_kernel_sm_mhz() > 100.0


def test_concurrent_streams_do_not_share_scratch(eng):
    """Two pairing launches on two CUDA streams overlap at the tail of the persistent grids; each stream has its own
    far-slot scratch, so both must still reproduce the reference's fixtures (regression test: the scratch used to be
    shared between streams)."""
    import torch
    from noble_bls12_381_b200 import synth
    n = 296 * 32 + 1000
    g1, g2 = synth.multiples_wire(n)
    gold = open(os.path.join(GOLDEN, "pairing_kilic_1000.bin"), "rb").read()
    d1 = torch.frombuffer(bytearray(g1), dtype=torch.uint8).cuda()
    d2 = torch.frombuffer(bytearray(g2), dtype=torch.uint8).cuda()
    # second launch: the same points in reverse order
    r1 = torch.flip(d1.view(n, 96), dims=[0]).contiguous()
    r2 = torch.flip(d2.view(n, 192), dims=[0]).contiguous()
    oa = torch.zeros(576 * n, dtype=torch.uint8, device="cuda")
    ob = torch.zeros(576 * n, dtype=torch.uint8, device="cuda")
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    for _ in range(2):
        eng.pairing_batch_dev(d1.data_ptr(), d2.data_ptr(), n, True, oa.data_ptr(), sa.cuda_stream)
        eng.pairing_batch_dev(r1.data_ptr(), r2.data_ptr(), n, True, ob.data_ptr(), sb.cuda_stream)
    torch.cuda.synchronize()
    a = bytes(oa.cpu().numpy().tobytes())
    b = bytes(torch.flip(ob.view(n, 576), dims=[0]).contiguous().cpu().numpy().tobytes())
    assert a[: 576 * 1000] == gold
    assert a == b
