"""CPU tests of the product's host logic: Fp carry-chain core (portable backend), tower-VM program
builder / scheduler / slot allocator / encoder, validated through the CPU emulation of the interpreter
(tests/emu) against the oracle.  No GPU needed."""
import ctypes
import os
import random

import pytest

from noble_bls12_381_b200 import synth
from noble_bls12_381_b200.vmprog import compile as vmcompile
from oracle import noble_oracle as O
from tests.emu import emu

R = 1 << 384
RINV = pow(R, -1, O.P)


def _limbs(x):
    return (ctypes.c_uint32 * 12)(*[(x >> (32 * i)) & 0xFFFFFFFF for i in range(12)])


def _val(a):
    return sum(int(a[i]) << (32 * i) for i in range(12))


def test_fp_core_mont_mul_and_lazy_mac():
    L = emu.lib()
    rng = random.Random(1)
    edge = [0, 1, O.P - 1, O.P - 2, (1 << 381) - 1, 2**32 - 1, 2**64, O.P // 2]
    cases = [(a, b) for a in edge for b in edge] + [(rng.randrange(O.P), rng.randrange(O.P)) for _ in range(500)]
    for a, b in cases:
        r = (ctypes.c_uint32 * 12)()
        L.emu_mont_mul(_limbs(a), _limbs(b), r)
        assert _val(r) == a * b * RINV % O.P
    for _ in range(200):
        n = rng.randint(1, 12)
        bound = rng.choice([1, 2, 4, 8])
        A = [rng.choice([rng.randrange(bound * O.P), bound * O.P - 1, (1 << 192) - 1, ((1 << 192) - 1) << 192, (1 << 383) - 1]) for _ in range(n)]
        B = [rng.choice([rng.randrange(bound * O.P), bound * O.P - 1, (1 << 192) - 1, ((1 << 192) - 1) << 192, (1 << 383) - 1]) for _ in range(n)]
        S = sum(x * y for x, y in zip(A, B))
        K = S // (R * O.P) + 2
        rounds = (K - 1).bit_length()
        if S >= 80 * O.P * O.P or rounds > 3:
            continue
        aa = (ctypes.c_uint32 * (12 * n))(*[(A[i] >> (32 * k)) & 0xFFFFFFFF for i in range(n) for k in range(12)])
        bb = (ctypes.c_uint32 * (12 * n))(*[(B[i] >> (32 * k)) & 0xFFFFFFFF for i in range(n) for k in range(12)])
        r = (ctypes.c_uint32 * 12)()
        L.emu_mac_redc(n, aa, bb, rounds, r)
        assert _val(r) == S * RINV % O.P


def _random_pairs(n, seed):
    rng = random.Random(seed)
    pts = []
    for _ in range(n):
        a, c = rng.randrange(1, O.R_ORDER), rng.randrange(1, O.R_ORDER)
        pts.append((O.pt_to_affine(O.G1, O.pt_multiply_unsafe(O.G1, O.G1_BASE, a)),
                    O.pt_to_affine(O.G2, O.pt_multiply_unsafe(O.G2, O.G2_BASE, c))))
    g1 = bytearray(b"".join(p[0].to_bytes(48, "big") + p[1].to_bytes(48, "big") for p, _ in pts))
    g2 = bytearray(b"".join(b"".join(v.to_bytes(48, "big") for v in (q[0][0], q[0][1], q[1][0], q[1][1])) for _, q in pts))
    return pts, g1, g2


def _miller(p, q):
    return O.miller_loop(O.calc_pairing_precomputes(*q), p)


@pytest.fixture(scope="module")
def programs():
    return {n: vmcompile.compile_program(n) for n in ("pairing", "miller", "final_exp", "miller_product", "f12_product")}


def test_programs_schedule_and_slots(programs):
    for name, b in programs.items():
        b.check_hazards()
        assert b.peak_slots <= b.nslots and b.nslots + b.nfar <= 255
        assert b.sched_stats["efficiency"] > 0.5
        # image size is consistent
        img = vmcompile.image(b)
        nrec = int.from_bytes(img[12:16], "little")
        assert len(img) == 32 + 48 * len(b.consts) + b.warps * nrec * 256


def test_pairing_program_bit_exact_on_emulator(programs):
    n = 35  # two batches, the second one ragged
    pts, g1, g2 = _random_pairs(n, 5)
    out = bytearray(576 * n)
    emu.run_program(programs["pairing"], {0: (g1, 96), 1: (g2, 192), 2: (out, 576)}, n)
    for policy in (0, 1):  # other warp interleavings must give the same bytes (dataflow waits are sufficient)
        out2 = bytearray(576 * n)
        emu.run_program(programs["pairing"], {0: (g1, 96), 1: (g2, 192), 2: (out2, 576)}, n, policy=policy)
        assert out2 == out
    mil = bytearray(576 * n)
    emu.run_program(programs["miller"], {0: (g1, 96), 1: (g2, 192), 2: (mil, 576)}, n)
    for i, (p, q) in enumerate(pts):
        f = _miller(p, q)
        assert bytes(mil[576 * i : 576 * i + 576]) == O.fp12_to_bytes(f)  # pairing(P, Q, false)
        assert bytes(out[576 * i : 576 * i + 576]) == O.fp12_to_bytes(O.fp12_final_exponentiate(f))


def test_kilic_prefix_on_emulator(programs, golden_dir):
    n = 8
    g1, g2 = synth.multiples_wire(n)
    out = bytearray(576 * n)
    emu.run_program(programs["pairing"], {0: (bytearray(g1), 96), 1: (bytearray(g2), 192), 2: (out, 576)}, n)
    gold = open(os.path.join(golden_dir, "pairing_kilic_1000.bin"), "rb").read()
    assert bytes(out) == gold[: 576 * n]


def test_final_exp_kat_on_emulator(programs, golden_dir):
    import json
    k = json.load(open(os.path.join(golden_dir, "pairing_kats.json")))
    fin = O.fp12_from_twelve([int(x, 16) for x in k["final_exp_in"]])
    out = bytearray(576)
    emu.run_program(programs["final_exp"], {2: (out, 576), 3: (bytearray(O.fp12_to_bytes(fin)), 576)}, 1)
    assert [int.from_bytes(out[48 * i : 48 * i + 48], "big") for i in range(12)] == [int(x, 16) for x in k["final_exp_out"]]


def test_product_tree_programs_on_emulator(programs):
    n = 37
    pts, g1, g2 = _random_pairs(n, 7)
    part = bytearray(576 * 2)
    emu.run_program(programs["miller_product"], {0: (g1, 96), 1: (g2, 192), 2: (part, 576)}, n)
    mill = [_miller(p, q) for p, q in pts]

    def prod(fs):
        r = O.FP12_ONE
        for f in fs:
            r = O.fp12_mul(r, f)
        return r

    assert bytes(part[:576]) == O.fp12_to_bytes(prod(mill[:32]))
    assert bytes(part[576:]) == O.fp12_to_bytes(prod(mill[32:]))
    out = bytearray(576)
    emu.run_program(programs["f12_product"], {2: (out, 576), 3: (part, 576)}, 2)
    assert bytes(out) == O.fp12_to_bytes(prod(mill))


@pytest.mark.parametrize("k", [2, 3])
def test_pairs_per_lane_miller_product_on_emulator(k):
    """miller_product<k>: k consecutive items per lane with shared Fp12 squarings == product of the individual Miller
    loops, bit for bit (raw, before any final exponentiation); 35 lane-items = two batches, the second ragged."""
    b = vmcompile.compile_program("miller_product%d" % k)
    b.check_hazards()
    n = 35 * k
    pts, g1, g2 = _random_pairs(n, 9)
    part = bytearray(576 * 2)
    emu.run_program(b, {0: (g1, 96 * k), 1: (g2, 192 * k), 2: (part, 576)}, n // k)
    mill = [_miller(p, q) for p, q in pts]

    def prod(fs):
        r = O.FP12_ONE
        for f in fs:
            r = O.fp12_mul(r, f)
        return r

    assert bytes(part[:576]) == O.fp12_to_bytes(prod(mill[: 32 * k]))
    assert bytes(part[576:]) == O.fp12_to_bytes(prod(mill[32 * k :]))


@pytest.mark.parametrize("warps", [10, 12])
def test_wide_cta_requirement_fields_on_emulator(warps, golden_dir):
    """10- and 12-warp CTAs pack their progress requirements as 12- / 10-bit fields (vm_kernel.cu, builder.encode): the
    final-exponentiation KAT (pairing.test.ts:65-96) through programs scheduled for those widths."""
    import json
    b = vmcompile.compile_program("final_exp", warps=warps)
    b.check_hazards()
    k = json.load(open(os.path.join(golden_dir, "pairing_kats.json")))
    fin = O.fp12_from_twelve([int(x, 16) for x in k["final_exp_in"]])
    out = bytearray(576)
    emu.run_program(b, {2: (out, 576), 3: (bytearray(O.fp12_to_bytes(fin)), 576)}, 1)
    assert [int.from_bytes(out[48 * i : 48 * i + 48], "big") for i in range(12)] == [int(x, 16) for x in k["final_exp_out"]]


def test_synth_matches_oracle_multiples():
    g1, g2 = synth.multiples_wire(5)
    for i in range(5):
        p = O.pt_to_affine(O.G1, O.pt_multiply_unsafe(O.G1, O.G1_BASE, i + 1))
        q = O.pt_to_affine(O.G2, O.pt_multiply_unsafe(O.G2, O.G2_BASE, i + 1))
        assert g1[96 * i : 96 * i + 96] == p[0].to_bytes(48, "big") + p[1].to_bytes(48, "big")
        assert g2[192 * i : 192 * i + 192] == b"".join(v.to_bytes(48, "big") for v in (q[0][0], q[0][1], q[1][0], q[1][1]))


@pytest.mark.parametrize("name", ["pairing", "final_exp"])
def test_two_output_record_format_on_emulator(name, golden_dir):
    """The experimental two-output record format (shared products, OP_MAC2: csrc/vm.cuh) computes the same bytes: the
    pairing program against the reference's kilic vectors, finalExponentiate against pairing.test.ts:65-96."""
    b = vmcompile.compile_program(name, dual=True)
    b.check_hazards()
    assert any(op.kind == "mac2" for op in b.ops)
    assert sum(op.n_products() for op in b.ops) < 0.8 * sum(op.n_products() for op in vmcompile.compile_program(name, dual=False).ops)
    if name == "pairing":
        n = 33
        g1, g2 = synth.multiples_wire(n)
        out = bytearray(576 * n)
        emu.run_program(b, {0: (bytearray(g1), 96), 1: (bytearray(g2), 192), 2: (out, 576)}, n)
        assert bytes(out) == open(os.path.join(golden_dir, "pairing_kilic_1000.bin"), "rb").read()[: 576 * n]
    else:
        import json
        k = json.load(open(os.path.join(golden_dir, "pairing_kats.json")))
        fin = O.fp12_to_bytes(O.fp12_from_twelve([int(x, 16) for x in k["final_exp_in"]]))
        out = bytearray(576)
        emu.run_program(b, {3: (bytearray(fin), 576), 2: (out, 576)}, 1)
        assert bytes(out) == b"".join(int(x, 16).to_bytes(48, "big") for x in k["final_exp_out"])
