"""Pins the plain-C oracle port against the reference's golden vectors and the Python oracle."""
import json
import os
import random

from noble_bls12_381_b200 import synth
from oracle import c_oracle as C
from oracle import noble_oracle as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_c_oracle_kilic_1000():
    g1, g2 = synth.multiples_wire(1000)
    out = C.pairing_batch(g1, g2, 1000, True)
    assert out == open(os.path.join(GOLDEN, "pairing_kilic_1000.bin"), "rb").read()


def test_c_oracle_final_exp_kat_and_random():
    k = json.load(open(os.path.join(GOLDEN, "pairing_kats.json")))
    fin = O.fp12_to_bytes(O.fp12_from_twelve([int(x, 16) for x in k["final_exp_in"]]))
    assert C.final_exp(fin) == b"".join(int(x, 16).to_bytes(48, "big") for x in k["final_exp_out"])
    rng = random.Random(9)
    for _ in range(5):
        a = O.fp12_from_twelve([rng.randrange(O.P) for _ in range(12)])
        b = O.fp12_from_twelve([rng.randrange(O.P) for _ in range(12)])
        assert C.fp12_mul(O.fp12_to_bytes(a), O.fp12_to_bytes(b)) == O.fp12_to_bytes(O.fp12_mul(a, b))
        assert C.final_exp(O.fp12_to_bytes(a)) == O.fp12_to_bytes(O.fp12_final_exponentiate(a))


def test_c_oracle_miller_matches_python_oracle():
    rng = random.Random(4)
    for _ in range(4):
        p = O.pt_to_affine(O.G1, O.pt_multiply_unsafe(O.G1, O.G1_BASE, rng.randrange(1, O.R_ORDER)))
        q = O.pt_to_affine(O.G2, O.pt_multiply_unsafe(O.G2, O.G2_BASE, rng.randrange(1, O.R_ORDER)))
        g1 = p[0].to_bytes(48, "big") + p[1].to_bytes(48, "big")
        g2 = b"".join(v.to_bytes(48, "big") for v in (q[0][0], q[0][1], q[1][0], q[1][1]))
        f = O.miller_loop(O.calc_pairing_precomputes(*q), p)
        assert C.pairing_batch(g1, g2, 1, False) == O.fp12_to_bytes(f)
