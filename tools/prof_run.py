#!/usr/bin/env python3
"""Runs one device program family a few times on a small synthetic batch (for `ncu`): tools/prof_run.py <what> [n] [program_dir]
what: pairing | sign | verify | aggregate"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import noble_bls12_381_b200 as bls  # noqa: E402
from noble_bls12_381_b200 import synth  # noqa: E402
from noble_bls12_381_b200._lib import Engine  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "pairing"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 9472
pdir = sys.argv[3] if len(sys.argv) > 3 else None
eng = Engine(0, pdir)
dst = b"BLS_SIG_BLS12381G2_XMD:SHA-256_SSWU_RO_NUL_"
reps = 3
if what == "pairing":
    g1, g2 = synth.multiples_wire(n)
    for _ in range(reps):
        out = eng.pairing_batch(g1, g2, n, True)
    gold = open(os.path.join(ROOT, "tests", "golden", "pairing_kilic_1000.bin"), "rb").read()
    k = min(n, 1000)
    assert out[: 576 * k] == gold[: 576 * k]
    print("pairing", n, "kernel ms", eng.last_kernel_ms())
else:
    import hashlib
    sks = b"".join(hashlib.sha256(b"sk%d" % i).digest() for i in range(n))
    msgs = [hashlib.sha256(b"msg%d" % i).digest() for i in range(n)]
    for _ in range(reps if what == "sign" else 1):
        sigs = eng.sign_batch(sks, msgs, dst)
    print("sign", n, "kernel ms", eng.last_kernel_ms())
    if what in ("verify", "aggregate"):
        agg, st = eng.aggregate_g2(sigs, n)
        print("aggregate", n, "kernel ms", eng.last_kernel_ms())
    if what == "verify":
        from noble_bls12_381_b200 import api
        pks = eng.get_public_key_batch(sks) if hasattr(eng, "get_public_key_batch") else b"".join(api.getPublicKey(sks[32 * i: 32 * i + 32]) for i in range(n))
        for _ in range(reps):
            v, st = eng.verify_batch(agg, msgs, pks, dst)
        assert v == 1, v
        print("verify", n, "kernel ms", eng.last_kernel_ms())
