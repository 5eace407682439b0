#!/usr/bin/env python3
"""A/B timing of the sign path (hash-to-curve + scalar multiplication) for alternative program directories:
    tools/ab_sign.py [n] dir1 dir2 ...   -> kernel ms per directory (first 8 signatures checked against the KAT file)"""
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import noble_bls12_381_b200 as bls  # noqa: E402

n = int(sys.argv[1])
dst = b"BLS_SIG_BLS12381G2_XMD:SHA-256_SSWU_RO_NUL_"
kats = [l.split(":") for l in open(os.path.join(ROOT, "tests", "golden", "sign_g2_vectors.txt")).read().split("\n") if l][:8]
msgs = [bytes.fromhex(m) for _, m, _ in kats] + [hashlib.sha256(b"m%d" % i).digest() for i in range(n - len(kats))]
sks = b"".join((int(sk, 16) % 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001).to_bytes(32, "big") for sk, _, _ in kats) + b"".join((i + 1).to_bytes(32, "big") for i in range(n - len(kats)))
for d in sys.argv[2:]:
    eng = bls.Engine(0, d)
    eng.sign_batch(sks[: 32 * 64], msgs[:64], dst)
    best = None
    for _ in range(2):
        sigs = eng.sign_batch(sks, msgs, dst)
        ms = eng.last_kernel_ms()
        best = ms if best is None else min(best, ms)
    ok = all(sigs[96 * i: 96 * i + 96].hex() == kats[i][2].strip().lower() for i in range(len(kats)))
    print(f"{d}: sign kernel {best:.2f} ms for {n} = {n / best * 1e3:.0f} sigs/s, KATs {'ok' if ok else 'MISMATCH'}", flush=True)
    del eng
