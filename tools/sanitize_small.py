#!/usr/bin/env python3
"""Small end-to-end pass over every device program family (pairing, Miller product with 1..4 items per lane, sign, aggregate,
verifyBatch) for `compute-sanitizer --tool memcheck python tools/sanitize_small.py`; checks results against the fixtures."""
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import noble_bls12_381_b200 as bls  # noqa: E402
from noble_bls12_381_b200 import synth  # noqa: E402

eng = bls.engine()
dst = b"BLS_SIG_BLS12381G2_XMD:SHA-256_SSWU_RO_NUL_"
n = 40
g1, g2 = synth.multiples_wire(n)
gold = open(os.path.join(ROOT, "tests", "golden", "pairing_kilic_1000.bin"), "rb").read()
assert eng.pairing_batch(g1, g2, n, True) == gold[: 576 * n]
prods = []
for k in (1, 2, 3, 4):
    eng.set_option("pairs_per_lane", k)
    prods.append(eng.miller_product(g1, g2, n, False))
eng.set_option("pairs_per_lane", 3)
assert prods[0] == prods[1] == prods[2] == prods[3]
kats = [l.split(":") for l in open(os.path.join(ROOT, "tests", "golden", "sign_g2_vectors.txt")).read().split("\n") if l][:33]
r = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
msgs = [bytes.fromhex(m) for _, m, _ in kats]
sks = [int(sk, 16) % r for sk, _, _ in kats]
sigs = eng.sign_batch(b"".join(k.to_bytes(32, "big") for k in sks), msgs, dst)
assert all(sigs[96 * i: 96 * i + 96].hex() == kats[i][2].strip().lower() for i in range(len(kats)))
# public keys sk*G1 through the device scalar multiplication + compression path of the API mirror
from noble_bls12_381_b200 import api  # noqa: E402
pks = [api.getPublicKey(k.to_bytes(32, "big")) for k in sks[:5]]
agg, st = eng.aggregate_g2(sigs[: 96 * 5], 5)
v, st = eng.verify_batch(agg, msgs[:5], b"".join(pks), dst)
assert v == 1, (v, st)
bad = list(msgs[:5])
bad[2] = bad[2] + b"!"
assert eng.verify_batch(agg, bad, b"".join(pks), dst)[0] == 0
print("sanitize_small OK")
