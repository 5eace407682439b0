#!/usr/bin/env python3
"""In-process multi-GPU verifyBatch (bls381_init_devices + bls381_verify_batch_multi): ONE process, one context and one host
thread per GPU, the path the Node addon uses.  tools/bench_multi.py [total_sigs] [reps] -> one JSON line with the rate at
every device count 1, 2, 4, ... up to the visible GPUs (strong scaling: the batch is fixed)."""
import ctypes
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from noble_bls12_381_b200 import synth  # noqa: E402
from noble_bls12_381_b200._lib import Engine  # noqa: E402

total = int(sys.argv[1]) if len(sys.argv) > 1 else 2097152
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dst = b"BLS_SIG_BLS12381G2_XMD:SHA-256_SSWU_RO_NUL_"
ndev = torch.cuda.device_count()
eng = Engine(0)
sks, msgs = synth.signing_inputs(total, 0xBA7C4)
packed, off = eng._pack(msgs)
pks = eng.get_public_key_batch(sks)
sigs = eng.sign_batch(sks, msgs, dst)
agg, _ = eng.aggregate_g2(sigs, total)
st = (ctypes.c_int32 * (total + 1))()
v = ctypes.c_int(0)
rows = []
w = 1
while w <= ndev:
    assert eng.init_devices((1 << w) - 1) >= w
    # contexts are added, never removed: run with exactly w devices by masking is not possible once more exist, so the
    # device counts are visited in increasing order
    for _ in range(2):  # warm-up: program load + staging growth on the new devices
        rc = eng.lib.bls381_verify_batch_multi(agg, packed, off, pks, total, dst, len(dst), ctypes.byref(v), st)
        assert rc == 0, eng.lib.bls381_last_error()
    assert v.value == 1
    t0 = time.perf_counter()
    for _ in range(reps):
        rc = eng.lib.bls381_verify_batch_multi(agg, packed, off, pks, total, dst, len(dst), ctypes.byref(v), st)
        assert rc == 0 and v.value == 1
    dt = (time.perf_counter() - t0) / reps
    bad = bytearray(packed)
    bad[32 * (total - 1)] ^= 1
    rc = eng.lib.bls381_verify_batch_multi(agg, bytes(bad), off, pks, total, dst, len(dst), ctypes.byref(v), st)
    rows.append({"devices": w, "sigs_per_s": total / dt, "ms": dt * 1e3, "transport": eng.multi_transport(), "flipped_byte_verdict": v.value})
    w *= 2
print(json.dumps({"metric": "verifyBatch sigs/s, in-process multi-GPU (bls381_verify_batch_multi), strong scaling", "total_sigs": total,
                  "reps": reps, "rows": rows}))
