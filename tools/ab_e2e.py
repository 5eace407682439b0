#!/usr/bin/env python3
"""A/B of the chunked copy / compute pipeline of bls381_pairing_batch (host buffers in, host buffers out):
    tools/ab_e2e.py [n]   ->  pairings/s end to end with pipeline_copies = 0 / 1, outputs compared byte for byte"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import noble_bls12_381_b200 as bls  # noqa: E402
from noble_bls12_381_b200 import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
eng = bls.engine()
g1, g2 = synth.random_pairs_wire(eng, n, seed=0xB200, prefix=1000)
h1 = torch.frombuffer(bytearray(g1), dtype=torch.uint8).pin_memory()
h2 = torch.frombuffer(bytearray(g2), dtype=torch.uint8).pin_memory()
outs = []
for mode in (0, 1, 0, 1):
    eng.set_option("pipeline_copies", mode)
    hout = torch.zeros(576 * n, dtype=torch.uint8).pin_memory()
    for _ in range(2):
        assert eng.lib.bls381_pairing_batch(h1.data_ptr(), h2.data_ptr(), n, 1, hout.data_ptr(), None) == 0
    reps = 8
    t0 = time.perf_counter()
    for _ in range(reps):
        assert eng.lib.bls381_pairing_batch(h1.data_ptr(), h2.data_ptr(), n, 1, hout.data_ptr(), None) == 0
    dt = (time.perf_counter() - t0) / reps
    outs.append(bytes(hout.numpy().tobytes()))
    print(f"pipeline_copies={mode}: {dt * 1e3:.3f} ms per call, {n / dt:,.0f} pairings/s end to end, kernel span {eng.last_kernel_ms():.3f} ms", flush=True)
print("outputs identical:", all(o == outs[0] for o in outs))

