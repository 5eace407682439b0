#!/usr/bin/env python
"""Summarise a tower-VM trace (tools/build_trace_lib.sh): per-warp wait / phase totals, per-op-class means."""
import collections
import struct
import sys

import numpy as np


def main(trace_path, program_path):
    raw = np.fromfile(trace_path, dtype=np.uint32)
    ctas, W, nrec, nf = (int(x) for x in raw[:4])
    t = raw[4:].reshape(ctas, W, nrec, nf).astype(np.int64)
    img = open(program_path, "rb").read()
    _, _, warps, nr, nconst, _, _, _ = struct.unpack_from("<8I", img, 0)
    assert warps == W and nr == nrec
    base = 32 + nconst * 48
    names = ["operand loads", "multiply-accumulate", "reduction", "epilogue+correct", "store"]
    for c in range(ctas):
        start, end = t[c, :, :, 0].min(), t[c, :, :, 2].max()
        span = (end - start) & 0xFFFFFFFF
        print(f"CTA {c}: batch span {span} cycles")
        for w in range(W):
            wait = ((t[c, w, :, 1] - t[c, w, :, 0]) & 0xFFFFFFFF).sum()
            ex = ((t[c, w, :, 2] - t[c, w, :, 1]) & 0xFFFFFFFF).sum()
            ph = t[c, w, :, 3:8].sum(axis=0) if nf >= 8 else []
            gaps = span - wait - ex
            print(f"  warp {w}: wait {wait / span:6.1%} exec {ex / span:6.1%} between-records {gaps / span:6.1%} | " +
                  " ".join(f"{n} {v / span:5.1%}" for n, v in zip(names, ph)))
    cls = collections.defaultdict(list)
    for w in range(W):
        for r in range(nrec):
            hdr, = struct.unpack_from("<I", img, base + (w * nrec + r) * 256)
            key = (hdr & 0xFF, (hdr >> 16) & 0xF, (hdr >> 20) & 3, (hdr >> 22) & 7)
            cls[key].append(t[0, w, r])
    print("op class (opcode, T, E, ncorr): count, mean exec, mean phases")
    for k in sorted(cls):
        v = np.array(cls[k])
        ex = ((v[:, 2] - v[:, 1]) & 0xFFFFFFFF).mean()
        ph = v[:, 3:8].mean(axis=0) if nf >= 8 else []
        print(" ", k, len(v), f"{ex:8.0f}", " ".join(f"{x:7.0f}" for x in ph))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
