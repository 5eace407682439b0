"""Compiles the tower-VM programs into the binary images the C-ABI library loads (*.b2vm).

image := header (8 x u32: magic 'B2VM', version, warps, nrec, nconst, nslots, nfar, staged-buffer mask | format bit 31)
         constant table (nconst x 12 u32, Montgomery form)    program (warps x nrec x 64 u32)
"""
from __future__ import annotations

import os
import struct
import sys

from . import curves, tower

MAGIC = 0x4D563242  # 'B2VM'
VERSION = 2  # 64-word records, two-output records
DEFAULT_WARPS = 8
DEFAULT_SLOTS = 66  # + up to 9 KB of TMA input staging per CTA, two CTAs per SM


def image(b) -> bytes:
    prog, nrec = b.encode()
    staged = 0  # bit mask of the wire-format input buffers (TMA-staged into shared memory per batch)
    for op in b.ops:
        for x, y in op.terms:
            for o in (x, y):
                if o.flags & 2:
                    staged |= 1 << o.gl[0]
        if op.kind == "bit":
            staged |= 1 << op.bit[0]
    for op in b.ops:  # wire-format operands of two-output records (entries of the H_SINGLE form are op.terms above)
        for g in op.groups.values():
            for ent in g:
                for o in ent:
                    if o is not None and o.flags & 2:
                        staged |= 1 << o.gl[0]
    if b.dual:
        staged |= 1 << 31  # two-output record format: the library picks the MODE_MAC2 interpreter
    hdr = struct.pack("<8I", MAGIC, VERSION, b.warps, nrec, len(b.consts), b.nslots, b.nfar, staged)
    return hdr + b.const_table() + prog


# ingest programs are serial Fp2 chains: fewer warps per CTA, more CTAs per SM
INGEST_WARPS = {"g1_decompress": 2, "g2_decompress": 4, "hash_to_g2": 4, "sign": 4, "g1_sum_affine": 4, "g1_sum_proj": 4,
                "g2_sum_affine": 4, "g2_sum_proj": 4, "g1_compress": 2, "g2_compress": 2, "g1_validate": 2, "g2_validate": 4,
                "g2_scalar_mul": 4, "g1_scalar_mul": 4, "g1_from_uncompressed": 2, "g2_from_uncompressed": 4, "hash_to_g1": 2,
                "h2g2_tail": 4, "sign_tail": 4}
# shared-memory slots per CTA: the serial ingest programs run more CTAs per SM with fewer slots each (cold values go to the
# L2-resident far slots); measured on a B200: hash_to_g2 / sign / g2_decompress 4 CTAs x 4 warps instead of 2 x 4
# (sign +13 %, verifyBatch +7 %), g1_decompress 8 CTAs x 2 warps instead of 3 x 2 (verifyBatch +9 %)
# the scalar-multiplication window tables (16 points) are cold: they live in far slots
INGEST_SLOTS = {"hash_to_g1": 16, "hash_to_g2": 28, "sign": 28, "h2g2_tail": 28, "sign_tail": 28, "g2_decompress": 33, "g1_decompress": 16, "g2_scalar_mul": 33, "g1_scalar_mul": 24}
ALL_PROGRAMS = dict(tower.PROGRAMS)
ALL_PROGRAMS.update(curves.PROGRAMS)


def compile_program(name: str, warps=None, nslots=None, dual=None):
    """dual: None = the build default (builder.DUAL_DEFAULT); True / False = two-output / one-output record format."""
    from . import builder as _b
    if warps is None:
        warps = INGEST_WARPS.get(name, DEFAULT_WARPS)
    if nslots is None:
        nslots = INGEST_SLOTS.get(name, DEFAULT_SLOTS)
    prev = _b.DUAL_DEFAULT
    if dual is not None:
        _b.DUAL_DEFAULT = dual
    try:
        b = ALL_PROGRAMS[name](warps)
    finally:
        _b.DUAL_DEFAULT = prev
    b.schedule()
    b.allocate(nslots)
    b.check_hazards()
    return b


def build_all(outdir: str, warps=None, nslots=None, names=None, verbose=True):
    os.makedirs(outdir, exist_ok=True)
    for name in names or ALL_PROGRAMS:
        b = compile_program(name, warps, nslots)
        path = os.path.join(outdir, name + ".b2vm")
        with open(path, "wb") as f:
            f.write(image(b))
        if verbose:
            st = b.sched_stats
            print(f"{name:16s} ops={st['ops']:6d} steps={st['steps']:5d} cost={st['total_cost']:9.0f} "
                  f"sched_eff={st['efficiency']:.3f} slots={b.peak_slots} far={b.nfar}")


if __name__ == "__main__":
    build_all(sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(__file__)), "programs"))
