#!/usr/bin/env python3
"""Builds alternative tower-VM program directories for A/B runs on the GPU (programs are data: bench.py --program-dir).

    tools/ab_programs.py <outdir> name=warps:slots[:window] ...      e.g.  gpurun_in/p_w3 sign=4:33:3
Every program not named is copied from the default directory."""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from noble_bls12_381_b200.vmprog import compile as vmcompile  # noqa: E402
from noble_bls12_381_b200.vmprog import curves  # noqa: E402

out = sys.argv[1]
os.makedirs(out, exist_ok=True)
src = os.path.join(ROOT, "noble_bls12_381_b200", "programs")
for f in os.listdir(src):
    shutil.copy(os.path.join(src, f), os.path.join(out, f))
for spec in sys.argv[2:]:
    name, cfg = spec.split("=")
    parts = [int(x) for x in cfg.split(":")]
    if len(parts) > 2:
        curves.SECRET_WINDOW = parts[2]
    b = vmcompile.compile_program(name, parts[0], parts[1])
    with open(os.path.join(out, name + ".b2vm"), "wb") as f:
        f.write(vmcompile.image(b))
    st = b.sched_stats
    print(f"{out}: {name} warps={parts[0]} slots={b.peak_slots} far={b.nfar} cost={st['total_cost']:.0f} makespan={st['critical_cost']:.0f}")
