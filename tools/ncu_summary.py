#!/usr/bin/env python3
"""Summarises an `ncu --set full` report of tower-VM launches: tools/ncu_summary.py <report.ncu-rep> [launch log]
The launch log is the stderr of the profiled process run with BLS381_B200_LOG_LAUNCHES=1: its i-th `[vm_run]` line names the
program of the i-th vm_kernel launch.  Prints one row per launch (program, duration, fmaheavy pipe, issue slots, registers,
shared memory, DRAM bytes, top stall reasons) plus the issued-IMAD fraction when the program image is found."""
import csv
import io
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rep = sys.argv[1]
names = []
if len(sys.argv) > 2 and os.path.exists(sys.argv[2]):
    names = [(m.group(1), int(m.group(2))) for m in re.finditer(r"\[vm_run\] (\S+) n=(\d+)", open(sys.argv[2]).read())]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
col = {k: i for i, k in enumerate(hdr)}
K = {"dur": "gpu__time_duration.sum", "fma": "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
     "issue": "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "regs": "launch__registers_per_thread",
     "smem": "launch__shared_mem_per_block_dynamic", "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum",
     "warps": "sm__warps_active.avg.pct_of_peak_sustained_active", "grid": "launch__grid_size", "block": "launch__block_size"}
units = rows[1]
import bench  # issued_imad_per_item
print("%-3s %-22s %8s %5s %9s %7s %7s %5s %8s %12s %12s  %s" % ("#", "program", "items", "grid", "dur", "fmahvy%", "issue%", "regs", "smem KB", "dram rd", "dram wr", "stalls per issue (top 3) / issued IMAD.WIDE fraction of 9.05 T/s"))
vm = 0
row_no = 0
PLAIN = {"swu_g2_kernel": "swu_g2_kernel", "h2g2_tail_kernel": "h2g2_tail_kernel", "sign_kernel": "sign_kernel", "g1_decompress_kernel": "g1_decompress_kernel"}
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    kn = r[col["Kernel Name"]]
    plain = next((k for k in PLAIN if k in kn), None)
    if "vm_kernel" not in kn and plain is None:
        continue
    g = lambda k: r[col[K[k]]] if K[k] in col else "?"
    if plain is None:
        prog, items = names[vm] if vm < len(names) else ("?", 0)
        vm += 1
    else:  # hand-written per-item kernels: a thread per item (swu: per field element, two per message)
        prog = plain
        try:
            items = int(g("grid").replace(",", "")) * int(g("block").replace(",", ""))
        except ValueError:
            items = 0
        if plain == "swu_g2_kernel":
            items //= 2
    row_no += 1
    stalls = sorted(((float(r[i]), h.split("issue_stalled_")[1].split("_per_issue")[0]) for h, i in col.items()
                     if "smsp__average_warps_issue_stalled" in h and "per_issue_active" in h and r[i] not in ("", "n/a")), reverse=True)
    top = ", ".join("%s %.2f" % (n, v) for v, n in stalls[:4] if n != "selected")
    frac = ""
    img = os.path.join(ROOT, "noble_bls12_381_b200", "programs", prog + ".b2vm")
    try:
        dur_s = float(g("dur")) * {"ms": 1e-3, "us": 1e-6, "s": 1.0, "ns": 1e-9}[units[col[K["dur"]]]]
        if plain is not None and PLAIN[plain] and items:
            iw, _, _ = bench.kernel_imad_per_item(PLAIN[plain])
            frac = " | issued %.1f %% (items <= threads launched)" % (100.0 * iw * items / dur_s / 9.05e12)
        elif os.path.exists(img) and items:
            iw, _, _ = bench.issued_imad_per_item(img)
            frac = " | issued %.1f %%" % (100.0 * iw * (items / 1.0) / dur_s / 9.05e12)
    except Exception as e:  # noqa: BLE001
        frac = " | (%s)" % e
    print("%-3d %-22s %8d %5s %7s%-2s %7s %7s %5s %8s %10s%-2s %10s%-2s  %s%s" % (
        row_no, prog, items, g("grid"), g("dur"), units[col[K["dur"]]], g("fma")[:6], g("issue")[:6], g("regs"), g("smem")[:7], g("rd")[:9], units[col[K["rd"]]],
        g("wr")[:9], units[col[K["wr"]]], top, frac))
