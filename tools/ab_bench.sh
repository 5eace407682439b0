#!/bin/sh
# A/B helper for one gpurun call: tools/ab_bench.sh <tag> [ENV=VAL ...] -- [bench args]; result line -> gpurun_out/<tag>.json
tag="$1"; shift
envs=""
while [ "$1" != "--" ] && [ $# -gt 0 ]; do envs="$envs $1"; shift; done
[ "$1" = "--" ] && shift
env $envs python bench.py --verify-n 0 --no-cpu-baseline --steps 10 "$@" > "gpurun_out/$tag.json" 2> "gpurun_out/$tag.err"
python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/{tag}.json").read().strip().split("\n")[-1])
    print(tag, round(d["value"]), "parity", d["parity_first_1000_vs_reference_fixtures"], "issued_frac", round(d["roofline"]["issued_frac_sustained"], 3))
except Exception as e:
    print(tag, "ERR", e, open(f"gpurun_out/{tag}.err").read()[-300:])
PY
