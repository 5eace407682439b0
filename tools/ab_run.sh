#!/bin/sh
# One bench.py run for an A/B comparison inside a gpurun call: tools/ab_run.sh <tag> [ENV=VAL ...] -- [bench args]
# The JSON line goes to gpurun_out/<tag>.json; a one-line summary is printed.
tag="$1"; shift
envs=""
while [ "$1" != "--" ] && [ $# -gt 0 ]; do envs="$envs $1"; shift; done
[ "$1" = "--" ] && shift
env $envs python bench.py --no-cpu-baseline "$@" > "gpurun_out/$tag.json" 2> "gpurun_out/$tag.err"
python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/{tag}.json").read().strip().split("\n")[-1])
    vb = d.get("verify_batch") or {}
    print(tag, "pairings/s", round(d["value"]), "parity", d.get("parity_first_1000_vs_reference_fixtures", d.get("parity")),
          "verify", round(vb.get("value", 0)), "sign", round((vb.get("sign") or {}).get("value", 0)), "ok", vb.get("verdict_true"))
except Exception as e:
    print(tag, "ERR", e, open(f"gpurun_out/{tag}.err").read()[-600:])
PY
