#!/bin/sh
# A/B helper: verify/sign section only.  tools/ab_verify.sh <tag> [ENV=VAL ...] -- [bench args]
tag="$1"; shift
envs=""
while [ "$1" != "--" ] && [ $# -gt 0 ]; do envs="$envs $1"; shift; done
[ "$1" = "--" ] && shift
env $envs python bench.py --no-cpu-baseline --steps 3 --n 9472 "$@" > "gpurun_out/$tag.json" 2> "gpurun_out/$tag.err"
python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/{tag}.json").read().strip().split("\n")[-1])
    v = d["verify_batch"]
    print(tag, "verify", round(v["value"]), "sign", round(v["sign"]["value"]), "aggregate", round(v["aggregate_signatures"]["value"]), v["verdict_true"])
except Exception as e:
    print(tag, "ERR", e, open(f"gpurun_out/{tag}.err").read()[-300:])
PY
