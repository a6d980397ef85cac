#!/bin/bash
# bench step under several environment settings (development aid).  usage: scripts/ab_env.sh <tag> "VAR=val ..." ...
tag=$1; shift
mkdir -p gpurun_out; : > gpurun_out/abenv_$tag.log
i=0
for envs in "$@"; do
  i=$((i+1))
  env $envs timeout 400 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/abenv_${tag}_$i.json 2>> gpurun_out/abenv_$tag.log
  python - <<PY >> gpurun_out/abenv_$tag.log
import json
try:
    d = json.loads(open("gpurun_out/abenv_${tag}_$i.json").read().strip().splitlines()[-1])
    print("$envs: %.2f ms/step %.2f Gpts/s " % (d["ms_per_step"], d["value"]), {k: round(x["ms_per_step"], 2) for k, x in d["kernels"].items()})
except Exception as e:
    print("$envs: no result", e)
PY
done
cat gpurun_out/abenv_$tag.log
