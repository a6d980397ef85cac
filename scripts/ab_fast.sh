#!/bin/bash
# A/B of the interior-kernel generations on one B200 (development aid): parity tests with the candidate, then the
# bench step with every variant.  usage: scripts/ab_fast.sh <tag> <candidate> <variants...>
tag=$1; cand=$2; shift 2
mkdir -p gpurun_out
echo "== parity tests with SW4B200_FAST_GEN=$cand" > gpurun_out/ab_$tag.log
SW4B200_FAST_GEN=$cand timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_step.py -x -q -m gpu >> gpurun_out/ab_$tag.log 2>&1
echo "exit $?" >> gpurun_out/ab_$tag.log
for v in "$@"; do
  echo "== bench SW4B200_FAST_GEN=$v" >> gpurun_out/ab_$tag.log
  SW4B200_FAST_GEN=$v timeout 400 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ab_${tag}_$v.json 2>> gpurun_out/ab_$tag.log
  echo "exit $?" >> gpurun_out/ab_$tag.log
  python - <<PY >> gpurun_out/ab_$tag.log
import json
try:
    d = json.loads(open("gpurun_out/ab_${tag}_$v.json").read().strip().splitlines()[-1])
    print("gen $v: %.2f ms/step %.2f Gpts/s " % (d["ms_per_step"], d["value"]), {k: round(x["ms_per_step"], 2) for k, x in d["kernels"].items()}, d["clocks"])
except Exception as e:
    print("gen $v: no result", e)
PY
done
cat gpurun_out/ab_$tag.log
