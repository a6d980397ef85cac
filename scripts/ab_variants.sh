#!/bin/bash
# A/B of variant builds of the library on one B200 (development aid): the bench step with every variant
# (python -m sw4lite_b200.build -o variants/<name>.so -D...), then the fast-path parity tests with the product library.
# usage: scripts/ab_variants.sh <tag> [--test] <variant.so ...>     ("default" = sw4lite_b200/libsw4b200.so)
tag=$1; shift
test=0; if [ "$1" == "--test" ]; then test=1; shift; fi
mkdir -p gpurun_out; log=gpurun_out/abv_$tag.log; : > $log
for v in "$@"; do
  n=$(basename $v .so)
  if [ "$v" == "default" ]; then unset SW4B200_LIB; else export SW4B200_LIB=$PWD/$v; fi
  timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/abv_${tag}_$n.json 2>> $log
  python - <<PY >> $log
import json
try:
    d = json.loads(open("gpurun_out/abv_${tag}_$n.json").read().strip().splitlines()[-1])
    print("$n: %.2f ms/step %.2f Gpts/s " % (d["ms_per_step"], d["value"]), {k: round(x.get("ms_per_step", x.get("plain_frac", 0)), 2) for k, x in d["kernels"].items()}, d["clocks"]["sm_mhz"])
except Exception as e:
    print("$n: no result", e)
PY
done
unset SW4B200_LIB
if [ $test == 1 ]; then
  timeout 900 python -m pytest tests/test_gpu_fastpath.py tests/test_gpu_step.py -x -q -m gpu 2>&1 | tail -3 >> $log
fi
cat $log
