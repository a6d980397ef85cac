set -x
timeout 900 python -m pytest tests/test_gpu_fastpath.py tests/test_gpu_step.py tests/test_gpu_kernels.py -m gpu -q 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02s_bench_n1.json 2> gpurun_out/r02s_bench_n1.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r02s_bench_n1.json") if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"]); print({k:(round(v["ms_per_step"],3),v["launches_per_step"]) for k,v in d["kernels"].items()})
PY
