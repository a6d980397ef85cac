#!/bin/bash
# single-GPU measurements of the round (run under gpurun):  bash scripts/run_final_n1.sh TAG
TAG=${1:-r02p}
set -x
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
timeout 900 python bench.py --impl reference --steps 6 --warmup 2 > gpurun_out/${TAG}_reference_cpu.json 2> gpurun_out/${TAG}_reference_cpu.err
timeout 900 python bench.py --impl reference-cuda --steps 10 > gpurun_out/${TAG}_refcuda.json 2> gpurun_out/${TAG}_refcuda.err
timeout 900 python bench.py --config host --steps 10 > gpurun_out/${TAG}_host.json 2> gpurun_out/${TAG}_host.err
timeout 600 python bench.py --config testil256 --steps 10 --warmup 3 > gpurun_out/${TAG}_testil256.json 2> gpurun_out/${TAG}_testil256.err
timeout 600 python bench.py --config strong --steps 10 --warmup 3 > gpurun_out/${TAG}_strong_n1.json 2> gpurun_out/${TAG}_strong_n1.err
timeout 600 python bench.py --config loh1-h100 > gpurun_out/${TAG}_loh1_h100_n1.json 2> gpurun_out/${TAG}_loh1_h100_n1.err
timeout 900 python bench.py --config loh1-h50 > gpurun_out/${TAG}_loh1_h50_n1.json 2> gpurun_out/${TAG}_loh1_h50_n1.err
timeout 600 host/run_slabs.sh 1 --nx 2048 --ny 2048 --nzl 128 --steps 10 --warmup 3 > gpurun_out/${TAG}_cxx_weak_n1.json 2> gpurun_out/${TAG}_cxx_weak_n1.err
timeout 600 host/run_slabs.sh 1 --nx 2048 --ny 2048 --nz-total 256 --steps 10 --warmup 3 > gpurun_out/${TAG}_cxx_strong_n1.json 2> gpurun_out/${TAG}_cxx_strong_n1.err
for f in bench_n1 reference_cpu refcuda host testil256 strong_n1 loh1_h100_n1 loh1_h50_n1 cxx_weak_n1 cxx_strong_n1; do python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/${TAG}_$f.json") if l.startswith("{")][-1])
    print("$f", round(d.get("value",0),4), d.get("unit"), round(d.get("ms_per_step",0),3), "ms/step", d.get("station_ok",""), d.get("unavailable",""))
except Exception as e:
    print("$f FAILED", e)
PY
done
