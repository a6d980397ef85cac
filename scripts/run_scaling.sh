#!/bin/bash
# scaling measurements on N GPUs of one box (run under gpurun --gpus N):  bash scripts/run_scaling.sh N TAG
N=${1:-2}; TAG=${2:-r02}
set -x
mkdir -p gpurun_out
if [ "$N" = "1" ]; then TR="python"; else TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"; fi
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${TAG}_weak_n$N.json 2> gpurun_out/${TAG}_weak_n$N.err
timeout 600 $TR bench.py --gpus $N --config strong --steps 10 --warmup 3 > gpurun_out/${TAG}_strong_n$N.json 2> gpurun_out/${TAG}_strong_n$N.err
timeout 600 $TR bench.py --gpus $N --config loh1-h50 > gpurun_out/${TAG}_loh1_h50_n$N.json 2> gpurun_out/${TAG}_loh1_h50_n$N.err
timeout 600 host/run_slabs.sh $N --nx 2048 --ny 2048 --nzl 128 --steps 10 --warmup 3 > gpurun_out/${TAG}_cxx_weak_n$N.json 2> gpurun_out/${TAG}_cxx_weak_n$N.err
[ "$N" = "8" ] || timeout 600 host/run_slabs.sh $N --nx 2048 --ny 2048 --nz-total 256 --steps 10 --warmup 3 > gpurun_out/${TAG}_cxx_strong_n$N.json 2> gpurun_out/${TAG}_cxx_strong_n$N.err
if [ "$N" != "1" ]; then
timeout 600 $TR scripts/check_topo_multigpu.py 42 gaussianHill-rev.in 0 > gpurun_out/${TAG}_topo_n$N.log 2>&1
timeout 600 $TR scripts/check_topo_multigpu.py 42 gaussianHill-rev.in 1 > gpurun_out/${TAG}_topo_bal_n$N.log 2>&1
timeout 300 $TR scripts/check_slabs_multigpu.py $((24*N)) 71 > gpurun_out/${TAG}_check_odd_n$N.log 2>&1
fi
for f in weak strong loh1_h50 cxx_weak cxx_strong; do python - <<PY
import json
try:
    lines=[l for l in open("gpurun_out/${TAG}_${f}_n$N.json") if l.startswith("{")]
    d=json.loads(lines[-1]); print("$f N=$N", round(d["value"],3), "Gpts/s", round(d["ms_per_step"],3), "ms/step", d.get("station_ok",""), d.get("checksum",""))
except Exception as e:
    print("$f N=$N FAILED", e)
PY
done
grep -h "TIMING\|DIFFERENT" gpurun_out/${TAG}_topo*_n$N.log gpurun_out/${TAG}_check_odd_n$N.log 2>/dev/null | cut -c1-400
grep -c "bit-identical" gpurun_out/${TAG}_topo*_n$N.log gpurun_out/${TAG}_check_odd_n$N.log 2>/dev/null
