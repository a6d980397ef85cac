set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR scripts/check_slabs_multigpu.py 48 70 > gpurun_out/r02n_check_even.log 2>&1; grep -i "rank" gpurun_out/r02n_check_even.log | tail -4; tail -n 5 gpurun_out/r02n_check_even.log | cut -c1-300
timeout 300 $TR scripts/check_slabs_multigpu.py 48 71 > gpurun_out/r02n_check_odd.log 2>&1; grep -i "rank" gpurun_out/r02n_check_odd.log | tail -4
timeout 600 $TR bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/r02n_weak_n2.json 2> gpurun_out/r02n_weak_n2.err
timeout 600 $TR bench.py --gpus 2 --config strong --nz-total 64 --steps 6 --warmup 3 > gpurun_out/r02n_strong64_n2.json 2> gpurun_out/r02n_strong64_n2.err
for f in weak_n2 strong64_n2; do python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r02n_$f.json") if l.startswith("{")][-1])
print("$f", d["value"], d["ms_per_step"], d["config"].get("exchange")); print({k:(round(v["ms_per_step"],3),v["launches_per_step"]) for k,v in d["kernels"].items()})
PY
done
tail -n 4 gpurun_out/r02n_weak_n2.err | cut -c1-300
