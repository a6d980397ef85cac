set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for c in 8 4 16; do
SW4B200_NCCL_MAX_CTAS=$c timeout 600 $TR bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/r02l_weak_n2_c$c.json 2> gpurun_out/r02l_weak_n2.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r02l_weak_n2_c$c.json") if l.startswith("{")][-1])
print("maxCTAs $c:", d["value"], d["ms_per_step"]); print({k:(round(v["ms_per_step"],3),v["launches_per_step"]) for k,v in d["kernels"].items()})
PY
done
SW4B200_NCCL_MAX_CTAS=8 timeout 600 $TR bench.py --gpus 2 --config strong --nz-total 64 --steps 6 --warmup 3 > gpurun_out/r02l_strong64_n2.json 2> gpurun_out/r02l_strong64_n2.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r02l_strong64_n2.json") if l.startswith("{")][-1])
print("thin slabs 32 planes/GPU:", d["value"], d["ms_per_step"]); print({k:(round(v["ms_per_step"],3),v["launches_per_step"]) for k,v in d["kernels"].items()})
PY
tail -n 3 gpurun_out/r02l_weak_n2.err
