set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR scripts/check_slabs_multigpu.py 48 70 > gpurun_out/r02e_check_even.log 2>&1; grep -i "rank" gpurun_out/r02e_check_even.log | tail -4
timeout 300 $TR scripts/check_slabs_multigpu.py 48 71 > gpurun_out/r02e_check_odd.log 2>&1; grep -i "rank" gpurun_out/r02e_check_odd.log | tail -4
timeout 300 $TR scripts/check_topo_multigpu.py 8 > gpurun_out/r02e_check_topo.log 2>&1; grep -i "rank" gpurun_out/r02e_check_topo.log | tail -4
timeout 600 $TR bench.py --gpus 2 --config strong --steps 5 --warmup 3 > gpurun_out/r02e_strong_n2.json 2> gpurun_out/r02e_strong_n2.err; python -c "
import json; d=json.load(open('gpurun_out/r02e_strong_n2.json')); print(d['value'], d['ms_per_step'], d['kernels'])"
tail -n 3 gpurun_out/r02e_strong_n2.err
