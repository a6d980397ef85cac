set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --config loh1-h50 > gpurun_out/r02h_loh1_h50_n2.json 2> gpurun_out/r02h_loh1_h50_n2.err; cut -c1-700 gpurun_out/r02h_loh1_h50_n2.json; tail -n 3 gpurun_out/r02h_loh1_h50_n2.err
timeout 600 $TR bench.py --gpus 2 --config strong --steps 5 --warmup 3 > gpurun_out/r02h_strong_n2.json 2> gpurun_out/r02h_strong_n2.err; cut -c1-300 gpurun_out/r02h_strong_n2.json
timeout 300 host/run_slabs.sh 1 --nx 512 --ny 512 --nz-total 128 --steps 5 --warmup 2 > gpurun_out/r02h_cxx_n1.json 2> gpurun_out/r02h_cxx_n1.err; cat gpurun_out/r02h_cxx_n1.json; tail -n 3 gpurun_out/r02h_cxx_n1.err
timeout 300 host/run_slabs.sh 2 --nx 512 --ny 512 --nz-total 128 --steps 5 --warmup 2 > gpurun_out/r02h_cxx_n2.json 2> gpurun_out/r02h_cxx_n2.err; cat gpurun_out/r02h_cxx_n2.json; tail -n 3 gpurun_out/r02h_cxx_n2.err
timeout 600 host/run_slabs.sh 2 --nx 2048 --ny 2048 --nzl 128 --steps 5 --warmup 3 > gpurun_out/r02h_cxx_weak_n2.json 2> gpurun_out/r02h_cxx_weak_n2.err; cat gpurun_out/r02h_cxx_weak_n2.json; tail -n 3 gpurun_out/r02h_cxx_weak_n2.err
timeout 600 $TR scripts/check_topo_multigpu.py 20 gaussianHill-rev.in 0 > gpurun_out/r02h_topo_n2.log 2>&1; grep "rank\|TIMING" gpurun_out/r02h_topo_n2.log
timeout 600 $TR scripts/check_topo_multigpu.py 20 gaussianHill-rev.in 1 > gpurun_out/r02h_topo_n2_bal.log 2>&1; grep "rank\|TIMING" gpurun_out/r02h_topo_n2_bal.log
