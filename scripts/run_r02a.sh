set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv
free -g | head -2; nproc
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r02a_pytest.log; tail -5 gpurun_out/r02a_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; tail -c 1500 gpurun_out/r02a_bench.json
timeout 900 python bench.py --impl reference-cuda --steps 10 > gpurun_out/r02a_refcuda.json 2> gpurun_out/r02a_refcuda.err; cat gpurun_out/r02a_refcuda.json
timeout 900 python bench.py --config host --steps 10 > gpurun_out/r02a_host.json 2> gpurun_out/r02a_host.err; cat gpurun_out/r02a_host.json
timeout 600 python bench.py --config testil256 --steps 10 --warmup 3 > gpurun_out/r02a_testil.json 2> gpurun_out/r02a_testil.err; cat gpurun_out/r02a_testil.json
