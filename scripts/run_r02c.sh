set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_host.py -m gpu -x -q -s 2>&1 | tail -40 > gpurun_out/r02c_pytest_host.log; tail -12 gpurun_out/r02c_pytest_host.log
timeout 900 python bench.py --config host --steps 10 > gpurun_out/r02c_host.json 2> gpurun_out/r02c_host.err; cat gpurun_out/r02c_host.json
