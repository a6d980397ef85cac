set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_curvilinear.py tests/test_gpu_host.py -m gpu -q -s 2>&1 | grep -v "^$" | tail -25 | cut -c1-400 > gpurun_out/r02i_pytest.log; cat gpurun_out/r02i_pytest.log
timeout 300 python scripts/quick_bench.py 256 256 2>&1 | tail -8
timeout 600 python bench.py --config strong --steps 5 --warmup 3 > gpurun_out/r02i_strong_n1.json 2> gpurun_out/r02i_strong_n1.err; cut -c1-400 gpurun_out/r02i_strong_n1.json; tail -n 3 gpurun_out/r02i_strong_n1.err
