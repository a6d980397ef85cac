set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r02b_pytest.log; tail -8 gpurun_out/r02b_pytest.log
timeout 900 python bench.py --impl reference-cuda --steps 10 > gpurun_out/r02b_refcuda.json 2> gpurun_out/r02b_refcuda.err; cat gpurun_out/r02b_refcuda.json
timeout 900 python bench.py --config host --steps 10 > gpurun_out/r02b_host.json 2> gpurun_out/r02b_host.err; cat gpurun_out/r02b_host.json
