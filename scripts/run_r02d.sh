set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_host.py tests/test_checkpoint.py -m gpu -q -s 2>&1 | grep -v "^$" | tail -30 > gpurun_out/r02d_pytest.log; grep -i "through\|passed\|failed\|error" gpurun_out/r02d_pytest.log | cut -c1-300
