set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_host.py -m gpu -q -s -k "slab_driver" 2>&1 | tail -8 | cut -c1-400
timeout 1500 bash scripts/profile.sh r02a
