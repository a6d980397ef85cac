set -x
NCCL_DEBUG=INFO timeout 300 host/run_slabs.sh 2 --nx 512 --ny 512 --nz-total 128 --steps 3 --warmup 2 2>&1 | grep -i "via\|P2P\|SHM\|NVLS\|channel" | head -20 | cut -c1-300
nvidia-smi topo -m | head -12
