#!/bin/bash
# one ncu --set full capture of the two fused passes of the bench step (run under gpurun).  $1 = tag
TAG=${1:-x}
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --nx 2048 --ny 2048 --nzl 128"
ncu --set full --clock-control none --import-source on -k regex:k_rhs_fast4 -s 2 -c 2 -f -o gpurun_out/prof_fast_$TAG $B > gpurun_out/prof_fast_$TAG.log 2>&1
tail -3 gpurun_out/prof_fast_$TAG.log
