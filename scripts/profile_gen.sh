#!/bin/bash
# ncu --set full capture of the interior kernel of one generation (run under gpurun).  $1 = tag, $2 = SW4B200_FAST_GEN, $3 = kernel regex
TAG=$1; GEN=$2; KRE=${3:-k_rhs_fast}
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --nx 768 --ny 768 --nzl 96"
SW4B200_FAST_GEN=$GEN ncu --set full --clock-control none --import-source on -k regex:$KRE -s 2 -c 2 -f -o gpurun_out/prof_$TAG $B > gpurun_out/prof_$TAG.log 2>&1
tail -3 gpurun_out/prof_$TAG.log
