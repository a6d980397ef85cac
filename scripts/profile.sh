#!/bin/bash
# ncu captures of the bench step on one B200 (run under gpurun).  $1 = tag, $2.. = bench shape options (default: the bench shape)
TAG=${1:-r02}; shift
SHAPE=${@:---nx 2048 --ny 2048 --nzl 128}
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline $SHAPE"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $B > gpurun_out/launches_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_rhs_fast4 -s 2 -c 2 -f -o gpurun_out/prof_fast_$TAG $B > gpurun_out/prof_fast_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_closure_fast|k_addsgd4" -s 11 -c 11 -f -o gpurun_out/prof_small_$TAG $B > gpurun_out/prof_small_$TAG.log 2>&1
ls -la gpurun_out/ | tail -5
