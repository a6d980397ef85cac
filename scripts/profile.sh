#!/bin/bash
# ncu captures of the bench step on one B200 (run under gpurun).  $1 = tag, $2 = fast|small|all (what to capture with --set full;
# gpurun brings back at most 64 MiB, one --set full report of the step's kernels is 30-50 MB), $3.. = bench shape options
TAG=${1:-r02}; WHAT=${2:-fast}; shift; shift
SHAPE=${@:---nx 2048 --ny 2048 --nzl 128}
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline $SHAPE"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $B > gpurun_out/launches_$TAG.log 2>&1
# the two fused passes of the second step: plain tiles + tiles with stretching, predictor and corrector
if [ "$WHAT" != "small" ]; then
ncu --set full --clock-control none --import-source on -k regex:k_rhs_fast4 -s 4 -c 4 -f -o gpurun_out/prof_fast_$TAG $B > gpurun_out/prof_fast_$TAG.log 2>&1
fi
if [ "$WHAT" != "fast" ]; then
ncu --set full --clock-control none --import-source on -k regex:"k_closure_fast|k_addsgd4" -s 7 -c 7 -f -o gpurun_out/prof_small_$TAG $B > gpurun_out/prof_small_$TAG.log 2>&1
fi
ls -la gpurun_out/ | tail -5
