#!/bin/bash
# compute-sanitizer passes over the GPU parity tests (run under gpurun; a few minutes).  Round 1: 0 errors, 0 hazards.
set -x
compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "tma_path or leaves_ghost" || exit 1
compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_step.py -x -q -m gpu || exit 1
compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "tma_path_sizes and dims1" || exit 1
compute-sanitizer --tool racecheck --error-exitcode 1 python __graft_entry__.py smoke || exit 1
