#!/bin/bash
# Round 2, GPU call 30 (1 GPU): checkpoint of the final traversal kernel (quantised nodes, sign-selected planes, three
# inner nodes per pair of votes, 9 CTAs per SM): tests, smoke, both bench arms, launch list, traverse traffic + capture;
# then the parity suite on the bounds-checked build.
set -u
OUT=gpurun_out
mkdir -p $OUT
bash scripts/gpu_checkpoint.sh r2i
timeout 600 bash scripts/run_bounds_check.sh > $OUT/r2i_bounds_checked_suite.log 2>&1
tail -3 $OUT/r2i_bounds_checked_suite.log
