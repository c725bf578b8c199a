#!/bin/bash
# Round 2, GPU call 20 (1 GPU): bounds-checked debug build on the final code (lanes, finish kernel, sphere trees) + the release suite.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 bash scripts/run_bounds_check.sh > $OUT/r2c20_bounds.log 2>&1
tail -4 $OUT/r2c20_bounds.log
PT_LANES=1 PT_FINISH=0 timeout 900 python -m pytest tests -m gpu -q -x -k "parity or edge or full_size or full_configs" > $OUT/r2c20_tests_plain.log 2>&1
tail -3 $OUT/r2c20_tests_plain.log
timeout 1500 python -m pytest tests -m gpu -q > $OUT/r2c20_tests.log 2>&1
tail -3 $OUT/r2c20_tests.log
