#!/bin/bash
# Round 2, GPU call 4a (1 GPU): parity suite (release), the same suite on the bounds-checked debug
# build, A/B of the per-warp TMA staging.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q -x > $OUT/r2c4_tests.log 2>&1
tail -5 $OUT/r2c4_tests.log
timeout 1200 bash scripts/run_bounds_check.sh > $OUT/r2c4_bounds.log 2>&1
tail -5 $OUT/r2c4_bounds.log
for WL in bunny many_materials terrain; do
  timeout 900 python scripts/ab.py $WL "PT_CHAIN_TMA=0" "PT_CHAIN_TMA=2" >> $OUT/r2c4_ab.log 2>&1
done
cat $OUT/r2c4_ab.log
