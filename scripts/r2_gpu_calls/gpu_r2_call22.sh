#!/bin/bash
# Round 2, GPU call 22 (1 GPU): re-tune the traversal scheduling thresholds and chain grids under two lanes.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python scripts/ab.py bunny "PT_X=0" "PT_REFILL=12" "PT_REFILL=20" "PT_REFILL=24" "PT_INNER_MIN=4" "PT_INNER_MIN=12" "PT_CHAIN_GRID=3" "PT_CHAIN_GRID=12" "PT_CHAIN_GRID0=12" "PT_CHAIN_GRID0=48" > $OUT/r2c22_ab.log 2>&1
timeout 900 python scripts/ab.py terrain "PT_X=0" "PT_REFILL=12" "PT_REFILL=24" "PT_INNER_MIN=4" "PT_INNER_MIN=12" "PT_TRAV=10,0" >> $OUT/r2c22_ab.log 2>&1
cat $OUT/r2c22_ab.log
