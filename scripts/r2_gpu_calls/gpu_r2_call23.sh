#!/bin/bash
# Round 2, GPU call 23 (1 GPU): 10 CTAs per SM as the big-scene default, L2 prefetch-size hints
# (ld.global.nc.L2::256B on triangles, ::128B on nodes) and the device's L2 fetch granularity.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python scripts/ab.py terrain "PT_X=0" "PT_TRAV=12,0" "PT_L2HINT=1" "PT_L2HINT=2" "PT_L2_FETCH=128" "PT_L2_FETCH=128 PT_L2HINT=2" "PT_L2_FETCH=32" "PT_REFILL=12" "PT_REFILL=12 PT_L2HINT=1" > $OUT/r2c23_ab.log 2>&1
timeout 900 python scripts/ab.py bunny_1m "PT_X=0" "PT_TRAV=10,0" "PT_TRAV=10,0 PT_L2HINT=1" "PT_TRAV=10,0 PT_L2HINT=2" "PT_L2_FETCH=128" "PT_L2_FETCH=32" >> $OUT/r2c23_ab.log 2>&1
timeout 300 python scripts/ab.py bunny "PT_X=0" "PT_L2_FETCH=128" "PT_L2_FETCH=32" >> $OUT/r2c23_ab.log 2>&1
cat $OUT/r2c23_ab.log
