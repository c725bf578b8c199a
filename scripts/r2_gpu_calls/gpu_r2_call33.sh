#!/bin/bash
# Round 2, GPU call 33 (1 GPU): the final quantised-node kernel at 8 CTAs per SM (64 registers, no spill) against 9.
set -u
OUT=gpurun_out
mkdir -p $OUT
for wl in bunny many_materials bunny_1m; do
timeout 600 python scripts/ab.py $wl "PT_X=0" "PT_TRAV=8,1" >> $OUT/r2c33_ab.log 2>&1
done
cat $OUT/r2c33_ab.log
