#!/bin/bash
# Round 2, GPU call 29 (1 GPU): single 256-bit fetch of a quantised node on the mid-size scenes, quantised nodes on
# the terrain once more (now with the sign-selected planes), CTAs per SM for the exact-node variant under 3 steps.
set -u
OUT=gpurun_out
mkdir -p $OUT
for wl in many_materials bunny_1m; do
timeout 600 python scripts/ab.py $wl "PT_X=0" "PT_TRAV=9,1" >> $OUT/r2c29_ab.log 2>&1
done
timeout 600 python scripts/ab.py terrain "PT_X=0" "PT_QNODES=1" "PT_QNODES=1 PT_TRAV=9,1" "PT_TRAV=8,0" >> $OUT/r2c29_ab.log 2>&1
timeout 600 python scripts/ab.py bunny "PT_QNODES=0" "PT_QNODES=0 PT_TRAV=10,1" >> $OUT/r2c29_ab.log 2>&1
cat $OUT/r2c29_ab.log
