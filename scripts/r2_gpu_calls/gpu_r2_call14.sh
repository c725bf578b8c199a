#!/bin/bash
# Round 2, GPU call 14 (1 GPU): where does finish_kernel stop paying? passes of 8 / 4 / 2 spp at 1080p (16.6 / 8.3 / 4.1 M paths).
set -u
OUT=gpurun_out
mkdir -p $OUT
for S in 8 4 2; do
  timeout 900 python scripts/ab.py bunny "PT_SPP_PASS=$S PT_FINISH=0" "PT_SPP_PASS=$S PT_FINISH=3 PT_FINISH_MAX=40000000" "PT_SPP_PASS=$S PT_FINISH=4 PT_FINISH_MAX=40000000" "PT_SPP_PASS=$S PT_FINISH=5 PT_FINISH_MAX=40000000" >> $OUT/r2c14_ab.log 2>&1
done
timeout 900 python scripts/ab.py terrain "PT_SPP_PASS=1 PT_FINISH=0" "PT_SPP_PASS=1 PT_FINISH=3 PT_FINISH_MAX=40000000" "PT_SPP_PASS=1 PT_FINISH=5 PT_FINISH_MAX=40000000" >> $OUT/r2c14_ab.log 2>&1
cat $OUT/r2c14_ab.log
