#!/bin/bash
# Round 2, GPU call 15 (1 GPU): adaptive finish_kernel: suite, frames, workloads at several pass sizes.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/r2c15_tests.log 2>&1
tail -5 $OUT/r2c15_tests.log
for S in "PT_FINISH=0" "PT_FINISH=2" "PT_FINISH=2 PT_FINISH_RAYS=32768" "PT_FINISH=2 PT_FINISH_RAYS=524288" "PT_FINISH=1" "PT_FINISH=3"; do
  env $S timeout 300 python scripts/frame_ab.py 2>&1 | grep interactive >> $OUT/r2c15_frames.log
done
cat $OUT/r2c15_frames.log
timeout 900 python scripts/ab.py bunny "PT_FINISH=0" "PT_FINISH=2" "PT_SPP_PASS=8 PT_FINISH=0" "PT_SPP_PASS=8 PT_FINISH=2" "PT_SPP_PASS=2 PT_FINISH=0" "PT_SPP_PASS=2 PT_FINISH=2" "PT_SPP_PASS=2 PT_FINISH=2 PT_FINISH_RAYS=524288" >> $OUT/r2c15_ab.log 2>&1
timeout 900 python scripts/ab.py terrain "PT_FINISH=0" "PT_FINISH=2" "PT_SPP_PASS=1 PT_FINISH=0" "PT_SPP_PASS=1 PT_FINISH=2" >> $OUT/r2c15_ab.log 2>&1
cat $OUT/r2c15_ab.log
