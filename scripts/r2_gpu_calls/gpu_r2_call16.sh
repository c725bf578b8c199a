#!/bin/bash
# Round 2, GPU call 16 (1 GPU): history-guided finish_kernel: suite, frames, workloads at several pass sizes, FFMA2 micro-benchmark.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/r2c16_tests.log 2>&1
tail -5 $OUT/r2c16_tests.log
for S in "PT_FINISH=0" "PT_FINISH=2" "PT_FINISH=2 PT_FINISH_RAYS=49152" "PT_FINISH=2 PT_FINISH_RAYS=196608" "PT_FINISH=1" "PT_FINISH=3"; do
  env $S timeout 300 python scripts/frame_ab.py 2>&1 | grep frame >> $OUT/r2c16_frames.log
done
cat $OUT/r2c16_frames.log
timeout 900 python scripts/ab.py bunny "PT_FINISH=0" "PT_FINISH=2" "PT_SPP_PASS=8 PT_FINISH=0" "PT_SPP_PASS=8 PT_FINISH=2" "PT_SPP_PASS=2 PT_FINISH=0" "PT_SPP_PASS=2 PT_FINISH=2" "PT_SPP_PASS=2 PT_FINISH=2 PT_FINISH_RAYS=196608" >> $OUT/r2c16_ab.log 2>&1
timeout 900 python scripts/ab.py terrain "PT_FINISH=0" "PT_FINISH=2" "PT_SPP_PASS=1 PT_FINISH=0" "PT_SPP_PASS=1 PT_FINISH=2" >> $OUT/r2c16_ab.log 2>&1
timeout 600 python scripts/ab.py many_materials "PT_FINISH=0" "PT_FINISH=2" >> $OUT/r2c16_ab.log 2>&1
cat $OUT/r2c16_ab.log
./scripts/micro/ffma2 > $OUT/r2c16_ffma2.log 2>&1; cat $OUT/r2c16_ffma2.log
