#!/bin/bash
# Round 2, GPU call 13 (1 GPU): finish_kernel (late bounces of small passes in one launch): parity suite + frame A/B.
set -u
OUT=gpurun_out
mkdir -p $OUT
PT_FINISH=3 timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/r2c13_tests_finish3.log 2>&1
tail -6 $OUT/r2c13_tests_finish3.log
PT_FINISH=1 timeout 900 python -m pytest tests -m gpu -q -x -k "parity or edge or full_size or group or shim" > $OUT/r2c13_tests_finish1.log 2>&1
tail -4 $OUT/r2c13_tests_finish1.log
for F in 0 1 2 3 4 5; do
  PT_FINISH=$F timeout 300 python scripts/frame_ab.py >> $OUT/r2c13_frames.log 2>&1
done
cat $OUT/r2c13_frames.log
