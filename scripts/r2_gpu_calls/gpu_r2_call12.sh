#!/bin/bash
# Round 2, GPU call 12 (1 GPU): packed-FP32 A-Trous variants + denoiser parity under each.
set -u
OUT=gpurun_out
mkdir -p $OUT
for P in 0 1; do for R in 2 3 4; do
  echo "PT_ATR_PACK=$P PT_ATR_R=$R" >> $OUT/r2c12_denoise.log
  PT_ATR_PACK=$P PT_ATR_R=$R timeout 300 python scripts/denoise_bench.py 50 >> $OUT/r2c12_denoise.log 2>&1
  PT_ATR_PACK=$P PT_ATR_R=$R timeout 300 python -m pytest tests -m gpu -q -x -k "denois or interactive" 2>&1 | tail -1 >> $OUT/r2c12_denoise.log
done; done
cat $OUT/r2c12_denoise.log
