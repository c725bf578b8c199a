#!/bin/bash
# Round 2, GPU call 6 (1 GPU): full suite (terrain golden frame, sphere trees, merged A-Trous
# launch), denoiser variants, sphere-tree A/B, headline check.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 > $OUT/r2c6_tests.log 2>&1
tail -14 $OUT/r2c6_tests.log
for R in 2 3 4; do
  echo "PT_ATR_R=$R" >> $OUT/r2c6_denoise.log
  PT_ATR_R=$R timeout 300 python scripts/denoise_bench.py 50 >> $OUT/r2c6_denoise.log 2>&1
done
cat $OUT/r2c6_denoise.log
timeout 900 python scripts/ab.py many_spheres "PT_SPHERE_BVH=0" "PT_SPHERE_BVH=1" > $OUT/r2c6_ab.log 2>&1
timeout 900 python scripts/ab.py bunny "PT_X=0" >> $OUT/r2c6_ab.log 2>&1
timeout 900 python scripts/ab.py three_balls "PT_X=0" >> $OUT/r2c6_ab.log 2>&1
cat $OUT/r2c6_ab.log
