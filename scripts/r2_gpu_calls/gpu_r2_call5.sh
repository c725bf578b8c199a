#!/bin/bash
# Round 2, GPU call 5 (1 GPU): denoiser variants (interior/edge split, ex2.approx, outputs per
# thread 2/3/4) + denoiser parity tests + SAH cost-ratio knob on the terrain + ncu of atrous.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -q -x -k "denois or interactive or band or smoke" > $OUT/r2c5_tests.log 2>&1
tail -4 $OUT/r2c5_tests.log
for R in 2 3 4; do
  echo "PT_ATR_R=$R" >> $OUT/r2c5_denoise.log
  PT_ATR_R=$R timeout 300 python scripts/denoise_bench.py 50 >> $OUT/r2c5_denoise.log 2>&1
  PT_ATR_R=$R timeout 300 python -m pytest tests -m gpu -q -x -k "denois" 2>&1 | tail -1 >> $OUT/r2c5_denoise.log
done
cat $OUT/r2c5_denoise.log
timeout 900 python scripts/ab.py terrain "PT_SAH_ISECT=1.0" "PT_SAH_ISECT=0.6" "PT_SAH_ISECT=1.7" "PT_SAH_ISECT=0.6 PT_SAH_LEAF=8" > $OUT/r2c5_ab_sah.log 2>&1
cat $OUT/r2c5_ab_sah.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:atrous_kernel -c 6 \
  -o $OUT/r2c5_atrous -f python scripts/denoise_bench.py 1 > $OUT/r2c5_ncu_atrous.log 2>&1
if [ -f $OUT/r2c5_atrous.ncu-rep ]; then
  ncu -i $OUT/r2c5_atrous.ncu-rep --page raw --csv > $OUT/r2c5_atrous_raw.csv 2>/dev/null
  python scripts/ncu_summary.py $OUT/r2c5_atrous_raw.csv > $OUT/r2c5_atrous_summary.csv
  cut -c1-160 $OUT/r2c5_atrous_summary.csv | head -40
fi
