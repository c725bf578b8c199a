#!/bin/bash
# Round 2, GPU call 3 (1 GPU): parity suite on the TMA-staged chain kernel + ray binning, A/B of both.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q -x > $OUT/r2c3_tests.log 2>&1
tail -8 $OUT/r2c3_tests.log
for WL in bunny many_materials terrain bunny_1m; do
  timeout 900 python scripts/ab.py $WL "PT_CHAIN_TMA=0" "PT_CHAIN_TMA=1" "PT_SORT_RAYS=1" >> $OUT/r2c3_ab.log 2>&1
done
cat $OUT/r2c3_ab.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chain_kernel -s 6 -c 4 \
  -o $OUT/r2c3_chain -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary > $OUT/r2c3_ncu_chain.log 2>&1
if [ -f $OUT/r2c3_chain.ncu-rep ]; then
  ncu -i $OUT/r2c3_chain.ncu-rep --page raw --csv > $OUT/r2c3_chain_raw.csv 2>/dev/null
  python scripts/ncu_summary.py $OUT/r2c3_chain_raw.csv > $OUT/r2c3_chain_summary.csv
  cat $OUT/r2c3_chain_summary.csv | cut -c1-200 | head -40
fi
