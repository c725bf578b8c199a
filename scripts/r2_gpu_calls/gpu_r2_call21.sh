#!/bin/bash
# Round 2, GPU call 21 (1 GPU): traverse_kernel on the 10 M-triangle terrain as it runs at the end of the round
# (item order 2, 12 CTAs per SM / 40 registers; one lane so that the kernel is alone on the GPU).
set -u
OUT=gpurun_out
mkdir -p $OUT
PT_LANES=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:traverse_kernel -s 8 -c 3 \
  -o $OUT/r2c21_traverse_terrain -f python bench.py --workload terrain --steps 1 --warmup 1 --no-cpu-baseline --no-secondary > $OUT/r2c21_ncu.log 2>&1
ncu -i $OUT/r2c21_traverse_terrain.ncu-rep --page raw --csv > $OUT/r2c21_traverse_terrain_raw.csv 2>/dev/null
python scripts/ncu_summary.py $OUT/r2c21_traverse_terrain_raw.csv > $OUT/r2c21_traverse_terrain_summary.csv
cut -c1-140 $OUT/r2c21_traverse_terrain_summary.csv
