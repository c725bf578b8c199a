#!/bin/bash
# Round 2, GPU call 24 (1 GPU): the checkpoint of the final code (tests, smoke, both bench arms, launch list, traverse
# traffic + capture on the bunny) and a --set full capture of traverse on the terrain under the new big-scene
# defaults (10 CTAs per SM / 48 registers, refill 12; one lane so that the kernel is alone on the GPU).
set -u
OUT=gpurun_out
mkdir -p $OUT
bash scripts/gpu_checkpoint.sh r2h
PT_LANES=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:traverse_kernel -s 8 -c 3 \
  -o $OUT/r2c24_traverse_terrain -f python bench.py --workload terrain --steps 1 --warmup 1 --no-cpu-baseline --no-secondary > $OUT/r2c24_ncu.log 2>&1
ncu -i $OUT/r2c24_traverse_terrain.ncu-rep --page raw --csv > $OUT/r2c24_traverse_terrain_raw.csv 2>/dev/null
python scripts/ncu_summary.py $OUT/r2c24_traverse_terrain_raw.csv > $OUT/r2c24_traverse_terrain_summary.csv
cut -c1-140 $OUT/r2c24_traverse_terrain_summary.csv
