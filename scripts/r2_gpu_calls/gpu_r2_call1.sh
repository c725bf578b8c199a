#!/bin/bash
# Round 2, GPU call 1: parity suite (old + new full-config tests), baseline bench, A/B of the
# batch-1 switches, fresh ncu captures of traverse_kernel on the headline and terrain workloads.
set -u
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > $OUT/r2c1_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q --durations=15 > $OUT/r2c1_tests.log 2>&1
tail -25 $OUT/r2c1_tests.log
timeout 300 python bench.py > $OUT/r2c1_bench.json 2> $OUT/r2c1_bench.err || tail -5 $OUT/r2c1_bench.err
timeout 900 python scripts/ab.py bunny "PT_X=0" "PT_ORDER=1" "PT_ORDER=2" "PT_TRAV=8,1" "PT_TRAV=10,0" "PT_TRAV=12,0" "PT_TRAV=12,1" "PT_SPHERE_PREREJECT=0" "PT_ORDER=1 PT_TRAV=12,0" > $OUT/r2c1_ab_bunny.log 2>&1
cat $OUT/r2c1_ab_bunny.log
timeout 900 python scripts/ab.py terrain "PT_X=0" "PT_ORDER=1" "PT_ORDER=2" "PT_TRAV=8,1" "PT_TRAV=10,0" "PT_TRAV=12,0" "PT_ORDER=2 PT_TRAV=12,0" > $OUT/r2c1_ab_terrain.log 2>&1
cat $OUT/r2c1_ab_terrain.log
timeout 600 python scripts/ab.py bunny_1m "PT_X=0" "PT_ORDER=1" "PT_TRAV=12,0" "PT_TRAV=8,1" > $OUT/r2c1_ab_bunny1m.log 2>&1
cat $OUT/r2c1_ab_bunny1m.log
for WL in bunny terrain; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:traverse_kernel -s 8 -c 3 \
    -o $OUT/r2c1_traverse_$WL -f python bench.py --workload $WL --steps 1 --warmup 1 --no-cpu-baseline > $OUT/r2c1_ncu_$WL.log 2>&1
  if [ -f $OUT/r2c1_traverse_$WL.ncu-rep ]; then
    ncu -i $OUT/r2c1_traverse_$WL.ncu-rep --page raw --csv > $OUT/r2c1_traverse_${WL}_raw.csv 2>/dev/null
    python scripts/ncu_summary.py $OUT/r2c1_traverse_${WL}_raw.csv > $OUT/r2c1_traverse_${WL}_summary.csv
  fi
done
ls -la $OUT | tail -20
