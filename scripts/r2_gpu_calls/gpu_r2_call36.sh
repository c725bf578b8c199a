#!/bin/bash
# Round 2, GPU call 36 (1 GPU): --set full of the chain kernels as they run at the end of the round
# (launch_bounds(256, 3); one lane so that each launch is alone on the GPU).
set -u
OUT=gpurun_out
mkdir -p $OUT
PT_LANES=1 timeout 200 ncu --set full --clock-control none --import-source on -k regex:chain_kernel -s 8 -c 3 \
  -o $OUT/r2c36_chain -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary > $OUT/r2c36_ncu.log 2>&1
ncu -i $OUT/r2c36_chain.ncu-rep --page raw --csv > $OUT/r2c36_chain_raw.csv 2>/dev/null
python scripts/ncu_summary.py $OUT/r2c36_chain_raw.csv > $OUT/r2c36_chain_summary.csv
cut -c1-150 $OUT/r2c36_chain_summary.csv | head -40
