#!/bin/bash
# Round 2, GPU call 32 (1 GPU): length of the stack-free prefix walk in classify() (2 / 4 / 6 / 8 nodes) now that
# traversal is cheaper; launch_bounds(256, 3) for the chain kernels.
set -u
OUT=gpurun_out
mkdir -p $OUT
P=$PWD/cuda_path_tracer_b200
for wl in bunny many_materials terrain; do
timeout 600 python scripts/ab.py $wl "PT_X=0" "B200PT_LIB=$P/libb200pt_prefix2.so" "B200PT_LIB=$P/libb200pt_prefix4.so" "B200PT_LIB=$P/libb200pt_prefix8.so" "B200PT_LIB=$P/libb200pt_chainb3.so" >> $OUT/r2c32_ab.log 2>&1
done
sed -e "s#$P/##g" $OUT/r2c32_ab.log
