#!/bin/bash
# Round 2, GPU call 31 (1 GPU): register bounds of the chain kernels (4 / 5 CTAs of 256 threads per SM instead of 3).
set -u
OUT=gpurun_out
mkdir -p $OUT
P=$PWD/cuda_path_tracer_b200
for wl in bunny many_materials terrain; do
timeout 600 python scripts/ab.py $wl "PT_X=0" "B200PT_LIB=$P/libb200pt_chain4.so" "B200PT_LIB=$P/libb200pt_chain4_4.so" "B200PT_LIB=$P/libb200pt_chain5.so" >> $OUT/r2c31_ab.log 2>&1
done
sed -e "s#$P/##g" $OUT/r2c31_ab.log
