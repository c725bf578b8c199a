#!/bin/bash
# Round 2, GPU call 26 (1 GPU): compile-time variants of traverse_kernel on top of the quantised nodes
# (libb200pt_<variant>.so, selected with B200PT_LIB): two inner nodes per pair of votes, 9 CTAs per SM
# (56 registers, no spill), both; and the 10-CTA instantiation through PT_TRAV.
set -u
OUT=gpurun_out
mkdir -p $OUT
P=$PWD/cuda_path_tracer_b200
for wl in bunny many_materials bunny_1m; do
timeout 900 python scripts/ab.py $wl "PT_X=0" "B200PT_LIB=$P/libb200pt_steps2.so" "B200PT_LIB=$P/libb200pt_minb9.so" "B200PT_LIB=$P/libb200pt_steps2_minb9.so" "PT_TRAV=10,0" >> $OUT/r2c26_ab.log 2>&1
done
timeout 300 python scripts/ab.py terrain "PT_X=0" "B200PT_LIB=$P/libb200pt_steps2.so" >> $OUT/r2c26_ab.log 2>&1
sed -e "s#$P/##" $OUT/r2c26_ab.log
