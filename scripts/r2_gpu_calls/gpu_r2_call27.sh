#!/bin/bash
# Round 2, GPU call 27 (1 GPU): inner nodes per pair of votes 2 / 3 / 4 at 9 and 10 CTAs per SM, inner-phase minimum.
set -u
OUT=gpurun_out
mkdir -p $OUT
P=$PWD/cuda_path_tracer_b200
L2=$P/libb200pt_steps2_minb9.so; L3=$P/libb200pt_steps3_minb9.so; L4=$P/libb200pt_steps4_minb9.so
timeout 900 python scripts/ab.py bunny "B200PT_LIB=$L2" "B200PT_LIB=$L3" "B200PT_LIB=$L4" "B200PT_LIB=$L2 PT_TRAV=10,0" "B200PT_LIB=$L3 PT_TRAV=10,0" "B200PT_LIB=$L4 PT_TRAV=10,0" "B200PT_LIB=$L3 PT_INNER_MIN=4" "B200PT_LIB=$L3 PT_INNER_MIN=12" "B200PT_LIB=$L3 PT_REFILL=12" "B200PT_LIB=$L3 PT_REFILL=20" >> $OUT/r2c27_ab.log 2>&1
for wl in many_materials bunny_1m; do
timeout 900 python scripts/ab.py $wl "B200PT_LIB=$L2" "B200PT_LIB=$L3" "B200PT_LIB=$L4" "B200PT_LIB=$L3 PT_TRAV=10,0" "B200PT_LIB=$L4 PT_TRAV=10,0" >> $OUT/r2c27_ab.log 2>&1
done
timeout 300 python scripts/ab.py terrain "B200PT_LIB=$L3" "B200PT_LIB=$L4" >> $OUT/r2c27_ab.log 2>&1
sed -e "s#$P/##g" $OUT/r2c27_ab.log
