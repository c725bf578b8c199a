#!/bin/bash
# Round 2, GPU call 28 (1 GPU): new traverse defaults (quantised nodes, 3 inner nodes per pair of votes, 9 / 10 CTAs per
# SM, near/far planes picked by the ray's direction signs): parity suite, A/B against the variant without the sign
# selection, inner-phase minimum and refill threshold on top.
set -u
OUT=gpurun_out
mkdir -p $OUT
P=$PWD/cuda_path_tracer_b200
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/r2c28_tests.log 2>&1
tail -4 $OUT/r2c28_tests.log
N=$P/libb200pt_nosignsel.so
timeout 900 python scripts/ab.py bunny "PT_X=0" "B200PT_LIB=$N" "PT_INNER_MIN=12" "PT_INNER_MIN=16" "PT_INNER_MIN=20" "PT_INNER_MIN=12 PT_REFILL=12" "PT_INNER_MIN=12 PT_REFILL=20" "PT_TRAV=10,0" >> $OUT/r2c28_ab.log 2>&1
for wl in many_materials bunny_1m; do
timeout 900 python scripts/ab.py $wl "PT_X=0" "B200PT_LIB=$N" "PT_INNER_MIN=12" "PT_INNER_MIN=16" "PT_TRAV=9,0" >> $OUT/r2c28_ab.log 2>&1
done
timeout 300 python scripts/ab.py terrain "PT_X=0" "PT_INNER_MIN=12" "PT_INNER_MIN=16" >> $OUT/r2c28_ab.log 2>&1
sed -e "s#$P/##g" $OUT/r2c28_ab.log
