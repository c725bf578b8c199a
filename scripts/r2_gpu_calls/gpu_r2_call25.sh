#!/bin/bash
# Round 2, GPU call 25 (1 GPU): 32-byte quantised nodes for the traversal kernels (DevScene::qnodes):
# parity suite with them on (the default for host-built trees up to 512 MB), A/B against the exact nodes.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/r2c25_tests.log 2>&1
tail -6 $OUT/r2c25_tests.log
timeout 600 python scripts/ab.py bunny "PT_QNODES=0" "PT_QNODES=1" "PT_QNODES=1 PT_TRAV=8,0" > $OUT/r2c25_ab.log 2>&1
timeout 600 python scripts/ab.py many_materials "PT_QNODES=0" "PT_QNODES=1" >> $OUT/r2c25_ab.log 2>&1
timeout 600 python scripts/ab.py bunny_1m "PT_QNODES=0" "PT_QNODES=1" >> $OUT/r2c25_ab.log 2>&1
timeout 600 python scripts/ab.py terrain "PT_QNODES=0" "PT_QNODES=1" >> $OUT/r2c25_ab.log 2>&1
cat $OUT/r2c25_ab.log
