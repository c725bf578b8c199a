#!/bin/bash
# Round 2, GPU call 19 (1 GPU): how repeatable is the reference arm? three runs, per-step times.
set -u
OUT=gpurun_out
mkdir -p $OUT
for i in 1 2 3; do
  timeout 300 python bench.py --impl reference --no-secondary 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('run $i', round(j['value'],1), 'Mrays/s step_ms', j.get('step_ms'), j['reference_build']['library'])" | tee -a $OUT/r2c19_ref_runs.log
done
timeout 300 python bench.py --impl reference --no-secondary --steps 2 --warmup 0 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cold 2 steps', round(j['value'],1), j.get('step_ms'))" | tee -a $OUT/r2c19_ref_runs.log
