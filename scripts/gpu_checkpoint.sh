#!/bin/bash
# One gpurun call that refreshes every measured artefact of the headline path (run from the repo
# root ON THE GPU BOX, e.g.  gpurun --timeout 600 -- 'bash scripts/gpu_checkpoint.sh r2a'):
#   1. pytest -m gpu                        -> gpurun_out/<tag>_tests.log
#   2. bench.py (both arms)                 -> gpurun_out/<tag>_bench.json, <tag>_bench_ref.json
#   3. ncu launch list of bench.py          -> gpurun_out/<tag>_launches.csv    (only after 2 exited 0)
#   4. ncu --set full of traverse_kernel    -> gpurun_out/<tag>_traverse.ncu-rep + _raw.csv + summary
# Numbers printed by a run under ncu are never bench values; read shares from the launch list.
set -u
TAG=${1:-ckpt}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_tests.log 2>&1
tail -3 $OUT/${TAG}_tests.log
timeout 300 python bench.py --impl reference > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err
timeout 300 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err || { echo "bench failed"; tail -5 $OUT/${TAG}_bench.err; exit 1; }
python - <<PY
import json
j = json.load(open("$OUT/${TAG}_bench.json"))
print("bench", round(j["value"], 1), j["unit"], "e2e", round(j["e2e"]["value"], 1), "ms/step", round(j["ms_per_step"], 3),
      "roofline.frac", round(j["roofline"]["frac"], 4), "clocks", j["clocks"])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:traverse_kernel -s 8 -c 3 \
  -o $OUT/${TAG}_traverse -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1
if [ -f $OUT/${TAG}_traverse.ncu-rep ]; then
  ncu -i $OUT/${TAG}_traverse.ncu-rep --page raw --csv > $OUT/${TAG}_traverse_raw.csv 2>/dev/null
  python scripts/ncu_summary.py $OUT/${TAG}_traverse_raw.csv > $OUT/${TAG}_traverse_summary.txt 2>&1
  head -30 $OUT/${TAG}_traverse_summary.txt
fi
