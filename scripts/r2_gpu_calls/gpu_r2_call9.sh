#!/bin/bash
# Round 2, GPU call 9 (1 GPU): suite after the lane refactor (1 lane and 2 lanes), A/B of two lanes, frame launch list.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/r2c9_tests.log 2>&1
tail -6 $OUT/r2c9_tests.log
PT_LANES=2 timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/r2c9_tests_lanes2.log 2>&1
tail -6 $OUT/r2c9_tests_lanes2.log
for WL in bunny terrain bunny_1m; do
  timeout 900 python scripts/ab.py $WL "PT_LANES=1" "PT_LANES=2" >> $OUT/r2c9_ab.log 2>&1
done
cat $OUT/r2c9_ab.log
PT_LANES=1 timeout 300 python scripts/frame_ab.py >> $OUT/r2c9_frames.log 2>&1
PT_LANES=2 timeout 300 python scripts/frame_ab.py >> $OUT/r2c9_frames.log 2>&1
cat $OUT/r2c9_frames.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/r2c9_frame_launches.csv \
  python scripts/frame_launches.py 3 > $OUT/r2c9_ncu_frame.log 2>&1
python - <<'PY'
import csv
rows = [l for l in open("gpurun_out/r2c9_frame_launches.csv") if l.startswith('"')]
r = list(csv.DictReader(rows))
# last frame only: find the last chain_kernel<1 (first chain) occurrence
idx = [i for i, x in enumerate(r) if "chain_kernel<1" in x["Kernel Name"] or "chain_kernel<(bool)1" in x["Kernel Name"]]
start = idx[-1] if idx else 0
tot = 0.0
for x in r[start - 1:]:
    v = float(x["Metric Value"].replace(",", "")); tot += v
    print(f"{x['Kernel Name'][:60]:60s} grid {x['Grid Size']:16s} {v/1000:8.1f} us")
print("sum of kernel durations of one frame:", tot / 1000, "us")
PY
