#!/bin/bash
# One gpurun call that refreshes every measured artefact of the headline path (run from the repo
# root ON THE GPU BOX, e.g.  gpurun --timeout 1800 -- 'bash scripts/gpu_checkpoint.sh r2_final'):
#   1. pytest -m gpu                        -> gpurun_out/<tag>_tests.log
#   2. smoke()                              -> gpurun_out/<tag>_smoke.log
#   3. bench.py (both arms)                 -> gpurun_out/<tag>_bench.json, <tag>_bench_ref.json
#   4. ncu launch list of bench.py          -> gpurun_out/<tag>_launches.csv    (only after 3 exited 0)
#   5. DRAM traffic of every traverse launch-> gpurun_out/<tag>_traverse_traffic.{csv,json}
#   6. ncu --set full of traverse_kernel    -> gpurun_out/<tag>_traverse.ncu-rep + _raw.csv + summary
# Numbers printed by a run under ncu are never bench values; read shares from the launch list.
set -u
TAG=${1:-ckpt}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q --durations=10 > $OUT/${TAG}_tests.log 2>&1
tail -16 $OUT/${TAG}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1
tail -2 $OUT/${TAG}_smoke.log
timeout 600 python bench.py --impl reference > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err || { echo "bench failed"; tail -5 $OUT/${TAG}_bench.err; exit 1; }
python - <<PY
import json
j = json.loads(open("$OUT/${TAG}_bench.json").read().strip().splitlines()[-1])
r = json.loads(open("$OUT/${TAG}_bench_ref.json").read().strip().splitlines()[-1])
print("bench", round(j["value"], 1), j["unit"], "e2e", round(j["e2e"]["value"], 1), "ms/step", round(j["ms_per_step"], 3),
      "roofline.frac", round(j["roofline"]["frac"], 4), "launches", j["gpu_launches"], "clocks", j["clocks"])
print("kernel_ms", j["kernel_ms"])
print("reference", round(r["value"], 1), r.get("reference_build", {}).get("library"), "ratio e2e", round(j["e2e"]["value"] / r["value"], 1))
for s in j.get("secondary", []):
    print("  ", s["workload"], round(s["value"], 3), s["unit"], "e2e", s["e2e"].get("value"), "frac", round(s["roofline"]["frac"], 4),
          "| ref", {k: (round(v, 2) if isinstance(v, float) else v) for k, v in (s.get("reference") or {}).items() if k in ("value", "frame_ms_min", "gpu_ms_min", "scene_create_s", "reference_fails")})
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary > $OUT/${TAG}_ncu_launches.log 2>&1
# (PT_LANES=1 for the two kernel captures: every launch alone on the GPU, as the roofline's event times)
PT_LANES=1 timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:traverse --clock-control none \
  --csv --log-file $OUT/${TAG}_traffic_raw.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary > $OUT/${TAG}_ncu_traffic.log 2>&1
python scripts/traffic_json.py $OUT/${TAG}_traffic_raw.csv bunny $OUT/${TAG}_traverse_traffic
PT_LANES=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:traverse_kernel -s 8 -c 3 \
  -o $OUT/${TAG}_traverse -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary > $OUT/${TAG}_ncu_full.log 2>&1
if [ -f $OUT/${TAG}_traverse.ncu-rep ]; then
  ncu -i $OUT/${TAG}_traverse.ncu-rep --page raw --csv > $OUT/${TAG}_traverse_raw.csv 2>/dev/null
  python scripts/ncu_summary.py $OUT/${TAG}_traverse_raw.csv > $OUT/${TAG}_traverse_summary.csv 2>&1
  cut -c1-150 $OUT/${TAG}_traverse_summary.csv | head -12
fi
python - <<PY
# kernel-time shares from the ncu launch list (serialised, cold cache: shares, not absolutes)
import csv, collections
rows = list(csv.DictReader([l for l in open("$OUT/${TAG}_launches.csv") if l.startswith('"')]))
tot = collections.Counter()
for r in rows:
    n = r["Kernel Name"]
    key = "traverse" if "traverse_kernel" in n else "chain<true>" if "chain_kernel<1" in n or "chain_kernel<(bool)1" in n else \
          "chain<false>" if "chain_kernel" in n else "accumulate" if "accumulate" in n else "other"
    tot[key] += float(r["Metric Value"].replace(",", ""))
s = sum(v for k, v in tot.items() if k != "other")
print("ncu launch-list shares:", {k: round(100 * v / s, 1) for k, v in tot.items() if k != "other"}, "other(us)", round(tot["other"] / 1e3))
PY
