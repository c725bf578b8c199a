#!/bin/bash
# Round 2, GPU call 2a (1 GPU): whole parity suite, both bench arms with the secondary configs.
set -u
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q --durations=10 > $OUT/r2c2_tests.log 2>&1
tail -25 $OUT/r2c2_tests.log
timeout 600 python bench.py --impl reference > $OUT/r2c2_bench_ref.json 2> $OUT/r2c2_bench_ref.err || tail -5 $OUT/r2c2_bench_ref.err
timeout 600 python bench.py > $OUT/r2c2_bench.json 2> $OUT/r2c2_bench.err || tail -20 $OUT/r2c2_bench.err
python - <<'PY'
import json
for f in ("gpurun_out/r2c2_bench_ref.json", "gpurun_out/r2c2_bench.json"):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f, round(j["value"], 1), j["unit"], "ms/step", j.get("ms_per_step"), "e2e", j["e2e"]["value"])
    for s in j.get("secondary", []):
        print("   ", s.get("workload"), {k: (round(v, 3) if isinstance(v, float) else v) for k, v in s.items()
                                        if k in ("value", "unit", "ms_per_step", "frame_ms_median", "frame_ms_min", "gpu_ms_min", "reference_fails", "denoise_ms", "mrays_per_s")},
              "ref:", s.get("reference"))
PY
