#!/bin/bash
# Round 2, GPU call 8 (8 GPUs): group tests at 3/4/8 devices, the driver's scaling run: bench.py at N = 1, 2, 4, 8.
set -u
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L > $OUT/r2c8_gpus.txt
timeout 900 python -m pytest tests/test_gpu_group.py tests/test_gpu_full_configs.py -m gpu -q -k "group or two_gpu or terrain_4k" > $OUT/r2c8_tests.log 2>&1
tail -5 $OUT/r2c8_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/r2c8_bench_n1.json 2> $OUT/r2c8_bench_n1.err || tail -5 $OUT/r2c8_bench_n1.err
for N in 2 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N \
    bench.py --gpus $N --steps 5 --warmup 3 > $OUT/r2c8_bench_n$N.json 2> $OUT/r2c8_bench_n$N.err || tail -20 $OUT/r2c8_bench_n$N.err
done
python - <<'PY'
import json
base = None
for n in (1, 2, 4, 8):
    f = f"gpurun_out/r2c8_bench_n{n}.json"
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    t = [s for s in j.get("secondary", []) if s.get("workload") == "terrain"]
    if n == 1:
        base = (j["value"], t[0]["value"] if t else None)
    print("N", n, "bunny", round(j["value"], 1), "ms/step", round(j["ms_per_step"], 3), "eff", round(j["value"] / (n * base[0]), 3),
          "e2e", round(j["e2e"]["value"], 1), "rmse", j.get("image_rmse_vs_single"),
          "| terrain", round(t[0]["value"], 1) if t else None, "eff", round(t[0]["value"] / (n * base[1]), 3) if t and base[1] else None,
          "rmse", t[0].get("image_rmse_vs_single") if t else None, "| build ms", j["scene"]["scene_build_ms"], "kernel_ms", {k: (round(v, 2) if isinstance(v, float) else v) for k, v in j["kernel_ms"].items()})
PY
# DRAM traffic of every one-lane traverse launch (for roofline.traffic)
PT_LANES=1 timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:traverse --clock-control none \
  --csv --log-file $OUT/r2c8_traffic_raw.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary > $OUT/r2c8_ncu_traffic.log 2>&1
python scripts/traffic_json.py $OUT/r2c8_traffic_raw.csv bunny $OUT/r2_traverse_traffic
