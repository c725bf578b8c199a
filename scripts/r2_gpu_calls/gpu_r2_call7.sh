#!/bin/bash
# Round 2, GPU call 7 (2 GPUs): whole suite incl. the pt_group_* / 2-GPU tests, strong-scaling bench at N=1 and N=2.
set -u
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L > $OUT/r2c7_gpus.txt
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > $OUT/r2c7_tests.log 2>&1
tail -16 $OUT/r2c7_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $OUT/r2c7_bench_n1.json 2> $OUT/r2c7_bench_n1.err || tail -5 $OUT/r2c7_bench_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 5 --warmup 3 > $OUT/r2c7_bench_n2.json 2> $OUT/r2c7_bench_n2.err || tail -20 $OUT/r2c7_bench_n2.err
python - <<'PY'
import json
for f in ("gpurun_out/r2c7_bench_n1.json", "gpurun_out/r2c7_bench_n2.json"):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f, "N", j["n_gpus"], round(j["value"], 1), j["unit"], "ms/step", round(j["ms_per_step"], 3), "e2e", round(j["e2e"]["value"], 1),
          "scaling", j["scaling"], "rmse", j.get("image_rmse_vs_single"))
    for s in j.get("secondary", []):
        print("   ", s.get("workload"), round(s["value"], 2), s["unit"], "ms/step", s.get("ms_per_step"), "rmse", s.get("image_rmse_vs_single"))
PY
