#!/bin/bash
# Round 2, GPU call 35 (2 GPUs): the multi-GPU tests (pt_group_*, cuda_pt --gpus) and the strong-scaling bench at
# N = 2 on the round's final code (quantised nodes are uploaded per device).
set -u
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L > $OUT/r2c35_gpus.txt
timeout 300 python -m pytest tests/test_gpu_group.py -m gpu -q > $OUT/r2c35_tests.log 2>&1
tail -4 $OUT/r2c35_tests.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > $OUT/r2c35_bench_n2.json 2> $OUT/r2c35_bench_n2.err || tail -20 $OUT/r2c35_bench_n2.err
python - <<'PY'
import json
f = "gpurun_out/r2c35_bench_n2.json"
try:
    j = json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "N", j["n_gpus"], round(j["value"], 1), j["unit"], "ms/step", round(j["ms_per_step"], 3), "e2e", round(j["e2e"]["value"], 1),
          "scaling", j["scaling"], "rmse", j.get("image_rmse_vs_single"))
    for s in j.get("secondary", []):
        print("   ", s.get("workload"), round(s["value"], 2), s["unit"], "ms/step", s.get("ms_per_step"), "rmse", s.get("image_rmse_vs_single"))
except Exception as e:
    print(f, "unreadable", e)
PY
