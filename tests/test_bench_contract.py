"""bench.py's reference arm prints the contract's JSON line on any box: with a device it times the
reference's own CUDA build (cpu_baseline.kind == "reference"), without one it falls back to the
CPU oracle port (kind == "port") and says why."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload",
                        "three_balls", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "Mrays/s (all bounces)" and line["unit"] == "Mrays/s"
    assert line["higher_is_better"] is True and line["n_gpus"] == 1 and line["vs_baseline"] is None
    assert line["config"]["workload"] == "three_balls" and line["value"] > 0
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    import torch
    if not torch.cuda.is_available():
        assert cb["kind"] == "port" and "unavailable" in cb["sample"]


def test_reference_arm_does_no_work_on_other_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--workload", "three_balls", "--steps", "1", "--warmup", "0"], env=env, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip() == ""
