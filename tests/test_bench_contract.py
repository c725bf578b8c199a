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


def test_strong_scaling_shares_partition_the_sample_range(monkeypatch):
    """bench.py --gpus N splits the frame's spp into N disjoint sample ranges that cover it."""
    sys.path.insert(0, ROOT)
    import bench
    for spp in (1, 16, 64, 1024):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for rank in range(world):
                monkeypatch.setenv("WORLD_SIZE", str(world))
                monkeypatch.setenv("RANK", str(rank))
                first, n = bench.Dist().share(spp)
                seen.extend(range(first, first + n))
            assert seen == list(range(spp)), (spp, world)


def test_both_arms_describe_the_workload_with_the_same_config_keys():
    sys.path.insert(0, ROOT)
    import bench
    for name in bench.WORKLOADS:
        c = bench.config_of(name)
        assert c["workload"] == name and "l2" in c and {"w", "h", "spp", "depth"} <= set(c)
    src = open(os.path.join(ROOT, "bench.py")).read()
    # one function builds `config` for both arms; neither arm adds keys to it afterwards
    assert src.count('"config": config_of(') >= 3 and 'line["config"][' not in src
