"""A/B driver: runs bench.py under different environment settings and prints one summary line each.
usage: python scripts/ab.py <workload> [--steps K] "ENV1=a ENV2=b" "ENV1=c" ..."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
wl = sys.argv[1]
steps = "3"
args = sys.argv[2:]
if args and args[0] == "--steps":
    steps = args[1]; args = args[2:]
for setting in args:
    env = dict(os.environ)
    for kv in setting.split():
        k, v = kv.split("=", 1); env[k] = v
    try:
      r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", wl, "--steps", steps, "--warmup", "3",
                        "--no-cpu-baseline", "--no-secondary"], env=env, capture_output=True, text=True, timeout=150)
    except subprocess.TimeoutExpired:
      print(wl, setting, "TIMEOUT", flush=True); continue
    try:
        j = json.loads(r.stdout.strip().splitlines()[-1])
        k = j["kernel_ms"]
        print(f"{wl:10s} {setting:40s} {j['value']:9.1f} Mrays/s  step {j['ms_per_step']:8.2f} ms  traverse {k['traverse']:8.2f}  "
              f"chain0 {k['raygen_classify']:6.2f}  chain {k['shade_classify_compact']:6.2f}  img {j.get('image_mean'):.4f}", flush=True)
    except Exception as e:
        print(wl, setting, "FAILED", r.stdout[-300:], r.stderr[-600:], flush=True)
