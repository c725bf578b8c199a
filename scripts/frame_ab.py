"""Device and wall time of the two interactive frame configs (bench.py's bench_frame), one line
each — for A/B runs under different environment settings:  PT_LANES=2 python scripts/frame_ab.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cuda_path_tracer_b200 as pt
import bench
torch.cuda.set_device(0)
tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("PT_"))
for name, sd, w, h, spp, depth, fs in [("three_balls_frame", pt.three_balls(800, 800), 800, 800, 1, 5, 0),
                                       ("interactive_frame", pt.bunny_scene(pt.bunny_like(4), 1920, 1080), 1920, 1080, 1, 8, 16)]:
    e = bench.bench_frame(name, sd, w, h, spp, depth, fs, reps=40)
    print(f"{tag or 'default':24s} {name:18s} device {e['value']:.4f} ms (min {e['value_min']:.4f})  wall {e['e2e']['value']:.4f} ms  "
          f"denoise {e.get('denoise_ms')}  launches {e['gpu_launches_per_frame']}", flush=True)
