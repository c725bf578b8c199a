"""BASELINE.json configs[4]: synthetic 10M-triangle scene, 3840x2160, 1024 spp progressive, STRONG
scaling by sample range over N GPUs (rank r renders iterations iteration_range(r, N, 1024)), one
NCCL reduce of the sums, mean + tonemap on rank 0.  Rank 0 then renders the same 1024 spp alone and
reports the image difference (should be float re-association only).
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/progressive_shard_bench.py [--spp 1024] [--n 2236]"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import cuda_path_tracer_b200 as pt
from cuda_path_tracer_b200 import DisplayBufferType as DB
from cuda_path_tracer_b200 import sharding

ap = argparse.ArgumentParser()
ap.add_argument("--spp", type=int, default=1024)
ap.add_argument("--n", type=int, default=2236)
ap.add_argument("--skip-single", action="store_true")
args = ap.parse_args()
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
W, H, depth = 3840, 2160, 8
sd = pt.terrain_scene(args.n, W, H, args.spp)
t0 = time.perf_counter()
scene = pt.Scene.from_description(sd, device=local)
scene_s = time.perf_counter() - t0
stream = torch.cuda.Stream()
tr = pt.PathTracer(max_depth=depth, stream=stream.cuda_stream)
tr.max_iterations = 1 << 30
tr.create_buffers((W, H), scene)
sums = torch.zeros(2 * W * H * 4, dtype=torch.float32, device="cuda")
tr.bind_sums(sums.data_ptr())


def render(first, n, reduce):
    with torch.cuda.stream(stream):
        sums.zero_()
        tr.reset_stats()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        tr.render_range(sd.camera, first, n)
        if reduce and world > 1:
            sharding.reduce_sums(sums, dst=0)
        e1.record(stream)
        torch.cuda.synchronize()
    return e0.elapsed_time(e1), int(tr.stats().rays)


first, n = sharding.iteration_range(rank, world, args.spp)
render(first, min(n, 16), True)                       # warm-up
ms, rays = render(first, n, True)
t = torch.tensor([ms, float(rays)], dtype=torch.float64, device="cuda")
if world > 1:
    tmax, tsum = t.clone(), t.clone()
    dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    ms, rays = float(tmax[0]), int(tsum[1])
out = {"config": f"terrain {sd.meshes[sorted(sd.meshes)[0]].triangle_count} triangles, {W}x{H}, {args.spp} spp progressive, depth {depth}",
       "n_gpus": world, "sharded_ms": ms, "rays": rays, "mrays_per_s": rays / ms * 1e-3,
       "spp_per_s": args.spp / ms * 1e3, "scene_build_s_per_rank": scene_s}
if rank == 0:
    tr.set_sample_count(args.spp)
    img_n = tr.download(DB.color)
    if not args.skip_single and world > 1:
        ms1, rays1 = render(0, args.spp, False)
        tr.set_sample_count(args.spp)
        img_1 = tr.download(DB.color)
        out.update(one_gpu_ms=ms1, speedup=ms1 / ms, efficiency=ms1 / ms / world,
                   rmse_vs_one_gpu=float(np.sqrt(np.mean((img_n - img_1) ** 2))),
                   max_abs_vs_one_gpu=float(np.abs(img_n - img_1).max()), rays_one_gpu=rays1)
    print(json.dumps(out), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
