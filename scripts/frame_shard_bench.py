"""Interactive frame (BASELINE.json configs[2]: 1 spp + A-Trous, 1080p) sharded by ROW BANDS over N
GPUs: every rank renders its band, exchanges the denoiser halo rows with its neighbours, denoises
its band, tonemaps it and sends the RGBA8 rows to rank 0.
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/frame_shard_bench.py
Prints one JSON line: frame time (device events, max over ranks), the same on one GPU, and
whether rank 0's assembled frame equals the unsharded frame."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import cuda_path_tracer_b200 as pt
from cuda_path_tracer_b200 import DisplayBufferType as DB
from cuda_path_tracer_b200 import sharding

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
W, H, depth, spp, reps = 1920, 1080, 8, 1, 30
sd = pt.bunny_scene(pt.bunny_like(4), W, H)
scene = pt.Scene.from_description(sd, device=local)
stream = torch.cuda.Stream()


def make(rows=None):
    tr = pt.PathTracer(max_depth=depth, stream=stream.cuda_stream)
    tr.max_iterations = 1 << 30
    tr.create_buffers((W, H), scene)
    tr.atrous_denoiser.filter_size = 16
    buf = torch.zeros(2 * H * W * 4, dtype=torch.float32, device="cuda")
    tr.bind_sums(buf.data_ptr())
    if rows:
        tr.set_rows(*rows)
    return tr, buf


def frame(tr, buf, img, sharded):
    buf.zero_()
    tr.set_sample_count(0)
    tr.render_range(sd.camera, 0, spp)
    tr.set_sample_count(spp)
    if sharded:
        sharding.exchange_halo(buf, W, H, halo)
    tr.denoise()
    tr.send_to_preview(dev_pbo=img.data_ptr())
    if sharded:
        sharding.gather_rows(img, H, dst=0)


def timed(tr, buf, img, sharded):
    with torch.cuda.stream(stream):
        for _ in range(5):
            frame(tr, buf, img, sharded)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            frame(tr, buf, img, sharded)
        e1.record(stream)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


img_full = torch.zeros(H, W, 4, dtype=torch.uint8, device="cuda")
full, fbuf = make()
halo = full.halo_rows()
ms_one = timed(full, fbuf, img_full, False)
out = {"config": "bunny 1080p 1 spp + a-trous 5 iterations", "n_gpus": world, "frame_ms_one_gpu": ms_one}
if world > 1:
    band, bbuf = make(sharding.band_rows(rank, world, H))
    img = torch.zeros(H, W, 4, dtype=torch.uint8, device="cuda")
    out["frame_ms_sharded"] = timed(band, bbuf, img, True)
    out["speedup"] = ms_one / out["frame_ms_sharded"]
    out["halo_rows"] = halo
    out["halo_bytes_per_neighbour"] = halo * W * 32
    if rank == 0:
        d = (img.int() - img_full.int()).abs()
        out["max_abs_rgba_diff_vs_one_gpu"] = int(d.max())
        out["pixels_differing"] = int((d.amax(dim=2) > 0).sum())
if rank == 0:
    print(json.dumps(out), flush=True)
if world > 1:
    dist.destroy_process_group()
