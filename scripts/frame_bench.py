"""Interactive-frame configs (BASELINE.json configs[0] and [2]): ours vs the reference CUDA build.
A frame = restart + spp x path_trace (+ denoise) + tonemap to a host RGBA8 image, timed by wall
clock through each implementation's own API; min and median over the repetitions are reported
(the reference synchronises with the host and allocates inside thrust every bounce, so its
frame time is noisy on a shared box — its MIN is the number compared against)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import cuda_path_tracer_b200 as pt
from tests import ref_lib


def ours(sd, w, h, spp, depth, filter_size, reps=30):
    tr = pt.PathTracer(max_depth=depth)
    tr.max_iterations = 1 << 30
    tr.create_buffers((w, h), sd)
    tr.atrous_denoiser.filter_size = max(1, filter_size)

    import torch
    pinned = torch.empty((h, w, 4), dtype=torch.uint8, pin_memory=True).numpy()   # D2H target

    def frame():
        tr.restart()
        tr.render(sd.camera, spp)
        if filter_size:
            tr.denoise()
        return tr.send_to_preview(out=pinned)

    for _ in range(5):
        frame()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        frame()
        ts.append(time.perf_counter() - t0)
    tr.reset_stats()
    frame()
    rays = int(tr.stats().rays)
    dn = None
    if filter_size:
        tr.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            tr.denoise()
        tr.synchronize()
        dn = (time.perf_counter() - t0) / reps
    return min(ts), float(np.median(ts)), rays, dn


def ref(sd, w, h, spp, depth, filter_size, reps=15):
    R = ref_lib.load_ref_cuda()
    rt = R.tracer(sd, w, h, depth)
    gpu_ms, dn_ms, ts, rays = [], [], [], 0
    for i in range(reps + 3):
        t0 = time.perf_counter()
        rt.restart()
        ms, rays = rt.render_timed(sd.camera, spp, depth)
        dms = rt.denoise(filter_size) if filter_size else 0.0
        rt.preview(0)
        if i >= 3:
            ts.append(time.perf_counter() - t0)
            gpu_ms.append(ms)
            dn_ms.append(dms)
    return min(ts), float(np.median(ts)), rays, min(gpu_ms), (min(dn_ms) if filter_size else None)


for name, sd, spp, depth, fs in [
    ("config0_three_balls_800x800_1spp_d5", pt.three_balls(800, 800), 1, 5, 0),
    ("config2_bunny_1080p_1spp_denoise5", pt.bunny_scene(pt.bunny_like(4), 1920, 1080), 1, 8, 16),
]:
    w, h = sd.resolution
    o = ours(sd, w, h, spp, depth, fs)
    row = {"ours_frame_ms_min": o[0] * 1e3, "ours_frame_ms_median": o[1] * 1e3, "ours_rays": o[2],
           "ours_mrays_s": o[2] / o[0] * 1e-6, "ours_denoise_ms": None if o[3] is None else o[3] * 1e3}
    if ref_lib.have_ref_cuda():
        r = ref(sd, w, h, spp, depth, fs)
        row.update({"ref_frame_ms_min": r[0] * 1e3, "ref_frame_ms_median": r[1] * 1e3, "ref_rays": r[2],
                    "ref_render_gpu_ms_min": r[3], "ref_denoise_ms": r[4],
                    "speedup_frame_min_vs_min": r[0] / o[0]})
    print(name, json.dumps(row), flush=True)
