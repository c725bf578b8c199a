"""Interactive-frame configs (BASELINE.json configs[0] and [2]): ours vs the reference CUDA build."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cuda_path_tracer_b200 as pt
from cuda_path_tracer_b200 import DisplayBufferType as DB
from tests import ref_lib

def ours(sd, w, h, spp, depth, filter_size, reps=20):
    tr = pt.PathTracer(max_depth=depth)
    tr.max_iterations = 1 << 30
    tr.create_buffers((w, h), sd)
    tr.atrous_denoiser.filter_size = filter_size
    def frame():
        tr.restart()
        tr.render(sd.camera, spp)
        if filter_size: tr.denoise()
        return tr.send_to_preview()
    for _ in range(3): frame()
    t0 = time.perf_counter()
    for _ in range(reps): img = frame()
    dt = (time.perf_counter() - t0) / reps
    tr.reset_stats(); frame(); st = tr.stats()
    # denoise alone
    dn = None
    if filter_size:
        tr.synchronize(); t0 = time.perf_counter()
        for _ in range(reps): tr.denoise()
        tr.synchronize(); dn = (time.perf_counter() - t0) / reps
    return dt, int(st.rays), dn, img

def ref(sd, w, h, spp, depth, filter_size, reps=5):
    R = ref_lib.load_ref_cuda()
    rt = R.tracer(sd, w, h, depth)
    def frame():
        rt.restart()
        ms, rays = rt.render_timed(sd.camera, spp, depth)
        dms = rt.denoise(filter_size) if filter_size else 0.0
        img = rt.preview(0)
        return ms, rays, dms
    for _ in range(2): frame()
    t0 = time.perf_counter()
    for _ in range(reps): ms, rays, dms = frame()
    dt = (time.perf_counter() - t0) / reps
    return dt, rays, dms * 1e-3

out = {}
for name, sd, spp, depth, fs in [
    ("config0_three_balls_800x800_1spp_d5", pt.three_balls(800, 800), 1, 5, 0),
    ("config2_bunny_1080p_1spp_denoise5", pt.bunny_scene(pt.bunny_like(4), 1920, 1080), 1, 8, 16),
]:
    w, h = sd.resolution
    o = ours(sd, w, h, spp, depth, fs)
    row = {"ours_frame_ms": o[0] * 1e3, "ours_rays": o[1], "ours_mrays_s": o[1] / o[0] * 1e-6,
           "ours_denoise_ms": None if o[2] is None else o[2] * 1e3}
    if ref_lib.have_ref_cuda():
        r = ref(sd, w, h, spp, depth, fs)
        row.update({"ref_frame_ms": r[0] * 1e3, "ref_rays": r[1], "ref_mrays_s": r[1] / r[0] * 1e-6,
                    "ref_denoise_ms": r[2] * 1e3 if fs else None, "speedup_frame": r[0] / o[0]})
    out[name] = row
    print(name, json.dumps(row), flush=True)
