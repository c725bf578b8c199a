"""Small run touching every kernel, for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cuda_path_tracer_b200 as pt
from cuda_path_tracer_b200 import DisplayBufferType as DB
for sd, method in [(pt.bunny_scene(pt.bunny_like(2), 75, 41), pt.GPUMethod.megakernel),
                   (pt.three_balls(33, 17), pt.GPUMethod.streaming)]:
    w, h = sd.resolution
    tr = pt.PathTracer(max_depth=6)
    tr.current_gpu_method = method
    tr.max_iterations = 100
    tr.create_buffers((w, h), sd)
    tr.render(sd.camera, 3)
    tr.path_trace(sd.camera)
    tr.atrous_denoiser.filter_size = 10
    tr.denoise()
    for k in (DB.final, DB.color, DB.normal, DB.depth):
        tr.send_to_preview(type=k)
    tr.download(DB.denoised)
    tr.resize_image((40, 24))
    tr.render(sd.camera, 2)
    print(method, tr.stats().rays, float(tr.download(DB.color).mean()))
    sc = pt.Scene.from_description(sd)
    rays = np.zeros((1000, 8), np.float32); rays[:, 3] = 1e-4; rays[:, 7] = 3e38
    rays[:, 4:7] = np.random.default_rng(0).normal(size=(1000, 3)); rays[:, 2] = 1.0
    print((sc.trace_batch(rays)["t"] > 0).sum())
print("sanitize target OK")
