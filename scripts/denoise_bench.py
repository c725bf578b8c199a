import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cuda_path_tracer_b200 as pt
sd = pt.bunny_scene(pt.bunny_like(4), 1920, 1080)
tr = pt.PathTracer(max_depth=8, profile=True)
tr.max_iterations = 1 << 30
tr.create_buffers((1920, 1080), sd)
tr.render(sd.camera, 1)
tr.atrous_denoiser.filter_size = 16
for _ in range(3): tr.denoise()
tr.synchronize(); tr.reset_stats()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
for _ in range(n): tr.denoise()
st = tr.stats()
print("denoise ms per call (5 iterations + prepare):", st.ms_denoise / n)
