"""One interactive frame (1 spp + A-Trous 5 iterations + tonemap at 1080p) a few times, for an ncu
launch list: python scripts/frame_launches.py [frames]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cuda_path_tracer_b200 as pt
sd = pt.bunny_scene(pt.bunny_like(4), 1920, 1080)
tr = pt.PathTracer(max_depth=8)
tr.max_iterations = 1 << 30
tr.create_buffers((1920, 1080), sd)
tr.atrous_denoiser.filter_size = 16
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    tr.restart()
    tr.render(sd.camera, 1)
    tr.denoise()
    tr.send_to_preview()
