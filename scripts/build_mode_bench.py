"""Scene creation + render rate for the two BVH builders: host SAH (default) and device LBVH
(PT_BUILD=lbvh), on the 2.6-M-triangle bunny scene and the 10-M-triangle terrain."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cuda_path_tracer_b200 as pt

for name, sd, spp in [("bunny_1m 1080p", pt.bunny_scene(pt.bunny_like(8), 1920, 1080), 16),
                      ("terrain 10M 4K", pt.terrain_scene(2236, 3840, 2160), 4)]:
    w, h = sd.resolution
    for mode in ("sah", "lbvh"):
        if mode == "lbvh":
            os.environ["PT_BUILD"] = "lbvh"
        else:
            os.environ.pop("PT_BUILD", None)
        t0 = time.perf_counter()
        scene = pt.Scene.from_description(sd)
        create_s = time.perf_counter() - t0
        info = scene.info
        tr = pt.PathTracer(max_depth=8)
        tr.max_iterations = 1 << 30
        tr.create_buffers((w, h), scene)
        tr.render(sd.camera, spp); tr.synchronize()
        tr.restart(); tr.reset_stats()
        t0 = time.perf_counter()
        tr.render(sd.camera, spp); tr.synchronize()
        dt = time.perf_counter() - t0
        print(json.dumps({"scene": name, "builder": mode, "device_build": int(info.device_build),
                          "scene_create_s": round(create_s, 3), "build_ms": round(info.build_ms, 1),
                          "upload_ms": round(info.upload_ms, 1), "bvh_nodes": int(info.n_bvh_nodes),
                          "bvh_depth": int(info.bvh_depth), "mrays_per_s": round(int(tr.stats().rays) / dt * 1e-6, 1)}),
              flush=True)
        del tr, scene
        torch.cuda.empty_cache()
