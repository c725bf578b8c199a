"""Host BVH build times only (no device, no reference run): best of 3 per mesh.
usage: python scripts/host_build_quick.py [--small]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cuda_path_tracer_b200 as pt
from cuda_path_tracer_b200.api import HostBVH

meshes = [("terrain_10m", lambda: pt.heightfield(2236)), ("bunny_1.3m", lambda: pt.bunny_like(8)),
          ("bunny_82k", lambda: pt.bunny_like(6))]
if "--small" in sys.argv:
    meshes = meshes[1:]
for name, make in meshes:
    mesh = make()
    sd = pt.SceneDescription()
    sd.add_material("m", pt.Material.lambertian((.5, .5, .5)))
    sd.add_mesh("mesh", mesh)
    sd.add_mesh_object("mesh", pt.translate((0, 0, 0)), "m")
    best = 1e30
    for _ in range(3):
        hb = HostBVH(sd, wide=False)
        best = min(best, float(hb.info.build_ms))
        nodes, depth = int(hb.info.n_bvh_nodes), int(hb.info.bvh_depth)
        hb.close()
    print(json.dumps({"mesh": name, "triangles": mesh.triangle_count, "build_ms_best_of_3": round(best, 1),
                      "bvh_nodes": nodes, "bvh_depth": depth, "threads": os.cpu_count()}), flush=True)
