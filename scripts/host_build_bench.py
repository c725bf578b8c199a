"""B2 baseline: host-side BVH build, ours (OpenMP) vs the reference's bvh_from_mesh (1 thread)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cuda_path_tracer_b200 as pt
from tests import ref_lib
rows = []
for name, mesh in [("bunny_5k", pt.bunny_like(4)), ("bunny_82k", pt.bunny_like(6)), ("bunny_1.3m", pt.bunny_like(8)),
                   ("terrain_10m", pt.heightfield(2236))]:
    sd = pt.SceneDescription(); sd.add_material("m", pt.Material.lambertian((.5, .5, .5)))
    sd.add_mesh("mesh", mesh); sd.add_mesh_object("mesh", pt.translate((0, 0, 0)), "m")
    t0 = time.perf_counter(); sc = pt.Scene.from_description(sd); dt = time.perf_counter() - t0
    info = sc.info
    row = {"mesh": name, "triangles": mesh.triangle_count, "ours_build_ms": info.build_ms, "ours_upload_ms": info.upload_ms,
           "ours_total_create_s": dt, "bvh_nodes": info.n_bvh_nodes, "bvh_depth": info.bvh_depth,
           "threads": os.cpu_count()}
    if ref_lib.have_ref_host() and mesh.triangle_count <= 2_000_000 or (ref_lib.have_ref_host() and "--big" in sys.argv):
        nodes, secs = ref_lib.load_ref_host().bvh_from_mesh(mesh.positions, mesh.indices)
        row.update(ref_build_s=secs, ref_nodes=len(nodes), speedup=secs / (info.build_ms * 1e-3))
    rows.append(row); print(json.dumps(row), flush=True)
