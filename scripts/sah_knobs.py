"""SAH cost-model knobs on the CPU (no GPU): inner-node visits / triangle tests per ray of the binary
tree for PT_SAH_LEAF x PT_SAH_ISECT, and the traversal kernel's instruction estimate
71 x inner visits + 119 x triangle tests (SASS counts of traverse_kernel, profiles/README.md).
Usage: python scripts/sah_knobs.py [bunny|bunny_82k|bunny_1m|terrain_small]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cuda_path_tracer_b200 as pt
from cuda_path_tracer_b200.api import HostBVH

which = sys.argv[1] if len(sys.argv) > 1 else "bunny"
sd = {"bunny": lambda: pt.bunny_scene(pt.bunny_like(4), 1920, 1080),
      "bunny_82k": lambda: pt.bunny_scene(pt.bunny_like(6), 1920, 1080),
      "bunny_1m": lambda: pt.bunny_scene(pt.bunny_like(8), 1920, 1080),
      "terrain_small": lambda: pt.terrain_scene(700, 3840, 2160)}[which]()
rng = np.random.default_rng(0)


def bounce_rays(tris, n):
    t = tris[rng.integers(0, tris.shape[0], n)]
    a, b = rng.uniform(0, 1, n), rng.uniform(0, 1, n)
    flip = a + b > 1
    a[flip], b[flip] = 1 - a[flip], 1 - b[flip]
    p = t[:, 0:3] + a[:, None] * t[:, 4:7] + b[:, None] * t[:, 8:11]
    nrm = np.cross(t[:, 4:7], t[:, 8:11])
    nrm /= np.maximum(np.linalg.norm(nrm, axis=1, keepdims=True), 1e-30)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d = nrm + d
    d /= np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-30)
    r = np.zeros((n, 8), np.float32)
    r[:, 0:3] = p + 1e-4 * nrm
    r[:, 3] = 1e-4
    r[:, 4:7] = d
    r[:, 7] = np.finfo(np.float32).max
    return r


os.environ.pop("PT_SAH_LEAF", None)
os.environ.pop("PT_SAH_ISECT", None)
base = HostBVH(sd, wide=False)
rays = bounce_rays(base.arrays()[2], 200_000)
for leaf in (2, 4, 8):
    for isect in (0.5, 1.0, 1.7, 3.0):
        os.environ["PT_SAH_LEAF"], os.environ["PT_SAH_ISECT"] = str(leaf), str(isect)
        hb = HostBVH(sd, wide=False)
        st = hb.trace_stats(rays, wide=0)
        inner, tri = st["inner_per_ray"], st["tri_tests_per_ray"]
        print(json.dumps({"scene": which, "leaf_max": leaf, "isect_cost": isect, "nodes": int(hb.info.n_bvh_nodes),
                          "depth": int(hb.info.bvh_depth), "inner_per_ray": round(inner, 2), "tri_per_ray": round(tri, 2),
                          "leaf_visits_per_ray": round(st["leaves_per_ray"], 2), "max_stack": st["max_stack"],
                          "instr_estimate": round(71 * inner + 119 * tri)}), flush=True)
        hb.close()
