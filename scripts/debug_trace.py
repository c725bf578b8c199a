import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cuda_path_tracer_b200 as pt
from tests.oracle_lib import load_oracle
from tests.test_gpu_parity import _rays_for, _secondary
o = load_oracle()
for name in ["bunny", "terrain"]:
    sd = pt.bunny_scene(pt.bunny_like(3), 96, 54) if name == "bunny" else pt.terrain_scene(24, 64, 36)
    w, h = sd.resolution
    osc = o.scene(sd); scene = pt.Scene.from_description(sd)
    prim, rng = _rays_for(o, sd, w, h, n_random=20000)
    ref = osc.trace_batch(prim, 0)
    sec = _secondary(prim, ref, rng)
    for label, rays in [("primary", prim), ("secondary", sec)]:
        r0 = osc.trace_batch(rays, 0); r1 = osc.trace_batch(rays, 1); g = scene.trace_batch(rays)
        def bad(a, b):
            m = (a["t"] < 0) != (b["t"] < 0)
            both = (a["t"] > 0) & (b["t"] > 0)
            m |= both & (np.abs(a["t"] - b["t"]) > 1e-5 * np.abs(b["t"]))
            return m
        b_g0, b_g1, b_01 = bad(g, r0), bad(g, r1), bad(r0, r1)
        print(name, label, len(rays), "gpu!=bvh", b_g0.sum(), "gpu!=brute", b_g1.sum(), "bvh!=brute", b_01.sum())
        for i in np.flatnonzero(b_g1)[:6]:
            print("  ray", rays[i], "\n   gpu", g[i]["t"], g[i]["object"], g[i]["prim"], " bvh", r0[i]["t"], r0[i]["object"], r0[i]["prim"], " brute", r1[i]["t"], r1[i]["object"], r1[i]["prim"])
