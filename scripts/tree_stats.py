"""Tree-layout analysis on the CPU (no GPU): node visits / triangle tests per ray for the SAH
binary tree, the compressed 8-wide tree derived from it and the LBVH tree, on primary rays and on
diffuse bounce rays of the bench scenes (host walks in the device kernels' visiting order,
pt_host_bvh_trace_stats).  Usage: python scripts/tree_stats.py [bunny|bunny_82k|terrain_small]"""
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
w, h = sd.resolution
rng = np.random.default_rng(0)


def camera_rays(n):
    """Pinhole rays of the scene camera (ray_gen.cu:34-61 semantics, numpy float64 is fine here)."""
    cam = sd.camera
    x, y = rng.uniform(0, w, n), rng.uniform(0, h, n)
    aspect = w / h
    vh = 2.0 * np.tan(cam.vfov / 2.0)
    vw = aspect * vh
    u, v = x / (w - 1), (h - y) / (h - 1)
    d_cam = np.stack([-vw / 2 + u * vw, -vh / 2 + v * vh, -np.ones(n)], axis=1)
    qw, qx, qy, qz = cam.rotation
    R = np.array([[1 - 2 * (qy * qy + qz * qz), 2 * (qx * qy - qz * qw), 2 * (qx * qz + qy * qw)],
                  [2 * (qx * qy + qz * qw), 1 - 2 * (qx * qx + qz * qz), 2 * (qy * qz - qx * qw)],
                  [2 * (qx * qz - qy * qw), 2 * (qy * qz + qx * qw), 1 - 2 * (qx * qx + qy * qy)]])
    d = d_cam @ R.T
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    r = np.zeros((n, 8), np.float32)
    r[:, 0:3] = np.asarray(cam.position, np.float32)
    r[:, 3] = 1e-4
    r[:, 4:7] = d
    r[:, 7] = np.finfo(np.float32).max
    return r


def bounce_rays(tris, n):
    """Diffuse rays leaving random points of random triangles (cosine-ish hemisphere)."""
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


n = 200_000
_sah = HostBVH(sd, wide=False)
trees = {"sah_binary": (_sah, 0), "sah_virtual4": (_sah, 2), "sah_wide8": (HostBVH(sd, wide=True), 1),
         "lbvh_binary": (HostBVH(sd, lbvh=True), 0)}
prim = camera_rays(n)
sec = bounce_rays(trees["sah_binary"][0].arrays()[2], n)
for kind, rays in (("primary", prim), ("diffuse bounce", sec)):
    for name, (hb, wide) in trees.items():
        st = hb.trace_stats(rays, wide=wide)
        st.update(scene=which, rays_kind=kind, tree=name, nodes=int(hb.info.n_bvh8_nodes if wide == 1 else hb.info.n_bvh_nodes))
        print(json.dumps(st), flush=True)
