"""CPU checks of the host builder (csrc/bvh_build.cpp): both trees are structurally valid, and the
compressed 8-wide tree — decoded with the DEVICE's arithmetic (byte dropped into the mantissa of
2^23, one FMA per plane; csrc/wavefront.cu node8_test), restated here in float32 — finds the
closest hits the oracle's brute-force reference traversal finds (intersection semantics:
reference path_tracer.cu:36-76, intersections.cuh:49-85)."""
import numpy as np
import pytest

import cuda_path_tracer_b200 as pt
from cuda_path_tracer_b200.api import HostBVH
from tests.oracle_lib import load_oracle

f32 = np.float32


def _mesh_scene(mesh, xf=None):
    sd = pt.SceneDescription()
    sd.add_material("m", pt.Material.lambertian((0.5, 0.5, 0.5)))
    sd.add_mesh("mesh", mesh)
    sd.add_mesh_object("mesh", xf if xf is not None else pt.translate((0, 0, 0)), "m")
    return sd


def _flat_quads(n):
    """n x n axis-aligned quads in the plane y = 0: every box has zero thickness."""
    xs = np.linspace(-1, 1, n + 1, dtype=np.float32)
    gx, gz = np.meshgrid(xs, xs, indexing="ij")
    pos = np.stack([gx.ravel(), np.zeros(gx.size, np.float32), gz.ravel()], axis=1)
    idx = []
    for i in range(n):
        for j in range(n):
            a, b, c, d = i * (n + 1) + j, (i + 1) * (n + 1) + j, (i + 1) * (n + 1) + j + 1, i * (n + 1) + j + 1
            idx += [a, b, c, a, c, d]
    return pt.Mesh(pos, np.array(idx, dtype=np.uint32))


MESHES = {
    "one_triangle": lambda: pt.Mesh(np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32),
                                    np.array([0, 1, 2], np.uint32)),
    "three_coincident": lambda: pt.Mesh(np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32),
                                        np.array([0, 1, 2] * 5, np.uint32)),
    "bunny_320": lambda: pt.bunny_like(2),
    "bunny_5k": lambda: pt.bunny_like(4),
    "bunny_20k": lambda: pt.bunny_like(5),
    "terrain_5k": lambda: pt.heightfield(50),
    "flat_quads": lambda: _flat_quads(24),
}


@pytest.mark.parametrize("name", sorted(MESHES))
@pytest.mark.parametrize("wide", [False, True])
def test_trees_are_structurally_valid(name, wide):
    hb = HostBVH(_mesh_scene(MESHES[name]()), wide=wide)
    i = hb.info
    assert hb.violations() == 0
    assert int(i.n_bvh_nodes) >= 1
    if wide:
        assert int(i.n_bvh8_nodes) >= 1 and 1 <= int(i.bvh8_depth) <= 32
        assert int(i.n_bvh8_nodes) <= max(1, int(i.n_bvh_nodes))
    else:
        assert int(i.n_bvh8_nodes) == 0


def test_instances_are_baked_into_one_tree():
    sd = pt.bunny_scene(pt.bunny_like(3), 64, 36)          # two instances of one mesh + a sphere
    hb = HostBVH(sd)
    assert int(hb.info.n_world_triangles) == 2 * int(hb.info.n_triangles)
    assert hb.violations() == 0


# ---------------------------------------------------------------- device arithmetic, restated
def _fma(a, b, c):
    return f32(np.float64(a) * np.float64(b) + np.float64(c))


def _safe_inv(x):
    e = f32(8.271806125530277e-25)
    x = f32(x)
    return f32(1.0) / (x if abs(x) > e else f32(np.copysign(e, x)))


def _tri_test(tris, slot, o, d, tmin, tbest):
    t0, t1, t2 = tris[slot, 0:3], tris[slot, 4:7], tris[slot, 8:11]

    def cross(a, b):
        return np.array([a[1] * b[2] - b[1] * a[2], a[2] * b[0] - b[2] * a[0], a[0] * b[1] - b[0] * a[1]], f32)

    def dot(a, b):
        return f32(f32(f32(a[0] * b[0]) + f32(a[1] * b[1])) + f32(a[2] * b[2]))

    h = cross(d, t2)
    a = dot(t1, h)
    if -1e-7 < a < 1e-7:
        return None
    f = f32(1.0) / a
    s = (o - t0).astype(f32)
    u = f32(f * dot(s, h))
    if u < 0 or u > 1:
        return None
    q = cross(s, t1)
    v = f32(f * dot(d, q))
    if v < 0 or f32(u + v) > 1:
        return None
    t = f32(f * dot(t2, q))
    return t if tmin <= t <= tbest else None


def _trace_wide(nodes8, tris, o, d, tmin, tmax):
    """Closest hit through the wide tree: slab arithmetic as node8_test, unordered stack walk."""
    o, d = o.astype(f32), d.astype(f32)
    idir = np.array([_safe_inv(x) for x in d], f32)
    tbest, best = f32(tmax), -1
    stack = [0]
    visited = 0
    while stack:
        w = nodes8[stack.pop()]
        visited += 1
        p = w[0:3].view(f32)
        eb = [(int(w[3]) >> (8 * a)) & 0xFF for a in range(3)]
        imask = int(w[3]) >> 24
        A = [f32(np.array([e << 23], np.uint32).view(f32)[0] * idir[a]) for a, e in enumerate(eb)]
        B = [_fma(f32(-8388608.0), A[a], f32(f32(p[a] - o[a]) * idir[a])) for a in range(3)]
        meta = w[6:8].view(np.uint8)
        q = w[8:20].view(np.uint8).reshape(6, 8)
        rel = 0
        for s in range(8):
            m = int(meta[s])
            inner = (imask >> s) & 1
            if m:
                cmin, cmax = f32(tmin), tbest
                for a in range(3):
                    lo = np.array([0x4B000000 | (int(q[a, s]) << 8)], np.uint32).view(f32)[0]
                    hi = np.array([0x4B000000 | (int(q[3 + a, s]) << 8)], np.uint32).view(f32)[0]
                    tlo, thi = _fma(lo, A[a], B[a]), _fma(hi, A[a], B[a])
                    near, far = (thi, tlo) if idir[a] < 0 else (tlo, thi)
                    cmin, cmax = max(cmin, near), min(cmax, far)
                if f32(cmax * f32(1.0000004)) >= cmin:
                    if inner:
                        stack.append(int(w[4]) + rel)
                    else:
                        count = {1: 1, 3: 2, 7: 3}[m >> 5]
                        for k in range(count):
                            slot = int(w[5]) + (m & 31) + k
                            t = _tri_test(tris, slot, o, d, f32(tmin), tbest)
                            if t is not None:
                                tbest, best = t, slot
            rel += inner
    return tbest, best, visited


@pytest.mark.parametrize("name", ["bunny_5k", "terrain_5k", "flat_quads"])
def test_wide_tree_with_device_arithmetic_matches_oracle(name):
    mesh = MESHES[name]()
    xf = pt.compose(pt.translate((0.3, -0.2, -3.0)), pt.rotate(0.7, (0.2, 1.0, 0.1)), pt.scale((1.3, 0.9, 1.1)))
    sd = _mesh_scene(mesh, xf)
    hb = HostBVH(sd)
    _, nodes8, tris = hb.arrays()
    rng = np.random.default_rng(7)
    n = 160
    # rays from a shell around the object towards jittered points near it (most hit), a few
    # axis-parallel ones and a few that start inside the bounds
    lo = tris[:, 0:3].min(axis=0) - 0.5
    hi = tris[:, 0:3].max(axis=0) + 0.5
    cen, rad = 0.5 * (lo + hi), 0.5 * np.linalg.norm(hi - lo)
    dirs = rng.normal(size=(n, 3))
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    org = cen + dirs * rad * rng.uniform(0.05, 2.0, size=(n, 1))
    tgt = cen + rng.uniform(-0.5, 0.5, size=(n, 3)) * (hi - lo)
    d = tgt - org
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d[:12] = np.eye(3)[rng.integers(0, 3, 12)] * rng.choice([-1.0, 1.0], size=(12, 1))
    rays = np.concatenate([org, np.full((n, 1), 1e-4), d, np.full((n, 1), 3.4e38)], axis=1).astype(np.float32)
    ref = load_oracle().scene(sd).trace_batch(rays)
    hits = 0
    coord = float(np.abs(tris[:, 0:3]).max())
    for i in range(n):
        t, slot, visited = _trace_wide(nodes8, tris, rays[i, 0:3], rays[i, 4:7], rays[i, 3], rays[i, 7])
        if ref["t"][i] < 0:
            assert slot < 0, (i, t, slot)
            continue
        hits += 1
        assert slot >= 0, (i, ref["t"][i])
        # 1e-5 relative (north star); hits a few millimetres from the origin only reach the
        # absolute rounding of world-space coordinates (the reference intersects in object space)
        assert abs(t - ref["t"][i]) <= max(1e-5 * abs(ref["t"][i]), 4e-7 * coord), (i, t, ref["t"][i])
        prim = int(tris[slot, 3:4].view(np.uint32)[0])
        if prim != ref["prim"][i]:      # ties between coincident/adjacent triangles only
            assert abs(t - ref["t"][i]) <= 1e-6 * abs(ref["t"][i])
    assert hits > n // 4


# ---------------------------------------------------------------- quantised nodes
# The 32-byte nodes traverse_kernel / finish_kernel read (DevScene::qnodes, wavefront.cu trav_init<QN>
# / trav_inner<QN>): the reference walks exact boxes un-culled (path_tracer.cu:36-84); a conservative
# box may only change which subtrees are skipped, never the closest hit.
_Q_SRC = [(0, 1), (2, 3), (8, 9), (4, 5), (6, 7), (10, 11)]   # plane slots of the 64-byte node per word
_Q_AXIS = [0, 1, 2, 0, 1, 2]


@pytest.mark.parametrize("name", ["bunny_320", "bunny_5k", "bunny_20k", "terrain_5k", "flat_quads", "three_coincident"])
def test_quantised_nodes_enclose_the_exact_boxes_with_a_cell_in_reserve(name):
    hb = HostBVH(_mesh_scene(MESHES[name]()), wide=False)
    nodes, _, _ = hb.arrays()
    q, org, cell = hb.quantised()
    assert q.shape == (nodes.shape[0], 8)
    assert np.all(cell > 0)
    # child references are the 64-byte node's, bit for bit
    assert np.array_equal(q[:, 6:8], nodes[:, 12:14].view(np.uint32))
    org64, cell64 = org.astype(np.float64), cell.astype(np.float64)
    for k in range(6):
        a = _Q_AXIS[k]
        lo = org64[a] + (q[:, k] & 0xFFFF).astype(np.float64) * cell64[a]
        hi = org64[a] + (q[:, k] >> 16).astype(np.float64) * cell64[a]
        ex_lo = nodes[:, _Q_SRC[k][0]].astype(np.float64)
        ex_hi = nodes[:, _Q_SRC[k][1]].astype(np.float64)
        # at least one whole cell outside (the device's rounding takes half of it), at most two
        assert np.all(ex_lo - lo >= cell64[a] * (1 - 1e-9)), k
        assert np.all(hi - ex_hi >= cell64[a] * (1 - 1e-9)), k
        assert np.all(ex_lo - lo <= cell64[a] * (2 + 1e-9)), k
        assert np.all(hi - ex_hi <= cell64[a] * (2 + 1e-9)), k


def _q_float(h16):
    return np.array([0x4B000000 | int(h16)], np.uint32).view(f32)[0]


def _trace_quantised(q, org, cell, tris, o, d, tmin, tmax):
    """Closest hit through the quantised nodes with the device's arithmetic (trav_init<QN>,
    trav_inner<QN>, trav_leaf): near/far planes picked by the direction signs, nearer child first."""
    o, d = o.astype(f32), d.astype(f32)
    inv = np.array([_safe_inv(x) for x in d], f32)
    A = np.array([f32(cell[a] * inv[a]) for a in range(3)], f32)
    B = np.array([_fma(f32(8388608.0), A[a], -f32(f32(org[a] - o[a]) * inv[a])) for a in range(3)], f32)
    near_lo = [A[a] >= 0 for a in range(3)]
    tbest, best, visited = f32(tmax), -1, 0
    stack, node = [], 0
    SENT = 0x7FFFFFFF

    def child(w, base):
        cmin, cmax = f32(tmin), tbest
        for a in range(3):
            lo, hi = _q_float(w[base + a] & 0xFFFF), _q_float(w[base + a] >> 16)
            tn = _fma(lo if near_lo[a] else hi, A[a], -B[a])
            tf = _fma(hi if near_lo[a] else lo, A[a], -B[a])
            cmin, cmax = max(cmin, tn), min(cmax, tf)
        return bool(f32(cmax * f32(1.0000004)) >= cmin), cmin

    while node != SENT:
        if node >= 0:
            w = q[node]
            visited += 1
            t0, m0 = child(w, 0)
            t1, m1 = child(w, 3)
            c0, c1 = int(np.int32(w[6])), int(np.int32(w[7]))
            if not t0 and not t1:
                node = stack.pop() if stack else SENT
            else:
                swap = t1 and (not t0 or m1 < m0)
                node = c1 if swap else c0
                if t0 and t1:
                    stack.append(c0 if swap else c1)
        else:
            code = ~node & 0xFFFFFFFF
            first, count = code >> 3, (code & 7) + 1
            for k in range(count):
                t = _tri_test(tris, first + k, o, d, f32(tmin), tbest)
                if t is not None:
                    tbest, best = t, first + k
            node = stack.pop() if stack else SENT
    return tbest, best, visited


@pytest.mark.parametrize("name", ["bunny_5k", "terrain_5k", "flat_quads"])
def test_quantised_nodes_with_device_arithmetic_match_oracle(name):
    mesh = MESHES[name]()
    xf = pt.compose(pt.translate((0.3, -0.2, -3.0)), pt.rotate(0.7, (0.2, 1.0, 0.1)), pt.scale((1.3, 0.9, 1.1)))
    sd = _mesh_scene(mesh, xf)
    hb = HostBVH(sd, wide=False)
    _, _, tris = hb.arrays()
    q, org, cell = hb.quantised()
    rng = np.random.default_rng(11)
    n = 160
    lo = tris[:, 0:3].min(axis=0) - 0.5
    hi = tris[:, 0:3].max(axis=0) + 0.5
    cen, rad = 0.5 * (lo + hi), 0.5 * np.linalg.norm(hi - lo)
    dirs = rng.normal(size=(n, 3))
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    # origins from inside the bounds out to 40 radii away (the far ones exercise the part of the
    # error that grows with the distance and is covered by the widened comparison)
    org_r = cen + dirs * rad * np.concatenate([rng.uniform(0.05, 2.0, size=(n - 24, 1)), rng.uniform(10.0, 40.0, size=(24, 1))])
    tgt = cen + rng.uniform(-0.5, 0.5, size=(n, 3)) * (hi - lo)
    d = tgt - org_r
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d[:12] = np.eye(3)[rng.integers(0, 3, 12)] * rng.choice([-1.0, 1.0], size=(12, 1))
    rays = np.concatenate([org_r, np.full((n, 1), 1e-4), d, np.full((n, 1), 3.4e38)], axis=1).astype(np.float32)
    ref = load_oracle().scene(sd).trace_batch(rays)
    hits = 0
    coord = float(np.abs(tris[:, 0:3]).max())
    for i in range(n):
        t, slot, _ = _trace_quantised(q, org, cell, tris, rays[i, 0:3], rays[i, 4:7], rays[i, 3], rays[i, 7])
        if ref["t"][i] < 0:
            assert slot < 0, (i, t, slot)
            continue
        hits += 1
        assert slot >= 0, (i, ref["t"][i])
        assert abs(t - ref["t"][i]) <= max(1e-5 * abs(ref["t"][i]), 4e-7 * coord * max(1.0, np.linalg.norm(rays[i, 0:3] - cen) / rad)), (i, t, ref["t"][i])
        prim = int(tris[slot, 3:4].view(np.uint32)[0])
        if prim != ref["prim"][i]:      # ties between coincident/adjacent triangles only
            assert abs(t - ref["t"][i]) <= 1e-6 * abs(ref["t"][i])
    assert hits > n // 4


@pytest.mark.parametrize("name", ["bunny_320", "bunny_5k", "bunny_20k", "terrain_5k", "flat_quads", "three_coincident"])
def test_lbvh_restatement_is_structurally_valid(name):
    """The host restatement of the device LBVH builder (lbvh.h shared with lbvh.cu): every triangle
    referenced once, every stored child box contains its content, depth fits the traversal stack.
    Scenes it declines (<= 4 triangles) fall back to the SAH builder."""
    hb = HostBVH(_mesh_scene(MESHES[name]()), lbvh=True)
    assert hb.violations() == 0
    assert 1 <= int(hb.info.bvh_depth) <= 61
    nodes, _, tris = hb.arrays()
    assert nodes.shape[0] == int(hb.info.n_bvh_nodes) and tris.shape[0] == int(hb.info.n_bvh_triangles)


def test_lbvh_orders_triangles_along_the_morton_curve():
    mesh = MESHES["terrain_5k"]()
    hb = HostBVH(_mesh_scene(mesh), lbvh=True)
    _, _, tris = hb.arrays()
    prim = tris[:, 3].copy().view(np.uint32)
    assert sorted(prim.tolist()) == list(range(mesh.triangle_count))       # a permutation
    cen = tris[:, 0:3] + (tris[:, 4:7] + tris[:, 8:11]) / 3.0
    step = np.linalg.norm(np.diff(cen, axis=0), axis=1)
    rng = np.random.default_rng(0)
    shuffled = np.linalg.norm(np.diff(cen[rng.permutation(len(cen))], axis=0), axis=1)
    assert step.mean() < 0.2 * shuffled.mean()                              # neighbours stay close


def test_trace_stats_walks_agree_across_tree_layouts():
    """The host walks of the analysis tool (binary, virtual 4-wide, compressed 8-wide, LBVH) find a
    hit for exactly the rays the oracle's reference traversal hits, and the wider layouts visit
    fewer nodes."""
    sd = _mesh_scene(MESHES["bunny_5k"](), pt.translate((0.0, 0.0, -3.0)))
    rng = np.random.default_rng(3)
    n = 4000
    o = rng.normal(size=(n, 3)) * 0.2
    d = np.array([0, 0, -1.0]) + rng.normal(size=(n, 3)) * 0.35
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([o, np.full((n, 1), 1e-4), d, np.full((n, 1), 3.4e38)], axis=1).astype(np.float32)
    ref_hits = float((load_oracle().scene(sd).trace_batch(rays)["t"] > 0).mean())
    sah, wide, lb = HostBVH(sd, wide=False), HostBVH(sd, wide=True), HostBVH(sd, lbvh=True)
    stats = {"binary": sah.trace_stats(rays, 0), "virtual4": sah.trace_stats(rays, 2),
             "wide8": wide.trace_stats(rays, 1), "lbvh": lb.trace_stats(rays, 0)}
    assert 0.1 < ref_hits < 1.0
    for name, st in stats.items():
        assert abs(st["hit_fraction"] - ref_hits) <= 2.0 / n, (name, st["hit_fraction"], ref_hits)
    assert stats["wide8"]["inner_per_ray"] < stats["virtual4"]["inner_per_ray"] < stats["binary"]["inner_per_ray"]
    assert stats["wide8"]["max_stack"] <= 32 and stats["binary"]["max_stack"] <= 60


_DIGEST_SCRIPT = r"""
import hashlib, sys
import numpy as np
import cuda_path_tracer_b200 as pt
from cuda_path_tracer_b200.api import HostBVH
sd = pt.SceneDescription()
sd.add_material("m", pt.Material.lambertian((.5, .5, .5)))
sd.add_mesh("mesh", pt.heightfield(int(sys.argv[1])))
sd.add_mesh_object("mesh", pt.translate((0, 0, 0)), "m")
for wide in (False, True):
    hb = HostBVH(sd, wide=wide)
    nodes, nodes8, tris = hb.arrays()
    h = hashlib.md5()
    for a in (nodes, tris) + ((nodes8,) if wide else ()):
        h.update(np.ascontiguousarray(a).tobytes())
    print(h.hexdigest(), int(hb.info.n_bvh_nodes), int(hb.info.bvh_depth))
"""


def test_host_tree_does_not_depend_on_the_thread_count(tmp_path):
    """Node slots are handed out by range and the flattener derives DFS indices from subtree
    counts, so the built arrays (binary nodes, triangle order, wide nodes) are byte-identical for
    any number of OpenMP threads — 90 000 triangles is enough for subtrees to become tasks."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for threads in ("1", "3", "8"):
        env = dict(os.environ, OMP_NUM_THREADS=threads, PYTHONPATH=root)
        r = subprocess.run([sys.executable, "-c", _DIGEST_SCRIPT, "212"], env=env, capture_output=True, text=True,
                           timeout=300)
        assert r.returncode == 0, r.stderr
        outs.append(r.stdout)
    assert outs[0] == outs[1] == outs[2], outs
    assert len(outs[0].split()) == 6


def _scene_check(sd):
    import ctypes as C
    d, keep = sd.to_desc()
    out4 = (C.c_uint64 * 4)()
    h = C.c_uint64(0)
    rc = pt.load_library().pt_host_scene_check(C.byref(d), out4, C.byref(h))
    assert rc == 0, pt.load_library().pt_last_error()
    return [int(x) for x in out4], int(h.value)


def test_sphere_group_trees_are_structurally_valid_and_the_fingerprint_tracks_the_description():
    """Host only: groups of more than 32 rigidly placed spheres get a tree whose leaves reference
    every sphere of the group exactly once inside nested boxes; small or non-rigid groups keep the
    reference's linear scan; the description fingerprint (progressive-state files) changes with
    the description and only with it."""
    sd = pt.many_spheres_scene(400, 64, 64)
    (n, nodes, violations, trees), h1 = _scene_check(sd)
    assert n == 401 and trees == 2 and nodes >= 2 * (200 // 4 // 2) and violations == 0
    (_, _, _, _), h1b = _scene_check(pt.many_spheres_scene(400, 64, 64))
    assert h1 == h1b
    (_, _, v2, t2), h2 = _scene_check(pt.many_spheres_scene(400, 64, 64, seed=1))
    assert h2 != h1 and v2 == 0 and t2 == 2
    # few spheres: no tree (the reference's object loop as it is)
    (n3, nodes3, v3, t3), _ = _scene_check(pt.three_balls(64, 64))
    assert (n3, nodes3, v3, t3) == (4, 0, 0, 0)
    # a scaled (non-rigid) sphere in a big group switches that group back to the scan
    sd4 = pt.many_spheres_scene(100, 64, 64, with_mesh=False)
    sd4.add_sphere(0.5, pt.compose(pt.scale(2.0), pt.translate((0.0, 1.0, -3.0))), "a")
    (n4, nodes4, v4, t4), _ = _scene_check(sd4)
    assert n4 == 102 and t4 == 0 and nodes4 == 0 and v4 == 0
