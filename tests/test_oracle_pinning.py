"""Pins the CPU oracle (oracle/oracle.c) BEFORE it is trusted as the checker:

 (a) the known-answer vectors of the reference's own unit tests
     (test/aabb_test.cpp:6-59, test/transform_test.cpp:8-45, glm_test_helper.hpp eps = 100*FLT_EPSILON);
 (b) the reference's own host code compiled from /root/reference into oracle/_ref/libref_host.so
     (bvh_from_mesh, ray_triangle/sphere/aabb tests, inverse_transform_ray, transform_aabb, hash);
 (c) committed golden vectors generated from (b) — tests/golden/ref_host_vectors.npz — so the pin
     also holds where /root/reference and oracle/_ref are absent.
The reference has no test for intersection, traversal, shading, RNG, accumulation or the
denoiser; those are pinned on the GPU box against its CUDA build (test_ref_cuda_parity.py).
"""
import ctypes as C
import math
import os

import numpy as np
import pytest

import cuda_path_tracer_b200 as pt
from cuda_path_tracer_b200.api import HIT_DTYPE
from tests import ref_lib

EPS = 100 * np.finfo(np.float32).eps
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ref_host_vectors.npz")


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _aabb_props(lib, fn, mn, mx, p):
    ext = np.zeros(3, np.float32)
    off = np.zeros(3, np.float32)
    me = C.c_int(0)
    sa = C.c_float(0)
    a, b, c = _f(mn), _f(mx), _f(p)  # keep the buffers alive across the call
    fnp = getattr(lib, fn)
    fnp.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_float),
                    C.c_void_p]
    fnp(a.ctypes.data, b.ctypes.data, c.ctypes.data, ext.ctypes.data, C.byref(me), C.byref(sa), off.ctypes.data)
    return ext, me.value, sa.value, off


# ---------------------------------------------------------------- (a) reference unit-test KATs
def test_aabb_known_answers(oracle):
    """test/aabb_test.cpp: extent (6,4,2); max_extent 0/1/2; surface_area 88; offset 0 / 1 / 0.5."""
    ext, _, sa, off = _aabb_props(oracle.lib, "orc_aabb_props", [1, 2, 3], [7, 6, 5], [1, 2, 3])
    assert np.allclose(ext, [6, 4, 2], atol=EPS)
    assert sa == 88
    assert np.allclose(off, [0, 0, 0], atol=EPS)
    assert _aabb_props(oracle.lib, "orc_aabb_props", [1, 2, 3], [100, 6, 5], [0, 0, 0])[1] == 0
    assert _aabb_props(oracle.lib, "orc_aabb_props", [1, 2, 3], [7, 100, 5], [0, 0, 0])[1] == 1
    assert _aabb_props(oracle.lib, "orc_aabb_props", [1, 2, 3], [7, 6, 100], [0, 0, 0])[1] == 2
    assert np.allclose(_aabb_props(oracle.lib, "orc_aabb_props", [1, 2, 3], [7, 6, 5], [7, 6, 5])[3], [1, 1, 1], atol=EPS)
    assert np.allclose(_aabb_props(oracle.lib, "orc_aabb_props", [1, 2, 3], [7, 6, 5], [4, 4, 4])[3], [.5, .5, .5], atol=EPS)


def _inv_ray(lib, fn, m, ray):
    m = _f(m)
    inv = np.zeros(16, np.float32)
    # Transform{m} computes glm::inverse(m): use the oracle's restatement of it
    from tests.oracle_lib import load_oracle
    mcol = _f(m.T.reshape(-1))
    load_oracle().lib.orc_mat4_inverse(mcol.ctypes.data, inv.ctypes.data)
    out = np.zeros(8, np.float32)
    mc, rc = _f(m.T.reshape(-1)), _f(ray)
    fnp = getattr(lib, fn)
    fnp.argtypes = [C.c_void_p] * 4
    fnp(mc.ctypes.data, inv.ctypes.data, rc.ctypes.data, out.ctypes.data)
    return out


def test_inverse_transform_ray_known_answers(oracle):
    """test/transform_test.cpp: translate(1,1,1), scale 2, rotate pi about Y on ray (1,2,3)->(1,0,0)."""
    ray = [1, 2, 3, 0, 1, 0, 0, 100]
    t = _inv_ray(oracle.lib, "orc_inverse_transform_ray", pt.translate((1, 1, 1)), ray)
    assert np.allclose(t[0:3], [0, 1, 2], atol=EPS) and np.allclose(t[4:7], [1, 0, 0], atol=EPS)
    assert t[3] == 0 and t[7] == 100
    s = _inv_ray(oracle.lib, "orc_inverse_transform_ray", pt.scale(2.0), ray)
    assert np.allclose(s[0:3], [0.5, 1, 1.5], atol=EPS) and np.allclose(s[4:7], [1, 0, 0], atol=EPS)
    r = _inv_ray(oracle.lib, "orc_inverse_transform_ray", pt.rotate(180.0, (0, 1, 0)), ray)
    assert np.allclose(r[0:3], [-1, 2, -3], atol=2e-6) and np.allclose(r[4:7], [-1, 0, 0], atol=2e-6)
    assert r[3] == 0 and r[7] == 100


def test_rng_known_answers(oracle):
    """minstd_rand: x0 = 1 -> 48271, 182605794, ...; the 10000th value is 399268537 (C++ standard
    [rand.predef]); uniform = float(x - 1) / 2^31 (thrust uniform_real_distribution.inl:62-79)."""
    st = C.c_uint32(oracle.lib.orc_rng_seed(1))
    u = oracle.lib.orc_rng_uniform(C.byref(st))
    assert st.value == 48271 and u == np.float32(48270) / np.float32(2147483648.0)
    oracle.lib.orc_rng_uniform(C.byref(st))
    assert st.value == 182605794
    st = C.c_uint32(oracle.lib.orc_rng_seed(1))
    oracle.lib.orc_rng_discard(C.byref(st), 10000)
    assert st.value == 399268537
    assert oracle.lib.orc_rng_seed(0) == 1 and oracle.lib.orc_rng_seed(2147483647) == 1
    assert oracle.lib.orc_rng_seed(2147483648) == 1 + 0  # 2^31 mod (2^31 - 1) = 1


# ---------------------------------------------------------------- (b)/(c) reference host code
def _cases(seed=0, n=400):
    rng = np.random.default_rng(seed)
    tri = rng.uniform(-1, 1, size=(n, 3, 3)).astype(np.float32)
    org = rng.uniform(-2, 2, size=(n, 3)).astype(np.float32)
    tgt = tri.mean(axis=1) + rng.normal(scale=0.3, size=(n, 3)).astype(np.float32)
    d = tgt - org
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.zeros((n, 8), np.float32)
    rays[:, 0:3], rays[:, 3], rays[:, 4:7], rays[:, 7] = org, 1e-4, d, np.finfo(np.float32).max
    rays[::7, 4:7] *= 1.7  # non-unit directions (fuzzy metal)
    rays[::11, 7] = 1.0     # finite t_max
    rays[::13, 4] = 0.0     # axis-parallel components
    centers = rng.uniform(-1, 1, size=(n, 3)).astype(np.float32)
    radii = rng.uniform(0.1, 1.5, size=n).astype(np.float32)
    bmin = rng.uniform(-1, 0, size=(n, 3)).astype(np.float32)
    bmax = bmin + rng.uniform(0, 1.5, size=(n, 3)).astype(np.float32)
    return tri, rays, centers, radii, bmin, bmax


def _run_prims(lib, prefix):
    tri, rays, centers, radii, bmin, bmax = _cases()
    n = len(rays)
    ht = np.zeros(n, HIT_DTYPE)
    hs = np.zeros(n, HIT_DTYPE)
    ab = np.zeros(n, np.int32)
    for i in range(n):
        getattr(lib, prefix + "ray_triangle")(rays[i].ctypes.data, tri[i, 0].ctypes.data, tri[i, 1].ctypes.data,
                                              tri[i, 2].ctypes.data, ht[i:i + 1].ctypes.data)
        getattr(lib, prefix + "ray_sphere")(rays[i].ctypes.data, centers[i].ctypes.data,
                                            C.c_float(float(radii[i])), hs[i:i + 1].ctypes.data)
        ab[i] = getattr(lib, prefix + "ray_aabb")(rays[i].ctypes.data, bmin[i].ctypes.data, bmax[i].ctypes.data)
    return ht, hs, ab


def _hits_equal(a, b, tol=2e-6):
    assert np.array_equal(a["t"] < 0, b["t"] < 0)
    m = a["t"] >= 0
    assert np.allclose(a["t"][m], b["t"][m], rtol=tol, atol=tol)
    assert np.allclose(a["point"][m], b["point"][m], rtol=tol, atol=1e-5)
    assert np.allclose(a["normal"][m], b["normal"][m], rtol=tol, atol=1e-5)
    assert np.array_equal(a["side"][m], b["side"][m])


def _mesh_cases():
    return {"blob320": pt.bunny_like(2), "blob1280": pt.bunny_like(3), "grid": pt.heightfield(12, seed=3)}


@pytest.mark.skipif(not ref_lib.have_ref_host(), reason="oracle/_ref/libref_host.so not built")
def test_oracle_matches_reference_host_code(oracle):
    ref = ref_lib.load_ref_host().lib
    for fn in ("ray_triangle", "ray_sphere"):
        getattr(oracle.lib, "orc_" + fn).restype = C.c_int
    ot, os_, oa = _run_prims(oracle.lib, "orc_")
    rt, rs, ra = _run_prims(ref, "ref_")
    _hits_equal(ot, rt)
    _hits_equal(os_, rs)
    assert np.array_equal(oa, ra)
    assert (ot["t"] >= 0).sum() > 20 and (os_["t"] >= 0).sum() > 20 and 0 < ra.sum() < len(ra)
    for a in (0, 1, 12345, 0xFFFFFFFF, 0x7ED55D16):
        assert oracle.hash(a) == ref.ref_hash(a)
    # AABB helpers + transforms on random inputs
    rng = np.random.default_rng(5)
    for _ in range(50):
        mn = rng.uniform(-3, 0, 3)
        mx = mn + rng.uniform(0, 4, 3)
        p = rng.uniform(-3, 4, 3)
        a = _aabb_props(oracle.lib, "orc_aabb_props", mn, mx, p)
        b = _aabb_props(ref, "ref_aabb_props", mn, mx, p)
        assert np.array_equal(a[0], b[0]) and a[1] == b[1] and a[2] == b[2] and np.array_equal(a[3], b[3])
    for m in (pt.translate((1, -2, 3)), pt.compose(pt.scale(0.5), pt.translate((-1, -0.5, -2))),
              pt.compose(pt.rotate(150, (0, 1, 0)), pt.scale((1, 2, 3)), pt.translate((0, 0, -0.25)))):
        for ray in _cases(2, 20)[1]:
            a = _inv_ray(oracle.lib, "orc_inverse_transform_ray", m, ray)
            b = _inv_ray(ref, "ref_inverse_transform_ray", m, ray)
            assert np.allclose(a, b, rtol=1e-6, atol=1e-6)


@pytest.mark.skipif(not ref_lib.have_ref_host(), reason="oracle/_ref/libref_host.so not built")
def test_oracle_bvh_equals_reference_builder(oracle):
    """The oracle's restatement of bvh_from_mesh yields the reference's flattened tree node for node."""
    ref = ref_lib.load_ref_host()
    for name, mesh in _mesh_cases().items():
        sd = pt.SceneDescription()
        sd.add_material("m", pt.Material.lambertian((0.5, 0.5, 0.5)))
        sd.add_mesh("mesh", mesh)
        sd.add_mesh_object("mesh", pt.translate((0, 0, 0)), "m")
        ours = oracle.scene(sd).bvh()
        theirs, _ = ref.bvh_from_mesh(mesh.positions, mesh.indices)
        assert len(ours) == len(theirs) == 2 * mesh.triangle_count - 1, name
        assert np.array_equal(ours["count"], theirs["count"]), name
        assert np.array_equal(ours["first"], theirs["first"]), name
        assert np.array_equal(ours["min"], theirs["min"]) and np.array_equal(ours["max"], theirs["max"]), name


def test_golden_vectors_from_reference_host_code(oracle):
    """Committed outputs of the reference's host code (generated by tests/golden/make_golden.py in the
    authoring container) — the pin that travels to boxes without /root/reference."""
    if not os.path.exists(GOLDEN):
        pytest.skip("golden vectors not generated yet")
    g = np.load(GOLDEN)
    ot, os_, oa = _run_prims(oracle.lib, "orc_")
    for name, arr in (("tri", ot), ("sph", os_)):
        assert np.array_equal(arr["t"] < 0, g[name + "_t"] < 0)
        m = arr["t"] >= 0
        assert np.allclose(arr["t"][m], g[name + "_t"][m], rtol=2e-6, atol=2e-6)
        assert np.allclose(arr["normal"][m], g[name + "_normal"][m], rtol=2e-6, atol=1e-5)
        assert np.array_equal(arr["side"][m], g[name + "_side"][m])
    assert np.array_equal(oa, g["aabb"])
    for name, mesh in _mesh_cases().items():
        sd = pt.SceneDescription()
        sd.add_material("m", pt.Material.lambertian((0.5, 0.5, 0.5)))
        sd.add_mesh("mesh", mesh)
        sd.add_mesh_object("mesh", pt.translate((0, 0, 0)), "m")
        ours = oracle.scene(sd).bvh()
        assert np.array_equal(ours["first"], g["bvh_first_" + name])
        assert np.array_equal(ours["count"], g["bvh_count_" + name])
        assert np.array_equal(ours["min"], g["bvh_min_" + name])
    assert [oracle.hash(a) for a in (0, 1, 12345, 0xFFFFFFFF)] == list(g["hash"])


# ---------------------------------------------------------------- oracle self-consistency
def test_oracle_bvh_traversal_equals_brute_force(oracle):
    sd = pt.bunny_scene(pt.bunny_like(2), 64, 36)
    osc = oracle.scene(sd)
    rng = np.random.default_rng(3)
    rays = oracle.primary_rays(sd.camera, 64, 36, rng.uniform(0, 64, 600), rng.uniform(0, 36, 600))
    a, b = osc.trace_batch(rays, 0), osc.trace_batch(rays, 1)
    assert np.array_equal(a["t"] < 0, b["t"] < 0)
    m = a["t"] > 0
    assert m.sum() > 100
    assert np.allclose(a["t"][m], b["t"][m], rtol=1e-6)
    assert np.array_equal(a["object"], b["object"])


def test_oracle_modes_agree_statistically(oracle):
    """Megakernel and streaming RNG disciplines estimate the same image."""
    sd = pt.three_balls(48, 32)
    osc = oracle.scene(sd)
    a = osc.render(sd.camera, 48, 32, 48, 50, mode="megakernel")[0]
    b = osc.render(sd.camera, 48, 32, 48, 50, mode="streaming")[0]
    assert abs(a.mean() - b.mean()) < 0.01
    assert math.sqrt(np.mean((a - b) ** 2)) < 0.12
