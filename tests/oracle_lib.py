"""ctypes loader for the CPU oracle (oracle/liboracle.so) — test infrastructure only.
Nothing under cuda_path_tracer_b200/ imports this module."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from cuda_path_tracer_b200 import _abi
from cuda_path_tracer_b200.api import HIT_DTYPE
from cuda_path_tracer_b200.scene_description import Camera, SceneDescription

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "liboracle.so")

BVH_NODE_DTYPE = np.dtype([("min", "<f4", 3), ("max", "<f4", 3), ("first", "<u4"), ("count", "<u4")])
VP = C.c_void_p


def build_oracle():
    src = os.path.join(ORACLE_DIR, "oracle.c")
    if (not os.path.exists(ORACLE_LIB)) or os.path.getmtime(src) > os.path.getmtime(ORACLE_LIB):
        subprocess.run(["make", "-C", ORACLE_DIR], check=True, capture_output=True)
    return ORACLE_LIB


class Oracle:
    def __init__(self, lib):
        self.lib = lib
        L = lib
        L.orc_hash.restype = C.c_uint32
        L.orc_hash.argtypes = [C.c_uint32]
        L.orc_rng_seed.restype = C.c_uint32
        L.orc_rng_seed.argtypes = [C.c_uint32]
        L.orc_rng_uniform.restype = C.c_float
        L.orc_rng_uniform.argtypes = [C.POINTER(C.c_uint32)]
        L.orc_rng_discard.argtypes = [C.POINTER(C.c_uint32), C.c_uint64]
        L.orc_aabb_props.argtypes = [VP, VP, VP, VP, C.POINTER(C.c_int), C.POINTER(C.c_float), VP]
        L.orc_inverse_transform_ray.argtypes = [VP, VP, VP, VP]
        L.orc_mat4_inverse.argtypes = [VP, VP]
        L.orc_ray_triangle.restype = C.c_int
        L.orc_ray_triangle.argtypes = [VP, VP, VP, VP, VP]
        L.orc_ray_sphere.restype = C.c_int
        L.orc_ray_sphere.argtypes = [VP, VP, C.c_float, VP]
        L.orc_ray_aabb.restype = C.c_int
        L.orc_ray_aabb.argtypes = [VP, VP, VP]
        L.orc_generate_ray.argtypes = [C.POINTER(_abi.pt_camera), C.c_uint32, C.c_uint32, C.c_float, C.c_float, VP]
        L.orc_scene_create.restype = VP
        L.orc_scene_create.argtypes = [C.POINTER(_abi.pt_scene_desc)]
        L.orc_scene_destroy.argtypes = [VP]
        L.orc_scene_bvh_size.restype = C.c_uint32
        L.orc_scene_bvh_size.argtypes = [VP]
        L.orc_scene_bvh.restype = VP
        L.orc_scene_bvh.argtypes = [VP]
        L.orc_scene_object_aabb.argtypes = [VP, C.c_uint32, VP, VP]
        L.orc_trace_batch.argtypes = [VP, VP, C.c_uint64, VP, C.c_int]
        render_args = [VP, C.POINTER(_abi.pt_camera), C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_int,
                       VP, VP, VP, C.POINTER(C.c_uint64)]
        L.orc_render_megakernel.argtypes = render_args
        L.orc_render_streaming.argtypes = render_args
        L.orc_denoise.argtypes = [C.c_uint32, C.c_uint32, C.POINTER(_abi.pt_camera), VP, VP, VP, C.c_int,
                                  C.c_float, C.c_float, C.c_float, C.c_int, VP, VP]
        L.orc_tonemap.argtypes = [C.c_int, C.c_uint32, VP, VP]
        L.orc_num_threads.restype = C.c_int

    # ---- scalars
    def hash(self, a):
        return int(self.lib.orc_hash(C.c_uint32(a & 0xFFFFFFFF)))

    def rng_stream(self, seed, n, discard=0):
        st = C.c_uint32(self.lib.orc_rng_seed(C.c_uint32(seed & 0xFFFFFFFF)))
        self.lib.orc_rng_discard(C.byref(st), discard)
        return [float(self.lib.orc_rng_uniform(C.byref(st))) for _ in range(n)], int(st.value)

    def generate_ray(self, cam: Camera, w, h, x, y):
        out = np.zeros(8, dtype=np.float32)
        c = cam.to_c()
        self.lib.orc_generate_ray(C.byref(c), w, h, x, y, out.ctypes.data)
        return out

    def primary_rays(self, cam: Camera, w, h, xs, ys):
        return np.stack([self.generate_ray(cam, w, h, float(x), float(y)) for x, y in zip(xs, ys)])

    # ---- scenes
    def scene(self, desc: SceneDescription) -> "OracleScene":
        d, keep = desc.to_desc()
        h = self.lib.orc_scene_create(C.byref(d))
        return OracleScene(self, h)

    def denoise(self, cam: Camera, color, normal, depth, filter_size=10, cw=0.45, nw=0.30, pw=0.25,
                clamp_fix=False):
        h, w = depth.shape
        c = np.ascontiguousarray(color, dtype=np.float32)
        n = np.ascontiguousarray(normal, dtype=np.float32)
        d = np.ascontiguousarray(depth, dtype=np.float32)
        out = np.zeros((h, w, 3), dtype=np.float32)
        taint = np.zeros((h, w), dtype=np.uint8)
        cc = cam.to_c()
        self.lib.orc_denoise(w, h, C.byref(cc), c.ctypes.data, n.ctypes.data, d.ctypes.data, filter_size,
                             cw, nw, pw, 1 if clamp_fix else 0, out.ctypes.data, taint.ctypes.data)
        return out, taint.astype(bool)

    def tonemap(self, kind, src):
        s = np.ascontiguousarray(src, dtype=np.float32)
        n = s.size if kind == 3 else s.size // 3
        out = np.zeros((n, 4), dtype=np.uint8)
        self.lib.orc_tonemap(kind, n, s.ctypes.data, out.ctypes.data)
        return out

    def num_threads(self):
        return int(self.lib.orc_num_threads())


class OracleScene:
    def __init__(self, oracle: Oracle, handle):
        self.o = oracle
        self.h = handle

    def bvh(self) -> np.ndarray:
        n = self.o.lib.orc_scene_bvh_size(self.h)
        if n == 0:
            return np.zeros(0, dtype=BVH_NODE_DTYPE)
        p = self.o.lib.orc_scene_bvh(self.h)
        buf = (C.c_char * (n * BVH_NODE_DTYPE.itemsize)).from_address(p)
        return np.frombuffer(buf, dtype=BVH_NODE_DTYPE).copy()

    def object_aabb(self, i):
        mn = np.zeros(3, dtype=np.float32)
        mx = np.zeros(3, dtype=np.float32)
        self.o.lib.orc_scene_object_aabb(self.h, i, mn.ctypes.data, mx.ctypes.data)
        return mn, mx

    def trace_batch(self, rays8, mode=0) -> np.ndarray:
        rays = np.ascontiguousarray(rays8, dtype=np.float32).reshape(-1, 8)
        out = np.zeros(rays.shape[0], dtype=HIT_DTYPE)
        self.o.lib.orc_trace_batch(self.h, rays.ctypes.data, rays.shape[0], out.ctypes.data, mode)
        return out

    def render(self, cam: Camera, w, h, n_iterations, max_bounces=50, first_iteration=0, mode="megakernel",
               state=None):
        if state is None:
            color = np.zeros((h, w, 3), dtype=np.float32)
            normal = np.zeros((h, w, 3), dtype=np.float32)
            depth = np.zeros((h, w), dtype=np.float32)
        else:
            color, normal, depth = state
        rays = C.c_uint64(0)
        c = cam.to_c()
        fn = self.o.lib.orc_render_megakernel if mode == "megakernel" else self.o.lib.orc_render_streaming
        fn(self.h, C.byref(c), w, h, first_iteration, n_iterations, max_bounces, color.ctypes.data,
           normal.ctypes.data, depth.ctypes.data, C.byref(rays))
        return color, normal, depth, int(rays.value)

    def close(self):
        if self.h:
            self.o.lib.orc_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_oracle = None


def load_oracle() -> Oracle:
    global _oracle
    if _oracle is None:
        _oracle = Oracle(C.CDLL(build_oracle()))
    return _oracle
