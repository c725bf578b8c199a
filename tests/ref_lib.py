"""ctypes loaders for the reference's own code compiled into oracle/_ref/ (build_ref.sh).
Test infrastructure / baseline arm only; never imported by the product package."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from cuda_path_tracer_b200 import _abi
from cuda_path_tracer_b200.api import HIT_DTYPE
from cuda_path_tracer_b200.scene_description import Camera, SceneDescription

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
REF_HOST = os.path.join(REF_DIR, "libref_host.so")
REF_CUDA = os.path.join(REF_DIR, "libref_cuda.so")              # traversal stack 64: the parity checker
REF_CUDA_STOCK = os.path.join(REF_DIR, "libref_cuda_stock.so")  # stack 24 = the reference's own: the timed baseline
VP = C.c_void_p
BVH_NODE_DTYPE = np.dtype([("min", "<f4", 3), ("max", "<f4", 3), ("first", "<u4"), ("count", "<u4")])


class RefHost:
    def __init__(self):
        L = C.CDLL(REF_HOST)
        L.ref_hash.restype = C.c_uint32
        L.ref_hash.argtypes = [C.c_uint32]
        L.ref_bvh_from_mesh.restype = C.c_uint64
        L.ref_bvh_from_mesh.argtypes = [VP, C.c_uint64, VP, C.c_uint64, VP, C.c_uint64, C.POINTER(C.c_double)]
        for n in ("ref_ray_triangle",):
            getattr(L, n).restype = C.c_int
            getattr(L, n).argtypes = [VP, VP, VP, VP, VP]
        L.ref_ray_sphere.restype = C.c_int
        L.ref_ray_sphere.argtypes = [VP, VP, C.c_float, VP]
        L.ref_ray_aabb.restype = C.c_int
        L.ref_ray_aabb.argtypes = [VP, VP, VP]
        L.ref_inverse_transform_ray.argtypes = [VP, VP, VP, VP]
        L.ref_transform_aabb.argtypes = [VP, VP, VP, VP, VP, VP]
        L.ref_aabb_props.argtypes = [VP, VP, VP, VP, C.POINTER(C.c_int), C.POINTER(C.c_float), VP]
        self.lib = L

    def bvh_from_mesh(self, positions, indices):
        pos = np.ascontiguousarray(positions, dtype=np.float32).reshape(-1, 3)
        idx = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1)
        cap = 2 * (idx.size // 3)
        out = np.zeros(cap, dtype=BVH_NODE_DTYPE)
        secs = C.c_double(0)
        n = self.lib.ref_bvh_from_mesh(pos.ctypes.data, pos.shape[0], idx.ctypes.data, idx.size,
                                       out.ctypes.data, cap, C.byref(secs))
        return out[:n].copy(), secs.value


class RefCudaTracer:
    def __init__(self, lib, desc: SceneDescription, w, h, max_bounces=50, megakernel=False):
        self.lib = lib
        d, keep = desc.to_desc()
        self.w, self.h = w, h
        self.max_bounces = max_bounces
        self.h_ = lib.ref_tracer_create(C.byref(d), w, h, max_bounces, 1 if megakernel else 0)
        del keep

    def render_timed(self, cam: Camera, n_iterations, max_bounces=None):
        c = cam.to_c()
        rays = C.c_ulonglong(0)
        ms = self.lib.ref_tracer_render(self.h_, C.byref(c), n_iterations,
                                        max_bounces or self.max_bounces, C.byref(rays))
        return float(ms), int(rays.value)

    def restart(self):
        self.lib.ref_tracer_restart(self.h_)

    def download(self, kind):
        out = np.zeros((self.h, self.w) if kind == 3 else (self.h, self.w, 3), dtype=np.float32)
        rc = self.lib.ref_tracer_download(self.h_, kind, out.ctypes.data)
        assert rc == 0, rc
        return out

    def upload_frame(self, color, normal, depth, cam: Camera):
        c = cam.to_c()
        a = np.ascontiguousarray(color, dtype=np.float32)
        b = np.ascontiguousarray(normal, dtype=np.float32)
        d = np.ascontiguousarray(depth, dtype=np.float32)
        assert self.lib.ref_tracer_upload_frame(self.h_, a.ctypes.data, b.ctypes.data, d.ctypes.data, C.byref(c)) == 0

    def denoise(self, filter_size=10, cw=0.45, nw=0.30, pw=0.25):
        return float(self.lib.ref_tracer_denoise(self.h_, filter_size, cw, nw, pw))

    def preview(self, kind=0):
        out = np.zeros((self.h, self.w, 4), dtype=np.uint8)
        assert self.lib.ref_tracer_preview(self.h_, kind, out.ctypes.data) == 0
        return out

    def trace_batch(self, rays8):
        rays = np.ascontiguousarray(rays8, dtype=np.float32).reshape(-1, 8)
        out = np.zeros(rays.shape[0], dtype=HIT_DTYPE)
        assert self.lib.ref_trace_batch(self.h_, rays.ctypes.data, rays.shape[0], out.ctypes.data) == 0
        return out

    def close(self):
        if self.h_:
            self.lib.ref_tracer_destroy(self.h_)
            self.h_ = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class RefCuda:
    def __init__(self, path=REF_CUDA):
        L = C.CDLL(path)
        self.path = path
        L.ref_tracer_create.restype = VP
        L.ref_tracer_create.argtypes = [C.POINTER(_abi.pt_scene_desc), C.c_uint32, C.c_uint32, C.c_int, C.c_int]
        L.ref_tracer_destroy.argtypes = [VP]
        L.ref_tracer_restart.argtypes = [VP]
        L.ref_tracer_render.restype = C.c_float
        L.ref_tracer_render.argtypes = [VP, C.POINTER(_abi.pt_camera), C.c_int, C.c_int, C.POINTER(C.c_ulonglong)]
        L.ref_tracer_download.restype = C.c_int
        L.ref_tracer_download.argtypes = [VP, C.c_int, VP]
        L.ref_tracer_upload_frame.restype = C.c_int
        L.ref_tracer_upload_frame.argtypes = [VP, VP, VP, VP, C.POINTER(_abi.pt_camera)]
        L.ref_tracer_denoise.restype = C.c_float
        L.ref_tracer_denoise.argtypes = [VP, C.c_int, C.c_float, C.c_float, C.c_float]
        L.ref_tracer_preview.restype = C.c_int
        L.ref_tracer_preview.argtypes = [VP, C.c_int, VP]
        L.ref_trace_batch.restype = C.c_int
        L.ref_trace_batch.argtypes = [VP, VP, C.c_uint64, VP]
        L.ref_bvh_build_seconds.restype = C.c_double
        L.ref_bvh_build_seconds.argtypes = [VP, C.c_uint64, VP, C.c_uint64, C.POINTER(C.c_uint64)]
        L.ref_bvh_depth.restype = C.c_int
        L.ref_bvh_depth.argtypes = [VP, C.c_uint64, VP, C.c_uint64]
        L.ref_stack_size.restype = C.c_int
        self.lib = L
        self.stack_size = int(L.ref_stack_size())

    def tracer(self, desc, w, h, max_bounces=50, megakernel=False):
        return RefCudaTracer(self.lib, desc, w, h, max_bounces, megakernel)

    def bvh_build_seconds(self, mesh):
        pos = np.ascontiguousarray(mesh.positions, dtype=np.float32).reshape(-1, 3)
        idx = np.ascontiguousarray(mesh.indices, dtype=np.uint32).reshape(-1)
        n = C.c_uint64(0)
        s = self.lib.ref_bvh_build_seconds(pos.ctypes.data, pos.shape[0], idx.ctypes.data, idx.size, C.byref(n))
        return float(s), int(n.value)


    def bvh_depth(self, mesh):
        """Depth of the reference's own tree for `mesh` (host only; root = 1).  The reference's
        unchecked 24-entry traversal stack is correct exactly when this is <= 23."""
        pos = np.ascontiguousarray(mesh.positions, dtype=np.float32).reshape(-1, 3)
        idx = np.ascontiguousarray(mesh.indices, dtype=np.uint32).reshape(-1)
        return int(self.lib.ref_bvh_depth(pos.ctypes.data, pos.shape[0], idx.ctypes.data, idx.size))


_host = None
_cuda = None
_cuda_stock = None


def have_ref_host():
    return os.path.exists(REF_HOST)


def have_ref_cuda():
    return os.path.exists(REF_CUDA)


def load_ref_host() -> RefHost:
    global _host
    if _host is None:
        if not have_ref_host():
            raise FileNotFoundError(f"{REF_HOST} not built (oracle/build_ref.sh needs /root/reference)")
        _host = RefHost()
    return _host


def have_ref_cuda_stock():
    return os.path.exists(REF_CUDA_STOCK)


def load_ref_cuda_stock() -> RefCuda:
    """The reference with its own traversal stack size (24): the timed baseline."""
    global _cuda_stock
    if _cuda_stock is None:
        if not have_ref_cuda_stock():
            raise FileNotFoundError(f"{REF_CUDA_STOCK} not built (oracle/build_ref.sh needs /root/reference)")
        _cuda_stock = RefCuda(REF_CUDA_STOCK)
    return _cuda_stock


def load_ref_cuda() -> RefCuda:
    global _cuda
    if _cuda is None:
        if not have_ref_cuda():
            raise FileNotFoundError(f"{REF_CUDA} not built (oracle/build_ref.sh needs /root/reference)")
        _cuda = RefCuda()
    return _cuda
