"""Python mirror of the reference's integrator interface — class PathTracer
(src/lib/path_tracer.hpp:60-99) — over the C ABI in include/b200pt.h.

Names, argument meaning and defaults follow the reference so the parity tests read
like tests of the reference: create_buffers / resize_image / path_trace / denoise /
send_to_preview / restart / iteration(), fields max_iterations, current_gpu_method,
atrous_denoiser.{filter_size,color_weight,normal_weight,position_weight}.
Everything that touches rays or pixels runs in libb200pt.so on the GPU.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _abi
from ._abi import check, load_library
from .scene_description import Camera, SceneDescription


class GPUMethod:
    """GPUMethod (path_tracer.hpp:58).  Both run on the same wavefront kernels; the
    value selects which of the reference's two RNG disciplines is reproduced."""
    megakernel = _abi.RNG_PIXEL_STREAM
    streaming = _abi.RNG_SLOT_RESEED


class DisplayBufferType:
    final, color, normal, depth, denoised = 0, 1, 2, 3, 4


@dataclass
class EdgeAvoidingATrousDenoiser:
    """denoising/edge_avoiding_a_trous_denoiser.hpp:7-22"""
    filter_size: int = 10
    color_weight: float = 0.45
    normal_weight: float = 0.30
    position_weight: float = 0.25
    clamp_fix: bool = False


class HostBVH:
    """Host-only build of both trees (no CUDA call): structure checks and build timing."""

    def __init__(self, desc: SceneDescription, wide: bool = True, lbvh: bool = False):
        """wide: also derive the compressed 8-wide tree; lbvh: build with the host restatement of
        the device LBVH builder instead of the SAH builder."""
        lib = load_library()
        d, keep = desc.to_desc()
        self._h = C.c_void_p()
        self.info = _abi.pt_scene_info()
        mode = 2 if lbvh else (1 if wide else 0)
        check(lib.pt_host_bvh_build(C.byref(d), mode, C.byref(self._h), C.byref(self.info)))
        del keep

    def violations(self) -> int:
        n = C.c_uint64()
        check(load_library().pt_host_bvh_validate(self._h, C.byref(n)))
        return int(n.value)

    def arrays(self):
        """(binary nodes [n,16] f32, wide nodes [n8,20] u32, triangles [t,12] f32) as numpy copies."""
        a, b, t = C.c_void_p(), C.c_void_p(), C.c_void_p()
        check(load_library().pt_host_bvh_arrays(self._h, C.byref(a), C.byref(b), C.byref(t)))
        i = self.info

        def view(ptr, n, w, ct, dt):
            if not ptr.value or n == 0:
                return np.zeros((0, w), dtype=dt)
            return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(n, w)).astype(dt, copy=True)

        n_tris = int(i.n_bvh_triangles)
        return (view(a, int(i.n_bvh_nodes), 16, C.c_float, np.float32),
                view(b, int(i.n_bvh8_nodes), 20, C.c_uint32, np.uint32),
                view(t, n_tris, 12, C.c_float, np.float32))

    def quantised(self):
        """(quantised nodes [n,8] u32, grid origin [3] f32, cell size [3] f32): the 32-byte nodes the
        traversal kernels read, as numpy copies."""
        q = C.c_void_p()
        org = (C.c_float * 3)()
        cell = (C.c_float * 3)()
        check(load_library().pt_host_bvh_quantised(self._h, C.byref(q), org, cell))
        n = int(self.info.n_bvh_nodes)
        arr = np.ctypeslib.as_array(C.cast(q, C.POINTER(C.c_uint32)), shape=(n, 8)).astype(np.uint32, copy=True)
        return arr, np.array(list(org), np.float32), np.array(list(cell), np.float32)

    def trace_stats(self, rays8: np.ndarray, wide=False) -> dict:
        """Host walk in the device kernels' order: what the rays cost in this tree (analysis tool).
        wide: False/0 binary tree, True/1 compressed 8-wide tree, 2 a virtual 4-wide tree (every
        other level of the binary tree collapsed)."""
        rays = np.ascontiguousarray(rays8, dtype=np.float32).reshape(-1, 8)
        out = np.zeros(5, dtype=np.uint64)
        check(load_library().pt_host_bvh_trace_stats(self._h, rays.ctypes.data, rays.shape[0], int(wide),
                                                     out.ctypes.data))
        n = max(1, rays.shape[0])
        return {"rays": rays.shape[0], "inner_per_ray": float(out[0]) / n, "leaves_per_ray": float(out[1]) / n,
                "tri_tests_per_ray": float(out[2]) / n, "hit_fraction": float(out[3]) / n, "max_stack": int(out[4])}

    def close(self):
        if self._h:
            load_library().pt_host_bvh_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Scene:
    """Device-resident scene == Scene/Aggregate (scene.hpp:23-67) after build_scene()."""

    def __init__(self, handle, file_info=None):
        self._h = handle
        self.file_info = file_info

    @staticmethod
    def from_description(desc: SceneDescription, device: int = 0) -> "Scene":
        lib = load_library()
        d, keep = desc.to_desc()
        h = C.c_void_p()
        check(lib.pt_scene_create(C.byref(d), device, C.byref(h)))
        del keep
        return Scene(h)

    @staticmethod
    def from_file(json_path: str, device: int = 0) -> "Scene":
        """read_scene (assets/scene_parser.cpp:6-22)."""
        lib = load_library()
        h = C.c_void_p()
        info = _abi.pt_scene_file_info()
        check(lib.pt_scene_load_file(json_path.encode(), device, C.byref(h), C.byref(info)))
        return Scene(h, info)

    @property
    def info(self) -> _abi.pt_scene_info:
        i = _abi.pt_scene_info()
        check(load_library().pt_scene_get_info(self._h, C.byref(i)))
        return i

    def copy_bvh(self):
        """(nodes [n,16] f32, triangles [t,12] f32) copied back from the device."""
        i = self.info
        nodes = np.zeros((int(i.n_bvh_nodes), 16), dtype=np.float32)
        tris = np.zeros((int(i.n_bvh_triangles), 12), dtype=np.float32)
        check(load_library().pt_scene_copy_bvh(self._h, nodes.ctypes.data, tris.ctypes.data))
        return nodes, tris

    def trace_batch(self, rays8: np.ndarray) -> np.ndarray:
        """Closest hits of n rays {o, t_min, d, t_max} -> structured array of pt_hit."""
        rays = np.ascontiguousarray(rays8, dtype=np.float32).reshape(-1, 8)
        out = np.zeros(rays.shape[0], dtype=HIT_DTYPE)
        check(load_library().pt_trace_batch(self._h, rays.ctypes.data, rays.shape[0], out.ctypes.data))
        return out

    def close(self):
        if self._h:
            load_library().pt_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


HIT_DTYPE = np.dtype([("t", "<f4"), ("point", "<f4", 3), ("normal", "<f4", 3), ("material", "<u4"),
                      ("side", "<u4"), ("object", "<i4"), ("prim", "<i4"), ("pad", "<u4")])
assert HIT_DTYPE.itemsize == C.sizeof(_abi.pt_hit)


class PathTracer:
    def __init__(self, max_depth: int = 50, samples_per_pass: int = 0, profile: bool = False,
                 stream: int | None = None, sort_rays: bool = False, lanes: int = 0):
        self.sort_rays = sort_rays
        self.lanes = lanes
        self.max_iterations = 1
        self.current_gpu_method = GPUMethod.megakernel
        self.atrous_denoiser = EdgeAvoidingATrousDenoiser()
        self.max_depth = max_depth
        self.samples_per_pass = samples_per_pass
        self.profile = profile
        self._stream = stream
        self._ctx = None
        self._scene = None
        self._res = (0, 0)

    # -- PathTracer::create_buffers (path_tracer.cu:559-564)
    def create_buffers(self, resolution, scene):
        lib = load_library()
        if isinstance(scene, SceneDescription):
            scene = Scene.from_description(scene)
        self._destroy_ctx()
        self._scene = scene
        p = _abi.pt_params()
        lib.pt_params_default(C.byref(p))
        p.max_depth = self.max_depth
        p.rng_mode = self.current_gpu_method
        p.samples_per_pass = self.samples_per_pass
        p.profile = 1 if self.profile else 0
        p.sort_rays = 1 if self.sort_rays else 0
        p.lanes = int(self.lanes)
        h = C.c_void_p()
        w, hh = int(resolution[0]), int(resolution[1])
        check(lib.pt_ctx_create(scene._h, w, hh, C.byref(p), C.c_void_p(self._stream or 0), C.byref(h)))
        self._ctx = h
        self._res = (w, hh)

    # -- PathTracer::resize_image (path_tracer.cu:527-545)
    def resize_image(self, resolution):
        check(load_library().pt_ctx_resize(self._ctx, int(resolution[0]), int(resolution[1])))
        self._res = (int(resolution[0]), int(resolution[1]))

    # -- PathTracer::restart / iteration
    def restart(self):
        check(load_library().pt_ctx_restart(self._ctx))

    def iteration(self) -> int:
        return load_library().pt_ctx_iteration(self._ctx)

    # -- PathTracer::path_trace (path_tracer.cu:389-477): one spp per call, no-op at max_iterations
    def path_trace(self, camera: Camera, resolution=None):
        lib = load_library()
        check(lib.pt_ctx_set_max_iterations(self._ctx, int(self.max_iterations)))
        cam = camera.to_c()
        check(lib.pt_path_trace(self._ctx, C.byref(cam)))

    def render(self, camera: Camera, n_iterations: int):
        """The CLI loop `for i < spp: path_trace` (cli.cpp:96-99) as one batched call."""
        lib = load_library()
        check(lib.pt_ctx_set_max_iterations(self._ctx, int(self.max_iterations)))
        cam = camera.to_c()
        check(lib.pt_render(self._ctx, C.byref(cam), int(n_iterations)))

    def render_range(self, camera: Camera, first_iteration: int, n_iterations: int):
        cam = camera.to_c()
        check(load_library().pt_render_range(self._ctx, C.byref(cam), int(first_iteration), int(n_iterations)))

    def synchronize(self):
        check(load_library().pt_sync(self._ctx))

    # -- PathTracer::denoise (path_tracer.cu:479-485)
    def denoise(self, resolution=None):
        d = self.atrous_denoiser
        p = _abi.pt_denoise_params()
        load_library().pt_denoise_params_default(C.byref(p))
        p.filter_size = int(d.filter_size)
        p.color_weight = float(d.color_weight)
        p.normal_weight = float(d.normal_weight)
        p.position_weight = float(d.position_weight)
        p.clamp_fix = 1 if d.clamp_fix else 0
        check(load_library().pt_denoise(self._ctx, C.byref(p)))

    # -- row-band sharding of one frame (multi-GPU interactive path)
    def set_rows(self, row_begin: int, row_end: int):
        """Render / accumulate / denoise rows [row_begin, row_end) only (multiples of 4)."""
        check(load_library().pt_ctx_set_rows(self._ctx, int(row_begin), int(row_end)))

    def halo_rows(self) -> int:
        """Rows of valid neighbour data the denoiser needs beyond either end of a band."""
        p = _abi.pt_denoise_params()
        load_library().pt_denoise_params_default(C.byref(p))
        p.filter_size = int(self.atrous_denoiser.filter_size)
        n = C.c_uint32()
        check(load_library().pt_denoise_halo_rows(C.byref(p), C.byref(n)))
        return int(n.value)

    # -- PathTracer::send_to_preview (path_tracer.cu:487-520)
    def send_to_preview(self, dev_pbo=None, resolution=None, type: int = DisplayBufferType.final, out=None):
        """Tonemap into an RGBA8 image.  dev_pbo: device pointer (int); otherwise the image comes
        back as numpy [H,W,4] — into `out` when given (e.g. a view of PINNED host memory, which
        makes the device->host copy a plain DMA instead of a staged pageable copy)."""
        lib = load_library()
        w, h = self._res
        if dev_pbo is not None:
            check(lib.pt_resolve_rgba8(self._ctx, int(type), C.c_void_p(int(dev_pbo)), 1))
            return None
        if out is None:
            out = np.empty((h, w, 4), dtype=np.uint8)
        elif out.shape != (h, w, 4) or out.dtype != np.uint8 or not out.flags["C_CONTIGUOUS"]:
            raise ValueError("out must be a C-contiguous uint8 array of shape [H, W, 4]")
        check(lib.pt_resolve_rgba8(self._ctx, int(type), out.ctypes.data, 0))
        return out

    def download(self, type: int) -> np.ndarray:
        """Raw float means of a frame buffer: [H,W,3], or [H,W] for depth."""
        w, h = self._res
        out = np.empty((h, w) if type == DisplayBufferType.depth else (h, w, 3), dtype=np.float32)
        check(load_library().pt_download_f32(self._ctx, int(type), out.ctypes.data))
        return out

    def upload_frame(self, color, normal, depth, camera: Camera):
        c = np.ascontiguousarray(color, dtype=np.float32)
        n = np.ascontiguousarray(normal, dtype=np.float32)
        d = np.ascontiguousarray(depth, dtype=np.float32)
        cam = camera.to_c()
        check(load_library().pt_ctx_upload_frame(self._ctx, c.ctypes.data, n.ctypes.data, d.ctypes.data,
                                                 C.byref(cam)))

    def sums_ptr(self):
        p = C.c_void_p()
        n = C.c_uint64()
        check(load_library().pt_ctx_sums(self._ctx, C.byref(p), C.byref(n)))
        return p.value, n.value

    def bind_sums(self, device_ptr: int | None):
        check(load_library().pt_ctx_bind_sums(self._ctx, C.c_void_p(device_ptr or 0)))

    def save_state(self, path: str):
        """Progressive state (running sums + iteration) to disk."""
        check(load_library().pt_ctx_save_state(self._ctx, str(path).encode()))

    def load_state(self, path: str):
        """Resume from a saved progressive state of the same resolution."""
        check(load_library().pt_ctx_load_state(self._ctx, str(path).encode()))

    def set_sample_count(self, n: int):
        check(load_library().pt_ctx_set_sample_count(self._ctx, int(n)))

    def set_stream(self, stream: int | None):
        check(load_library().pt_ctx_set_stream(self._ctx, C.c_void_p(stream or 0)))

    def stats(self) -> _abi.pt_stats:
        s = _abi.pt_stats()
        check(load_library().pt_get_stats(self._ctx, C.byref(s)))
        return s

    def reset_stats(self):
        check(load_library().pt_reset_stats(self._ctx))

    def _destroy_ctx(self):
        if self._ctx and not getattr(self, "_borrowed", False):
            load_library().pt_ctx_destroy(self._ctx)
        self._ctx = None

    def close(self):
        self._destroy_ctx()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class PathTracerGroup:
    """One process, N GPUs (pt_group_*): a scene replica and a PathTracer context per device,
    sample-range sharding + one NCCL reduce (render), or row bands of one frame (render_bands).
    The frame ends up in `root`, an ordinary PathTracer view of device 0's context."""

    def __init__(self, desc: SceneDescription, resolution, devices=None, n_devices: int = 0,
                 max_depth: int = 50, samples_per_pass: int = 0, profile: bool = False):
        lib = load_library()
        d, keep = desc.to_desc()
        p = _abi.pt_params()
        lib.pt_params_default(C.byref(p))
        p.max_depth = max_depth
        p.samples_per_pass = samples_per_pass
        p.profile = 1 if profile else 0
        devs = None
        if devices is not None:
            n_devices = len(devices)
            devs = (C.c_int * n_devices)(*[int(x) for x in devices])
        h = C.c_void_p()
        w, hh = int(resolution[0]), int(resolution[1])
        check(lib.pt_group_create(C.byref(d), devs, int(n_devices), w, hh, C.byref(p), C.byref(h)))
        del keep
        self._g = h
        self._res = (w, hh)
        self.root = PathTracer(max_depth=max_depth)
        self.root._ctx = C.c_void_p(lib.pt_group_ctx(self._g, 0))
        self.root._res = (w, hh)
        self.root._borrowed = True

    def __len__(self):
        return int(load_library().pt_group_size(self._g))

    def devices(self):
        return [int(load_library().pt_group_device(self._g, i)) for i in range(len(self))]

    def restart(self):
        check(load_library().pt_group_restart(self._g))

    def iteration(self) -> int:
        return int(load_library().pt_group_iteration(self._g))

    def render(self, camera: Camera, first_iteration: int, n_iterations: int):
        cam = camera.to_c()
        check(load_library().pt_group_render(self._g, C.byref(cam), int(first_iteration), int(n_iterations)))

    def render_bands(self, camera: Camera, first_iteration: int, n_iterations: int):
        cam = camera.to_c()
        check(load_library().pt_group_render_bands(self._g, C.byref(cam), int(first_iteration), int(n_iterations)))

    def synchronize(self):
        check(load_library().pt_group_sync(self._g))

    def stats(self) -> _abi.pt_stats:
        s = _abi.pt_stats()
        check(load_library().pt_group_get_stats(self._g, C.byref(s)))
        return s

    def close(self):
        if self._g:
            self.root._ctx = None
            load_library().pt_group_destroy(self._g)
            self._g = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def write_image_file(filename: str, rgba: np.ndarray):
    """write_image_file (src/lib/image.cpp:9-22)."""
    img = np.ascontiguousarray(rgba, dtype=np.uint8)
    h, w = img.shape[:2]
    check(load_library().pt_write_png_rgba8(filename.encode(), img.ctypes.data, w, h))


def cli_main(argv: list[str]) -> int:
    """`cuda_pt [options] <filename>` (src/main.cpp, src/cli/cli.cpp)."""
    args = [b"cuda_pt"] + [a.encode() for a in argv]
    arr = (C.c_char_p * len(args))(*args)
    return int(load_library().pt_cli_main(len(args), arr))
