"""ctypes view of include/b200pt.h.  Loading fails loudly when the CUDA library
has not been built — there is no CPU fallback anywhere in this package."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# B200PT_LIB selects another build of the same library (the bounds-checked debug build)
LIB_PATH = os.environ.get("B200PT_LIB") or os.path.join(HERE, "libb200pt.so")

PT_OK = 0
MAT_DIFFUSE, MAT_METAL, MAT_DIELECTRIC = 0, 1, 2
OBJ_SPHERE, OBJ_MESH = 0, 1
RNG_PIXEL_STREAM, RNG_SLOT_RESEED = 0, 1
BUF_FINAL, BUF_COLOR, BUF_NORMAL, BUF_DEPTH, BUF_DENOISED = 0, 1, 2, 3, 4


class pt_material(C.Structure):
    _fields_ = [("type", C.c_int32), ("albedo", C.c_float * 3), ("fuzz", C.c_float),
                ("refraction_index", C.c_float)]


class pt_sphere(C.Structure):
    _fields_ = [("center", C.c_float * 3), ("radius", C.c_float)]


class pt_object(C.Structure):
    _fields_ = [("type", C.c_int32), ("prim_index", C.c_uint32), ("material", C.c_uint32),
                ("m", C.c_float * 16), ("inv", C.c_float * 16)]


class pt_scene_desc(C.Structure):
    _fields_ = [("positions", C.POINTER(C.c_float)), ("n_vertices", C.c_uint64),
                ("indices", C.POINTER(C.c_uint32)), ("n_indices", C.c_uint64),
                ("objects", C.POINTER(pt_object)), ("n_objects", C.c_uint32),
                ("spheres", C.POINTER(pt_sphere)), ("n_spheres", C.c_uint32),
                ("materials", C.POINTER(pt_material)), ("n_materials", C.c_uint32),
                ("n_meshes", C.c_uint32), ("mesh_first_index", C.POINTER(C.c_uint64))]


class pt_camera(C.Structure):
    _fields_ = [("position", C.c_float * 3), ("rotation", C.c_float * 4), ("vfov", C.c_float)]


class pt_params(C.Structure):
    _fields_ = [("max_depth", C.c_int32), ("rng_mode", C.c_int32), ("max_iterations", C.c_int32),
                ("samples_per_pass", C.c_int32), ("profile", C.c_int32), ("sort_rays", C.c_int32),
                ("lanes", C.c_int32), ("reserved", C.c_int32 * 1)]


class pt_denoise_params(C.Structure):
    _fields_ = [("filter_size", C.c_int32), ("color_weight", C.c_float),
                ("normal_weight", C.c_float), ("position_weight", C.c_float),
                ("clamp_fix", C.c_int32), ("reserved", C.c_int32 * 3)]


class pt_hit(C.Structure):
    _fields_ = [("t", C.c_float), ("point", C.c_float * 3), ("normal", C.c_float * 3),
                ("material", C.c_uint32), ("side", C.c_uint32), ("object", C.c_int32),
                ("prim", C.c_int32), ("pad", C.c_uint32)]


class pt_stats(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("samples", C.c_uint64), ("iterations", C.c_uint32),
                ("passes", C.c_uint32), ("kernel_launches", C.c_uint64),
                ("ms_raygen_extend0", C.c_double), ("ms_extend", C.c_double),
                ("ms_shade", C.c_double), ("ms_compact", C.c_double),
                ("ms_accumulate", C.c_double), ("ms_denoise", C.c_double),
                ("ms_resolve", C.c_double), ("n_extend_launches", C.c_uint64),
                ("n_shade_launches", C.c_uint64), ("max_bounce_reached", C.c_uint32),
                ("reserved", C.c_uint32), ("rays_traversed", C.c_uint64)]


class pt_scene_info(C.Structure):
    _fields_ = [("n_triangles", C.c_uint64), ("n_world_triangles", C.c_uint64),
                ("n_bvh_nodes", C.c_uint64), ("bvh_depth", C.c_uint32), ("n_objects", C.c_uint32),
                ("n_spheres", C.c_uint32), ("n_materials", C.c_uint32), ("build_ms", C.c_double),
                ("upload_ms", C.c_double), ("device_bytes", C.c_uint64),
                ("n_bvh8_nodes", C.c_uint64), ("bvh8_depth", C.c_uint32), ("device_build", C.c_uint32),
                ("n_bvh_triangles", C.c_uint64)]


class pt_scene_file_info(C.Structure):
    _fields_ = [("camera", pt_camera), ("width", C.c_int32), ("height", C.c_int32),
                ("spp", C.c_int32), ("load_ms", C.c_double)]


# every symbol include/b200pt.h declares: name -> (restype, argtypes)
VP = C.c_void_p
SYMBOLS = {
    "pt_last_error": (C.c_char_p, []),
    "pt_version": (C.c_int, []),
    "pt_scene_create": (C.c_int, [C.POINTER(pt_scene_desc), C.c_int, C.POINTER(VP)]),
    "pt_scene_destroy": (C.c_int, [VP]),
    "pt_scene_get_info": (C.c_int, [VP, C.POINTER(pt_scene_info)]),
    "pt_scene_copy_bvh": (C.c_int, [VP, VP, VP]),
    "pt_host_bvh_build": (C.c_int, [C.POINTER(pt_scene_desc), C.c_int, C.POINTER(VP), C.POINTER(pt_scene_info)]),
    "pt_host_bvh_validate": (C.c_int, [VP, C.POINTER(C.c_uint64)]),
    "pt_host_bvh_arrays": (C.c_int, [VP, C.POINTER(VP), C.POINTER(VP), C.POINTER(VP)]),
    "pt_host_bvh_quantised": (C.c_int, [VP, C.POINTER(VP), C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "pt_host_bvh_trace_stats": (C.c_int, [VP, VP, C.c_uint64, C.c_int, VP]),
    "pt_host_bvh_free": (C.c_int, [VP]),
    "pt_host_scene_check": (C.c_int, [C.POINTER(pt_scene_desc), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "pt_scene_load_file": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(VP), C.POINTER(pt_scene_file_info)]),
    "pt_scene_file_read": (C.c_int, [C.c_char_p, C.POINTER(VP), C.POINTER(pt_scene_desc),
                                     C.POINTER(pt_scene_file_info)]),
    "pt_scene_file_free": (C.c_int, [VP]),
    "pt_params_default": (None, [C.POINTER(pt_params)]),
    "pt_denoise_params_default": (None, [C.POINTER(pt_denoise_params)]),
    "pt_ctx_create": (C.c_int, [VP, C.c_uint32, C.c_uint32, C.POINTER(pt_params), VP, C.POINTER(VP)]),
    "pt_ctx_destroy": (C.c_int, [VP]),
    "pt_ctx_resize": (C.c_int, [VP, C.c_uint32, C.c_uint32]),
    "pt_ctx_set_rows": (C.c_int, [VP, C.c_uint32, C.c_uint32]),
    "pt_denoise_halo_rows": (C.c_int, [C.POINTER(pt_denoise_params), C.POINTER(C.c_uint32)]),
    "pt_ctx_restart": (C.c_int, [VP]),
    "pt_ctx_iteration": (C.c_int, [VP]),
    "pt_ctx_set_max_iterations": (C.c_int, [VP, C.c_int]),
    "pt_ctx_set_stream": (C.c_int, [VP, VP]),
    "pt_path_trace": (C.c_int, [VP, C.POINTER(pt_camera)]),
    "pt_render": (C.c_int, [VP, C.POINTER(pt_camera), C.c_int]),
    "pt_render_range": (C.c_int, [VP, C.POINTER(pt_camera), C.c_int, C.c_int]),
    "pt_sync": (C.c_int, [VP]),
    "pt_denoise": (C.c_int, [VP, C.POINTER(pt_denoise_params)]),
    "pt_resolve_rgba8": (C.c_int, [VP, C.c_int, VP, C.c_int]),
    "pt_download_f32": (C.c_int, [VP, C.c_int, VP]),
    "pt_ctx_sums": (C.c_int, [VP, C.POINTER(VP), C.POINTER(C.c_uint64)]),
    "pt_ctx_set_sample_count": (C.c_int, [VP, C.c_int]),
    "pt_ctx_bind_sums": (C.c_int, [VP, VP]),
    "pt_ctx_upload_frame": (C.c_int, [VP, VP, VP, VP, C.POINTER(pt_camera)]),
    "pt_ctx_save_state": (C.c_int, [VP, C.c_char_p]),
    "pt_ctx_load_state": (C.c_int, [VP, C.c_char_p]),
    "pt_get_stats": (C.c_int, [VP, C.POINTER(pt_stats)]),
    "pt_reset_stats": (C.c_int, [VP]),
    "pt_group_create": (C.c_int, [C.POINTER(pt_scene_desc), C.POINTER(C.c_int), C.c_int, C.c_uint32, C.c_uint32,
                                  C.POINTER(pt_params), C.POINTER(VP)]),
    "pt_group_destroy": (C.c_int, [VP]),
    "pt_group_size": (C.c_int, [VP]),
    "pt_group_ctx": (VP, [VP, C.c_int]),
    "pt_group_device": (C.c_int, [VP, C.c_int]),
    "pt_group_scene_info": (C.c_int, [VP, C.POINTER(pt_scene_info)]),
    "pt_group_restart": (C.c_int, [VP]),
    "pt_group_iteration": (C.c_int, [VP]),
    "pt_group_render": (C.c_int, [VP, C.POINTER(pt_camera), C.c_int, C.c_int]),
    "pt_group_render_bands": (C.c_int, [VP, C.POINTER(pt_camera), C.c_int, C.c_int]),
    "pt_group_sync": (C.c_int, [VP]),
    "pt_group_get_stats": (C.c_int, [VP, C.POINTER(pt_stats)]),
    "pt_trace_batch": (C.c_int, [VP, VP, C.c_uint64, VP]),
    "pt_write_png_rgba8": (C.c_int, [C.c_char_p, VP, C.c_uint32, C.c_uint32]),
    "pt_cli_main": (C.c_int, [C.c_int, C.POINTER(C.c_char_p)]),
}

_lib = None


class LibraryMissing(RuntimeError):
    pass


def load_library():
    """Load libb200pt.so and bind every symbol of the public header."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraryMissing(
            f"{LIB_PATH} is missing: build it with `python -m cuda_path_tracer_b200.build` "
            "(nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class PTError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"b200pt error {code}: {msg}")
        self.code = code


def check(rc: int):
    if rc != PT_OK:
        raise PTError(rc, load_library().pt_last_error().decode(errors="replace"))
