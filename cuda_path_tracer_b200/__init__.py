"""B200-native wavefront path tracer — drop-in for the hot path of LesleyLai/cuda-path-tracer.

Product code: csrc/ (hand-written sm_100a CUDA + host C++ behind the C ABI in include/b200pt.h)
and this thin ctypes mirror of the reference's PathTracer interface.  No CPU fallback.
"""
from ._abi import (BUF_COLOR, BUF_DENOISED, BUF_DEPTH, BUF_FINAL, BUF_NORMAL, LIB_PATH,
                   LibraryMissing, PTError, load_library)
from .api import (DisplayBufferType, EdgeAvoidingATrousDenoiser, GPUMethod, HIT_DTYPE, PathTracer,
                  PathTracerGroup, Scene, cli_main, write_image_file)
from .scene_description import (Camera, Material, Mesh, SceneDescription, bunny_like, bunny_scene,
                                compose, heightfield, many_materials_scene, many_spheres_scene, rotate, scale, terrain_scene,
                                three_balls,
                                translate, write_obj)

__all__ = [n for n in dir() if not n.startswith("_")]
