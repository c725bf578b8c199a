"""In-tree build of libb200pt.so (sm_100a) and the cuda_pt executable.

nvcc cross-compiles without a GPU; the built .so travels to the GPU box with the
repo snapshot (it is git-ignored, not gpurun-ignored).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libb200pt.so")
LIB_DEBUG = os.path.join(HERE, "libb200pt_debug.so")
EXE = os.path.join(HERE, "cuda_pt")

SOURCES = ["wavefront.cu", "image_kernels.cu", "context.cpp", "bvh_build.cpp", "lbvh_host.cpp", "lbvh.cu", "hostmath.cpp", "scene_io.cpp",
           "cli.cpp", "group.cpp"]
HEADERS = ["common.cuh", "kernels.h", "internal.h", "bvh_build.h", "lbvh.h", "../../include/b200pt.h"]

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler", "-fPIC,-fopenmp,-fvisibility=hidden,-O3",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_debug(verbose: bool = False) -> str:
    """libb200pt_debug.so: the same sources with -DPT_BOUNDS_CHECK (device-side index checks that
    trap; common.cuh).  Select it with B200PT_LIB=<path> and run the GPU parity suite against it:
    scripts/run_bounds_check.sh.  Not built by default."""
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    cmd = [_nvcc(), *NVCC_FLAGS, "-DPT_BOUNDS_CHECK", "-shared", "-o", LIB_DEBUG, *srcs, "-lz", "-lgomp"]
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True, cwd=CSRC)
    return LIB_DEBUG


def build_variant(name: str, defines: list[str], verbose: bool = False) -> str:
    """libb200pt_<name>.so: the same sources with extra -D flags, for A/B runs of compile-time
    choices (select with B200PT_LIB=<path>, e.g. through scripts/ab.py).  Not built by default."""
    out = os.path.join(HERE, f"libb200pt_{name}.so")
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    cmd = [_nvcc(), *NVCC_FLAGS, *defines, "-shared", "-o", out, *srcs, "-lz", "-lgomp"]
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True, cwd=CSRC)
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    if force or _stale(LIB, deps):
        cmd = [_nvcc(), *NVCC_FLAGS, "-shared", "-o", LIB, *srcs, "-lz", "-lgomp"]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True, cwd=CSRC)
    main_src = os.path.join(CSRC, "cuda_pt_main.cpp")
    if force or _stale(EXE, [main_src, LIB]):
        subprocess.run(
            ["g++", "-O2", "-o", EXE, main_src, "-L" + HERE, "-lb200pt", "-Wl,-rpath,$ORIGIN"],
            check=True,
        )
    return LIB


if __name__ == "__main__":
    if "--debug" in sys.argv:
        print(build_debug(verbose="-v" in sys.argv))
    elif "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], [a for a in sys.argv[i + 2:] if a.startswith("-D")], verbose="-v" in sys.argv))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
