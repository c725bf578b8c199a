#!/usr/bin/env bash
# build_ref.sh — compile the REFERENCE (LesleyLai/cuda-path-tracer) from the sources where they
# lie under /root/reference into oracle/_ref/ (git-ignored; travels to the GPU box).
#   oracle/_ref/libref_host.so  g++   host code: BVH builder, intersection routines, transforms
#   oracle/_ref/libref_cuda.so  nvcc  sm_100: the reference's kernels + PathTracer class
# The reference's own build (CMake + Conan) is not run: its dependencies (glm, assimp, stb,
# cxxopts, fmt, spdlog) are absent and there is no network.  glm/fmt/spdlog are replaced by the
# minimal stand-ins in oracle/ref_shim; the asset/CLI layer (assimp, cxxopts, stb) is not built —
# scenes reach the reference through SceneDescription's public add_* API (ref_cuda_wrap.cu).
# No reference source is copied into the repository: one file, path_tracer.cu, is read through
# a patched temporary copy (three sed edits below); everything else is #included in place.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${REFERENCE_DIR:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF/src/lib" ]; then
  echo "build_ref.sh: $REF not present (GPU box): using prebuilt $OUT" >&2
  exit 0
fi
mkdir -p "$OUT"
TMP="$(mktemp -d /tmp/b200pt_ref.XXXXXX)"
trap 'rm -rf "$TMP"' EXIT
CXX="${REF_CXX:-/usr/bin/g++}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"

# ---- host library (runs anywhere): pins oracle/oracle.c against the real implementation
# -mfma -ffp-contract=fast: same contraction policy as the oracle and as nvcc on the device
"$CXX" -std=c++20 -O2 -mfma -ffp-contract=fast -fPIC -shared -fvisibility=hidden \
  -I "$HERE/ref_shim" -I "$REF/src/lib" -I "$REF/src" \
  -o "$OUT/libref_host.so" "$HERE/ref_host_wrap.cpp"

# ---- CUDA library: the reference's kernels and PathTracer, recompiled for sm_100
sed -e 's/i < max_bounces \&\& paths_count > 0/i < g_ref_max_bounces \&\& paths_count > 0/' \
    -e 's/for (int i = 0; i < max_bounces; ++i) {/for (int i = 0; i < c_ref_max_bounces; ++i) {/' \
    -e 's/StaticStack<unsigned int, 24> node_stack;/StaticStack<unsigned int, REF_STACK_SIZE> node_stack;/' \
    -e 's/^        const PathsView paths_view{paths_, paths_count};$/        const PathsView paths_view{paths_, paths_count}; g_ref_ray_count += paths_count;/' \
    "$REF/src/lib/path_tracer.cu" > "$TMP/path_tracer.cu"
for pat in g_ref_max_bounces c_ref_max_bounces REF_STACK_SIZE g_ref_ray_count; do
  grep -q "$pat" "$TMP/path_tracer.cu" || { echo "build_ref.sh: patch '$pat' did not apply" >&2; exit 1; }
done
# the reference's flags (cmake/compiler.cmake:65-72) + the arch it never sets (src/lib/CMakeLists.txt:69).
# -rdc=true stays: it IS the reference's configuration (CUDA_SEPARABLE_COMPILATION ON,
# src/lib/CMakeLists.txt:70, src/CMakeLists.txt:8) and its sources do not compile without it
# (constant_memory.cuh:6 declares `inline __constant__ GPUCamera`, which nvcc rejects in
# whole-program mode).  This is a unity build, so the device link sees one object and can inline
# across the reference's files exactly as its own build's device link does.
# Two libraries from the same sources:
#   libref_cuda.so        traversal stack 64: the PARITY CHECKER (deep trees, e.g. 10 M triangles)
#   libref_cuda_stock.so  traversal stack 24 = the reference's own value: the TIMED BASELINE of
#                         bench.py --impl reference wherever the reference's tree is at most 23 deep
build_cuda() { # $1 = stack size, $2 = output
  "$NVCC" -std=c++20 -O3 -DNDEBUG -arch=sm_100 -rdc=true --expt-relaxed-constexpr --extended-lambda -lineinfo \
    -Xcompiler -fPIC,-fvisibility=hidden -shared \
    -DREF_PATCHED_PATH_TRACER="\"$TMP/path_tracer.cu\"" -DREF_STACK_SIZE="$1" \
    -I "$HERE/ref_shim" -I "$REF/src/lib" -I "$REF/src" \
    -o "$2" "$HERE/ref_cuda_wrap.cu"
}
build_cuda "${REF_STACK_SIZE:-64}" "$OUT/libref_cuda.so" &
build_cuda 24 "$OUT/libref_cuda_stock.so" &
wait
[ -s "$OUT/libref_cuda.so" ] && [ -s "$OUT/libref_cuda_stock.so" ] || { echo "build_ref.sh: CUDA reference build failed" >&2; exit 1; }
echo "built $OUT/libref_host.so $OUT/libref_cuda.so $OUT/libref_cuda_stock.so"
