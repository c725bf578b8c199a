#!/usr/bin/env bash
# integration/build.sh — compiles integration/path_tracer_b200.cpp (the reference's PathTracer on
# the C ABI) against the reference's UNMODIFIED headers where they lie under /root/reference, plus
# the reference's own SceneDescription / Camera / prelude sources and the test driver, into
# integration/_build/ref_cli_driver (git-ignored; travels to the GPU box).  glm / fmt / spdlog come
# from oracle/ref_shim (absent third-party dependencies).  No reference source is copied.
# scene_description.cpp references bvh_from_mesh (its build_scene(), which the shim never calls):
# accelerators/bvh.cpp is linked to satisfy it.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(dirname "$HERE")"
REF="${REFERENCE_DIR:-/root/reference}"
OUT="$HERE/_build"
if [ ! -d "$REF/src/lib" ]; then
  echo "integration/build.sh: $REF not present (GPU box): using prebuilt $OUT" >&2
  exit 0
fi
mkdir -p "$OUT"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
LIBDIR="$ROOT/cuda_path_tracer_b200"
"$NVCC" -std=c++20 -O2 -DNDEBUG -x cu -arch=sm_100 --expt-relaxed-constexpr --extended-lambda \
  -I "$ROOT/oracle/ref_shim" -I "$REF/src/lib" -I "$REF/src" \
  "$HERE/path_tracer_b200.cpp" "$HERE/ref_cli_driver.cpp" \
  "$REF/src/lib/scene_description.cpp" "$REF/src/lib/camera.cpp" "$REF/src/lib/prelude.cpp" \
  "$REF/src/lib/accelerators/bvh.cpp" "$REF/src/lib/cuda_utils/cuda_check.cpp" \
  -L "$LIBDIR" -lb200pt -Xlinker -rpath -Xlinker '$ORIGIN/../../cuda_path_tracer_b200' \
  -o "$OUT/ref_cli_driver"
echo "built $OUT/ref_cli_driver"
