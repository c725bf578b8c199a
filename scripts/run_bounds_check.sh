#!/bin/bash
# The GPU parity suite against the bounds-checked debug build (stand-in for compute-sanitizer,
# which is closed on the GPU pool): every traversal-stack, tree, triangle, parked-state and
# bin-list index is checked on the device and a violation traps.  Build the library first, here or
# on the box:  python -m cuda_path_tracer_b200.build --debug
set -u
LIB="$(cd "$(dirname "$0")/.." && pwd)/cuda_path_tracer_b200/libb200pt_debug.so"
[ -f "$LIB" ] || python -m cuda_path_tracer_b200.build --debug || exit 1
B200PT_LIB="$LIB" python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_full_size.py \
  tests/test_gpu_full_configs.py -m gpu -q -x "$@"
