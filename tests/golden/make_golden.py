"""Generates tests/golden/ref_host_vectors.npz from the REFERENCE's own host code
(oracle/_ref/libref_host.so, built by oracle/build_ref.sh from /root/reference).
Run in the authoring container:  python tests/golden/make_golden.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from tests import ref_lib  # noqa: E402
from tests.test_oracle_pinning import _cases, _mesh_cases, _run_prims  # noqa: E402

ref = ref_lib.load_ref_host()
ht, hs, ab = _run_prims(ref.lib, "ref_")
out = {"tri_t": ht["t"], "tri_normal": ht["normal"], "tri_side": ht["side"],
       "sph_t": hs["t"], "sph_normal": hs["normal"], "sph_side": hs["side"], "aabb": ab,
       "hash": np.array([ref.lib.ref_hash(a) for a in (0, 1, 12345, 0xFFFFFFFF)], dtype=np.uint32)}
for name, mesh in _mesh_cases().items():
    nodes, _ = ref.bvh_from_mesh(mesh.positions, mesh.indices)
    out["bvh_first_" + name] = nodes["first"]
    out["bvh_count_" + name] = nodes["count"]
    out["bvh_min_" + name] = nodes["min"]
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_host_vectors.npz"), **out)
print("wrote", {k: v.shape for k, v in out.items()})
