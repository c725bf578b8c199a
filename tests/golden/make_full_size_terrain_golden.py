"""Generates tests/golden/full_size_terrain_4k_16spp.npz: BASELINE.json configs[3] (procedural
10 M-triangle height field, 3840x2160, 16 spp, max depth 8) rendered at FULL size by the CPU oracle
(oracle/liboracle.so: the reference's un-culled stack traversal over the restated bvh_from_mesh
tree, megakernel RNG discipline = the product's default), reduced to what a fixture can hold:

  color_box40 / normal_box40   40x40 box means of the colour and first-hit-normal means  [54,96,3]
  depth_box40                  40x40 box means of the first-hit depth                     [54,96]
  pixel_index / pixel_color    every 397th pixel of the frame, exact                      [20893], [20893,3]
  rays                         rays traced over all bounces

About 45-60 minutes on 8 cores (0.12 Mrays/s: the reference's traversal does not cull by distance).
Run in the authoring container:  python tests/golden/make_full_size_terrain_golden.py
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import cuda_path_tracer_b200 as pt  # noqa: E402  (scene description only; no device code runs)
from tests.oracle_lib import load_oracle  # noqa: E402

W, H, SPP, DEPTH, BOX, STRIDE, N = 3840, 2160, 16, 8, 40, 397, 2236


def main():
    sd = pt.terrain_scene(N, W, H, SPP)
    t0 = time.time()
    color, normal, depth, rays = load_oracle().scene(sd).render(sd.camera, W, H, SPP, DEPTH)
    print(f"oracle render {time.time() - t0:.0f} s", flush=True)
    box3 = lambda a: a.reshape(H // BOX, BOX, W // BOX, BOX, 3).mean(axis=(1, 3), dtype=np.float64).astype(np.float32)
    box1 = lambda a: a.reshape(H // BOX, BOX, W // BOX, BOX).mean(axis=(1, 3), dtype=np.float64).astype(np.float32)
    idx = np.arange(0, W * H, STRIDE, dtype=np.uint32)
    out = dict(color_box40=box3(color), normal_box40=box3(normal), depth_box40=box1(depth), pixel_index=idx,
               pixel_color=color.reshape(-1, 3)[idx].astype(np.float32), rays=np.uint64(rays),
               config=np.array([W, H, SPP, DEPTH, BOX, STRIDE, N], dtype=np.uint32))
    path = os.path.join(ROOT, "tests", "golden", "full_size_terrain_4k_16spp.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: np.shape(v) for k, v in out.items()}, "rays", rays)


if __name__ == "__main__":
    main()
