"""Generates tests/golden/full_size_bunny_1080p_64spp.npz: BASELINE.json configs[1] (bunny scene,
1920x1080, 64 spp, max depth 8) rendered at FULL size by the CPU oracle (oracle/liboracle.so,
megakernel RNG discipline = the product's default), reduced to what a fixture can hold:

  color_box20 / normal_box20   20x20 box means of the colour and first-hit-normal means  [54,96,3]
  pixel_index / pixel_color    every 97th pixel of the frame, exact                       [21378], [21378,3]
  rays                         rays traced over all bounces

About 35 s on 8 cores.  Run in the authoring container:  python tests/golden/make_full_size_golden.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import cuda_path_tracer_b200 as pt  # noqa: E402  (scene description only; no device code runs)
from tests.oracle_lib import load_oracle  # noqa: E402

W, H, SPP, DEPTH, BOX, STRIDE = 1920, 1080, 64, 8, 20, 97


def main():
    sd = pt.bunny_scene(pt.bunny_like(4), W, H, SPP)
    color, normal, depth, rays = load_oracle().scene(sd).render(sd.camera, W, H, SPP, DEPTH)
    box = lambda a: a.reshape(H // BOX, BOX, W // BOX, BOX, 3).mean(axis=(1, 3), dtype=np.float64).astype(np.float32)
    idx = np.arange(0, W * H, STRIDE, dtype=np.uint32)
    out = dict(color_box20=box(color), normal_box20=box(normal), pixel_index=idx,
               pixel_color=color.reshape(-1, 3)[idx].astype(np.float32), rays=np.uint64(rays),
               config=np.array([W, H, SPP, DEPTH, BOX, STRIDE], dtype=np.uint32))
    path = os.path.join(ROOT, "tests", "golden", "full_size_bunny_1080p_64spp.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: np.shape(v) for k, v in out.items()}, "rays", rays)


if __name__ == "__main__":
    main()
