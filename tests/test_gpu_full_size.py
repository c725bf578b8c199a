"""GPU tests at BASELINE.json's FULL sizes: against a committed fixture of the oracle's full-size
render (35 s of CPU work, done once by tests/golden/make_full_size_golden.py) and through
properties that need no oracle run of the same size:

  * configs[1] (bunny scene, 1920x1080, 64 spp, depth 8): the render is bit-deterministic, the
    running SUMS make sample ranges additive ([0,32) + [32,64) == [0,64), the property the
    multi-GPU sample-range sharding rests on), ray counts are equal and inside their bounds, and
    the frame equals the oracle's render of the same full-size frame (committed fixture: sampled
    pixels exactly, 20x20 box means everywhere, ray count).
  * configs[3] (10 M-triangle procedural mesh): closest hits are independent of the tree they were
    found with (host SAH build vs device LBVH build: different topology, leaf sizes and traversal
    order, same t).
  * configs[2] (1080p, A-Trous, 5 iterations): the filter's weights are normalised (a constant
    colour plane stays constant whatever the G-buffer holds) and it never leaves the input's range.
"""
import os

import numpy as np
import pytest

import cuda_path_tracer_b200 as pt
from cuda_path_tracer_b200 import DisplayBufferType as DB
from tests.test_gpu_parity import _rays_for, _secondary

pytestmark = pytest.mark.gpu

W, H, SPP, DEPTH = 1920, 1080, 64, 8


def test_full_size_frame_is_deterministic_additive_and_matches_the_golden_frame():
    sd = pt.bunny_scene(pt.bunny_like(4), W, H, SPP)
    scene = pt.Scene.from_description(sd)
    tr = pt.PathTracer(max_depth=DEPTH)
    tr.create_buffers((W, H), scene)

    def frame(ranges):
        tr.restart()
        tr.reset_stats()
        for first, n in ranges:
            tr.render_range(sd.camera, first, n)
        tr.synchronize()
        return tr.download(DB.color), int(tr.stats().rays)

    a, rays_a = frame([(0, SPP)])
    n = tr.download(DB.normal)
    b, rays_b = frame([(0, SPP)])
    assert np.array_equal(a, b) and rays_a == rays_b                      # bit-deterministic
    c, rays_c = frame([(SPP // 2, SPP // 2), (0, SPP // 2)])              # two shards, swapped order
    assert rays_c == rays_a
    assert np.abs(c - a).max() <= 1e-5 * max(1.0, float(np.abs(a).max()))  # float re-association only
    assert W * H * SPP <= rays_a <= W * H * SPP * DEPTH
    assert np.isfinite(a).all() and a.min() >= 0.0 and a.max() <= 1.0 + 1e-5  # albedos and sky <= 1
    # the oracle's render of the SAME full-size frame (tests/golden/make_full_size_golden.py):
    # every 97th pixel exactly, 20x20 box means of the whole frame, and the ray count.  Per-pixel
    # RNG streams are identical, so only rounding-flipped branches may differ.
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "full_size_bunny_1080p_64spp.npz"))
    assert list(g["config"]) == [W, H, SPP, DEPTH, 20, 97]
    diff = np.abs(a.reshape(-1, 3)[g["pixel_index"]] - g["pixel_color"]).max(axis=1)
    assert np.median(diff) < 1e-5 and (diff > 1e-3).mean() < 0.01, (np.median(diff), (diff > 1e-3).mean())
    box = a.reshape(H // 20, 20, W // 20, 20, 3).mean(axis=(1, 3), dtype=np.float64)
    assert np.abs(box - g["color_box20"]).max() < 2e-4, np.abs(box - g["color_box20"]).max()
    nbox = n.reshape(H // 20, 20, W // 20, 20, 3).mean(axis=(1, 3), dtype=np.float64)
    assert np.abs(nbox - g["normal_box20"]).max() < 2e-3, np.abs(nbox - g["normal_box20"]).max()
    assert abs(rays_a - int(g["rays"])) <= 1e-5 * int(g["rays"]), (rays_a, int(g["rays"]))


def test_ten_million_triangle_hits_do_not_depend_on_the_tree(oracle, monkeypatch):
    sd = pt.terrain_scene(2236, 3840, 2160, 16)
    w, h = sd.resolution
    monkeypatch.delenv("PT_BUILD", raising=False)
    sah = pt.Scene.from_description(sd)
    monkeypatch.setenv("PT_BUILD", "lbvh")
    lbvh = pt.Scene.from_description(sd)
    assert int(sah.info.n_world_triangles) == int(lbvh.info.n_world_triangles) >= 9_900_000
    assert int(lbvh.info.device_build) == 1 and int(sah.info.device_build) == 0
    assert int(lbvh.info.n_bvh_nodes) != int(sah.info.n_bvh_nodes)       # really two different trees
    prim, rng = _rays_for(oracle, sd, w, h, n_random=400_000, seed=11)
    first = sah.trace_batch(prim)
    assert (first["t"] > 0).mean() > 0.3
    rays = np.concatenate([prim, _secondary(prim, first, rng)])
    a, b = sah.trace_batch(rays), lbvh.trace_batch(rays)
    hit_a, hit_b = a["t"] > 0, b["t"] > 0
    assert (hit_a != hit_b).sum() <= max(2, len(rays) // 100_000)
    both = hit_a & hit_b
    rel = np.abs(a["t"][both] - b["t"][both]) / np.maximum(a["t"][both], 1e-6)
    assert (rel > 1e-5).sum() <= max(2, int(both.sum()) // 100_000), (rel > 1e-5).sum()
    # same triangle unless two triangles are hit at the same t (shared edges: tie order differs)
    assert (a["prim"][both] != b["prim"][both]).mean() < 2e-3
    assert (a["object"][both] == b["object"][both]).all()


def test_full_size_denoiser_is_normalised_and_range_preserving():
    sd = pt.bunny_scene(pt.bunny_like(2), W, H, 1)
    tr = pt.PathTracer(max_depth=DEPTH)
    tr.create_buffers((W, H), sd)
    tr.render(sd.camera, 1)                                               # a real G-buffer
    tr.synchronize()
    colour, normal, depth = tr.download(DB.color), tr.download(DB.normal), tr.download(DB.depth)
    tr.atrous_denoiser.filter_size = 16                                   # steps 1, 2, 4, 8, 16
    tr.atrous_denoiser.clamp_fix = True                                   # every pixel defined
    const = np.empty_like(colour)
    const[...] = np.float32([0.25, 0.5, 0.75])
    tr.upload_frame(const, normal, depth.reshape(H, W), sd.camera)
    tr.denoise()
    out = tr.download(DB.denoised)
    assert np.abs(out - const).max() <= 1e-5
    tr.upload_frame(colour, normal, depth.reshape(H, W), sd.camera)
    tr.denoise()
    out = tr.download(DB.denoised)
    assert np.isfinite(out).all()
    assert out.min() >= colour.min() - 1e-5 and out.max() <= colour.max() + 1e-5
    assert abs(float(out.mean()) - float(colour.mean())) < 0.05          # smoothing, not re-exposure
    assert float(np.abs(np.diff(out, axis=1)).mean()) < float(np.abs(np.diff(colour, axis=1)).mean())
