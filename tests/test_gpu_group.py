"""pt_group_* (C-level multi-GPU, one process / N devices / NCCL) against the single-context
path.  The one-device cases run on any GPU box; the N-device cases need >= 2 devices and are
skipped otherwise (the driver's 8-GPU step and `gpurun --gpus 2` run them)."""
import math
import os

import numpy as np
import pytest

import cuda_path_tracer_b200 as pt
from cuda_path_tracer_b200 import DisplayBufferType as DB

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_devices():
    import torch
    return torch.cuda.device_count()


def _single(sd, spp, depth=8):
    w, h = sd.resolution
    tr = pt.PathTracer(max_depth=depth)
    tr.create_buffers((w, h), sd)
    tr.render_range(sd.camera, 0, spp)
    tr.synchronize()
    return tr.download(DB.color), tr.download(DB.normal), int(tr.stats().rays), tr


def test_one_device_group_is_the_single_context_path():
    sd = pt.bunny_scene(pt.bunny_like(3), 320, 180, 8)
    ref_c, ref_n, ref_rays, _ = _single(sd, 8)
    g = pt.PathTracerGroup(sd, sd.resolution, n_devices=1, max_depth=8)
    assert len(g) == 1 and g.devices() == [0]
    g.render(sd.camera, 0, 5)
    g.render(sd.camera, 5, 3)                       # progressive: the root keeps accumulating
    g.synchronize()
    assert g.iteration() == 8
    c = g.root.download(DB.color)
    assert np.abs(c - ref_c).max() <= 1e-6
    assert int(g.stats().rays) == ref_rays
    g.root.atrous_denoiser.filter_size = 4          # the root is an ordinary context
    g.root.denoise()
    assert g.root.send_to_preview().shape == (180, 320, 4)
    g.restart()
    assert g.iteration() == 0
    g.render_bands(sd.camera, 0, 8)                 # one band == the whole frame
    g.synchronize()
    assert np.array_equal(g.root.download(DB.color), ref_c)
    g.close()


@pytest.mark.parametrize("n", [2, 4, 8])
def test_group_sample_ranges_equal_one_gpu(n):
    """N devices, NCCL reduce of the sums: the frame equals the single-GPU frame up to float
    re-association, ray counts add up exactly, repeated calls accumulate."""
    if _n_devices() < n:
        pytest.skip(f"needs {n} CUDA devices")
    sd = pt.bunny_scene(pt.bunny_like(3), 480, 270, 24)
    ref_c, ref_n, ref_rays, _ = _single(sd, 24)
    g = pt.PathTracerGroup(sd, sd.resolution, n_devices=n, max_depth=8)
    assert len(g) == n
    g.render(sd.camera, 0, 16)
    g.render(sd.camera, 16, 8)
    g.synchronize()
    c, nn = g.root.download(DB.color), g.root.download(DB.normal)
    assert int(g.stats().rays) == ref_rays
    assert math.sqrt(np.mean((c - ref_c) ** 2)) < 1e-6
    assert np.abs(nn - ref_n).max() < 1e-5
    g.close()


@pytest.mark.parametrize("n", [2, 3])
def test_group_row_bands_equal_one_gpu(n):
    """Row bands of ONE 1-spp frame gathered over NCCL: byte-identical to the single-GPU frame,
    before and after the denoiser (which then runs on the root over the whole frame)."""
    if _n_devices() < n:
        pytest.skip(f"needs {n} CUDA devices")
    sd = pt.bunny_scene(pt.bunny_like(3), 488, 274, 1)   # height not a multiple of 4 x n
    ref_c, _, ref_rays, tr = _single(sd, 1)
    tr.atrous_denoiser.filter_size = 16
    tr.denoise()
    ref_img = tr.send_to_preview()
    g = pt.PathTracerGroup(sd, sd.resolution, n_devices=n, max_depth=8)
    g.render_bands(sd.camera, 0, 1)
    g.synchronize()
    assert np.array_equal(g.root.download(DB.color), ref_c)
    assert int(g.stats().rays) == ref_rays
    g.root.atrous_denoiser.filter_size = 16
    g.root.denoise()
    assert np.array_equal(g.root.send_to_preview(), ref_img)
    g.close()


def test_cli_gpus_flag_renders_the_same_image(tmp_path):
    """`cuda_pt --gpus N` (additive flag; the reference hard-codes device 0, cli.cpp:71)."""
    from tests.test_cli import _assets, _read_png, _run
    n = min(2, _n_devices())
    _, cwd = _assets(tmp_path, res=(160, 92))
    imgs = []
    for gpus in (1, n, 0):
        out = str(tmp_path / f"o{gpus}.png")
        r = _run(["--gpus", str(gpus), "--spp", "8", "--max-depth", "6", "-o", out, "scenes/bunny.json"], cwd)
        assert r.returncode == 0, r.stderr
        assert ("gpus:" in r.stdout) == (gpus != 1)
        imgs.append(_read_png(out))
    # sums re-associate across devices: the tonemapped bytes may differ by one step at most
    for img in imgs[1:]:
        assert np.abs(img.astype(int) - imgs[0].astype(int)).max() <= 1
    # fewer samples than GPUs: the frame is split into row bands instead, byte-identical
    a, b = str(tmp_path / "a.png"), str(tmp_path / "b.png")
    assert _run(["--gpus", "1", "--spp", "1", "-o", a, "scenes/bunny.json"], cwd).returncode == 0
    assert _run(["--gpus", "0", "--spp", "1", "-o", b, "scenes/bunny.json"], cwd).returncode == 0
    assert np.array_equal(_read_png(a), _read_png(b))
