"""integration/path_tracer_b200.cpp: the reference's `PathTracer` class (src/lib/path_tracer.hpp:
60-99) implemented on the C ABI and COMPILED against the reference's unmodified headers.  The
driver (integration/ref_cli_driver.cpp) runs the reference CLI's loop (cli.cpp:86-105) —
create_buffers, max_iterations = spp, `for (i < spp) path_trace`, send_to_preview into a managed
buffer — through that class; the frame must equal what the C ABI's own pt_render produces, bit for
bit, in both RNG disciplines, with and without the denoiser."""
import os
import subprocess

import numpy as np
import pytest

import cuda_path_tracer_b200 as pt
from cuda_path_tracer_b200 import DisplayBufferType as DB

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "integration", "_build", "ref_cli_driver")
needs_driver = pytest.mark.skipif(not os.path.exists(DRIVER), reason="integration/_build/ref_cli_driver not built "
                                                                      "(integration/build.sh needs /root/reference)")


def test_shim_sources_cite_and_cover_the_reference_interface():
    """CPU: every public member of the reference class is implemented by the shim."""
    src = open(os.path.join(ROOT, "integration", "path_tracer_b200.cpp")).read()
    for member in ("PathTracer::PathTracer()", "PathTracer::create_buffers(", "PathTracer::resize_image(",
                   "PathTracer::restart()", "PathTracer::path_trace(", "PathTracer::denoise(",
                   "PathTracer::send_to_preview("):
        assert member in src, member
    assert "path_tracer.hpp:60-99" in src and '#include "path_tracer.hpp"' in src
    assert "oracle" not in src.replace("oracle/ref_shim", "")        # the product side never names the oracle


@needs_driver
@pytest.mark.gpu
@pytest.mark.parametrize("method,filter_size", [("megakernel", 0), ("streaming", 0), ("megakernel", 8)])
def test_reference_cli_loop_through_the_shim_equals_pt_render(tmp_path, method, filter_size):
    from tests.test_cli import _assets
    _assets(tmp_path, res=(160, 92))
    scene = str(tmp_path / "assets" / "scenes" / "bunny.json")
    out = str(tmp_path / "shim.rgba")
    spp, depth = 5, 8
    r = subprocess.run([DRIVER, scene, out, method, str(depth), str(spp), str(filter_size)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert f"spp {spp} iteration {spp}" in r.stdout
    shim = np.fromfile(out, dtype=np.uint8).reshape(92, 160, 4)

    s = pt.Scene.from_file(scene)
    assert (s.file_info.width, s.file_info.height) == (160, 92)
    tr = pt.PathTracer(max_depth=depth)
    tr.current_gpu_method = pt.GPUMethod.megakernel if method == "megakernel" else pt.GPUMethod.streaming
    tr.max_iterations = spp
    tr.create_buffers((160, 92), s)
    tr.render(pt.Camera.from_c(s.file_info.camera), spp)
    if filter_size:
        tr.atrous_denoiser.filter_size = filter_size
        tr.denoise()
    ours = tr.send_to_preview(type=DB.final)
    assert np.array_equal(shim, ours)
