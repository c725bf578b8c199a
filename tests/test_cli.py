"""The `cuda_pt` executable (csrc/cli.cpp) against the reference's command-line contract
(src/lib/configurations.cpp:9-41, src/cli/cli.cpp:86-115, README.md:57-60): grammar, asset
discovery, messages and exit codes on CPU; a full render against the API on the GPU."""
import json
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest

import cuda_path_tracer_b200 as pt
from tests.test_abi_and_host import BUNNY_JSON

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "cuda_path_tracer_b200", "cuda_pt")


def _assets(tmp_path, res=(96, 54), subdiv=2):
    (tmp_path / "assets" / "scenes").mkdir(parents=True)
    (tmp_path / "assets" / "models").mkdir(parents=True)
    mesh = pt.bunny_like(subdiv)
    pt.write_obj(str(tmp_path / "assets" / "models" / "bunny.obj"), mesh)
    js = json.loads(json.dumps(BUNNY_JSON))
    js["camera"] = {"vfov": 60, "resolution": list(res)}
    js["sampler"] = {"type": "independent", "samples": 3}
    (tmp_path / "assets" / "scenes" / "bunny.json").write_text(json.dumps(js))
    return mesh, str(tmp_path / "assets" / "scenes")      # cwd below assets/: discovery walks up


def _run(args, cwd):
    return subprocess.run([EXE, *args], cwd=cwd, capture_output=True, text=True, timeout=300)


def _read_png(path):
    data = open(path, "rb").read()
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, w, h = 8, b"", 0, 0
    while pos < len(data):
        n, kind = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + n]
        assert zlib.crc32(kind + body) == struct.unpack(">I", data[pos + 8 + n:pos + 12 + n])[0]
        if kind == b"IHDR":
            w, h, depth, ctype = struct.unpack(">IIBB", body[:10])
            assert (depth, ctype) == (8, 6)
        elif kind == b"IDAT":
            idat += body
        pos += 12 + n
    raw = np.frombuffer(zlib.decompress(idat), dtype=np.uint8).reshape(h, 1 + 4 * w)
    assert not raw[:, 0].any()                              # filter type 0 on every row
    return raw[:, 1:].reshape(h, w, 4)


def test_cli_grammar_messages_and_exit_codes(tmp_path, tmp_path_factory):
    _, cwd = _assets(tmp_path)
    r = _run([], cwd)
    assert r.returncode == 1 and "Usage: cuda_pt [options] <filename>" in r.stderr
    r = _run(["-h"], cwd)
    assert r.returncode == 0 and "A Path Tracer written in CUDA" in r.stdout and "--spp" in r.stdout
    assert "--gpus" in r.stdout
    r = _run(["--gpus", "-2", "-o", "x.png", "scenes/bunny.json"], cwd)
    assert r.returncode == 1 and "gpus" in r.stderr
    # an output the writer cannot produce is refused BEFORE anything is rendered, exit code 1
    r = _run(["-o", "x.jpeg", "scenes/bunny.json"], cwd)
    assert r.returncode == 1 and "unrecognized extension" in r.stderr and "Start path tracing" not in r.stdout
    r = _run(["--no-such-flag", "scenes/bunny.json"], cwd)
    assert r.returncode == 1 and "does not exist" in r.stderr
    r = _run(["--spp"], cwd)
    assert r.returncode == 1 and "missing an argument" in r.stderr
    r = _run(["scenes/bunny.json"], cwd)                    # the viewer is not part of this build
    assert r.returncode == 1 and "--output" in r.stderr
    r = _run(["--output", "x.png", "scenes/missing.json"], cwd)
    assert r.returncode == 1 and "Panic" in r.stderr
    r = _run(["--output", "x.png", "models/bunny.obj"], cwd)
    assert r.returncode == 1 and "Unsupported file extension" in r.stderr
    other = tmp_path_factory.mktemp("no_assets_above")     # discovery walks UP from the cwd only
    r = _run(["--output", "x.png", "scenes/bunny.json"], str(other))
    assert r.returncode == 1 and "Cannot find assets directory" in r.stderr


def test_cli_without_a_device_fails_loudly(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    _, cwd = _assets(tmp_path)
    r = _run(["--output", str(tmp_path / "o.png"), "scenes/bunny.json"], cwd)
    assert r.returncode == 1 and "Panic" in r.stderr and not (tmp_path / "o.png").exists()


@pytest.mark.gpu
def test_cli_render_equals_api_and_resume_equals_one_run(tmp_path):
    mesh, cwd = _assets(tmp_path, res=(160, 90), subdiv=3)
    out = str(tmp_path / "out.png")
    stats = str(tmp_path / "stats.json")
    r = _run(["--output", out, "--spp", "6", "--max-depth", "8", "--stats-json", stats, "scenes/bunny.json"], cwd)
    assert r.returncode == 0, r.stderr
    for line in ("Start path tracing", "spp: 6", "width: 160, height: 90", "Done path tracing scenes/bunny.json!",
                 "Elapsed time"):
        assert line in r.stdout, r.stdout
    img = _read_png(out)
    sd = pt.bunny_scene(mesh, 160, 90)
    tr = pt.PathTracer(max_depth=8)
    tr.max_iterations = 6
    tr.create_buffers((160, 90), sd)
    tr.render(sd.camera, 6)
    api = tr.send_to_preview()
    assert (np.abs(img.astype(int) - api.astype(int)) > 1).mean() < 0.01
    st = json.load(open(stats))
    assert st["spp"] == 6 and st["rays"] == int(tr.stats().rays) and st["triangles"] == 2 * mesh.triangle_count
    # scene default spp (3) + checkpoint, then resume up to 6: the same picture as the single run
    state = str(tmp_path / "state.bin")
    r = _run(["--output", str(tmp_path / "a.png"), "--max-depth", "8", "--checkpoint", state, "scenes/bunny.json"], cwd)
    assert r.returncode == 0 and "spp: 3" in r.stdout
    r = _run(["--output", str(tmp_path / "b.png"), "--max-depth", "8", "--spp", "6", "--resume", state,
              "scenes/bunny.json"], cwd)
    assert r.returncode == 0, r.stderr
    resumed = _read_png(str(tmp_path / "b.png"))
    assert (np.abs(resumed.astype(int) - img.astype(int)) > 1).mean() < 1e-3
    # denoiser flag produces a different (smoother) picture of the same size
    r = _run(["--output", str(tmp_path / "d.png"), "--spp", "2", "--max-depth", "8", "--filter-size", "8",
              "scenes/bunny.json"], cwd)
    assert r.returncode == 0 and "Denoising" in r.stdout
    assert _read_png(str(tmp_path / "d.png")).shape == img.shape


@pytest.mark.gpu
def test_cli_all_meshes_gives_every_mesh_object_its_own_mesh(tmp_path):
    """`--all-meshes` (extension): the scene file's mesh table reaches the scene upload, so the two
    mesh objects instance b.obj and a.obj; the default keeps the reference's first-mesh-only rule
    (scene_description.cpp:95)."""
    _, cwd = _assets(tmp_path, res=(96, 54))
    a, b = pt.bunny_like(1), pt.bunny_like(2)
    models = tmp_path / "assets" / "models"
    pt.write_obj(str(models / "a.obj"), a)
    pt.write_obj(str(models / "b.obj"), b)
    scene = tmp_path / "assets" / "scenes" / "bunny.json"
    js = json.loads(scene.read_text())
    js["surfaces"][1]["filename"] = "../models/b.obj"
    js["surfaces"][2]["filename"] = "../models/a.obj"
    scene.write_text(json.dumps(js))
    stats = str(tmp_path / "stats.json")
    r = _run(["--output", str(tmp_path / "one.png"), "--spp", "1", "--stats-json", stats, "scenes/bunny.json"], cwd)
    assert r.returncode == 0, r.stderr
    assert json.load(open(stats))["triangles"] == 2 * a.triangle_count
    r = _run(["--output", str(tmp_path / "all.png"), "--spp", "1", "--all-meshes", "--stats-json", stats,
              "scenes/bunny.json"], cwd)
    assert r.returncode == 0, r.stderr
    assert json.load(open(stats))["triangles"] == a.triangle_count + b.triangle_count
    assert (_read_png(str(tmp_path / "one.png")) != _read_png(str(tmp_path / "all.png"))).any()
