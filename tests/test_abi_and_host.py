"""CPU-side checks: the C-ABI library loads and exports every symbol include/b200pt.h declares,
the host scene reader (JSON + OBJ) follows the reference's grammar, host transform math."""
import ctypes as C
import json
import math
import os
import re

import numpy as np
import pytest

import cuda_path_tracer_b200 as pt
from cuda_path_tracer_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "b200pt.h")).read()
    declared = set(re.findall(r"PT_API\s+[\w\s\*]+?\b(pt_\w+)\s*\(", hdr))
    assert len(declared) >= 30
    lib = pt.load_library()
    for name in declared:
        assert hasattr(lib, name), f"libb200pt.so does not export {name}"
    assert declared == set(_abi.SYMBOLS), declared ^ set(_abi.SYMBOLS)
    assert lib.pt_version() >= 100


def test_struct_layouts_match_header():
    assert C.sizeof(_abi.pt_hit) == 48
    assert C.sizeof(_abi.pt_camera) == 32
    assert C.sizeof(_abi.pt_material) == 24
    assert C.sizeof(_abi.pt_object) == 12 + 128
    p = _abi.pt_params()
    pt.load_library().pt_params_default(C.byref(p))
    assert p.max_depth == 50 and p.rng_mode == 0          # max_bounces = 50 (path_tracer.cu:27)
    d = _abi.pt_denoise_params()
    pt.load_library().pt_denoise_params_default(C.byref(d))
    assert (d.filter_size, round(d.color_weight, 2), round(d.normal_weight, 2), round(d.position_weight, 2)) == \
        (10, 0.45, 0.30, 0.25)                             # edge_avoiding_a_trous_denoiser.hpp:9-12


def test_no_compute_without_gpu_fails_loudly():
    """The product must not fall back to a CPU path: without a device, scene creation errors."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pt.PTError):
        pt.Scene.from_description(pt.three_balls(8, 8))


def _read(path):
    lib = pt.load_library()
    h = C.c_void_p()
    desc = _abi.pt_scene_desc()
    info = _abi.pt_scene_file_info()
    rc = lib.pt_scene_file_read(path.encode(), C.byref(h), C.byref(desc), C.byref(info))
    return rc, h, desc, info


def _write_scene(tmp_path, scene_json, mesh=None):
    (tmp_path / "scenes").mkdir(exist_ok=True)
    (tmp_path / "models").mkdir(exist_ok=True)
    if mesh is not None:
        pt.write_obj(str(tmp_path / "models" / "bunny.obj"), mesh)
    p = tmp_path / "scenes" / "scene.json"
    p.write_text(json.dumps(scene_json))
    return str(p)


BUNNY_JSON = {
    "camera": {"vfov": 60, "resolution": [1920, 1080]},
    "sampler": {"type": "independent", "samples": 10},
    "background": [1, 1, 1],
    "materials": [{"type": "lambertian", "name": "ground", "albedo": [0.8, 0.8, 0.8]},
                  {"type": "lambertian", "name": "bunny", "albedo": [0.8, 0.8, 0.5]},
                  {"type": "lambertian", "name": "bunny2", "albedo": [0.6, 0.4, 0.8]}],
    "surfaces": [{"type": "sphere", "transform": {"translate": [0.0, -100.5, -1.0]}, "radius": 100.0,
                  "material": "ground"},
                 {"type": "mesh", "filename": "../models/bunny.obj",
                  "transform": {"translate": [1.0, -0.5, -2.0]}, "material": "bunny"},
                 {"type": "mesh", "filename": "../models/bunny.obj",
                  "transform": [{"scale": 0.5}, {"translate": [-1.0, -0.5, -2.0]}], "material": "bunny2"}],
}


def test_scene_json_matches_reference_semantics(tmp_path):
    """assets/scenes/bunny.json through the host reader == the Python SceneDescription mirror:
    alphabetical material table, transform lists composed left to right, de-indexed OBJ."""
    mesh = pt.bunny_like(1)
    rc, h, d, info = _read(_write_scene(tmp_path, BUNNY_JSON, mesh))
    assert rc == 0, pt.load_library().pt_last_error()
    assert (info.width, info.height, info.spp) == (1920, 1080, 10)
    assert math.isclose(info.camera.vfov, math.radians(60), rel_tol=1e-6)
    assert list(info.camera.rotation) == [1, 0, 0, 0] and list(info.camera.position) == [0, 0, 0]
    assert d.n_materials == 3 and d.n_objects == 3 and d.n_spheres == 1
    # std::map order: bunny, bunny2, ground
    assert [round(d.materials[i].albedo[2], 2) for i in range(3)] == [0.5, 0.8, 0.8]
    assert [d.objects[i].material for i in range(3)] == [2, 0, 1]
    assert d.n_indices == mesh.indices.size and d.n_vertices == mesh.positions.shape[0]
    pos = np.ctypeslib.as_array(d.positions, shape=(d.n_vertices * 3,)).reshape(-1, 3)
    assert np.allclose(pos, mesh.positions, rtol=1e-6, atol=1e-7)
    ref = pt.bunny_scene(mesh)
    rd, keep = ref.to_desc()
    for i in range(3):
        assert np.allclose(list(d.objects[i].m), list(rd.objects[i].m), atol=1e-6)
        assert np.allclose(list(d.objects[i].inv), list(rd.objects[i].inv), atol=1e-5)
        assert d.objects[i].type == rd.objects[i].type
    pt.load_library().pt_scene_file_free(h)


def test_camera_transform_decomposition(tmp_path):
    """ajax-white.json style {from, at, up} camera -> position + quaternion (glm::decompose)."""
    js = dict(BUNNY_JSON)
    js["camera"] = {"transform": {"from": [6, 5.5, 0], "at": [0, 3.5, 0], "up": [0, 1, 0]}, "vfov": 80,
                    "resolution": [720, 1280]}
    rc, h, d, info = _read(_write_scene(tmp_path, js, pt.bunny_like(0)))
    assert rc == 0
    assert np.allclose(list(info.camera.position), [6, 5.5, 0], atol=1e-6)
    ref = pt.Camera.look_at((6, 5.5, 0), (0, 3.5, 0), (0, 1, 0), 80)
    q, qr = np.array(list(info.camera.rotation)), np.array(ref.rotation)
    assert np.allclose(q, qr, atol=1e-5) or np.allclose(q, -qr, atol=1e-5)
    js["camera"] = {"transform": [{"rotate": 90, "axis": [0, 1, 0]}, {"translate": [0, 0, 4]}], "vfov": 45,
                    "resolution": [8, 8]}
    rc, h2, d2, info2 = _read(_write_scene(tmp_path, js, pt.bunny_like(0)))
    assert rc == 0 and np.allclose(list(info2.camera.position), [0, 0, 4], atol=1e-6)
    assert np.allclose(np.abs(list(info2.camera.rotation)), [math.sqrt(.5), 0, math.sqrt(.5), 0], atol=1e-6)


def test_scene_reader_errors(tmp_path):
    lib = pt.load_library()
    bad = dict(BUNNY_JSON)
    bad["camera"] = {"transform": {"o": [0, 0, 4]}, "vfov": 45}     # three_balls.json:3-9 is rejected
    rc, *_ = _read(_write_scene(tmp_path, bad, pt.bunny_like(0)))
    assert rc == 4 and b"Unrecognized transform" in lib.pt_last_error()
    bad = dict(BUNNY_JSON)
    bad["materials"] = [{"type": "plastic", "name": "x"}]
    rc, *_ = _read(_write_scene(tmp_path, bad, pt.bunny_like(0)))
    assert rc == 4 and b"Unsupported material type" in lib.pt_last_error()
    rc, *_ = _read(str(tmp_path / "nope.json"))
    assert rc == 3
    (tmp_path / "scenes" / "broken.json").write_text("{ \"camera\": ")
    rc, *_ = _read(str(tmp_path / "scenes" / "broken.json"))
    assert rc == 4
    # `file >> json` (json_parser.cpp:166-167) reads one value and ignores what follows; so do we
    pt.write_obj(str(tmp_path / "models" / "bunny.obj"), pt.bunny_like(0))
    (tmp_path / "scenes" / "trailing.json").write_text(json.dumps(BUNNY_JSON) + " x")
    rc, h, *_ = _read(str(tmp_path / "scenes" / "trailing.json"))
    assert rc == 0
    lib.pt_scene_file_free(h)
    # pathological nesting is an error, not a stack overflow
    (tmp_path / "scenes" / "deep.json").write_text("[" * 100000)
    rc, *_ = _read(str(tmp_path / "scenes" / "deep.json"))
    assert rc == 4 and b"nested deeper" in lib.pt_last_error()


def test_obj_reader_polygons_negative_indices_and_first_mesh(tmp_path):
    (tmp_path / "scenes").mkdir()
    (tmp_path / "models").mkdir()
    (tmp_path / "models" / "bunny.obj").write_text(
        "# quad + triangle, then a second object that must be ignored (aiScene::mMeshes[0])\n"
        "v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nvn 0 0 1\n"
        "f 1//1 2//1 3//1 4//1\nf -4 -3 -2\no second\nv 5 5 5\nf 1 2 5\n")
    (tmp_path / "scenes" / "scene.json").write_text(json.dumps(BUNNY_JSON))
    rc, h, d, info = _read(str(tmp_path / "scenes" / "scene.json"))
    assert rc == 0
    assert d.n_indices == 9 and d.n_vertices == 9          # fan-triangulated quad + 1 triangle
    pos = np.ctypeslib.as_array(d.positions, shape=(27,)).reshape(-1, 3)
    assert np.array_equal(pos[3:6], [[0, 0, 0], [1, 1, 0], [0, 1, 0]])
    assert np.array_equal(pos[6:9], [[0, 0, 0], [1, 0, 0], [1, 1, 0]])


def test_png_writer_roundtrip(tmp_path):
    from PIL import Image
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, size=(37, 53, 4), dtype=np.uint8)
    path = str(tmp_path / "out.png")
    pt.write_image_file(path, img)
    back = np.asarray(Image.open(path))
    assert back.shape == img.shape and np.array_equal(back, img)


def test_python_scene_description_mirror():
    s = pt.three_balls(64, 64)
    d, keep = s.to_desc()
    # alphabetical: blue, dielectric, ground, metal
    assert [d.materials[i].type for i in range(4)] == [0, 2, 0, 1]
    assert [d.objects[i].material for i in range(4)] == [2, 0, 1, 3]
    m = pt.compose(pt.scale(0.5), pt.translate((-1.0, -0.5, -2.0)))
    assert np.allclose(m @ np.array([2, 2, 2, 1], dtype=np.float32), [0, 0.5, -1, 1])
    s.add_material("ground", pt.Material.lambertian((0, 0, 0)))      # duplicate keeps the first
    assert s.materials["ground"].albedo == (0.8, 0.8, 0.0)
    with pytest.raises(KeyError):
        s.add_sphere(1.0, pt.translate((0, 0, 0)), "missing")


def test_scene_json_first_mesh_only_by_default_and_all_meshes_extension(tmp_path, monkeypatch):
    """Two different OBJs named by one scene: the reference uploads only the alphabetically-first
    mesh and every mesh object instances it (scene_description.cpp:95) — kept as the default;
    PT_ALL_MESHES=1 (cuda_pt --all-meshes) gives every mesh object its own mesh."""
    a, b = pt.bunny_like(0), pt.bunny_like(1)
    (tmp_path / "models").mkdir(exist_ok=True)
    pt.write_obj(str(tmp_path / "models" / "a.obj"), a)
    pt.write_obj(str(tmp_path / "models" / "b.obj"), b)
    js = json.loads(json.dumps(BUNNY_JSON))
    js["surfaces"][1]["filename"] = "../models/b.obj"
    js["surfaces"][2]["filename"] = "../models/a.obj"
    path = _write_scene(tmp_path, js)
    monkeypatch.delenv("PT_ALL_MESHES", raising=False)
    rc, h, d, info = _read(path)
    assert rc == 0 and d.n_meshes == 0 and d.n_indices == a.indices.size       # a.obj sorts first
    pt.load_library().pt_scene_file_free(h)
    monkeypatch.setenv("PT_ALL_MESHES", "1")
    rc, h, d, info = _read(path)
    assert rc == 0 and d.n_meshes == 2
    assert [d.mesh_first_index[i] for i in range(3)] == [0, a.indices.size, a.indices.size + b.indices.size]
    assert d.n_indices == a.indices.size + b.indices.size
    assert [d.objects[i].prim_index for i in (1, 2)] == [1, 0]                 # b.obj, a.obj
    idx = np.ctypeslib.as_array(d.indices, shape=(d.n_indices,))
    assert idx[:a.indices.size].max() < a.positions.shape[0] <= idx[a.indices.size:].min()
    pt.load_library().pt_scene_file_free(h)


def test_multi_mesh_description_bakes_every_mesh():
    from cuda_path_tracer_b200.api import HostBVH
    sd = pt.SceneDescription()
    sd.add_material("m", pt.Material.lambertian((0.5, 0.5, 0.5)))
    sd.add_mesh("a", pt.bunny_like(1))
    sd.add_mesh("b", pt.heightfield(6))
    sd.add_mesh_object("b", pt.translate((0, -1, -3)), "m")
    sd.add_mesh_object("a", pt.translate((0, 0, -3)), "m")
    sd.add_mesh_object("a", pt.translate((2, 0, -3)), "m")
    first_only = HostBVH(sd)                      # reference behaviour: three instances of "a"
    assert int(first_only.info.n_world_triangles) == 3 * pt.bunny_like(1).triangle_count
    sd.all_meshes = True
    hb = HostBVH(sd)
    assert int(hb.info.n_world_triangles) == 2 * pt.bunny_like(1).triangle_count + pt.heightfield(6).triangle_count
    assert hb.violations() == 0
    # a mesh object naming a mesh that does not exist is an error, not a crash
    d, keep = sd.to_desc()
    d.objects[0].prim_index = 7
    h = C.c_void_p()
    assert pt.load_library().pt_host_bvh_build(C.byref(d), 1, C.byref(h), None) != 0
    assert b"mesh" in pt.load_library().pt_last_error()


def test_binary_mesh_cache_round_trip(tmp_path, monkeypatch):
    """PT_MESH_CACHE=1: the parsed OBJ is kept as <obj>.b200mesh and reused while the source is
    unchanged; a modified source invalidates it; without the switch nothing is written."""
    mesh = pt.bunny_like(2)
    path = _write_scene(tmp_path, BUNNY_JSON, mesh)
    obj = tmp_path / "models" / "bunny.obj"
    side = tmp_path / "models" / "bunny.obj.b200mesh"
    monkeypatch.delenv("PT_MESH_CACHE", raising=False)
    rc, h, d, _ = _read(path)
    assert rc == 0 and not side.exists()
    first = np.ctypeslib.as_array(d.positions, shape=(d.n_vertices * 3,)).copy()
    pt.load_library().pt_scene_file_free(h)
    monkeypatch.setenv("PT_MESH_CACHE", "1")
    rc, h, d, _ = _read(path)
    assert rc == 0 and side.exists()
    pt.load_library().pt_scene_file_free(h)
    rc, h, d, _ = _read(path)                                  # served from the cache
    assert rc == 0
    assert np.array_equal(np.ctypeslib.as_array(d.positions, shape=(d.n_vertices * 3,)), first)
    assert d.n_indices == mesh.indices.size
    pt.load_library().pt_scene_file_free(h)
    pt.write_obj(str(obj), pt.bunny_like(1))                   # the source changes: cache is stale
    rc, h, d, _ = _read(path)
    assert rc == 0 and d.n_indices == pt.bunny_like(1).indices.size
    pt.load_library().pt_scene_file_free(h)
    side.write_bytes(b"garbage")                               # a corrupt side-car is ignored
    rc, h, d, _ = _read(path)
    assert rc == 0 and d.n_indices == pt.bunny_like(1).indices.size
    pt.load_library().pt_scene_file_free(h)


def _read_obj_scene(tmp_path, obj_text, monkeypatch, serial):
    (tmp_path / "models").mkdir(exist_ok=True)
    (tmp_path / "models" / "bunny.obj").write_text(obj_text)
    if serial:
        monkeypatch.setenv("PT_OBJ_SERIAL", "1")
    else:
        monkeypatch.delenv("PT_OBJ_SERIAL", raising=False)
    rc, h, d, _ = _read(_write_scene(tmp_path, BUNNY_JSON))
    if rc != 0:
        return rc, None, None
    pos = np.ctypeslib.as_array(d.positions, shape=(d.n_vertices * 3,)).copy()
    idx = np.ctypeslib.as_array(d.indices, shape=(d.n_indices,)).copy()
    pt.load_library().pt_scene_file_free(h)
    return rc, pos, idx


def test_parallel_obj_parser_equals_the_serial_reader(tmp_path, monkeypatch):
    """Files above 1 MB go through the chunked three-sweep parser; it must reproduce the serial
    reader (the Assimp-semantics restatement) byte for byte on a file that mixes polygons,
    negative indices, v/vt/vn corners, comments, CRLF line ends and a group statement that cuts
    the first mesh off."""
    rng = np.random.default_rng(11)
    lines = ["# header", "mtllib nothing.mtl"]
    n_v = 0
    for block in range(1500):
        k = int(rng.integers(3, 9))
        vs = rng.normal(size=(k, 3))
        for v in vs:
            lines.append(f"v {v[0]:.7g} {v[1]:+.7g} {v[2]:.7g}" + ("\r" if block % 7 == 0 else ""))
            lines.append(f"vn 0 1 0")
            if block % 3 == 0:
                lines.append("vt 0.5 0.5")
        n_v += k
        style = block % 4
        if style == 0:      # absolute polygon
            lines.append("f " + " ".join(str(n_v - k + 1 + i) for i in range(k)))
        elif style == 1:    # negative indices
            lines.append("f " + " ".join(str(-k + i) for i in range(k)) + "  # relative")
        elif style == 2:    # v/vt/vn corners, triangle fan of earlier vertices too
            lines.append("f " + " ".join(f"{n_v - k + 1 + i}/1/1" for i in range(k)))
            lines.append(f"f 1//1 {n_v}//1 {max(2, n_v - 1)}//1")
        else:               # v//vn with tabs
            lines.append("f\t" + "\t".join(f"{n_v - k + 1 + i}//{i + 1}" for i in range(k)))
        lines.append("s off")
    body = "\n".join(lines) + "\n"
    filler = "".join(f"# padding line {i} ................................................\n" for i in range(22000))
    tail = "g second_mesh\nv 9 9 9\nv 8 8 8\nv 7 7 7\nf -1 -2 -3\n"
    text = filler[: len(filler) // 2] + body + filler[len(filler) // 2:] + tail
    assert len(text) > (1 << 20)
    rc_s, pos_s, idx_s = _read_obj_scene(tmp_path, text, monkeypatch, serial=True)
    rc_p, pos_p, idx_p = _read_obj_scene(tmp_path, text, monkeypatch, serial=False)
    assert rc_s == 0 and rc_p == 0
    assert idx_s.size == idx_p.size and idx_s.size % 3 == 0 and idx_s.size > 3 * 1500
    assert np.array_equal(idx_s, idx_p) and np.array_equal(pos_s, pos_p)
    assert not np.any(np.all(pos_p.reshape(-1, 3) == [9, 9, 9], axis=1))      # cut off at "g"
    # errors surface the same way
    bad = text.replace("f -1 -2 -3", "f -1 -2 -3").replace("# header", "# header\nv 0 0 0\nf 1 2 99999999")
    rc_s, _, _ = _read_obj_scene(tmp_path, bad, monkeypatch, serial=True)
    rc_p, _, _ = _read_obj_scene(tmp_path, bad, monkeypatch, serial=False)
    assert rc_s != 0 and rc_p != 0


def test_product_never_touches_the_oracle():
    """The oracle and the reference copy are test infrastructure: nothing under the package (or
    the public header) may import, link, dlopen or even name them."""
    pkg = os.path.join(ROOT, "cuda_path_tracer_b200")
    offenders = []
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".h")):
                text = open(os.path.join(base, f), errors="replace").read()
                for needle in ("liboracle", "oracle_lib", "oracle/", "libref_", "ref_lib", "orc_"):
                    if needle in text:
                        offenders.append((f, needle))
    assert not offenders, offenders
    hdr = open(os.path.join(ROOT, "include", "b200pt.h")).read()
    assert "oracle" not in hdr.lower()
    # and the shared library itself does not depend on either
    import subprocess
    ldd = subprocess.run(["ldd", pt.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in ldd and "libref" not in ldd


def test_scene_create_rejects_bad_descriptors_before_touching_the_device():
    """Argument validation comes first and reports through pt_last_error (the reference would
    read out of bounds or panic): it needs no device, so it is checked here."""
    lib = pt.load_library()
    sd = pt.bunny_scene(pt.bunny_like(0), 16, 16)

    def create(mutate):
        d, keep = sd.to_desc()
        undo = mutate(d)
        h = C.c_void_p()
        rc = lib.pt_scene_create(C.byref(d), 0, C.byref(h))
        msg = lib.pt_last_error()
        if undo:
            undo()
        assert not h.value
        return rc, msg

    def null_indices(d):
        d.indices = None
    def null_positions(d):
        d.positions = None
    def ragged(d):
        d.n_indices -= 1
    def bad_index(d):
        idx = np.ctypeslib.as_array(d.indices, shape=(d.n_indices,))   # the mesh's own array: restored below
        saved = int(idx[5])
        idx[5] = d.n_vertices
        return lambda: idx.__setitem__(5, saved)
    def bad_material_type(d):
        d.materials[0].type = 7
    def bad_object_material(d):
        d.objects[1].material = d.n_materials
    def no_materials(d):
        d.n_materials = 0
    def null_spheres(d):
        d.spheres = None
    def projective(d):
        d.objects[1].m[3] = 0.5

    for mutate, needle in [(null_indices, b"null"), (null_positions, b"null"), (ragged, b"multiple of 3"),
                           (bad_index, b"out of range"), (bad_material_type, b"material type"),
                           (bad_object_material, b"material index"), (no_materials, b"material"),
                           (null_spheres, b"spheres"), (projective, b"projective")]:
        rc, msg = create(mutate)
        assert rc == 1 and needle in msg, (mutate.__name__, rc, msg)
    assert lib.pt_scene_create(None, 0, None) == 1
