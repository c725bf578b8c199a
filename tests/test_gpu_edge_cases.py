"""GPU edge cases: ragged resolutions, degenerate scenes, object-order quirks, transformed
instances — against the CPU oracle and, where built, the reference's own CUDA kernels."""
import math

import numpy as np
import pytest

import cuda_path_tracer_b200 as pt
from cuda_path_tracer_b200 import DisplayBufferType as DB
from tests import ref_lib
from tests.test_gpu_parity import _check_hits, _rays_for, _secondary

pytestmark = pytest.mark.gpu


def _render(sd, spp, depth, method=pt.GPUMethod.megakernel):
    w, h = sd.resolution
    tr = pt.PathTracer(max_depth=depth)
    tr.current_gpu_method = method
    tr.max_iterations = spp
    tr.create_buffers((w, h), sd)
    tr.render(sd.camera, spp)
    return tr.download(DB.color), tr.download(DB.normal), tr.download(DB.depth), int(tr.stats().rays), tr


def _image_close(c, rc, frac=0.01):
    diff = np.abs(c - rc).max(axis=2)
    assert np.median(diff) < 1e-5
    assert (diff > 1e-3).mean() < frac, (diff > 1e-3).mean()


@pytest.mark.parametrize("w,h", [(33, 17), (8, 4), (7, 3), (2, 2), (257, 65)])
def test_ragged_resolutions(oracle, w, h):
    """Widths/heights that are not multiples of the 8x4 warp tile; tiny frames."""
    sd = pt.three_balls(w, h)
    rc, rn, rd, rrays = oracle.scene(sd).render(sd.camera, w, h, 2, 6)
    c, n, d, rays, _ = _render(sd, 2, 6)
    assert c.shape == (h, w, 3)
    _image_close(c, rc, frac=0.03)
    assert abs(rays - rrays) <= max(4, 0.01 * rrays)


def _mixed_scene(w=80, h=48):
    """Object order sphere, mesh, sphere, mesh with rotated / non-uniformly scaled instances,
    a scaled sphere and all three material types."""
    s = pt.SceneDescription()
    s.add_material("a_ground", pt.Material.lambertian((0.7, 0.7, 0.6)))
    s.add_material("b_metal", pt.Material.metal((0.8, 0.6, 0.2), 0.3))
    s.add_material("c_glass", pt.Material.dielectric(1.5))
    s.add_material("d_red", pt.Material.lambertian((0.7, 0.2, 0.2)))
    s.add_mesh("blob", pt.bunny_like(2))
    s.add_sphere(100.0, pt.translate((0.0, -100.5, -1.0)), "a_ground")
    s.add_mesh_object("blob", pt.compose(pt.rotate(40, (0, 1, 0)), pt.scale((0.8, 1.2, 0.8)),
                                         pt.translate((0.9, -0.5, -2.2))), "b_metal")
    s.add_sphere(0.5, pt.compose(pt.scale(0.8), pt.translate((-0.2, -0.1, -1.4))), "c_glass")
    s.add_mesh_object("blob", pt.compose(pt.scale(0.6), pt.rotate(-30, (1, 0, 0)),
                                         pt.translate((-1.1, -0.4, -2.0))), "d_red")
    s.camera = pt.Camera((0.0, 0.2, 0.5), (1.0, 0.0, 0.0, 0.0), math.radians(70.0))
    s.resolution = (w, h)
    return s


def test_mixed_scene_hits_match_oracle(oracle):
    sd = _mixed_scene()
    w, h = sd.resolution
    osc = oracle.scene(sd)
    scene = pt.Scene.from_description(sd)
    prim, rng = _rays_for(oracle, sd, w, h, n_random=8000)
    ref = osc.trace_batch(prim, 0)
    ours = scene.trace_batch(prim)
    _check_hits(ours, ref, allow_frac=5e-4)
    sec = _secondary(prim, ref, rng)
    # secondary rays with non-unit directions (what a fuzzy metal bounce produces)
    sec[::3, 4:7] *= 1.3
    _check_hits(scene.trace_batch(sec), osc.trace_batch(sec, 0), allow_frac=2e-3)
    assert set(np.unique(ref["object"][ref["t"] > 0])) >= {0, 1, 2, 3}


def test_mixed_scene_image_matches_oracle(oracle):
    sd = _mixed_scene()
    w, h = sd.resolution
    rc, rn, rd, rrays = oracle.scene(sd).render(sd.camera, w, h, 3, 10)
    c, n, d, rays, _ = _render(sd, 3, 10)
    _image_close(c, rc, frac=0.03)
    assert abs(rays - rrays) <= 0.01 * rrays


@pytest.mark.skipif(not ref_lib.have_ref_cuda(), reason="oracle/_ref/libref_cuda.so not built")
def test_mixed_scene_matches_reference_cuda(oracle):
    """Same scene against the reference's kernels: hits, megakernel image, streaming image."""
    sd = _mixed_scene()
    w, h = sd.resolution
    ref = ref_lib.load_ref_cuda()
    scene = pt.Scene.from_description(sd)
    rt = ref.tracer(sd, w, h, 10, megakernel=True)
    prim, rng = _rays_for(oracle, sd, w, h, n_random=8000)
    theirs, ours = rt.trace_batch(prim), scene.trace_batch(prim)
    miss = (ours["t"] < 0) != (theirs["t"] < 0)
    both = (ours["t"] > 0) & (theirs["t"] > 0)
    rel = np.abs(ours["t"][both] - theirs["t"][both]) / np.abs(theirs["t"][both])
    assert miss.sum() + (rel > 1e-5).sum() <= 4
    rt.render_timed(sd.camera, 3, 10)
    c = _render(sd, 3, 10)[0]
    _image_close(c, rt.download(1), frac=0.03)
    rs = ref.tracer(sd, w, h, 10, megakernel=False)
    _, rrays = rs.render_timed(sd.camera, 2, 10)
    cs, _, _, rays, _ = _render(sd, 2, 10, method=pt.GPUMethod.streaming)
    assert np.median(np.abs(cs - rs.download(1)).max(axis=2)) < 1e-5
    assert abs(rays - rrays) <= 0.01 * rrays


def test_single_triangle_and_mesh_only_scenes(oracle):
    """Root-is-leaf BVH (1 and 3 triangles) and a scene without any sphere."""
    for n_tri in (1, 3):
        rng = np.random.default_rng(n_tri)
        pos = rng.uniform(-1, 1, size=(3 * n_tri, 3)).astype(np.float32)
        pos[:, 2] -= 3.0
        s = pt.SceneDescription()
        s.add_material("m", pt.Material.lambertian((0.5, 0.6, 0.7)))
        s.add_mesh("t", pt.Mesh(pos, np.arange(3 * n_tri, dtype=np.uint32)))
        s.add_mesh_object("t", pt.translate((0, 0, 0)), "m")
        s.camera = pt.Camera((0, 0, 0), (1, 0, 0, 0), math.radians(60))
        s.resolution = (64, 48)
        rc, _, rd, rrays = oracle.scene(s).render(s.camera, 64, 48, 2, 4)
        c, _, d, rays, _ = _render(s, 2, 4)
        _image_close(c, rc, frac=0.02)
        assert (rd < 1e5).sum() > 0 and abs(rays - rrays) <= max(2, 0.01 * rrays)


def test_depth_one_and_sky_only(oracle):
    sd = pt.three_balls(48, 32)
    rc = oracle.scene(sd).render(sd.camera, 48, 32, 2, 1)[0]
    c, _, _, rays, _ = _render(sd, 2, 1)
    _image_close(c, rc)
    assert rays == 48 * 32 * 2
    # a scene whose only object is behind the camera: every path leaves after one ray
    s = pt.SceneDescription()
    s.add_material("m", pt.Material.lambertian((0.5, 0.5, 0.5)))
    s.add_sphere(0.5, pt.translate((0, 0, 5)), "m")
    s.camera = pt.Camera()
    s.resolution = (32, 16)
    c, n, d, rays, _ = _render(s, 1, 50)
    assert rays == 32 * 16 and np.all(d == 1e6)
    assert np.allclose(c, oracle.scene(s).render(s.camera, 32, 16, 1, 50)[0], atol=1e-6)


def test_denoiser_on_ragged_frame(oracle):
    sd = pt.three_balls(75, 41)
    w, h = sd.resolution
    rc, rn, rd, _ = oracle.scene(sd).render(sd.camera, w, h, 1, 5)
    tr = pt.PathTracer(max_depth=5)
    tr.create_buffers((w, h), sd)
    tr.upload_frame(rc, rn, rd, sd.camera)
    # filter 31 taints all 41 rows in reference mode (2*(1+2+4+8+16) = 62), so it is also
    # checked with the sane clamp, where every pixel is defined
    for fs, fix in ((3, False), (31, True), (31, False)):
        tr.atrous_denoiser.filter_size = fs
        tr.atrous_denoiser.clamp_fix = fix
        tr.denoise()
        ours = tr.download(DB.denoised)
        ref, taint = oracle.denoise(sd.camera, rc, rn, rd, fs, clamp_fix=fix)
        assert np.abs(ours - ref).max(axis=2)[~taint].max(initial=0.0) <= 1e-4
        assert np.isfinite(ours).all()
        if fix:
            assert not taint.any()


@pytest.mark.gpu
@pytest.mark.parametrize("scene_name", ["bunny", "terrain", "single"])
def test_wide_tree_traversal_matches_binary_and_oracle(oracle, scene_name, monkeypatch):
    """PT_BVH=8 (compressed 8-wide tree, shared-memory staged top levels, traverse8_kernel): the
    same closest hits and the same image as the default binary tree."""
    if scene_name == "bunny":
        sd = pt.bunny_scene(pt.bunny_like(4), 96, 54)
    elif scene_name == "terrain":
        sd = pt.terrain_scene(40, 64, 36)
    else:
        sd = pt.SceneDescription()
        sd.add_material("m", pt.Material.lambertian((0.5, 0.5, 0.5)))
        sd.add_mesh("tri", pt.Mesh(np.array([[-1, -1, -3], [1, -1, -3], [0, 1, -3]], np.float32),
                                   np.array([0, 1, 2], np.uint32)))
        sd.add_mesh_object("tri", pt.translate((0, 0, 0)), "m")
        sd.resolution = (48, 32)
    w, h = sd.resolution
    rays, _ = _rays_for(oracle, sd, w, h, n_random=3000, seed=5)   # jittered: no edge-on alignments
    binary = pt.Scene.from_description(sd)
    assert int(binary.info.n_bvh8_nodes) == 0
    hb = binary.trace_batch(rays)
    monkeypatch.setenv("PT_BVH", "8")
    monkeypatch.setenv("PT_T8", "256,2,32")
    wide = pt.Scene.from_description(sd)
    assert int(wide.info.n_bvh8_nodes) >= 1
    hw = wide.trace_batch(rays)
    ref = oracle.scene(sd).trace_batch(rays, 0)
    for got in (hb, hw):
        _check_hits(got, ref, allow_frac=2e-3)
    # the two trees index one triangle array in different orders: compare what the hit means
    same = (hb["t"] > 0) & (hw["t"] > 0)
    assert np.array_equal(hb["t"] > 0, hw["t"] > 0)
    assert np.array_equal(hb["t"][same], hw["t"][same])          # bit-equal t: same triangle arithmetic
    assert (hb["prim"][same] != hw["prim"][same]).sum() <= max(1, same.sum() // 2000)   # exact ties only

    def image(scene):
        tr = pt.PathTracer(max_depth=5)
        tr.max_iterations = 2
        tr.create_buffers((w, h), scene)
        tr.render(sd.camera, 2)
        tr.synchronize()
        return tr.download(DB.color), int(tr.stats().rays)

    ib, rb = image(binary)
    iw, rw = image(wide)
    assert abs(rb - rw) <= max(2, rb // 2000)
    d = np.abs(ib - iw).max(axis=2)
    assert (d > 1e-5).mean() < 2e-3, (d > 1e-5).mean()


def test_multi_mesh_scene_matches_oracle_on_the_merged_mesh(oracle):
    """Several meshes (extension beyond the reference, which uploads one): the hits equal the
    oracle's on the equivalent single-mesh scene whose mesh is the union of the transformed parts."""
    a, b = pt.bunny_like(3), pt.heightfield(12)
    xa, xb = pt.compose(pt.scale(1.6), pt.translate((0.4, 0.1, -2.5))), pt.compose(pt.scale(2.0), pt.translate((0.0, -0.9, -2.5)))
    sd = pt.SceneDescription()
    sd.add_material("m", pt.Material.lambertian((0.5, 0.5, 0.5)))
    sd.add_mesh("a", a)
    sd.add_mesh("b", b)
    sd.add_mesh_object("a", xa, "m")
    sd.add_mesh_object("b", xb, "m")
    sd.resolution = (80, 60)
    sd.all_meshes = True

    def world(mesh, xf):
        p = np.concatenate([mesh.positions.astype(np.float32), np.ones((mesh.positions.shape[0], 1), np.float32)], 1)
        return (p @ np.asarray(xf, np.float32).T)[:, :3].astype(np.float32)

    merged = pt.Mesh(np.concatenate([world(a, xa), world(b, xb)]),
                     np.concatenate([a.indices, b.indices + np.uint32(a.positions.shape[0])]).astype(np.uint32))
    one = pt.SceneDescription()
    one.add_material("m", pt.Material.lambertian((0.5, 0.5, 0.5)))
    one.add_mesh("merged", merged)
    one.add_mesh_object("merged", pt.translate((0, 0, 0)), "m")
    one.resolution, one.camera = sd.resolution, sd.camera
    rays, rng = _rays_for(oracle, sd, 80, 60, n_random=3000, seed=3)
    ours = pt.Scene.from_description(sd).trace_batch(rays)
    ref = oracle.scene(one).trace_batch(rays, 0)
    assert (ref["t"] > 0).mean() > 0.1
    m = (ours["t"] > 0) & (ref["t"] > 0)
    assert ((ours["t"] > 0) != (ref["t"] > 0)).sum() <= 2
    rel = np.abs(ours["t"][m] - ref["t"][m]) / np.maximum(ref["t"][m], 1e-6)
    assert (rel > 1e-5).sum() <= 2
    # primitive ids are positions in the shared index buffer: identical to the merged mesh's
    agree = m & (rel.max(initial=0) <= 1e-5)
    assert (ours["prim"][agree] != ref["prim"][agree]).sum() <= max(2, agree.sum() // 500)
    # the two objects are told apart
    assert set(np.unique(ours["object"][m])) == {0, 1}


def test_checkpoint_and_resume_continue_the_same_render(tmp_path):
    sd = pt.bunny_scene(pt.bunny_like(3), 96, 54)
    w, h = sd.resolution
    scene = pt.Scene.from_description(sd)

    def tracer():
        tr = pt.PathTracer(max_depth=6)
        tr.max_iterations = 8
        tr.create_buffers((w, h), scene)
        return tr

    full = tracer()
    full.render(sd.camera, 8)
    ref = full.download(DB.color)
    part = tracer()
    part.render(sd.camera, 3)
    state = str(tmp_path / "state.b200pt")
    part.save_state(state)
    cont = tracer()
    cont.load_state(state)
    assert cont.iteration() == 3
    cont.render(sd.camera, 8)                     # clipped at max_iterations: renders 5 more
    assert cont.iteration() == 8
    got = cont.download(DB.color)
    assert np.allclose(got, ref, rtol=2e-6, atol=1e-6), np.abs(got - ref).max()
    other = pt.PathTracer(max_depth=6)
    other.create_buffers((w // 2, h), scene)
    with pytest.raises(pt.PTError):
        other.load_state(state)                   # resolution mismatch is an error
    with pytest.raises(pt.PTError):
        other.load_state(str(tmp_path / "missing"))


@pytest.mark.parametrize("world", [2, 3])
def test_row_band_contexts_reproduce_the_full_frame(world):
    """Row-band sharding on one device: `world` contexts each render + denoise their band of a
    1-spp frame; after the halo rows are copied between their sums buffers the assembled image
    equals the single-context frame (bit-equal radiance, denoised colour within 1e-6)."""
    import torch
    from cuda_path_tracer_b200 import sharding
    sd = pt.bunny_scene(pt.bunny_like(3), 200, 150)
    w, h = sd.resolution
    scene = pt.Scene.from_description(sd)

    def tracer():
        tr = pt.PathTracer(max_depth=6)
        tr.max_iterations = 1 << 20
        tr.create_buffers((w, h), scene)
        tr.atrous_denoiser.filter_size = 16
        return tr

    full = tracer()
    full.render(sd.camera, 2)
    full.denoise()
    ref_color, ref_final = full.download(DB.color), full.download(DB.denoised)
    ref_rays = int(full.stats().rays)

    trs, sums = [], []
    for r in range(world):
        tr = tracer()
        buf = torch.zeros(2, h, w * 4, dtype=torch.float32, device="cuda")
        tr.bind_sums(buf.data_ptr())
        tr.set_rows(*sharding.band_rows(r, world, h))
        tr.render(sd.camera, 2)
        tr.synchronize()
        trs.append(tr)
        sums.append(buf)
    assert sum(int(t.stats().rays) for t in trs) == ref_rays
    halo = trs[0].halo_rows()
    assert halo == 62
    for r in range(world):                                   # the halo exchange, done by hand
        _, recvs = sharding.halo_plan(r, world, h, halo)
        for peer, r0, r1 in recvs:
            sums[r][:, r0:r1] = sums[peer][:, r0:r1]
    torch.cuda.synchronize()
    color = np.zeros_like(ref_color)
    final = np.zeros_like(ref_final)
    for r, tr in enumerate(trs):
        b0, b1 = sharding.band_rows(r, world, h)
        tr.denoise()
        color[b0:b1] = tr.download(DB.color)[b0:b1]
        final[b0:b1] = tr.download(DB.denoised)[b0:b1]
    assert np.array_equal(color, ref_color)
    assert np.abs(final - ref_final).max() <= 1e-6, np.abs(final - ref_final).max()
    with pytest.raises(pt.PTError):
        trs[0].set_rows(2, 40)                               # not a multiple of 4
    with pytest.raises(pt.PTError):
        trs[0].set_rows(8, h + 4)


@pytest.mark.parametrize("scene_name", ["bunny", "terrain"])
def test_device_lbvh_build_matches_host_restatement_and_oracle(oracle, scene_name, monkeypatch):
    """PT_BUILD=lbvh: the tree built by lbvh.cu equals the host restatement node for node
    (same lbvh.h logic, stable radix sort == stable_sort), and its closest hits / image equal the
    SAH-built scene's and the oracle's."""
    from cuda_path_tracer_b200.api import HostBVH
    sd = pt.bunny_scene(pt.bunny_like(4), 96, 54) if scene_name == "bunny" else pt.terrain_scene(60, 64, 36)
    w, h = sd.resolution
    sah = pt.Scene.from_description(sd)
    monkeypatch.setenv("PT_BUILD", "lbvh")
    dev = pt.Scene.from_description(sd)
    assert int(dev.info.device_build) == 1 and int(sah.info.device_build) == 0
    host = HostBVH(sd, lbvh=True)
    h_nodes, _, h_tris = host.arrays()
    d_nodes, d_tris = dev.copy_bvh()
    assert d_nodes.shape == h_nodes.shape and d_tris.shape == h_tris.shape
    assert np.array_equal(d_tris.view(np.uint32), h_tris.view(np.uint32))
    assert np.array_equal(d_nodes.view(np.uint32), h_nodes.view(np.uint32))
    rays, rng = _rays_for(oracle, sd, w, h, n_random=3000, seed=9)
    ref = oracle.scene(sd).trace_batch(rays, 0)
    got, base = dev.trace_batch(rays), sah.trace_batch(rays)
    _check_hits(got, ref, allow_frac=2e-3)
    same = (got["t"] > 0) & (base["t"] > 0)
    assert np.array_equal(got["t"] > 0, base["t"] > 0)
    assert (got["t"][same] != base["t"][same]).sum() <= max(1, same.sum() // 2000)     # exact ties only

    def image(scene):
        tr = pt.PathTracer(max_depth=5)
        tr.max_iterations = 2
        tr.create_buffers((w, h), scene)
        tr.render(sd.camera, 2)
        tr.synchronize()
        return tr.download(DB.color)

    d = np.abs(image(dev) - image(sah)).max(axis=2)
    assert (d > 1e-5).mean() < 2e-3, (d > 1e-5).mean()


def test_ray_binning_changes_the_order_not_the_image():
    """pt_params.sort_rays (the reference's commented-out sort by material_id, path_tracer.cu:
    439-446): traversed rays are shaded bin by bin (miss / lambertian / metal / dielectric hit).
    Every path carries its own RNG stream, so the frame and the ray count are bit-identical."""
    sd = pt.many_materials_scene(320, 180, 4, subdiv=3)
    w, h = sd.resolution
    frames = []
    for sort in (False, True):
        tr = pt.PathTracer(max_depth=8, sort_rays=sort)
        tr.max_iterations = 4
        tr.create_buffers((w, h), sd)
        tr.render(sd.camera, 4)
        tr.synchronize()
        frames.append((tr.download(DB.color), tr.download(DB.normal), int(tr.stats().rays)))
    assert frames[0][2] == frames[1][2]
    assert np.array_equal(frames[0][0], frames[1][0]) and np.array_equal(frames[0][1], frames[1][1])
    assert frames[0][0].std() > 0.05                          # a real image, three material types


def test_progressive_state_records_what_it_holds(tmp_path):
    """a state file names the iterations its sums hold and what they were rendered from.
    A shard that rendered [10, 14) resumes at 14 (never re-using seeds); a different scene,
    max_depth or camera is refused instead of being blended in silently."""
    sd = pt.bunny_scene(pt.bunny_like(2), 96, 54)
    w, h = sd.resolution
    path = str(tmp_path / "shard.state")
    a = pt.PathTracer(max_depth=6)
    a.create_buffers((w, h), sd)
    a.render_range(sd.camera, 10, 4)
    a.save_state(path)
    b = pt.PathTracer(max_depth=6)
    b.max_iterations = 1 << 20
    b.create_buffers((w, h), sd)
    b.load_state(path)
    assert b.iteration() == 4
    b.render(sd.camera, 2)                                   # continues with iterations 14, 15
    whole = pt.PathTracer(max_depth=6)
    whole.create_buffers((w, h), sd)
    whole.render_range(sd.camera, 10, 6)
    assert np.abs(b.download(DB.color) - whole.download(DB.color)).max() < 1e-6
    # another camera would blend two images: refused until restart
    cam2 = pt.Camera((0.0, 0.2, 0.0), (1.0, 0.0, 0.0, 0.0), sd.camera.vfov)
    with pytest.raises(pt.PTError, match="camera"):
        b.render(cam2, 1)
    b.restart()
    b.render(cam2, 1)
    # another scene / another depth
    other = pt.PathTracer(max_depth=6)
    other.create_buffers((w, h), pt.bunny_scene(pt.bunny_like(3), 96, 54))
    with pytest.raises(pt.PTError, match="different scene"):
        other.load_state(path)
    deeper = pt.PathTracer(max_depth=7)
    deeper.create_buffers((w, h), sd)
    with pytest.raises(pt.PTError, match="max_depth"):
        deeper.load_state(path)
    # a merged, non-contiguous set of iterations is not continued by pt_render
    c = pt.PathTracer(max_depth=6)
    c.max_iterations = 1 << 20
    c.create_buffers((w, h), sd)
    c.render_range(sd.camera, 0, 2)
    c.render_range(sd.camera, 7, 2)
    with pytest.raises(pt.PTError, match="non-contiguous"):
        c.render(sd.camera, 1)


def test_failed_resize_leaves_a_context_that_refuses_work():
    """a refused size changes nothing; nothing is ever launched on freed buffers."""
    sd = pt.three_balls(64, 64)
    tr = pt.PathTracer(max_depth=4)
    tr.create_buffers((64, 64), sd)
    tr.render(sd.camera, 1)
    before = tr.download(DB.color)
    with pytest.raises(pt.PTError, match="resolution"):
        tr.resize_image((0, 64))
    with pytest.raises(pt.PTError, match="resolution"):
        tr.resize_image((1 << 14, 1 << 14))
    assert np.array_equal(tr.download(DB.color), before)        # untouched
    tr.resize_image((32, 48))
    tr.max_iterations = 4
    tr.render(sd.camera, 1)
    assert tr.download(DB.color).shape == (48, 32, 3)


def test_sphere_group_trees_match_the_linear_scan(oracle):
    """Scenes with many spheres (the reference scans its objects linearly per ray):
    groups of more than 32 rigidly placed spheres are walked through a small tree.  Closest hits
    and the image equal the oracle's linear object loop; ties aside, the order of the tests does
    not matter for rigid spheres."""
    sd = pt.many_spheres_scene(300, 160, 120)
    w, h = sd.resolution
    scene = pt.Scene.from_description(sd)
    osc = oracle.scene(sd)
    prim, rng = _rays_for(oracle, sd, w, h, n_random=6000, seed=3)
    ref = osc.trace_batch(prim, 0)
    ours = scene.trace_batch(prim)
    assert (ref["prim"][ref["t"] > 0] < 0).mean() > 0.5          # most hits are spheres
    _check_hits(ours, ref, allow_frac=5e-4)
    sec = _secondary(prim, ref, rng)
    _check_hits(scene.trace_batch(sec), osc.trace_batch(sec, 0), allow_frac=1e-3)
    tr = pt.PathTracer(max_depth=6)
    tr.max_iterations = 2
    tr.create_buffers((w, h), scene)
    tr.render(sd.camera, 2)
    c = tr.download(DB.color)
    rc, _, _, rrays = osc.render(sd.camera, w, h, 2, 6)
    diff = np.abs(c - rc).max(axis=2)
    assert np.median(diff) < 1e-5 and (diff > 1e-3).mean() < 0.02, (np.median(diff), (diff > 1e-3).mean())
    assert abs(int(tr.stats().rays) - rrays) <= 0.005 * rrays
    # a sphere-only scene (no mesh at all) takes the same path
    so = pt.many_spheres_scene(200, 96, 96, with_mesh=False)
    s2, o2 = pt.Scene.from_description(so), oracle.scene(so)
    p2, _ = _rays_for(oracle, so, 96, 96, n_random=3000, seed=4)
    _check_hits(s2.trace_batch(p2), o2.trace_batch(p2, 0), allow_frac=5e-4)
