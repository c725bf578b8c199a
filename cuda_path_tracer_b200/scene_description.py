"""Host-side mirror of the reference's SceneDescription (src/lib/scene_description.hpp:29-49,
scene_description.cpp:119-154) and Camera (src/lib/camera.hpp:17-23), plus procedural meshes
standing in for the reference's git-LFS OBJ models (assets/models/*.obj are pointer stubs).

Pure numpy; no rendering happens here.  `SceneDescription.to_desc()` yields the flat arrays the
C ABI (`pt_scene_create`) consumes — the same arrays `build_scene()` uploads in the reference.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from dataclasses import dataclass, field

import numpy as np

from . import _abi


# ----------------------------------------------------------------------------- transforms
def translate(v) -> np.ndarray:
    m = np.eye(4, dtype=np.float32)
    m[:3, 3] = np.asarray(v, dtype=np.float32)
    return m


def scale(s) -> np.ndarray:
    s = np.broadcast_to(np.asarray(s, dtype=np.float32), (3,))
    return np.diag(np.concatenate([s, [1.0]]).astype(np.float32))


def rotate(deg: float, axis) -> np.ndarray:
    """glm::rotate(radians(deg), axis) (json_parser.cpp:52-55)."""
    a = np.float32(deg) * np.float32(0.01745329251994329576923690768489)
    c, s = np.float32(math.cos(a)), np.float32(math.sin(a))
    ax = np.asarray(axis, dtype=np.float32)
    ax = ax / np.float32(np.sqrt(np.dot(ax, ax)))
    t = (np.float32(1) - c) * ax
    r = np.eye(4, dtype=np.float32)
    # columns of glm's Rotate matrix (row index = second subscript)
    r[0, 0] = c + t[0] * ax[0]
    r[1, 0] = t[0] * ax[1] + s * ax[2]
    r[2, 0] = t[0] * ax[2] - s * ax[1]
    r[0, 1] = t[1] * ax[0] - s * ax[2]
    r[1, 1] = c + t[1] * ax[1]
    r[2, 1] = t[1] * ax[2] + s * ax[0]
    r[0, 2] = t[2] * ax[0] + s * ax[1]
    r[1, 2] = t[2] * ax[1] - s * ax[0]
    r[2, 2] = c + t[2] * ax[2]
    return r


def compose(*commands) -> np.ndarray:
    """Transform command list, applied left to right: M = M_i @ M (json_parser.cpp:85-88)."""
    m = np.eye(4, dtype=np.float32)
    for c in commands:
        m = (np.asarray(c, dtype=np.float32) @ m).astype(np.float32)
    return m


@dataclass
class Camera:
    """Camera (camera.hpp:17-23): position, rotation quaternion (w,x,y,z), vfov in radians."""
    position: tuple = (0.0, 0.0, 0.0)
    rotation: tuple = (1.0, 0.0, 0.0, 0.0)
    vfov: float = math.pi / 2.0

    def to_c(self) -> _abi.pt_camera:
        c = _abi.pt_camera()
        c.position[:] = [float(x) for x in self.position]
        c.rotation[:] = [float(x) for x in self.rotation]
        c.vfov = float(self.vfov)
        return c

    @staticmethod
    def from_c(c: _abi.pt_camera) -> "Camera":
        return Camera(tuple(c.position), tuple(c.rotation), float(c.vfov))

    @staticmethod
    def look_at(frm, at, up, vfov_deg: float) -> "Camera":
        """The {from, at, up} transform command (json_parser.cpp:56-71) as a camera pose."""
        frm, at, up = (np.asarray(v, dtype=np.float64) for v in (frm, at, up))
        d = frm - at
        d /= np.linalg.norm(d)
        left = np.cross(up, d)
        left /= np.linalg.norm(left)
        nup = np.cross(d, left)
        r = np.stack([left, nup, d], axis=1)  # columns
        tr = np.trace(r)
        if tr > 0:
            s = math.sqrt(tr + 1.0) * 2
            q = (0.25 * s, (r[2, 1] - r[1, 2]) / s, (r[0, 2] - r[2, 0]) / s, (r[1, 0] - r[0, 1]) / s)
        else:
            i = int(np.argmax(np.diag(r)))
            j, k = (i + 1) % 3, (i + 2) % 3
            s = math.sqrt(r[i, i] - r[j, j] - r[k, k] + 1.0) * 2
            v = [0.0, 0.0, 0.0]
            v[i] = 0.25 * s
            v[j] = (r[j, i] + r[i, j]) / s
            v[k] = (r[k, i] + r[i, k]) / s
            q = ((r[k, j] - r[j, k]) / s, v[0], v[1], v[2])
        return Camera(tuple(frm), q, math.radians(vfov_deg))


@dataclass
class Material:
    type: int
    albedo: tuple = (0.0, 0.0, 0.0)
    fuzz: float = 0.0
    refraction_index: float = 1.0

    @staticmethod
    def lambertian(albedo):
        return Material(_abi.MAT_DIFFUSE, tuple(albedo))

    @staticmethod
    def metal(albedo, fuzz):
        return Material(_abi.MAT_METAL, tuple(albedo), float(fuzz))

    @staticmethod
    def dielectric(ior):
        return Material(_abi.MAT_DIELECTRIC, (0.0, 0.0, 0.0), 0.0, float(ior))


@dataclass
class Mesh:
    """Mesh (mesh.hpp:9-18): positions [V,3] float32, indices [3T] uint32."""
    positions: np.ndarray
    indices: np.ndarray

    @property
    def triangle_count(self) -> int:
        return int(self.indices.size // 3)


@dataclass
class _Object:
    type: int
    prim_index: int
    material: str
    m: np.ndarray
    mesh_name: str = ""


class SceneDescription:
    """add_material / add_mesh / add_object with the reference's semantics:
    the GPU material table is ordered alphabetically by name and duplicate names keep
    the first definition; only the alphabetically-first mesh is uploaded and every mesh
    object instances it (scene_description.cpp:59-66, 95, 151-154)."""

    def __init__(self):
        self.materials: dict[str, Material] = {}
        self.meshes: dict[str, Mesh] = {}
        self.objects: list[_Object] = []
        self.spheres: list[tuple] = []
        self.camera = Camera()
        self.resolution = (0, 0)
        self.spp = 1
        self.filename = ""
        # extension: True = every mesh is uploaded and each mesh object instances its own
        # (the reference, and the default here, uploads only the alphabetically-first mesh)
        self.all_meshes = False

    def add_material(self, name: str, material: Material):
        self.materials.setdefault(name, material)

    def add_mesh(self, name: str, mesh: Mesh) -> str:
        if name in self.meshes:
            raise ValueError("Cannot add the same mesh twice!")
        self.meshes[name] = mesh
        return name

    def add_sphere(self, radius: float, transform: np.ndarray, material: str, center=(0.0, 0.0, 0.0)):
        if material not in self.materials:
            raise KeyError(f"Cannot find material {material}")
        self.objects.append(_Object(_abi.OBJ_SPHERE, len(self.spheres), material,
                                    np.asarray(transform, dtype=np.float32)))
        self.spheres.append((tuple(center), float(radius)))

    def add_mesh_object(self, mesh_name: str, transform: np.ndarray, material: str):
        if material not in self.materials:
            raise KeyError(f"Cannot find material {material}")
        if mesh_name not in self.meshes:
            raise KeyError(f"Cannot find mesh {mesh_name}")
        o = _Object(_abi.OBJ_MESH, 0, material, np.asarray(transform, dtype=np.float32))
        o.mesh_name = mesh_name
        self.objects.append(o)

    # ---- flat arrays for the C ABI -------------------------------------------------
    def to_desc(self):
        """Returns (pt_scene_desc, keepalive) — keepalive owns the numpy/ctypes buffers."""
        names = sorted(self.materials)
        mat_index = {n: i for i, n in enumerate(names)}
        mats = (_abi.pt_material * max(1, len(names)))()
        for i, n in enumerate(names):
            m = self.materials[n]
            mats[i].type = m.type
            mats[i].albedo[:] = [float(x) for x in m.albedo]
            mats[i].fuzz = m.fuzz
            mats[i].refraction_index = m.refraction_index
        objs = (_abi.pt_object * max(1, len(self.objects)))()
        for i, o in enumerate(self.objects):
            objs[i].type = o.type
            objs[i].prim_index = o.prim_index
            objs[i].material = mat_index[o.material]
            m = np.ascontiguousarray(o.m, dtype=np.float32)
            inv = np.linalg.inv(m.astype(np.float64)).astype(np.float32)
            if np.allclose(m[3], [0, 0, 0, 1]):
                inv[3] = [0, 0, 0, 1]
            objs[i].m[:] = m.T.reshape(-1).tolist()      # column-major
            objs[i].inv[:] = inv.T.reshape(-1).tolist()
        sph = (_abi.pt_sphere * max(1, len(self.spheres)))()
        for i, (c, r) in enumerate(self.spheres):
            sph[i].center[:] = [float(x) for x in c]
            sph[i].radius = r
        first = None
        if self.meshes and self.all_meshes:
            names_m = sorted(self.meshes)
            rank = {n: i for i, n in enumerate(names_m)}
            ps, ix, first, base = [], [], [0], 0
            for n in names_m:
                mp = np.ascontiguousarray(self.meshes[n].positions, dtype=np.float32).reshape(-1, 3)
                mi = np.ascontiguousarray(self.meshes[n].indices, dtype=np.uint32).reshape(-1)
                ps.append(mp)
                ix.append(mi + np.uint32(base))
                base += mp.shape[0]
                first.append(first[-1] + mi.size)
            pos, idx = np.concatenate(ps), np.concatenate(ix).astype(np.uint32)
            first = np.array(first, dtype=np.uint64)
            for i, o in enumerate(self.objects):
                if o.type == _abi.OBJ_MESH:
                    objs[i].prim_index = rank[o.mesh_name]
        elif self.meshes:
            mesh = self.meshes[sorted(self.meshes)[0]]
            pos = np.ascontiguousarray(mesh.positions, dtype=np.float32).reshape(-1, 3)
            idx = np.ascontiguousarray(mesh.indices, dtype=np.uint32).reshape(-1)
        else:
            pos = np.zeros((0, 3), dtype=np.float32)
            idx = np.zeros((0,), dtype=np.uint32)
        d = _abi.pt_scene_desc()
        d.positions = pos.ctypes.data_as(C.POINTER(C.c_float))
        d.n_vertices = pos.shape[0]
        d.indices = idx.ctypes.data_as(C.POINTER(C.c_uint32))
        d.n_indices = idx.size
        d.objects = objs
        d.n_objects = len(self.objects)
        d.spheres = sph
        d.n_spheres = len(self.spheres)
        d.materials = mats
        d.n_materials = len(names)
        if first is not None:
            d.n_meshes = len(first) - 1
            d.mesh_first_index = first.ctypes.data_as(C.POINTER(C.c_uint64))
        return d, (pos, idx, objs, sph, mats, first)


# ----------------------------------------------------------------------------- procedural meshes
def icosphere(subdivisions: int) -> tuple[np.ndarray, np.ndarray]:
    """Unit icosphere: returns (vertices [V,3] float64, faces [T,3] int64). T = 20 * 4^s."""
    t = (1.0 + math.sqrt(5.0)) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t],
                  [0, -1, -t], [0, 1, -t], [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4],
                  [11, 10, 2], [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8],
                  [3, 8, 9], [4, 9, 5], [2, 4, 11], [6, 2, 10], [8, 6, 7], [9, 8, 1]], dtype=np.int64)
    for _ in range(subdivisions):
        edges = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], axis=0)
        edges.sort(axis=1)
        uniq, inv = np.unique(edges, axis=0, return_inverse=True)
        mid = v[uniq[:, 0]] + v[uniq[:, 1]]
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        base = v.shape[0]
        v = np.concatenate([v, mid], axis=0)
        n = f.shape[0]
        inv = np.asarray(inv).reshape(-1)
        m01, m12, m20 = base + inv[:n], base + inv[n:2 * n], base + inv[2 * n:]
        f = np.concatenate([
            np.stack([f[:, 0], m01, m20], axis=1), np.stack([f[:, 1], m12, m01], axis=1),
            np.stack([f[:, 2], m20, m12], axis=1), np.stack([m01, m12, m20], axis=1)], axis=0)
    return v, f


def deindex(v: np.ndarray, f: np.ndarray) -> Mesh:
    """One vertex per face corner, indices 0,1,2,... — the shape Assimp's OBJ importer
    produces so that any conforming loader yields identical Mesh data."""
    pos = v[f.reshape(-1)].astype(np.float32)
    return Mesh(pos, np.arange(pos.shape[0], dtype=np.uint32))


def bunny_like(subdivisions: int = 4, seed: int = 0, radius: float = 0.5) -> Mesh:
    """Deterministic closed 'blob' standing in for assets/models/bunny.obj (an LFS stub):
    an icosphere displaced by a few low-frequency lobes; T = 20 * 4^subdivisions
    (s=4: 5 120 triangles, size-matched to the 15 KB original; s=6: 81 920; s=8: 1.3 M)."""
    v, f = icosphere(subdivisions)
    rng = np.random.default_rng(seed)
    r = np.ones(v.shape[0])
    for _ in range(6):
        d = rng.normal(size=3)
        d /= np.linalg.norm(d)
        freq = rng.uniform(1.5, 4.0)
        r += 0.12 * np.sin(freq * (v @ d) * math.pi + rng.uniform(0, 2 * math.pi))
    v = v * (radius * r)[:, None]
    v[:, 1] += radius  # sits on y = 0 like a model on a ground plane
    return deindex(v, f)


def heightfield(n: int, seed: int = 0, size: float = 2.0, amplitude: float = 0.15) -> Mesh:
    """n x n quads (2 triangles each) of a displaced height field centred at the origin;
    n = 2236 gives ~10.0 M triangles (BASELINE.json configs[3])."""
    rng = np.random.default_rng(seed)
    xs = np.linspace(-size / 2, size / 2, n + 1)
    gx, gz = np.meshgrid(xs, xs, indexing="xy")
    y = np.zeros_like(gx)
    for _ in range(8):
        k = rng.uniform(2.0, 40.0, size=2)
        ph = rng.uniform(0, 2 * math.pi, size=2)
        y += amplitude / 8 * np.sin(k[0] * gx + ph[0]) * np.cos(k[1] * gz + ph[1]) * rng.uniform(0.3, 1.0)
    y += amplitude * 0.02 * rng.standard_normal(y.shape)
    v = np.stack([gx, y, gz], axis=-1).reshape(-1, 3)
    i = np.arange(n)[None, :] + (n + 1) * np.arange(n)[:, None]
    i = i.reshape(-1)
    f = np.concatenate([np.stack([i, i + n + 1, i + 1], axis=1),
                        np.stack([i + 1, i + n + 1, i + n + 2], axis=1)], axis=0)
    return deindex(v, f)


def write_obj(path: str, mesh: Mesh):
    """v/f-only OBJ so that every conforming loader reads the same triangles."""
    pos = mesh.positions.reshape(-1, 3)
    idx = mesh.indices.reshape(-1, 3) + 1
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "w") as fh:
        fh.write("# procedurally generated stand-in (the reference's models are git-LFS stubs)\n")
        np.savetxt(fh, pos, fmt="v %.9g %.9g %.9g")
        np.savetxt(fh, idx, fmt="f %d %d %d")


# ----------------------------------------------------------------------------- bundled scenes
def three_balls(width=800, height=800, spp=1) -> SceneDescription:
    """assets/scenes/three_balls.json with its unparseable camera transform {"o":[0,0,4]}
    read as {"translate":[0,0,4]}."""
    s = SceneDescription()
    s.filename = "scenes/three_balls.json"
    s.add_material("ground", Material.lambertian((0.8, 0.8, 0.0)))
    s.add_material("blue", Material.lambertian((0.1, 0.2, 0.5)))
    s.add_material("dielectric", Material.dielectric(1.5))
    s.add_material("metal", Material.metal((0.8, 0.6, 0.2), 1.0))
    s.add_sphere(100.0, translate((0.0, -100.5, -1.0)), "ground")
    s.add_sphere(0.5, translate((0.0, 0.0, -1.0)), "blue")
    s.add_sphere(0.5, translate((-1.0, 0.0, -1.0)), "dielectric")
    s.add_sphere(0.5, translate((1.0, 0.0, -1.0)), "metal")
    s.camera = Camera((0.0, 0.0, 4.0), (1.0, 0.0, 0.0, 0.0), math.radians(45.0))
    s.resolution = (width, height)
    s.spp = spp
    return s


def bunny_scene(mesh: Mesh | None = None, width=1920, height=1080, spp=10) -> SceneDescription:
    """assets/scenes/bunny.json: ground sphere + two instances of one mesh."""
    s = SceneDescription()
    s.filename = "scenes/bunny.json"
    s.add_material("ground", Material.lambertian((0.8, 0.8, 0.8)))
    s.add_material("bunny", Material.lambertian((0.8, 0.8, 0.5)))
    s.add_material("bunny2", Material.lambertian((0.6, 0.4, 0.8)))
    s.add_mesh("models/bunny.obj", mesh if mesh is not None else bunny_like())
    s.add_sphere(100.0, translate((0.0, -100.5, -1.0)), "ground")
    s.add_mesh_object("models/bunny.obj", translate((1.0, -0.5, -2.0)), "bunny")
    s.add_mesh_object("models/bunny.obj", compose(scale(0.5), translate((-1.0, -0.5, -2.0))), "bunny2")
    s.camera = Camera((0.0, 0.0, 0.0), (1.0, 0.0, 0.0, 0.0), math.radians(60.0))
    s.resolution = (width, height)
    s.spp = spp
    return s


def many_materials_scene(width=1920, height=1080, spp=16, subdiv=4) -> SceneDescription:
    """Material-divergence stress for the ray-binning experiment (reference hint: the commented-out
    sort by material_id, path_tracer.cu:439-446): a 5 x 3 grid of instances of one mesh, neighbours
    alternating between lambertian, metal and dielectric, over the ground sphere and under three
    free-standing spheres of the three types."""
    s = SceneDescription()
    s.filename = "synthetic/many_materials.json"
    s.add_material("ground", Material.lambertian((0.8, 0.8, 0.8)))
    palette = [Material.lambertian((0.8, 0.3, 0.3)), Material.metal((0.8, 0.8, 0.8), 0.05), Material.dielectric(1.5),
               Material.lambertian((0.3, 0.8, 0.3)), Material.metal((0.8, 0.6, 0.2), 0.4), Material.dielectric(1.3)]
    for i, m in enumerate(palette):
        s.add_material(f"m{i}", m)
    s.add_mesh("models/bunny.obj", bunny_like(subdiv))
    s.add_sphere(100.0, translate((0.0, -100.5, -1.0)), "ground")
    k = 0
    for row in range(3):
        for col in range(5):
            x, z = (col - 2) * 0.9, -1.6 - row * 0.9
            s.add_mesh_object("models/bunny.obj", compose(scale(0.6), translate((x, -0.5, z))), f"m{k % len(palette)}")
            k += 1
    for i, x in enumerate((-1.2, 0.0, 1.2)):
        s.add_sphere(0.25, translate((x, 0.55, -1.2)), f"m{(i * 2 + 1) % len(palette)}")
    s.camera = Camera((0.0, 0.4, 1.0), (1.0, 0.0, 0.0, 0.0), math.radians(60.0))
    s.resolution = (width, height)
    s.spp = spp
    return s


def many_spheres_scene(n: int = 400, width=800, height=800, spp=1, seed=0, with_mesh=True) -> SceneDescription:
    """Sphere-count stress (the reference scans its objects linearly per ray, path_tracer.cu:118):
    n small spheres of the three material types scattered over the ground sphere, half of them
    listed before the mesh object and half after it."""
    rng = np.random.default_rng(seed)
    s = SceneDescription()
    s.filename = f"synthetic/many_spheres_{n}.json"
    s.add_material("ground", Material.lambertian((0.8, 0.8, 0.8)))
    s.add_material("a", Material.lambertian((0.8, 0.3, 0.3)))
    s.add_material("b", Material.metal((0.8, 0.8, 0.8), 0.1))
    s.add_material("c", Material.dielectric(1.5))
    if with_mesh:
        s.add_mesh("models/bunny.obj", bunny_like(3))
    s.add_sphere(100.0, translate((0.0, -100.5, -1.0)), "ground")

    def scatter(count):
        for _ in range(count):
            x, z = rng.uniform(-4.0, 4.0), rng.uniform(-9.0, -1.0)
            r = float(rng.uniform(0.05, 0.18))
            y = -100.5 + math.sqrt(max(0.0, (100.0 + r) ** 2 - x * x - (z + 1.0) ** 2))   # resting on the ground
            s.add_sphere(r, translate((float(x), float(y), float(z))), "abc"[int(rng.integers(0, 3))])

    scatter(n // 2)
    if with_mesh:
        s.add_mesh_object("models/bunny.obj", translate((0.0, -0.5, -3.0)), "a")
    scatter(n - n // 2)
    s.camera = Camera((0.0, 0.6, 1.0), (1.0, 0.0, 0.0, 0.0), math.radians(60.0))
    s.resolution = (width, height)
    s.spp = spp
    return s


def terrain_scene(n: int = 2236, width=3840, height=2160, spp=16, seed=0) -> SceneDescription:
    """BASELINE.json configs[3]: one procedural ~2*n^2-triangle mesh, diffuse albedo 0.7."""
    s = SceneDescription()
    s.filename = f"synthetic/terrain_{n}.json"
    s.add_material("grey", Material.lambertian((0.7, 0.7, 0.7)))
    s.add_mesh("synthetic/terrain.obj", heightfield(n, seed=seed))
    s.add_mesh_object("synthetic/terrain.obj", translate((0.0, 0.0, 0.0)), "grey")
    s.camera = Camera.look_at((0.0, 0.9, 1.6), (0.0, 0.0, 0.0), (0.0, 1.0, 0.0), 50.0)
    s.resolution = (width, height)
    s.spp = spp
    return s
