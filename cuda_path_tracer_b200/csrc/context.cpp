// context.cpp — C-ABI entry points: scene upload and the PathTracer-shaped
// integrator context (include/b200pt.h).  Host orchestration only; every
// computation on rays/pixels happens in kernels.cu.
#include "bvh_build.h"
#include "internal.h"

#include <algorithm>
#include <chrono>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace pt {

static thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }
int fail(int code, const std::string& msg)
{
  g_last_error = msg;
  return code;
}
int cuda_fail(cudaError_t e, const char* what)
{
  g_last_error = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what;
  return PT_ERR_CUDA;
}

static double now_ms()
{
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

// transform_point (transform.hpp:37-42): (m0*x + m1*y) + (m2*z + m3*1), then / w
static void xform_point(const float* m, const float* p, float* out)
{
  float v[4];
  for (int r = 0; r < 4; ++r)
    v[r] = (m[0 * 4 + r] * p[0] + m[1 * 4 + r] * p[1]) + (m[2 * 4 + r] * p[2] + m[3 * 4 + r] * 1.0f);
  out[0] = v[0] / v[3];
  out[1] = v[1] / v[3];
  out[2] = v[2] / v[3];
}

static bool is_affine(const float* m)
{
  return m[3] == 0.f && m[7] == 0.f && m[11] == 0.f && m[15] == 1.f;
}

static void rows3x4(const float* colmajor, float* out12)
{
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 4; ++c) out12[r * 4 + c] = colmajor[c * 4 + r];
}

// FNV-1a fingerprint of a scene description: the tables in full, the mesh arrays sampled (every
// 257th vertex and index) together with their sizes.  Stored in progressive-state files so that a
// render is never resumed against a different scene.
static uint64_t description_hash(const pt_scene_desc* d)
{
  uint64_t h = 1469598103934665603ull;
  auto mix = [&](const void* p, size_t n) {
    const unsigned char* b = (const unsigned char*)p;
    for (size_t i = 0; i < n; ++i) h = (h ^ b[i]) * 1099511628211ull;
  };
  mix(&d->n_vertices, 8), mix(&d->n_indices, 8), mix(&d->n_objects, 4), mix(&d->n_spheres, 4);
  mix(&d->n_materials, 4), mix(&d->n_meshes, 4);
  for (uint32_t i = 0; i < d->n_objects; ++i) mix(&d->objects[i], sizeof(pt_object));
  for (uint32_t i = 0; i < d->n_spheres; ++i) mix(&d->spheres[i], sizeof(pt_sphere));
  for (uint32_t i = 0; i < d->n_materials; ++i) mix(&d->materials[i], sizeof(pt_material));
  for (uint64_t i = 0; i < d->n_vertices; i += 257) mix(d->positions + 3 * i, 12);
  for (uint64_t i = 0; i < d->n_indices; i += 257) mix(d->indices + i, 4);
  return h;
}

// World-space bounding sphere for sphere_test's conservative pre-reject.  Only for similarity
// transforms (rotation x uniform scale + translation, columns orthogonal and of equal length to
// 1e-5): there the object-space test's rounding noise has the same relative size in world space,
// which is what the margins in sphere_test assume.  Anything else keeps the exact test alone.
static void sphere_world_bound(const float* m, const pt_sphere& s, DevSphere& d)
{
  double col[3][3], len2[3];
  for (int c = 0; c < 3; ++c) {
    len2[c] = 0.0;
    for (int r = 0; r < 3; ++r) {
      col[c][r] = m[c * 4 + r];
      len2[c] += col[c][r] * col[c][r];
    }
  }
  const double sc2 = (len2[0] + len2[1] + len2[2]) / 3.0;
  bool ok = sc2 > 0.0 && std::isfinite(sc2) && s.radius > 0.f;
  for (int a = 0; ok && a < 3; ++a) {
    if (std::fabs(len2[a] - sc2) > 1e-5 * sc2) ok = false;
    const int b = (a + 1) % 3;
    const double dot = col[a][0] * col[b][0] + col[a][1] * col[b][1] + col[a][2] * col[b][2];
    if (std::fabs(dot) > 1e-5 * sc2) ok = false;
  }
  float cw[3];
  xform_point(m, s.center, cw);
  d.wx = cw[0], d.wy = cw[1], d.wz = cw[2];
  d.wr = (float)(std::sqrt(sc2) * (double)s.radius * (1.0 + 1e-5));
  // magnitudes the world->object transform of a point rounds against: the translation, the
  // world-space centre and the (scaled) object-space centre
  d.wl1 = (float)(std::fabs(cw[0]) + std::fabs(cw[1]) + std::fabs(cw[2]) + std::fabs(m[12]) + std::fabs(m[13]) +
                  std::fabs(m[14]) +
                  std::sqrt(sc2) * (std::fabs(s.center[0]) + std::fabs(s.center[1]) + std::fabs(s.center[2])));
  ok = ok && std::isfinite(d.wr) && std::isfinite(d.wl1);
  static const bool disabled = getenv("PT_SPHERE_PREREJECT") && atoi(getenv("PT_SPHERE_PREREJECT")) == 0;
  d.pre_ok = ok && !disabled ? 1u : 0u;
}

// Tree over one sphere group (DevScene::sph_nodes): median split of the centroids along the
// widest axis, <= 4 spheres per leaf, the group's spheres reordered into leaf order.  Returns the
// child reference of the range [a, b) and its bounds.
struct SphereTreeBuilder {
  std::vector<DevSphere>& sph;
  std::vector<float>& nodes;
  uint32_t offset; // index of the group's first sphere in the scene's sphere array
  struct B {
    float lo[3], hi[3];
  };
  static B box_of(const DevSphere& s)
  {
    B b;
    const float c[3] = {s.wx, s.wy, s.wz};
    for (int a = 0; a < 3; ++a) {
      const float lo = c[a] - s.wr, hi = c[a] + s.wr;
      b.lo[a] = lo - (std::fabs(lo) * 4e-7f + 1e-30f); // outwards: the node test is exact otherwise
      b.hi[a] = hi + (std::fabs(hi) * 4e-7f + 1e-30f);
    }
    return b;
  }
  int build(uint32_t a, uint32_t b, B& out)
  {
    out = box_of(sph[a]);
    for (uint32_t i = a + 1; i < b; ++i) {
      const B bi = box_of(sph[i]);
      for (int k = 0; k < 3; ++k) out.lo[k] = std::min(out.lo[k], bi.lo[k]), out.hi[k] = std::max(out.hi[k], bi.hi[k]);
    }
    if (b - a <= 4) return ~(int)(((offset + a) << 3) | (b - a - 1));
    float clo[3] = {sph[a].wx, sph[a].wy, sph[a].wz}, chi[3] = {sph[a].wx, sph[a].wy, sph[a].wz};
    for (uint32_t i = a + 1; i < b; ++i) {
      const float c[3] = {sph[i].wx, sph[i].wy, sph[i].wz};
      for (int k = 0; k < 3; ++k) clo[k] = std::min(clo[k], c[k]), chi[k] = std::max(chi[k], c[k]);
    }
    int axis = 0;
    for (int k = 1; k < 3; ++k)
      if (chi[k] - clo[k] > chi[axis] - clo[axis]) axis = k;
    const uint32_t mid = a + (b - a) / 2;
    std::nth_element(sph.begin() + a, sph.begin() + mid, sph.begin() + b, [axis](const DevSphere& x, const DevSphere& y) {
      const float cx[3] = {x.wx, x.wy, x.wz}, cy[3] = {y.wx, y.wy, y.wz};
      return cx[axis] < cy[axis];
    });
    const int me = (int)(nodes.size() / 16);
    nodes.resize(nodes.size() + 16, 0.f);
    B b0, b1;
    const int c0 = build(a, mid, b0), c1 = build(mid, b, b1);
    float* n = nodes.data() + (size_t)me * 16; // (after the recursion: the vector may have grown)
    n[0] = b0.lo[0], n[1] = b0.hi[0], n[2] = b0.lo[1], n[3] = b0.hi[1];
    n[4] = b1.lo[0], n[5] = b1.hi[0], n[6] = b1.lo[1], n[7] = b1.hi[1];
    n[8] = b0.lo[2], n[9] = b0.hi[2], n[10] = b1.lo[2], n[11] = b1.hi[2];
    std::memcpy(n + 12, &c0, 4);
    std::memcpy(n + 13, &c1, 4);
    return me;
  }
};

// Groups of more than PT_SPHERE_BVH_MIN spheres, all placed rigidly (unit scale: the sphere test
// then reports the same distance whatever else was hit first, so the order of the tests does not
// matter), get a tree; anything else keeps the reference's linear scan.
static void build_sphere_trees(SceneBuild& sb, const std::vector<uint8_t>& rigid)
{
  static const bool disabled = getenv("PT_SPHERE_BVH") && atoi(getenv("PT_SPHERE_BVH")) == 0;
  if (disabled) return;
  auto group = [&](uint32_t a, uint32_t b) -> int {
    if (b - a <= PT_SPHERE_BVH_MIN) return -1;
    for (uint32_t i = a; i < b; ++i)
      if (!rigid[i]) return -1;
    // the builder reorders a copy of the group into leaf order; `a` offsets its leaf references
    std::vector<DevSphere> part(sb.spheres.begin() + a, sb.spheres.begin() + b);
    SphereTreeBuilder builder{part, sb.sph_nodes, a};
    SphereTreeBuilder::B box;
    const int root = builder.build(0, (uint32_t)part.size(), box);
    std::copy(part.begin(), part.end(), sb.spheres.begin() + a);
    return root;
  };
  sb.sph_root_before = group(0, sb.n_spheres_before);
  sb.sph_root_after = group(sb.n_spheres_before, (uint32_t)sb.spheres.size());
}

} // namespace pt

using namespace pt;

// Index range [first, first + 3 * count) of the mesh a mesh object instances.
static bool mesh_range(const pt_scene_desc* desc, const pt_object& ob, uint64_t& first_tri, uint64_t& n_tri)
{
  if (desc->n_meshes == 0) {
    first_tri = 0;
    n_tri = desc->n_indices / 3;
    return true;
  }
  if (ob.prim_index >= desc->n_meshes || !desc->mesh_first_index) return false;
  const uint64_t a = desc->mesh_first_index[ob.prim_index], b = desc->mesh_first_index[ob.prim_index + 1];
  if (a > b || b > desc->n_indices || a % 3 != 0 || b % 3 != 0) return false;
  first_tri = a / 3;
  n_tri = (b - a) / 3;
  return true;
}

// Every mesh instance baked to world space (the reference re-transforms three vertices per
// leaf visit instead, path_tracer.cu:57-59).
static int bake_triangles(const pt_scene_desc* desc, const std::vector<uint32_t>& mesh_objects,
                          BuildTris& tris)
{
  std::vector<uint64_t> first(mesh_objects.size()), count(mesh_objects.size()), out_at(mesh_objects.size());
  uint64_t n_world = 0;
  for (size_t k = 0; k < mesh_objects.size(); ++k) {
    if (!mesh_range(desc, desc->objects[mesh_objects[k]], first[k], count[k]))
      return fail(PT_ERR_INVALID, "mesh object refers to a mesh that does not exist");
    out_at[k] = n_world;
    n_world += count[k];
  }
  if (n_world >= (1ull << 28)) return fail(PT_ERR_INVALID, "too many world-space triangles");
  try {
    tris.resize(n_world);
  } catch (...) {
    return fail(PT_ERR_NOMEM, "out of host memory baking mesh instances");
  }
  for (size_t k = 0; k < mesh_objects.size(); ++k) {
    const uint32_t oi = mesh_objects[k];
    const pt_object& ob = desc->objects[oi];
    BuildTri* dst = tris.data() + out_at[k];
    const uint64_t t0 = first[k];
#pragma omp parallel for schedule(static)
    for (long long t = 0; t < (long long)count[k]; ++t) {
      BuildTri& bt = dst[t];
      const uint64_t g = t0 + (uint64_t)t;
      xform_point(ob.m, desc->positions + 3 * (size_t)desc->indices[3 * g + 0], bt.v0);
      xform_point(ob.m, desc->positions + 3 * (size_t)desc->indices[3 * g + 1], bt.v1);
      xform_point(ob.m, desc->positions + 3 * (size_t)desc->indices[3 * g + 2], bt.v2);
      bt.prim = (uint32_t)g;
      bt.object = oi;
      bt.material = ob.material;
    }
  }
  return PT_OK;
}

extern "C" {

const char* pt_last_error(void) { return g_last_error.c_str(); }
int pt_version(void) { return 100; }

void pt_params_default(pt_params* p)
{
  if (!p) return;
  std::memset(p, 0, sizeof(*p));
  p->max_depth = 50; // reference: max_bounces (path_tracer.cu:27)
  p->rng_mode = PT_RNG_PIXEL_STREAM;
  p->max_iterations = 0;
  p->samples_per_pass = 0;
}

void pt_denoise_params_default(pt_denoise_params* p)
{
  if (!p) return;
  std::memset(p, 0, sizeof(*p));
  p->filter_size = 10;
  p->color_weight = 0.45f;
  p->normal_weight = 0.30f;
  p->position_weight = 0.25f;
  p->clamp_fix = 0;
}

// ------------------------------------------------------------------- scene
// Scene creation in two halves so that a multi-GPU group builds the tree once and uploads it to
// every device: scene_prepare (host only: validation, sphere/material tables, instance bake + SAH
// build unless the device builder is asked for) and scene_upload (one device).
} // extern "C"

namespace pt {

int scene_prepare(const pt_scene_desc* desc, bool host_build, SceneBuild& sb)
{
  if (!desc) return fail(PT_ERR_INVALID, "pt_scene_create: null argument");
  if (desc->n_indices % 3 != 0) return fail(PT_ERR_INVALID, "index count is not a multiple of 3");
  if (desc->n_objects && !desc->objects) return fail(PT_ERR_INVALID, "objects is null");
  if (desc->n_materials == 0 || !desc->materials)
    return fail(PT_ERR_INVALID, "scene needs at least one material");
  if (desc->n_indices && (!desc->indices || !desc->positions))
    return fail(PT_ERR_INVALID, "mesh arrays are null");
  if (desc->n_spheres && !desc->spheres) return fail(PT_ERR_INVALID, "spheres is null");
  for (uint64_t i = 0; i < desc->n_indices; ++i)
    if (desc->indices[i] >= desc->n_vertices)
      return fail(PT_ERR_INVALID, "vertex index out of range");
  for (uint32_t i = 0; i < desc->n_materials; ++i)
    if (desc->materials[i].type < 0 || desc->materials[i].type > 2)
      return fail(PT_ERR_INVALID, "unknown material type");

  const double t0 = now_ms();
  // ---- objects -> spheres (reference order quirks kept) + mesh instances
  std::vector<DevSphere> sph_after;
  std::vector<uint8_t> rigid_before, rigid_after;
  bool seen_mesh = false;
  for (uint32_t i = 0; i < desc->n_objects; ++i) {
    const pt_object& ob = desc->objects[i];
    if (ob.material >= desc->n_materials)
      return fail(PT_ERR_INVALID, "object material index out of range");
    if (!is_affine(ob.m) || !is_affine(ob.inv))
      return fail(PT_ERR_INVALID, "projective object transforms are not supported");
    if (ob.type == PT_OBJ_SPHERE) {
      if (ob.prim_index >= desc->n_spheres)
        return fail(PT_ERR_INVALID, "sphere index out of range");
      const pt_sphere& s = desc->spheres[ob.prim_index];
      DevSphere d{};
      rows3x4(ob.inv, d.inv);
      rows3x4(ob.m, d.m);
      d.cx = s.center[0], d.cy = s.center[1], d.cz = s.center[2];
      d.radius = s.radius;
      d.material = ob.material;
      d.object = (int32_t)i;
      sphere_world_bound(ob.m, s, d);
      // rigid = similarity (pre_ok) with unit scale
      const double len2 = (double)ob.m[0] * ob.m[0] + (double)ob.m[1] * ob.m[1] + (double)ob.m[2] * ob.m[2];
      (seen_mesh ? rigid_after : rigid_before).push_back(d.pre_ok && std::fabs(len2 - 1.0) < 1e-5 ? 1 : 0);
      (seen_mesh ? sph_after : sb.spheres).push_back(d);
    } else if (ob.type == PT_OBJ_MESH) {
      seen_mesh = true;
      sb.mesh_objects.push_back(i);
      uint64_t first, count;
      if (!mesh_range(desc, ob, first, count))
        return fail(PT_ERR_INVALID, "mesh object refers to a mesh that does not exist");
    } else {
      return fail(PT_ERR_INVALID, "unknown object type");
    }
  }
  sb.n_spheres_before = (uint32_t)sb.spheres.size();
  sb.spheres.insert(sb.spheres.end(), sph_after.begin(), sph_after.end());
  rigid_before.insert(rigid_before.end(), rigid_after.begin(), rigid_after.end());
  build_sphere_trees(sb, rigid_before);
  sb.content_hash = description_hash(desc);

  sb.mats.resize(desc->n_materials);
  for (uint32_t i = 0; i < desc->n_materials; ++i) {
    const pt_material& m = desc->materials[i];
    DevMaterial d{};
    d.type = m.type;
    d.r = m.albedo[0], d.g = m.albedo[1], d.b = m.albedo[2];
    d.param = m.type == PT_MAT_METAL ? m.fuzz : (m.type == PT_MAT_DIELECTRIC ? m.refraction_index : 0.f);
    sb.mats[i] = d;
  }
  // PT_BVH=8 additionally derives the compressed 8-wide tree and traverses it (opt-in: on B200
  // it trades the binary walk's L1 wavefront bound for an ALU-pipe bound and measures 10-30 %
  // slower, profiles/README.md); the default is the binary tree alone.
  const char* bvh_env = getenv("PT_BVH");
  sb.wide = bvh_env && atoi(bvh_env) == 8;
  if (host_build) {
    BuildTris tris;
    const int rc = bake_triangles(desc, sb.mesh_objects, tris);
    if (rc != PT_OK) return rc;
    sb.n_world = tris.size();
    build_bvh(tris, sb.bvh, sb.wide);
    sb.host_built = true;
  }
  sb.build_ms = now_ms() - t0;
  return PT_OK;
}

// The 32-byte quantised companion of the 64-byte nodes (DevScene::qnodes): every child plane on a
// 16-bit grid over the root box, rounded outwards and then moved one more cell outwards — the
// reserve the device's half-cell rounding needs (trav_init<QN>).  The grid spans 65 000 cells per
// axis (an axis thinner than 1/1024 of the longest one is given that much), 200 cells below the
// root box.  Returns false — the traversal then walks the exact nodes — if a plane does not fit.
static bool quantise_nodes(const FlatBVH& bvh, std::vector<uint32_t>& q, float org[3], float cell[3])
{
  const uint32_t n = bvh.n_nodes;
  if (n == 0 || bvh.nodes.size() < (size_t)n * 16) return false;
  float ext[3], longest = 0.f;
  for (int a = 0; a < 3; ++a) {
    ext[a] = bvh.root_hi[a] - bvh.root_lo[a];
    if (!(ext[a] >= 0.f) || !std::isfinite(ext[a])) return false;
    longest = std::max(longest, ext[a]);
  }
  if (!(longest > 0.f)) return false;
  for (int a = 0; a < 3; ++a) {
    cell[a] = std::max(ext[a], longest / 1024.f) / 65000.f;
    org[a] = bvh.root_lo[a] - 200.f * cell[a];
    if (!(cell[a] > 0.f) || !std::isfinite(org[a])) return false;
  }
  q.resize((size_t)n * 8);
  // plane slots of the 64-byte node (common.cuh): {lo, hi} x {x, y, z} x {child 0, child 1}
  static const int src[6][2] = {{0, 1}, {2, 3}, {8, 9}, {4, 5}, {6, 7}, {10, 11}};
  static const int axis[6] = {0, 1, 2, 0, 1, 2};
  bool ok = true;
#pragma omp parallel for schedule(static) reduction(&& : ok)
  for (long long i = 0; i < (long long)n; ++i) {
    const float* nd = &bvh.nodes[(size_t)i * 16];
    uint32_t* w = &q[(size_t)i * 8];
    for (int k = 0; k < 6; ++k) {
      const int a = axis[k];
      const double lo = std::floor(((double)nd[src[k][0]] - (double)org[a]) / (double)cell[a]) - 1.0;
      const double hi = std::ceil(((double)nd[src[k][1]] - (double)org[a]) / (double)cell[a]) + 1.0;
      if (!(lo >= 0.0 && hi <= 65535.0 && lo <= hi)) ok = false;
      const uint32_t ql = (uint32_t)std::min(65535.0, std::max(0.0, lo)), qh = (uint32_t)std::min(65535.0, std::max(0.0, hi));
      w[k] = ql | (qh << 16);
    }
    std::memcpy(&w[6], &nd[12], 4);
    std::memcpy(&w[7], &nd[13], 4);
  }
  return ok;
}

int scene_upload(const pt_scene_desc* desc, const SceneBuild& sb, DeviceLBVH* dl, int device, pt_scene** out)
{
  *out = nullptr;
  int dev_count = 0;
  PT_CUDA(cudaGetDeviceCount(&dev_count));
  if (device < 0 || device >= dev_count) return fail(PT_ERR_INVALID, "no such CUDA device");
  PT_CUDA(cudaSetDevice(device));
  const double t1 = now_ms();
  const FlatBVH& bvh = sb.bvh;
  const bool from_device = dl && dl->built;
  pt_scene* sc = new pt_scene();
  sc->device = device;
  auto upload = [&](const void* src, size_t bytes, void** dst) -> cudaError_t {
    *dst = nullptr;
    if (bytes == 0) return cudaSuccess;
    cudaError_t e = cudaMalloc(dst, bytes);
    if (e != cudaSuccess) return e;
    sc->info.device_bytes += bytes;
    return cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
  };
  cudaError_t e = cudaSuccess;
  if (from_device) {
    // ownership of the device-built arrays moves to the scene at once: any failure below frees them
    sc->d_nodes = dl->nodes;
    sc->d_tris = dl->tris;
    dl->nodes = nullptr;
    dl->tris = nullptr;
    sc->info.device_bytes += ((size_t)dl->n_nodes * 16 + (size_t)dl->n_tris * 12) * 4;
    sc->info.device_build = 1;
  } else {
    if (e == cudaSuccess) e = upload(bvh.nodes.data(), bvh.nodes.size() * 4, &sc->d_nodes);
    if (e == cudaSuccess) e = upload(bvh.tris.data(), bvh.tris.size() * 4, &sc->d_tris);
  }
  if (e == cudaSuccess) e = upload(bvh.nodes8.data(), bvh.nodes8.size() * 4, &sc->d_nodes8);
  // Quantised nodes for the traversal kernels: host-built binary trees of scenes that are not
  // 'big' (the > 512 MB variant of traverse_kernel keeps the exact nodes: its cells would be a
  // visible fraction of a triangle) and have no sphere trees.  PT_QNODES=0 / 1 forces off / on.
  float q_org[3] = {0.f, 0.f, 0.f}, q_cell[3] = {0.f, 0.f, 0.f};
  if (e == cudaSuccess && !from_device && bvh.n_nodes8 == 0 && bvh.n_nodes != 0) {
    static const int q_env = [] {
      const char* v = getenv("PT_QNODES");
      return v ? atoi(v) : -1;
    }();
    const size_t scene_bytes = (size_t)bvh.n_nodes * 64 + (size_t)bvh.n_tris * 48;
    const bool sphere_trees = sb.sph_root_before >= 0 || sb.sph_root_after >= 0;
    const bool want = q_env >= 0 ? q_env != 0 : scene_bytes <= (512ull << 20);
    std::vector<uint32_t> q;
    if (want && !sphere_trees && quantise_nodes(bvh, q, q_org, q_cell)) e = upload(q.data(), q.size() * 4, &sc->d_qnodes);
  }
  if (e == cudaSuccess) e = upload(sb.spheres.data(), sb.spheres.size() * sizeof(DevSphere), &sc->d_spheres);
  if (e == cudaSuccess) e = upload(sb.sph_nodes.data(), sb.sph_nodes.size() * 4, &sc->d_sph_nodes);
  if (e == cudaSuccess) e = upload(sb.mats.data(), sb.mats.size() * sizeof(DevMaterial), &sc->d_materials);
  if (e != cudaSuccess) {
    pt_scene_destroy(sc);
    return cuda_fail(e, "scene upload");
  }
  const double t2 = now_ms();

  const uint32_t n_nodes = from_device ? dl->n_nodes : bvh.n_nodes;
  const uint32_t n_tris = from_device ? dl->n_tris : bvh.n_tris;
  sc->dev.nodes = (const float4*)sc->d_nodes;
  sc->dev.tris = (const float4*)sc->d_tris;
  sc->dev.spheres = (const DevSphere*)sc->d_spheres;
  sc->dev.materials = (const DevMaterial*)sc->d_materials;
  sc->dev.n_spheres = (uint32_t)sb.spheres.size();
  sc->dev.n_spheres_before = sb.n_spheres_before;
  sc->dev.sph_nodes = (const float4*)sc->d_sph_nodes;
  sc->dev.sph_root_before = sb.sph_root_before;
  sc->dev.sph_root_after = sb.sph_root_after;
  sc->dev.n_nodes = n_nodes;
  sc->dev.qnodes = (const uint4*)sc->d_qnodes;
  for (int a = 0; a < 3; ++a) sc->dev.q_org[a] = q_org[a], sc->dev.q_cell[a] = q_cell[a];
  // the wide traversal keeps a PT_STACK8-entry stack: one node group per level at most
  const bool use_wide = !from_device && bvh.n_nodes8 != 0 && bvh.depth8 <= PT_STACK8;
  sc->dev.nodes8 = use_wide ? (const uint4*)sc->d_nodes8 : nullptr;
  sc->dev.n_nodes8 = use_wide ? bvh.n_nodes8 : 0u;
  sc->info.n_bvh8_nodes = bvh.n_nodes8;
  sc->info.bvh8_depth = bvh.depth8;
  sc->dev.n_tris = n_tris;
  for (int a = 0; a < 3; ++a) {
    sc->dev.root_lo[a] = from_device ? dl->root_lo[a] : bvh.root_lo[a];
    sc->dev.root_hi[a] = from_device ? dl->root_hi[a] : bvh.root_hi[a];
  }
  sc->info.n_triangles = desc->n_indices / 3;
  sc->info.n_world_triangles = sb.n_world;
  sc->info.n_bvh_nodes = n_nodes;
  sc->info.n_bvh_triangles = n_tris;
  sc->info.bvh_depth = from_device ? dl->depth : bvh.depth;
  sc->info.n_objects = desc->n_objects;
  sc->info.n_spheres = (uint32_t)sb.spheres.size();
  sc->info.n_materials = desc->n_materials;
  sc->info.build_ms = sb.build_ms;
  sc->info.upload_ms = t2 - t1;
  sc->content_hash = sb.content_hash;
  *out = sc;
  return PT_OK;
}

} // namespace pt

extern "C" {

static int pt_scene_create_impl(const pt_scene_desc* desc, int device, pt_scene** out)
{
  if (!desc || !out) return fail(PT_ERR_INVALID, "pt_scene_create: null argument");
  *out = nullptr;
  // PT_BUILD=lbvh (cuda_pt --fast-build): bake the instances and build the tree on the device
  // (lbvh.cu) instead of the host bake + SAH builder — tens of milliseconds at 10 M triangles, a
  // lower-quality tree.  The device builder may decline (tiny scene, tree deeper than the
  // traversal stack): the host path then runs as usual.
  const char* build_env = getenv("PT_BUILD");
  const char* bvh_env = getenv("PT_BVH");
  const bool want_lbvh = build_env && !strcmp(build_env, "lbvh") && !(bvh_env && atoi(bvh_env) == 8);
  SceneBuild sb;
  int rc = scene_prepare(desc, !want_lbvh, sb); // every argument check happens before the device is touched
  if (rc != PT_OK) return rc;
  DeviceLBVH dl;
  if (want_lbvh) {
    int dev_count = 0;
    PT_CUDA(cudaGetDeviceCount(&dev_count));
    if (device < 0 || device >= dev_count) return fail(PT_ERR_INVALID, "no such CUDA device");
    PT_CUDA(cudaSetDevice(device));
    const double t0 = now_ms();
    if (!sb.mesh_objects.empty()) {
      std::vector<MeshInstance> inst(sb.mesh_objects.size());
      uint64_t n_world = 0;
      for (size_t k = 0; k < sb.mesh_objects.size(); ++k) {
        const pt_object& ob = desc->objects[sb.mesh_objects[k]];
        MeshInstance& mi = inst[k];
        std::memcpy(mi.m, ob.m, sizeof(mi.m));
        mesh_range(desc, ob, mi.first_tri, mi.n_tri);
        mi.out_at = n_world;
        mi.object = sb.mesh_objects[k];
        mi.material = ob.material;
        n_world += mi.n_tri;
      }
      const int ce = build_lbvh_device_mesh_c(desc->positions, desc->n_vertices, desc->indices, desc->n_indices,
                                              inst.data(), (uint32_t)inst.size(), n_world, dl);
      if (ce != 0) return cuda_fail((cudaError_t)ce, "device LBVH build");
      if (dl.built) sb.n_world = n_world;
    }
    if (!dl.built) { // declined: host bake + SAH build
      BuildTris tris;
      rc = bake_triangles(desc, sb.mesh_objects, tris);
      if (rc != PT_OK) return rc;
      sb.n_world = tris.size();
      build_bvh(tris, sb.bvh, false);
      sb.host_built = true;
    }
    sb.build_ms += now_ms() - t0;
  }
  rc = scene_upload(desc, sb, &dl, device, out);
  if (dl.nodes) cudaFree(dl.nodes); // not taken over by a scene (upload failed early)
  if (dl.tris) cudaFree(dl.tris);
  return rc;
}

int pt_scene_create(const pt_scene_desc* desc, int device, pt_scene** out)
{
  return guarded("pt_scene_create", [&] { return pt_scene_create_impl(desc, device, out); });
}

// ---- host-only builder + structural validator (CPU tests, host build benchmark): no CUDA call
struct pt_host_bvh_impl {
  FlatBVH bvh;
  double build_ms = 0.0;
  // quantised companion of bvh.nodes, made on first request (pt_host_bvh_quantised)
  std::vector<uint32_t> qnodes;
  float q_org[3] = {0.f, 0.f, 0.f}, q_cell[3] = {0.f, 0.f, 0.f};
  int q_state = 0; // 0 = not tried, 1 = built, -1 = the quantiser declined
};

static int pt_host_bvh_build_impl(const pt_scene_desc* desc, int wide, pt_host_bvh** out, pt_scene_info* info)
{
  if (!desc || !out) return fail(PT_ERR_INVALID, "pt_host_bvh_build: null argument");
  *out = nullptr;
  if (desc->n_indices % 3 != 0) return fail(PT_ERR_INVALID, "index count is not a multiple of 3");
  for (uint64_t i = 0; i < desc->n_indices; ++i)
    if (desc->indices[i] >= desc->n_vertices) return fail(PT_ERR_INVALID, "vertex index out of range");
  std::vector<uint32_t> mesh_objects;
  for (uint32_t i = 0; i < desc->n_objects; ++i)
    if (desc->objects[i].type == PT_OBJ_MESH) mesh_objects.push_back(i);
  const double t0 = now_ms();
  BuildTris tris;
  const int rc = bake_triangles(desc, mesh_objects, tris);
  if (rc != PT_OK) return rc;
  auto* h = new pt_host_bvh_impl();
  // wide: 0 = binary SAH tree, 1 = + compressed 8-wide tree, 2 = LBVH (host restatement of the
  // device builder; falls back to the SAH builder for scenes it declines)
  if (wide != 2 || !build_lbvh_host(tris, h->bvh)) build_bvh(tris, h->bvh, wide == 1);
  h->build_ms = now_ms() - t0;
  if (info) {
    *info = pt_scene_info{};
    info->n_triangles = desc->n_indices / 3;
    info->n_world_triangles = tris.size();
    info->n_bvh_nodes = h->bvh.n_nodes;
    info->n_bvh_triangles = h->bvh.n_tris;
    info->bvh_depth = h->bvh.depth;
    info->n_bvh8_nodes = h->bvh.n_nodes8;
    info->bvh8_depth = h->bvh.depth8;
    info->build_ms = h->build_ms;
  }
  *out = reinterpret_cast<pt_host_bvh*>(h);
  return PT_OK;
}

int pt_host_bvh_build(const pt_scene_desc* desc, int wide, pt_host_bvh** out, pt_scene_info* info)
{
  return guarded("pt_host_bvh_build", [&] { return pt_host_bvh_build_impl(desc, wide, out, info); });
}

int pt_host_bvh_validate(const pt_host_bvh* hb, uint64_t* violations)
{
  if (!hb || !violations) return fail(PT_ERR_INVALID, "pt_host_bvh_validate: null argument");
  *violations = validate_bvh(reinterpret_cast<const pt_host_bvh_impl*>(hb)->bvh);
  return PT_OK;
}

int pt_host_bvh_arrays(const pt_host_bvh* hb, const float** nodes, const uint32_t** nodes8, const float** tris)
{
  if (!hb) return fail(PT_ERR_INVALID, "pt_host_bvh_arrays: null argument");
  const FlatBVH& b = reinterpret_cast<const pt_host_bvh_impl*>(hb)->bvh;
  if (nodes) *nodes = b.nodes.data();
  if (nodes8) *nodes8 = b.nodes8.data();
  if (tris) *tris = b.tris.data();
  return PT_OK;
}

int pt_host_bvh_quantised(pt_host_bvh* hb, const uint32_t** qnodes, float* org3, float* cell3)
{
  if (!hb || !qnodes || !org3 || !cell3) return fail(PT_ERR_INVALID, "pt_host_bvh_quantised: null argument");
  return guarded("pt_host_bvh_quantised", [&] {
    auto* h = reinterpret_cast<pt_host_bvh_impl*>(hb);
    if (h->q_state == 0) h->q_state = quantise_nodes(h->bvh, h->qnodes, h->q_org, h->q_cell) ? 1 : -1;
    if (h->q_state < 0) return fail(PT_ERR_INVALID, "the tree does not fit the 16-bit grid (degenerate or non-finite bounds)");
    *qnodes = h->qnodes.data();
    for (int a = 0; a < 3; ++a) org3[a] = h->q_org[a], cell3[a] = h->q_cell[a];
    return (int)PT_OK;
  });
}

int pt_host_bvh_trace_stats(const pt_host_bvh* hb, const float* rays8, uint64_t n_rays, int wide, uint64_t* out5)
{
  if (!hb || !rays8 || !out5) return fail(PT_ERR_INVALID, "pt_host_bvh_trace_stats: null argument");
  trace_stats(reinterpret_cast<const pt_host_bvh_impl*>(hb)->bvh, rays8, n_rays, wide, out5);
  return PT_OK;
}

static int pt_host_scene_check_impl(const pt_scene_desc* desc, uint64_t* out4, uint64_t* hash_out)
{
  if (!desc || !out4) return fail(PT_ERR_INVALID, "pt_host_scene_check: null argument");
  SceneBuild sb;
  const int rc = scene_prepare(desc, false, sb); // tables + sphere trees only: no mesh build
  if (rc != PT_OK) return rc;
  const uint32_t n = (uint32_t)sb.spheres.size();
  std::vector<uint32_t> seen(n, 0u);
  uint64_t violations = 0, trees = 0;
  struct Item {
    int node;
    float lo[3], hi[3];
  };
  auto walk = [&](int root, uint32_t lo_i, uint32_t hi_i) {
    if (root < 0) { // linear group: every sphere is "referenced" by the scan
      for (uint32_t i = lo_i; i < hi_i; ++i) seen[i]++;
      return;
    }
    ++trees;
    std::vector<Item> todo;
    Item r{root, {-FLT_MAX, -FLT_MAX, -FLT_MAX}, {FLT_MAX, FLT_MAX, FLT_MAX}};
    todo.push_back(r);
    while (!todo.empty()) {
      const Item it = todo.back();
      todo.pop_back();
      if (it.node >= 0) {
        if ((size_t)it.node * 16 + 16 > sb.sph_nodes.size()) {
          ++violations;
          continue;
        }
        const float* nd = sb.sph_nodes.data() + (size_t)it.node * 16;
        for (int k = 0; k < 2; ++k) {
          Item c;
          std::memcpy(&c.node, nd + 12 + k, 4);
          const float* xy = nd + 4 * k;
          c.lo[0] = xy[0], c.hi[0] = xy[1], c.lo[1] = xy[2], c.hi[1] = xy[3];
          c.lo[2] = nd[8 + 2 * k], c.hi[2] = nd[9 + 2 * k];
          for (int a = 0; a < 3; ++a) // a child box may not stick out of its parent's
            if (c.lo[a] < it.lo[a] || c.hi[a] > it.hi[a]) ++violations;
          todo.push_back(c);
        }
      } else {
        const uint32_t leaf = (uint32_t)(~it.node), first = leaf >> 3, count = (leaf & 7u) + 1u;
        for (uint32_t i = first; i < first + count; ++i) {
          if (i < lo_i || i >= hi_i) {
            ++violations;
            continue;
          }
          seen[i]++;
          const DevSphere& s = sb.spheres[i];
          const float c3[3] = {s.wx, s.wy, s.wz};
          for (int a = 0; a < 3; ++a)
            if (c3[a] - s.wr < it.lo[a] || c3[a] + s.wr > it.hi[a]) ++violations;
        }
      }
    }
  };
  walk(sb.sph_root_before, 0u, sb.n_spheres_before);
  walk(sb.sph_root_after, sb.n_spheres_before, n);
  for (uint32_t i = 0; i < n; ++i)
    if (seen[i] != 1u) ++violations;
  out4[0] = n;
  out4[1] = sb.sph_nodes.size() / 16;
  out4[2] = violations;
  out4[3] = trees;
  if (hash_out) *hash_out = sb.content_hash;
  return PT_OK;
}

int pt_host_scene_check(const pt_scene_desc* desc, uint64_t* out4, uint64_t* hash_out)
{
  return guarded("pt_host_scene_check", [&] { return pt_host_scene_check_impl(desc, out4, hash_out); });
}

int pt_host_bvh_free(pt_host_bvh* hb)
{
  delete reinterpret_cast<pt_host_bvh_impl*>(hb);
  return PT_OK;
}

int pt_scene_destroy(pt_scene* sc)
{
  if (!sc) return PT_OK;
  cudaSetDevice(sc->device);
  cudaFree(sc->d_nodes);
  cudaFree(sc->d_tris);
  cudaFree(sc->d_nodes8);
  cudaFree(sc->d_qnodes);
  cudaFree(sc->d_spheres);
  cudaFree(sc->d_sph_nodes);
  cudaFree(sc->d_materials);
  delete sc;
  return PT_OK;
}

int pt_scene_copy_bvh(const pt_scene* sc, float* nodes_out, float* tris_out)
{
  if (!sc) return fail(PT_ERR_INVALID, "pt_scene_copy_bvh: null scene");
  PT_CUDA(cudaSetDevice(sc->device));
  if (nodes_out && sc->dev.n_nodes)
    PT_CUDA(cudaMemcpy(nodes_out, sc->d_nodes, (size_t)sc->dev.n_nodes * 64, cudaMemcpyDeviceToHost));
  if (tris_out && sc->dev.n_tris)
    PT_CUDA(cudaMemcpy(tris_out, sc->d_tris, (size_t)sc->dev.n_tris * 48, cudaMemcpyDeviceToHost));
  return PT_OK;
}

int pt_scene_get_info(const pt_scene* sc, pt_scene_info* info)
{
  if (!sc || !info) return fail(PT_ERR_INVALID, "pt_scene_get_info: null argument");
  *info = sc->info;
  return PT_OK;
}

static void desc_from_file(const SceneFile& sf, pt_scene_desc& d)
{
  d = pt_scene_desc{};
  d.positions = sf.positions.data();
  d.n_vertices = sf.positions.size() / 3;
  d.indices = sf.indices.data();
  d.n_indices = sf.indices.size();
  d.objects = sf.objects.data();
  d.n_objects = (uint32_t)sf.objects.size();
  d.spheres = sf.spheres.data();
  d.n_spheres = (uint32_t)sf.spheres.size();
  d.materials = sf.materials.data();
  d.n_materials = (uint32_t)sf.materials.size();
  if (!sf.mesh_first_index.empty()) {
    d.n_meshes = (uint32_t)sf.mesh_first_index.size() - 1;
    d.mesh_first_index = sf.mesh_first_index.data();
  }
}

static int pt_scene_file_read_impl(const char* json_path, pt_scene_file** out, pt_scene_desc* desc,
                       pt_scene_file_info* info)
{
  if (!json_path || !out || !desc) return fail(PT_ERR_INVALID, "pt_scene_file_read: null argument");
  *out = nullptr;
  const double t0 = now_ms();
  auto* f = new pt_scene_file();
  const int rc = load_scene_file(json_path, f->sf);
  if (rc != PT_OK) {
    delete f;
    return rc;
  }
  desc_from_file(f->sf, *desc);
  if (info) {
    info->camera = f->sf.camera;
    info->width = f->sf.width;
    info->height = f->sf.height;
    info->spp = f->sf.spp;
    info->load_ms = now_ms() - t0;
  }
  *out = f;
  return PT_OK;
}

int pt_scene_file_read(const char* json_path, pt_scene_file** out, pt_scene_desc* desc,
                       pt_scene_file_info* info)
{
  return guarded("pt_scene_file_read", [&] { return pt_scene_file_read_impl(json_path, out, desc, info); });
}

int pt_scene_file_free(pt_scene_file* f)
{
  delete f;
  return PT_OK;
}

static int pt_scene_load_file_impl(const char* json_path, int device, pt_scene** out, pt_scene_file_info* info)
{
  if (!json_path || !out) return fail(PT_ERR_INVALID, "pt_scene_load_file: null argument");
  const double t0 = now_ms();
  SceneFile sf;
  int rc = load_scene_file(json_path, sf);
  if (rc != PT_OK) return rc;
  const double t1 = now_ms();
  pt_scene_desc d;
  desc_from_file(sf, d);
  rc = pt_scene_create(&d, device, out);
  if (rc != PT_OK) return rc;
  if (info) {
    info->camera = sf.camera;
    info->width = sf.width;
    info->height = sf.height;
    info->spp = sf.spp;
    info->load_ms = t1 - t0;
  }
  return PT_OK;
}

int pt_scene_load_file(const char* json_path, int device, pt_scene** out, pt_scene_file_info* info)
{
  return guarded("pt_scene_load_file", [&] { return pt_scene_load_file_impl(json_path, device, out, info); });
}

// ----------------------------------------------------------------- context
static void free_image_buffers(pt_ctx* c)
{
  if (c->lane2) free_image_buffers(c->lane2);
  if (c->is_lane) { // a lane shares the parent's frame buffers: only the path state is its own
    c->d_sums = nullptr;
    c->own_sums = false;
  }
  cudaFree(c->d_state);
  cudaFree(c->pb.queue[0]);
  cudaFree(c->pb.queue[1]);
  cudaFree(c->pb.tq);
  c->pb.tq = nullptr;
  cudaFree(c->pb.flags);
  cudaFree(c->pb.block_sums);
  cudaFree(c->pb.bin_list);
  c->pb.bin_list = nullptr;
  if (c->own_sums) cudaFree(c->d_sums);
  c->own_sums = !c->is_lane;
  for (auto& p : c->d_dn) {
    cudaFree(p);
    p = nullptr;
  }
  cudaFree(c->d_rgba);
  cudaFree(c->d_export);
  c->d_state = nullptr;
  c->pb.queue[0] = c->pb.queue[1] = nullptr;
  c->pb.flags = nullptr;
  c->pb.block_sums = nullptr;
  c->d_sums = nullptr;
  c->d_rgba = nullptr;
  c->d_export = nullptr;
  c->final_rgb = nullptr;
}

static int alloc_image_buffers(pt_ctx* c, uint32_t width, uint32_t height)
{
  if (width == 0 || height == 0 || (uint64_t)width * height > (1ull << 27))
    return fail(PT_ERR_INVALID, "unsupported resolution");
  c->width = width;
  c->height = height;
  c->row_begin = 0;
  c->row_end = height;
  c->pixels = width * height;
  uint32_t spp_pass = 1;
  if (c->params.rng_mode == PT_RNG_SLOT_RESEED) {
    spp_pass = 1; // slot indices are per iteration
  } else if (c->params.samples_per_pass > 0) {
    spp_pass = (uint32_t)c->params.samples_per_pass;
  } else {
    // Paths in flight per wavefront.  Big passes amortise the latency-bound tail bounces
    // (measured: 1080p, 64 spp: 4 -> 64 iterations per pass = +30 % Mrays/s); at 168 B per
    // path 2^27 paths are 22.5 GB of B200's 180 GB.  Never take more than a quarter of what is
    // free on the device.
    uint64_t target = 1ull << 27;
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
      const uint64_t by_mem = (uint64_t)(free_b / 4) / 168;
      target = std::max<uint64_t>(1ull << 21, std::min<uint64_t>(target, by_mem));
    }
    spp_pass = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(64, target / c->pixels));
  }
  c->samples_per_pass = spp_pass;
  // two lanes: each holds path state for half the frame's tile rows (rounded up)
  const uint32_t tile_rows = (height + 3) / 4;
  c->rows_cap = (c->lane2 || c->is_lane) ? std::min(height, ((tile_rows + 1) / 2) * 4) : height;
  // rounded up to whole 256-record tiles: the chain kernel stages the parked state tile by tile
  // with bulk copies of fixed size (and every plane stays 16-byte aligned)
  const size_t cap = (((size_t)width * c->rows_cap * spp_pass) + 255) & ~(size_t)255;
  if (cap > (1ull << 30)) return fail(PT_ERR_INVALID, "wavefront too large");
  c->pb.capacity = (uint32_t)cap;
  if (c->params.rng_mode == PT_RNG_SLOT_RESEED) {
    // bounce-synchronous kernels: path-id indexed planes + id queues
    PT_CUDA(cudaMalloc(&c->d_state, cap * (sizeof(float4) * 4 + sizeof(uint4))));
    float4* base = (float4*)c->d_state;
    c->pb.ps.ray = base + 0 * cap; // 2 float4 per path, interleaved
    c->pb.ps.thr = base + 2 * cap;
    c->pb.ps.gbuf = base + 3 * cap;
    c->pb.ps.aux = (uint4*)(base + 4 * cap);
    PT_CUDA(cudaMalloc((void**)&c->pb.queue[0], cap * 4));
    PT_CUDA(cudaMalloc((void**)&c->pb.queue[1], cap * 4));
    PT_CUDA(cudaMalloc((void**)&c->pb.tq, cap * 4));
    PT_CUDA(cudaMalloc((void**)&c->pb.flags, cap));
    PT_CUDA(cudaMalloc((void**)&c->pb.block_sums, ((cap + 2047) / 2048 + 1) * 4));
  } else {
    // chain scheduler: per-path result planes (contribution + first-hit G-buffer, 32 B) and two
    // compacted park buffers (68 B per slot each)
    PT_CUDA(cudaMalloc(&c->d_state, cap * (sizeof(float4) * 2 + 2 * (sizeof(float4) * 4 + sizeof(uint32_t)))));
    float4* base = (float4*)c->d_state;
    c->pb.ps.ray = nullptr;
    c->pb.ps.aux = nullptr;
    c->pb.ps.thr = base + 0 * cap;
    c->pb.ps.gbuf = base + 1 * cap;
    float4* p = base + 2 * cap;
    for (int k = 0; k < 2; ++k) {
      c->pb.park[k].ray = p;
      c->pb.park[k].thr = p + 2 * cap;
      c->pb.park[k].aux = (uint4*)(p + 3 * cap);
      p += 4 * cap;
    }
    uint32_t* ids = (uint32_t*)p;
    c->pb.park[0].pid = ids;
    c->pb.park[1].pid = ids + cap;
    if (c->params.sort_rays) PT_CUDA(cudaMalloc((void**)&c->pb.bin_list, cap * PT_BINS * sizeof(uint32_t)));
  }
  c->iteration = 0;
  c->final_rgb = nullptr;
  if (c->is_lane) return PT_OK; // frame buffers belong to the parent
  PT_CUDA(cudaMalloc((void**)&c->d_sums, (size_t)c->pixels * sizeof(float4) * 2));
  PT_CUDA(cudaMemsetAsync(c->d_sums, 0, (size_t)c->pixels * sizeof(float4) * 2, c->stream));
  PT_CUDA(cudaMalloc((void**)&c->d_rgba, (size_t)c->pixels * 4));
  PT_CUDA(cudaMalloc((void**)&c->d_export, (size_t)c->pixels * 3 * 4));
  if (c->lane2) {
    // same pass size as this context, half the rows, this context's sums
    c->lane2->params.samples_per_pass = (int32_t)c->samples_per_pass;
    const int rc = alloc_image_buffers(c->lane2, width, height);
    if (rc != PT_OK) return rc;
    c->lane2->d_sums = c->d_sums;
    c->lane2->own_sums = false;
  }
  return PT_OK;
}

// PT_LANES=1|2 (read once): 2 = every pixel-stream context renders its band as two half-bands on
// two streams (internal.h, pt_ctx::lane2).
// measured (profiles/README.md, round 2): two lanes +3.9 % on the headline frame, +3.0 % at 2.6 M
// triangles, +1.5 % on the 10 M-triangle terrain
#ifndef PT_DEFAULT_LANES
#define PT_DEFAULT_LANES 2
#endif
static int lanes_setting()
{
  static const int n = [] {
    const char* v = getenv("PT_LANES");
    const int k = v ? atoi(v) : PT_DEFAULT_LANES;
    return k == 2 ? 2 : 1;
  }();
  return n;
}

static int pt_ctx_create_core(const pt_scene* scene, uint32_t width, uint32_t height, const pt_params* params,
                              void* stream, bool as_lane, pt_ctx** out);

static int pt_ctx_create_impl(const pt_scene* scene, uint32_t width, uint32_t height, const pt_params* params,
                  void* stream, pt_ctx** out)
{
  return pt_ctx_create_core(scene, width, height, params, stream, false, out);
}

static int pt_ctx_create_core(const pt_scene* scene, uint32_t width, uint32_t height, const pt_params* params,
                              void* stream, bool as_lane, pt_ctx** out)
{
  if (!scene || !out) return fail(PT_ERR_INVALID, "pt_ctx_create: null argument");
  *out = nullptr;
  pt_params p;
  if (params)
    p = *params;
  else
    pt_params_default(&p);
  if (p.max_depth <= 0 || p.max_depth > 4096) return fail(PT_ERR_INVALID, "max_depth out of range");
  if (p.lanes < 0 || p.lanes > 2) return fail(PT_ERR_INVALID, "lanes must be 0 (default), 1 or 2");
  if (p.rng_mode != PT_RNG_PIXEL_STREAM && p.rng_mode != PT_RNG_SLOT_RESEED)
    return fail(PT_ERR_INVALID, "unknown rng_mode");
  PT_CUDA(cudaSetDevice(scene->device));
  pt_ctx* c = new pt_ctx();
  c->scene = scene;
  c->params = p;
  c->is_lane = as_lane;
  const int want_lanes = p.lanes ? p.lanes : lanes_setting();
  if (!as_lane && want_lanes == 2 && p.rng_mode == PT_RNG_PIXEL_STREAM && height >= 16 && !p.sort_rays) {
    // the second lane first: alloc_image_buffers below sizes both for half the rows
    pt_ctx* lane = nullptr;
    const int lrc = pt_ctx_create_core(scene, width, height, &p, nullptr, true, &lane);
    if (lrc != PT_OK) {
      delete c;
      return lrc;
    }
    c->lane2 = lane;
  }
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, scene->device);
  if (e != cudaSuccess) {
    pt_ctx_destroy(c); // (also the lane)
    return cuda_fail(e, "cudaGetDeviceProperties");
  }
  c->sms = prop.multiProcessorCount;
  if (stream) {
    c->stream = (cudaStream_t)stream;
    c->own_stream = false;
  } else {
    e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
      pt_ctx_destroy(c);
      return cuda_fail(e, "cudaStreamCreate");
    }
    c->own_stream = true;
  }
  const size_t n_ctr = (size_t)p.max_depth + 2;
  c->counters_bytes = n_ctr * (3 + PT_BINS) * sizeof(uint32_t); // + per-iteration bin counters (sort_rays)
  int rc = PT_OK;
  do {
    if ((e = cudaMalloc(&c->d_counters, c->counters_bytes + 24)) != cudaSuccess) break;
    c->pb.counters = (uint32_t*)c->d_counters;
    c->pb.work = c->pb.counters + n_ctr;
    c->pb.tcounters = c->pb.counters + 2 * n_ctr;
    c->pb.bin_counts = c->pb.counters + 3 * n_ctr;
    c->pb.total_rays = (unsigned long long*)((char*)c->d_counters + ((c->counters_bytes + 7) & ~7ull));
    if ((e = cudaMemset(c->d_counters, 0, c->counters_bytes + 24)) != cudaSuccess) break;
    if ((e = cudaHostAlloc((void**)&c->h_counts, n_ctr * sizeof(uint32_t), cudaHostAllocDefault)) !=
        cudaSuccess)
      break;
    c->bounce_events.resize(n_ctr);
    for (auto& ev : c->bounce_events)
      if ((e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) break;
  } while (0);
  if (e == cudaSuccess && c->lane2) {
    if ((e = cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming)) == cudaSuccess)
      e = cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming);
  }
  if (e != cudaSuccess) {
    rc = cuda_fail(e, "context allocation");
  } else if (!as_lane) {
    rc = alloc_image_buffers(c, width, height); // also the lane's
  }
  if (rc != PT_OK) {
    pt_ctx_destroy(c);
    return rc;
  }
  *out = c;
  return PT_OK;
}

int pt_ctx_create(const pt_scene* scene, uint32_t width, uint32_t height, const pt_params* params,
                  void* stream, pt_ctx** out)
{
  return guarded("pt_ctx_create", [&] { return pt_ctx_create_impl(scene, width, height, params, stream, out); });
}

int pt_ctx_destroy(pt_ctx* c)
{
  if (!c) return PT_OK;
  cudaSetDevice(c->scene->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->lane2 && c->lane2->stream) cudaStreamSynchronize(c->lane2->stream);
  free_image_buffers(c); // (the lane's too)
  if (c->lane2) {
    pt_ctx* lane = c->lane2;
    c->lane2 = nullptr;
    pt_ctx_destroy(lane);
  }
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  cudaFree(c->d_counters);
  cudaFreeHost(c->h_counts);
  for (auto ev : c->bounce_events)
    if (ev) cudaEventDestroy(ev);
  for (auto ev : c->prof_events) cudaEventDestroy(ev);
  for (auto ev : c->prof_pool) cudaEventDestroy(ev);
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return PT_OK;
}

int pt_ctx_resize(pt_ctx* c, uint32_t width, uint32_t height)
{
  if (!c) return fail(PT_ERR_INVALID, "pt_ctx_resize: null context");
  // validate BEFORE anything is freed: a refused size leaves the context as it was
  if (width == 0 || height == 0 || (uint64_t)width * height > (1ull << 27))
    return fail(PT_ERR_INVALID, "unsupported resolution");
  PT_CUDA(cudaSetDevice(c->scene->device));
  PT_CUDA(cudaStreamSynchronize(c->stream));
  if (c->lane2) PT_CUDA(cudaStreamSynchronize(c->lane2->stream));
  free_image_buffers(c);
  const int rc = alloc_image_buffers(c, width, height);
  if (rc != PT_OK) {
    // allocation failed half way: release what was allocated and refuse every later call on the
    // frame until a resize succeeds (no kernel is ever launched on a null buffer)
    const std::string why = pt_last_error();
    free_image_buffers(c);
    c->width = c->height = c->pixels = 0;
    c->row_begin = c->row_end = 0;
    c->usable = false;
    return fail(rc, why);
  }
  c->usable = true;
  return PT_OK;
}

// every entry point that touches the frame buffers starts here
#define PT_NEED_FRAME(c, what)                                                                     \
  do {                                                                                             \
    if (!(c)->usable) return fail(PT_ERR_INVALID, what ": the context has no frame buffers (a resize failed)"); \
  } while (0)

int pt_ctx_set_rows(pt_ctx* c, uint32_t row_begin, uint32_t row_end)
{
  if (!c) return fail(PT_ERR_INVALID, "pt_ctx_set_rows: null context");
  if (row_begin >= row_end || row_end > c->height) return fail(PT_ERR_INVALID, "row band outside the frame");
  if (row_begin % 4 != 0 || (row_end % 4 != 0 && row_end != c->height))
    return fail(PT_ERR_INVALID, "row bands start and end on multiples of 4 rows (8x4 primary-ray tiles)");
  c->row_begin = row_begin;
  c->row_end = row_end;
  return PT_OK;
}

int pt_denoise_halo_rows(const pt_denoise_params* dp, uint32_t* rows)
{
  if (!rows) return fail(PT_ERR_INVALID, "pt_denoise_halo_rows: null argument");
  pt_denoise_params d;
  if (dp)
    d = *dp;
  else
    pt_denoise_params_default(&d);
  uint32_t halo = 0;
  for (int step = 1; step <= d.filter_size; step *= 2) halo += 2u * (uint32_t)step;
  *rows = halo;
  return PT_OK;
}

int pt_ctx_restart(pt_ctx* c)
{
  if (!c) return fail(PT_ERR_INVALID, "pt_ctx_restart: null context");
  PT_NEED_FRAME(c, "pt_ctx_restart");
  // PathTracer::restart only zeroes the counter (final_gather overwrites at
  // iteration 0); with running sums the buffers are cleared instead.
  PT_CUDA(cudaMemsetAsync(c->d_sums, 0, (size_t)c->pixels * sizeof(float4) * 2, c->stream));
  c->iteration = 0;
  c->range_first = 0;
  c->range_contiguous = true;
  c->camera_locked = false;
  c->final_rgb = nullptr;
  return PT_OK;
}

int pt_ctx_iteration(const pt_ctx* c) { return c ? c->iteration : -1; }

int pt_ctx_set_max_iterations(pt_ctx* c, int n)
{
  if (!c) return fail(PT_ERR_INVALID, "null context");
  c->params.max_iterations = n;
  return PT_OK;
}

int pt_ctx_set_stream(pt_ctx* c, void* stream)
{
  if (!c) return fail(PT_ERR_INVALID, "null context");
  PT_CUDA(cudaStreamSynchronize(c->stream));
  if (c->own_stream) cudaStreamDestroy(c->stream);
  if (stream) {
    c->stream = (cudaStream_t)stream;
    c->own_stream = false;
  } else {
    PT_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->own_stream = true;
  }
  return PT_OK;
}

// profile=1 helpers: event pairs per launch, folded into stats at pass end
enum { TAG_EXT0 = 0, TAG_EXT, TAG_SHADE, TAG_COMPACT, TAG_ACC, TAG_DENOISE, TAG_RESOLVE };

static void prof_begin(pt_ctx* c, int tag)
{
  if (!c->params.profile) return;
  // events come from a per-context pool: no cudaEventCreate/Destroy on the launch path
  cudaEvent_t a = nullptr, b = nullptr;
  for (cudaEvent_t* e : {&a, &b}) {
    if (!c->prof_pool.empty()) {
      *e = c->prof_pool.back();
      c->prof_pool.pop_back();
    } else {
      cudaEventCreate(e);
    }
  }
  cudaEventRecord(a, c->stream);
  c->prof_events.push_back(a);
  c->prof_events.push_back(b);
  c->prof_tags.push_back(tag);
}
static void prof_end(pt_ctx* c)
{
  if (!c->params.profile) return;
  cudaEventRecord(c->prof_events.back(), c->stream);
}
static void prof_collect(pt_ctx* c)
{
  if (!c->params.profile || c->prof_tags.empty()) return;
  cudaStreamSynchronize(c->stream);
  for (size_t i = 0; i < c->prof_tags.size(); ++i) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c->prof_events[2 * i], c->prof_events[2 * i + 1]);
    switch (c->prof_tags[i]) {
    case TAG_EXT0: c->stats.ms_raygen_extend0 += ms; c->stats.n_extend_launches++; break;
    case TAG_EXT: c->stats.ms_extend += ms; c->stats.n_extend_launches++; break;
    case TAG_SHADE: c->stats.ms_shade += ms; c->stats.n_shade_launches++; break;
    case TAG_COMPACT: c->stats.ms_compact += ms; break;
    case TAG_ACC: c->stats.ms_accumulate += ms; break;
    case TAG_DENOISE: c->stats.ms_denoise += ms; break;
    case TAG_RESOLVE: c->stats.ms_resolve += ms; break;
    }
    c->prof_pool.push_back(c->prof_events[2 * i]);
    c->prof_pool.push_back(c->prof_events[2 * i + 1]);
  }
  c->prof_events.clear();
  c->prof_tags.clear();
}

// One wavefront pass over `samples` consecutive iterations of every pixel of rows
// [row_begin, row_end), on c's own stream and path-state buffers.
static int render_pass(pt_ctx* c, const pt_camera& cam, uint32_t first_iteration, uint32_t samples,
                       uint32_t row_begin, uint32_t row_end)
{
  const LaunchEnv env{c->stream, c->sms};
  PassParams pp{};
  pp.cam = make_dev_camera(cam, c->width, c->height);
  pp.pixels = c->pixels;
  pp.tiles_x = (c->width + 7) / 8;
  pp.tiles_y = (row_end - row_begin + 3) / 4;
  pp.tile_y0 = row_begin / 4;
  pp.pixel_begin = row_begin * c->width;
  pp.pixel_end = row_end * c->width;
  pp.band_pixels = pp.pixel_end - pp.pixel_begin;
  if ((uint64_t)samples * pp.band_pixels > c->pb.capacity)
    return fail(PT_ERR_INVALID, "render_pass: the band does not fit the path-state buffers");
  pp.fd_per_sample = make_fastdiv(pp.tiles_x * pp.tiles_y * 32u);
  pp.fd_tiles_x = make_fastdiv(pp.tiles_x);
  pp.fd_width = make_fastdiv(c->width);
  pp.samples = samples;
  pp.capacity = c->pb.capacity;
  pp.first_iteration = first_iteration;
  pp.rng_mode = (uint32_t)c->params.rng_mode;
  pp.stream_state = tunable_stream_state() ? 1u : 0u;
  pp.order = (uint32_t)tunable_order();
  pp.fd_samples = make_fastdiv(samples);
  pp.sbx = (pp.tiles_x + 7) / 8;
  pp.fd_sbx = make_fastdiv(pp.sbx);
  // bounce-0 work items: whole warps (8x4 tiles), order 2 pads the tile grid to 8x8-tile blocks
  const uint64_t n_tiles = pp.order == 2u ? (uint64_t)pp.sbx * ((pp.tiles_y + 7) / 8) * 64u
                                          : (uint64_t)pp.tiles_x * pp.tiles_y;
  const uint64_t n0_64 = (uint64_t)samples * n_tiles * 32u;
  if (n0_64 >= (1ull << 31)) return fail(PT_ERR_INVALID, "wavefront too large: lower samples_per_pass");
  const uint32_t n0 = (uint32_t)n0_64;
  const uint32_t max_depth = (uint32_t)c->params.max_depth;
  const bool stable = c->params.rng_mode == PT_RNG_SLOT_RESEED;
  if (stable && (row_begin != 0 || row_end != c->height))
    return fail(PT_ERR_INVALID, "row bands need PT_RNG_PIXEL_STREAM (slot re-seeding numbers the whole frame)");
  const uint32_t kLookBehind = 2;

  PT_CUDA(cudaMemsetAsync(c->d_counters, 0, c->counters_bytes, c->stream));
  int q = 0;
  uint32_t launched = 0;
  if (!stable) {
    // ---- PT_RNG_PIXEL_STREAM: asynchronous chains.  chain_kernel shades every path in
    // registers until its next ray needs the BVH; only those rays are queued, traversed and
    // handed back.  Iteration `it` traverses queue[it & 1] (length tcounters[it]).
    prof_begin(c, TAG_EXT0);
    launch_chain(env, c->scene->dev, c->pb, pp, 0, n0, max_depth);
    prof_end(c);
    launched += 1;
    // Where to hand the tail of the pass to finish_kernel: the first bounce >= PT_FINISH at which
    // the last pass whose counts are known had, scaled to this pass's size, at most PT_FINISH_RAYS
    // rays parked.  No history (first pass, or the previous one still in flight), ray binning or
    // the opt-in 8-wide tree: the whole pass runs as a wavefront.
    const uint64_t paths_now = (uint64_t)samples * pp.band_pixels;
    uint32_t finish_at = max_depth; // = never
    if (c->pend_event >= 0 && cudaEventQuery(c->bounce_events[c->pend_event]) == cudaSuccess) {
      c->hist_counts.assign(c->h_counts, c->h_counts + c->pend_known);
      c->hist_known = c->pend_known;
      c->hist_paths = c->pend_paths;
      c->pend_event = -1;
    }
    const uint32_t finish_first = (uint32_t)tunable_finish_after();
    if (finish_first != 0 && !c->pb.bin_list && c->scene->dev.n_nodes8 == 0u && c->hist_paths != 0) {
      for (uint32_t it = finish_first; it < c->hist_known && it < max_depth; ++it) {
        const double predicted = (double)c->hist_counts[it] * (double)paths_now / (double)c->hist_paths;
        if (predicted <= (double)tunable_finish_rays()) {
          finish_at = it;
          break;
        }
      }
    }
    c->pend_known = 0;
    c->pend_paths = paths_now;
    if (c->scene->dev.n_tris != 0) {
      for (uint32_t it = 0; it < max_depth; ++it) {
        PT_CUDA(cudaMemcpyAsync(&c->h_counts[it], c->pb.tcounters + it, sizeof(uint32_t),
                                cudaMemcpyDeviceToHost, c->stream));
        PT_CUDA(cudaEventRecord(c->bounce_events[it], c->stream));
        c->pend_known = it + 1;
        c->pend_event = (int)it;
        if (it == finish_at) {
          // few rays left: the paths still parked are finished in one launch (finish_kernel)
          prof_begin(c, TAG_SHADE);
          launch_finish(env, c->scene->dev, c->pb, it, max_depth);
          prof_end(c);
          launched += 1;
          c->stats.max_bounce_reached = std::max(c->stats.max_bounce_reached, max_depth);
          break;
        }
        if (it >= kLookBehind) {
          // look-behind early exit: stop enqueueing iterations once an older queue is known
          // to have been empty; never blocks the host.
          const uint32_t probe = it - kLookBehind;
          if (cudaEventQuery(c->bounce_events[probe]) == cudaSuccess && c->h_counts[probe] == 0) break;
        }
        prof_begin(c, TAG_EXT);
        launch_traverse_parked(env, c->scene->dev, c->pb, it);
        prof_end(c);
        prof_begin(c, TAG_SHADE);
        launch_chain(env, c->scene->dev, c->pb, pp, it + 1, n0, max_depth);
        prof_end(c);
        launched += 2;
        c->stats.max_bounce_reached = std::max(c->stats.max_bounce_reached, it + 1);
      }
    }
  }
  for (uint32_t b = 0; stable && b < max_depth; ++b) {
    if (b >= kLookBehind + 1) {
      // look-behind early exit: stop enqueueing bounces once an older bounce is
      // known to have produced no survivors; never blocks the host.
      const uint32_t probe = b - kLookBehind; // counters[probe] was copied after shade(probe-1)
      if (cudaEventQuery(c->bounce_events[probe]) == cudaSuccess && c->h_counts[probe] == 0) break;
    }
    const bool last = b + 1 == max_depth;
    if (b == 0) {
      prof_begin(c, TAG_EXT0);
      launch_raygen(env, c->scene->dev, c->pb, pp, n0);
      prof_end(c);
      launched += 1;
    }
    if (c->scene->dev.n_tris != 0) {
      prof_begin(c, TAG_EXT);
      launch_traverse(env, c->scene->dev, c->pb, c->pb.tq, b);
      prof_end(c);
      launched += 1;
    }
    prof_begin(c, TAG_SHADE);
    launch_shade(env, c->scene->dev, c->pb, pp, q, b, n0, last);
    prof_end(c);
    launched += 1;
    if (stable && !last) {
      prof_begin(c, TAG_COMPACT);
      launch_stable_compact(env, c->pb, pp, q, b);
      prof_end(c);
      launched += 3;
    }
    if (!last) {
      PT_CUDA(cudaMemcpyAsync(&c->h_counts[b + 1], c->pb.counters + b + 1, sizeof(uint32_t),
                              cudaMemcpyDeviceToHost, c->stream));
      PT_CUDA(cudaEventRecord(c->bounce_events[b + 1], c->stream));
    }
    q ^= 1;
    c->stats.max_bounce_reached = std::max(c->stats.max_bounce_reached, b + 1);
  }
  prof_begin(c, TAG_ACC);
  launch_accumulate(env, c->pb, pp, c->d_sums, c->d_sums + c->pixels, max_depth);
  prof_end(c);
  launched += 1;
  PT_CUDA(cudaGetLastError());
  c->stats.kernel_launches += launched;
  c->stats.passes += 1;
  c->stats.samples += (uint64_t)samples * (pp.pixel_end - pp.pixel_begin);
  return PT_OK;
}

// The pass of the context's band: one lane, or the band's two halves on two streams.  The second
// lane starts after everything already queued on the main stream (a restart's memset of the sums,
// a resolve still reading them) and the main stream continues after both halves have accumulated.
static int render_pass_lanes(pt_ctx* c, const pt_camera& cam, uint32_t first_iteration, uint32_t samples)
{
  pt_ctx* lane = c->lane2;
  const uint32_t t0 = c->row_begin / 4, t1 = (c->row_end + 3) / 4;
  const uint32_t mid = std::min(c->row_end, (t0 + (t1 - t0 + 1) / 2) * 4);
  if (!lane || mid >= c->row_end || mid <= c->row_begin)
    return render_pass(c, cam, first_iteration, samples, c->row_begin, c->row_end);
  PT_CUDA(cudaEventRecord(c->ev_fork, c->stream));
  PT_CUDA(cudaStreamWaitEvent(lane->stream, c->ev_fork, 0));
  int rc = render_pass(c, cam, first_iteration, samples, c->row_begin, mid);
  if (rc == PT_OK) rc = render_pass(lane, cam, first_iteration, samples, mid, c->row_end);
  PT_CUDA(cudaEventRecord(c->ev_join, lane->stream));
  PT_CUDA(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
  return rc;
}

int pt_render_range(pt_ctx* c, const pt_camera* cam, int first_iteration, int n_iterations)
{
  if (!c || !cam) return fail(PT_ERR_INVALID, "pt_render_range: null argument");
  if (first_iteration < 0 || n_iterations < 0) return fail(PT_ERR_INVALID, "negative iteration range");
  PT_NEED_FRAME(c, "pt_render_range");
  if (c->camera_locked && c->iteration > 0 && std::memcmp(cam, &c->last_camera, sizeof(pt_camera)) != 0)
    return fail(PT_ERR_INVALID, "pt_render: the camera differs from the camera of the loaded progressive state "
                                "(pt_ctx_restart first)");
  PT_CUDA(cudaSetDevice(c->scene->device));
  c->last_camera = *cam;
  c->have_camera = true;
  if (n_iterations > 0) {
    if (c->iteration == 0) {
      c->range_first = first_iteration;
      c->range_contiguous = true;
    } else if (first_iteration != c->range_first + c->iteration) {
      c->range_contiguous = false;
    }
  }
  int done = 0;
  while (done < n_iterations) {
    const uint32_t s = (uint32_t)std::min<int>(n_iterations - done, (int)c->samples_per_pass);
    int rc = render_pass_lanes(c, *cam, (uint32_t)(first_iteration + done), s);
    if (rc != PT_OK) return rc;
    done += (int)s;
  }
  c->iteration += n_iterations;
  c->final_rgb = nullptr; // path_trace_result_buffer_ = dev_color_buffer_
  return PT_OK;
}

int pt_render(pt_ctx* c, const pt_camera* cam, int n_iterations)
{
  if (!c) return fail(PT_ERR_INVALID, "pt_render: null context");
  int n = n_iterations;
  if (c->params.max_iterations > 0) n = std::min(n, std::max(0, c->params.max_iterations - c->iteration));
  if (!c->range_contiguous)
    return fail(PT_ERR_INVALID, "pt_render: the sums hold a non-contiguous set of iterations; continue with "
                                "pt_render_range or pt_ctx_restart");
  // continues after the iterations the sums already hold (a shard that started at iteration K,
  // or a resumed state of one, goes on at K + iteration(), never re-using seeds)
  return pt_render_range(c, cam, c->range_first + c->iteration, n);
}

int pt_path_trace(pt_ctx* c, const pt_camera* cam) { return pt_render(c, cam, 1); }

int pt_sync(pt_ctx* c)
{
  if (!c) return fail(PT_ERR_INVALID, "pt_sync: null context");
  PT_CUDA(cudaStreamSynchronize(c->stream));
  return PT_OK;
}

int pt_denoise(pt_ctx* c, const pt_denoise_params* dp_in)
{
  if (!c) return fail(PT_ERR_INVALID, "pt_denoise: null context");
  pt_denoise_params dp;
  if (dp_in)
    dp = *dp_in;
  else
    pt_denoise_params_default(&dp);
  if (dp.filter_size < 1) return fail(PT_ERR_INVALID, "filter_size must be >= 1");
  PT_NEED_FRAME(c, "pt_denoise");
  if (!c->have_camera) return fail(PT_ERR_INVALID, "pt_denoise before any path_trace / upload_frame");
  PT_CUDA(cudaSetDevice(c->scene->device));
  for (auto& p : c->d_dn)
    if (!p) PT_CUDA(cudaMalloc((void**)&p, (size_t)c->pixels * sizeof(float4)));
  const LaunchEnv env{c->stream, c->sms};
  const DevCamera cam = make_dev_camera(c->last_camera, c->width, c->height);
  // Row-band sharding: this context owns rows [row_begin, row_end) and holds valid sums for
  // `halo` more rows on either side (pt_denoise_halo_rows, exchanged by the caller).  Every
  // iteration runs over the widened window; whatever the missing rows beyond it spoil moves
  // inwards by 2*step per iteration and stops exactly at the band.
  uint32_t halo = 0;
  for (int step = 1; step <= dp.filter_size; step *= 2) halo += 2u * (uint32_t)step;
  const bool whole = c->row_begin == 0 && c->row_end == c->height;
  const uint32_t row_lo = whole ? 0u : (c->row_begin > halo ? c->row_begin - halo : 0u);
  const uint32_t row_hi = whole ? c->height : std::min(c->height, c->row_end + halo);
  prof_begin(c, TAG_DENOISE);
  launch_denoise_prepare(env, cam, c->d_sums, c->d_sums + c->pixels, c->d_dn[0], c->d_dn[1], c->d_dn[2],
                         row_lo, row_hi);
  DenoiseParams kp{dp.color_weight, dp.normal_weight, dp.position_weight, dp.clamp_fix};
  const float4* in = c->d_dn[0];
  float4* bufs[2] = {c->d_dn[3], c->d_dn[4]};
  int which = 0;
  float4* last = nullptr;
  uint32_t launches = 1;
  for (int step = 1; step <= dp.filter_size; step *= 2) {
    launch_atrous(env, cam, kp, in, c->d_dn[1], c->d_dn[2], bufs[which], step, row_lo, row_hi);
    last = bufs[which];
    in = last;
    which ^= 1;
    ++launches;
  }
  prof_end(c);
  PT_CUDA(cudaGetLastError());
  c->stats.kernel_launches += launches;
  c->final_rgb = last;
  return PT_OK;
}

int pt_resolve_rgba8(pt_ctx* c, int kind, void* dst, int dst_is_device)
{
  if (!c || !dst) return fail(PT_ERR_INVALID, "pt_resolve_rgba8: null argument");
  if (kind < 0 || kind > 4) return fail(PT_ERR_INVALID, "unknown buffer kind");
  if (kind == PT_BUF_DENOISED && !c->final_rgb) return fail(PT_ERR_INVALID, "no denoised buffer");
  PT_NEED_FRAME(c, "pt_resolve_rgba8");
  PT_CUDA(cudaSetDevice(c->scene->device));
  const LaunchEnv env{c->stream, c->sms};
  uchar4* target = dst_is_device ? (uchar4*)dst : c->d_rgba;
  prof_begin(c, TAG_RESOLVE);
  launch_resolve_rgba8(env, kind, c->d_sums, c->d_sums + c->pixels, c->final_rgb,
                       c->final_rgb != nullptr, c->pixels, target);
  prof_end(c);
  PT_CUDA(cudaGetLastError());
  c->stats.kernel_launches += 1;
  if (!dst_is_device) {
    PT_CUDA(cudaMemcpyAsync(dst, c->d_rgba, (size_t)c->pixels * 4, cudaMemcpyDeviceToHost, c->stream));
  }
  // send_to_preview ends with cudaDeviceSynchronize (path_tracer.cu:519)
  PT_CUDA(cudaStreamSynchronize(c->stream));
  prof_collect(c);
  return PT_OK;
}

int pt_download_f32(pt_ctx* c, int kind, float* dst)
{
  if (!c || !dst) return fail(PT_ERR_INVALID, "pt_download_f32: null argument");
  if (kind < 0 || kind > 4) return fail(PT_ERR_INVALID, "unknown buffer kind");
  if (kind == PT_BUF_DENOISED && !c->final_rgb) return fail(PT_ERR_INVALID, "no denoised buffer");
  PT_NEED_FRAME(c, "pt_download_f32");
  PT_CUDA(cudaSetDevice(c->scene->device));
  const LaunchEnv env{c->stream, c->sms};
  launch_export_f32(env, kind, c->d_sums, c->d_sums + c->pixels, c->final_rgb,
                    c->final_rgb != nullptr, c->pixels, c->d_export);
  PT_CUDA(cudaGetLastError());
  c->stats.kernel_launches += 1;
  const size_t n = (size_t)c->pixels * (kind == PT_BUF_DEPTH ? 1 : 3);
  PT_CUDA(cudaMemcpyAsync(dst, c->d_export, n * 4, cudaMemcpyDeviceToHost, c->stream));
  PT_CUDA(cudaStreamSynchronize(c->stream));
  return PT_OK;
}

int pt_ctx_sums(pt_ctx* c, void** device_ptr, uint64_t* n_floats)
{
  if (!c || !device_ptr || !n_floats) return fail(PT_ERR_INVALID, "pt_ctx_sums: null argument");
  *device_ptr = c->d_sums;
  *n_floats = (uint64_t)c->pixels * 8;
  return PT_OK;
}

int pt_ctx_bind_sums(pt_ctx* c, void* device_ptr)
{
  if (!c) return fail(PT_ERR_INVALID, "null context");
  PT_CUDA(cudaSetDevice(c->scene->device));
  PT_CUDA(cudaStreamSynchronize(c->stream));
  if (device_ptr) {
    if (c->own_sums) cudaFree(c->d_sums);
    c->d_sums = (float4*)device_ptr;
    c->own_sums = false;
    if (c->lane2) c->lane2->d_sums = c->d_sums;
  } else if (!c->own_sums) {
    c->d_sums = nullptr;
    PT_CUDA(cudaMalloc((void**)&c->d_sums, (size_t)c->pixels * sizeof(float4) * 2));
    c->own_sums = true;
    PT_CUDA(cudaMemsetAsync(c->d_sums, 0, (size_t)c->pixels * sizeof(float4) * 2, c->stream));
    if (c->lane2) c->lane2->d_sums = c->d_sums;
  }
  return PT_OK;
}

int pt_ctx_set_sample_count(pt_ctx* c, int n)
{
  if (!c) return fail(PT_ERR_INVALID, "null context");
  c->iteration = n;
  return PT_OK;
}

// ---- progressive state on disk: the running sums are associative, so a 1024-spp render can be
// stopped and resumed (or merged from shards) without changing the result.
namespace {
// version 2 (128 bytes).  Version 1 files (64-byte header: magic, version, width, height,
// iteration, n_floats) are still read, without the fingerprint checks.
struct StateHeader {
  char magic[8]; // "B200PTST"
  uint32_t version, width, height, iteration; // iteration = number of iterations in the sums
  uint64_t n_floats;
  // ---- version 2
  uint32_t range_first;      // the sums hold iterations [range_first, range_first + iteration)
  uint32_t range_contiguous; // 0: an arbitrary set (merged shards); pt_render will not continue it
  uint64_t scene_hash;       // description_hash of the scene the sums were rendered from
  int32_t max_depth, rng_mode;
  uint32_t have_camera;
  pt_camera camera;
  uint32_t reserved[9];
};
static_assert(sizeof(StateHeader) == 128, "state header is 128 bytes");
} // namespace

static int pt_ctx_save_state_impl(pt_ctx* c, const char* path)
{
  if (!c || !path) return fail(PT_ERR_INVALID, "pt_ctx_save_state: null argument");
  PT_NEED_FRAME(c, "pt_ctx_save_state");
  PT_CUDA(cudaSetDevice(c->scene->device));
  PT_CUDA(cudaStreamSynchronize(c->stream));
  const size_t n = (size_t)c->pixels * 8;
  std::vector<float> host(n);
  PT_CUDA(cudaMemcpy(host.data(), c->d_sums, n * sizeof(float), cudaMemcpyDeviceToHost));
  StateHeader h{};
  std::memcpy(h.magic, "B200PTST", 8);
  h.version = 2;
  h.width = c->width;
  h.height = c->height;
  h.iteration = (uint32_t)c->iteration;
  h.n_floats = n;
  h.range_first = (uint32_t)c->range_first;
  h.range_contiguous = c->range_contiguous ? 1u : 0u;
  h.scene_hash = c->scene->content_hash;
  h.max_depth = c->params.max_depth;
  h.rng_mode = c->params.rng_mode;
  h.have_camera = c->have_camera ? 1u : 0u;
  h.camera = c->last_camera;
  FILE* f = std::fopen(path, "wb");
  if (!f) return fail(PT_ERR_IO, std::string("cannot write ") + path);
  const bool ok = std::fwrite(&h, sizeof(h), 1, f) == 1 && std::fwrite(host.data(), sizeof(float), n, f) == n;
  std::fclose(f);
  if (!ok) return fail(PT_ERR_IO, std::string("short write to ") + path);
  return PT_OK;
}

int pt_ctx_save_state(pt_ctx* c, const char* path)
{
  return guarded("pt_ctx_save_state", [&] { return pt_ctx_save_state_impl(c, path); });
}

static int pt_ctx_load_state_impl(pt_ctx* c, const char* path)
{
  if (!c || !path) return fail(PT_ERR_INVALID, "pt_ctx_load_state: null argument");
  PT_NEED_FRAME(c, "pt_ctx_load_state");
  FILE* f = std::fopen(path, "rb");
  if (!f) return fail(PT_ERR_IO, std::string("cannot open ") + path);
  StateHeader h{};
  const size_t n = (size_t)c->pixels * 8;
  std::vector<float> host(n);
  bool ok = std::fread(&h, 64, 1, f) == 1; // the version-1 header is the first 64 bytes
  if (ok && (std::memcmp(h.magic, "B200PTST", 8) != 0 || (h.version != 1 && h.version != 2))) {
    std::fclose(f);
    return fail(PT_ERR_PARSE, std::string(path) + " is not a progressive-state file");
  }
  if (ok && h.version == 2) ok = std::fread((char*)&h + 64, sizeof(h) - 64, 1, f) == 1;
  if (ok && (h.width != c->width || h.height != c->height || h.n_floats != n)) {
    std::fclose(f);
    return fail(PT_ERR_INVALID, "progressive state has a different resolution");
  }
  if (ok && h.version == 2) {
    const char* why = nullptr;
    if (h.scene_hash != c->scene->content_hash) why = "progressive state was rendered from a different scene";
    else if (h.max_depth != c->params.max_depth) why = "progressive state was rendered with a different max_depth";
    else if (h.rng_mode != c->params.rng_mode) why = "progressive state was rendered with a different rng_mode";
    if (why) {
      std::fclose(f);
      return fail(PT_ERR_INVALID, why);
    }
  }
  ok = ok && std::fread(host.data(), sizeof(float), n, f) == n;
  std::fclose(f);
  if (!ok) return fail(PT_ERR_IO, std::string("short read from ") + path);
  PT_CUDA(cudaSetDevice(c->scene->device));
  PT_CUDA(cudaStreamSynchronize(c->stream));
  PT_CUDA(cudaMemcpy(c->d_sums, host.data(), n * sizeof(float), cudaMemcpyHostToDevice));
  c->iteration = (int)h.iteration;
  c->range_first = h.version == 2 ? (int)h.range_first : 0;
  c->range_contiguous = h.version == 2 ? h.range_contiguous != 0 : true;
  if (h.version == 2 && h.have_camera) {
    // rendering on with another camera would blend two different images
    c->last_camera = h.camera;
    c->have_camera = true;
    c->camera_locked = true;
  }
  c->final_rgb = nullptr;
  return PT_OK;
}

int pt_ctx_load_state(pt_ctx* c, const char* path)
{
  return guarded("pt_ctx_load_state", [&] { return pt_ctx_load_state_impl(c, path); });
}

int pt_ctx_upload_frame(pt_ctx* c, const float* color3, const float* normal3, const float* depth1,
                        const pt_camera* cam)
{
  if (!c || !color3 || !normal3 || !depth1 || !cam)
    return fail(PT_ERR_INVALID, "pt_ctx_upload_frame: null argument");
  PT_NEED_FRAME(c, "pt_ctx_upload_frame");
  PT_CUDA(cudaSetDevice(c->scene->device));
  float* tmp = nullptr;
  const size_t n = c->pixels;
  PT_CUDA(cudaMalloc((void**)&tmp, n * 7 * sizeof(float)));
  cudaError_t e = cudaMemcpyAsync(tmp, color3, n * 12, cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(tmp + 3 * n, normal3, n * 12, cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(tmp + 6 * n, depth1, n * 4, cudaMemcpyHostToDevice, c->stream);
  if (e == cudaSuccess) {
    const LaunchEnv env{c->stream, c->sms};
    launch_import_frame(env, tmp, tmp + 3 * n, tmp + 6 * n, c->pixels, c->d_sums, c->d_sums + c->pixels);
    e = cudaStreamSynchronize(c->stream);
  }
  cudaFree(tmp);
  if (e != cudaSuccess) return cuda_fail(e, "upload_frame");
  c->stats.kernel_launches += 1;
  c->last_camera = *cam;
  c->have_camera = true;
  c->iteration = 1;
  c->final_rgb = nullptr;
  return PT_OK;
}

// rays / launches / kernel times of one context (its own stream synchronised)
static int collect_stats(pt_ctx* c)
{
  PT_CUDA(cudaStreamSynchronize(c->stream));
  prof_collect(c);
  unsigned long long rays[2] = {0, 0};
  PT_CUDA(cudaMemcpy(rays, c->pb.total_rays, sizeof(rays), cudaMemcpyDeviceToHost));
  c->stats.rays = rays[0];
  c->stats.rays_traversed = rays[1];
  return PT_OK;
}

int pt_get_stats(pt_ctx* c, pt_stats* out)
{
  if (!c || !out) return fail(PT_ERR_INVALID, "pt_get_stats: null argument");
  PT_CUDA(cudaSetDevice(c->scene->device));
  int rc = collect_stats(c);
  if (rc != PT_OK) return rc;
  c->stats.iterations = (uint32_t)c->iteration;
  *out = c->stats;
  if (c->lane2) { // the second lane's share of every additive counter
    rc = collect_stats(c->lane2);
    if (rc != PT_OK) return rc;
    const pt_stats& l = c->lane2->stats;
    out->rays += l.rays;
    out->rays_traversed += l.rays_traversed;
    out->samples += l.samples;
    out->passes += l.passes;
    out->kernel_launches += l.kernel_launches;
    out->ms_raygen_extend0 += l.ms_raygen_extend0;
    out->ms_extend += l.ms_extend;
    out->ms_shade += l.ms_shade;
    out->ms_compact += l.ms_compact;
    out->ms_accumulate += l.ms_accumulate;
    out->n_extend_launches += l.n_extend_launches;
    out->n_shade_launches += l.n_shade_launches;
    out->max_bounce_reached = std::max(out->max_bounce_reached, l.max_bounce_reached);
  }
  return PT_OK;
}

int pt_reset_stats(pt_ctx* c)
{
  if (!c) return fail(PT_ERR_INVALID, "null context");
  PT_CUDA(cudaSetDevice(c->scene->device));
  for (pt_ctx* k : {c, c->lane2}) {
    if (!k) continue;
    PT_CUDA(cudaStreamSynchronize(k->stream));
    PT_CUDA(cudaMemset(k->pb.total_rays, 0, 2 * sizeof(unsigned long long)));
    prof_collect(k);
    k->stats = pt_stats{};
  }
  return PT_OK;
}

// -------------------------------------------------------------- trace batch
int pt_trace_batch(const pt_scene* sc, const float* rays8, uint64_t n, pt_hit* hits_out)
{
  if (!sc || (!rays8 && n) || (!hits_out && n)) return fail(PT_ERR_INVALID, "pt_trace_batch: null argument");
  if (n == 0) return PT_OK;
  if (n > (1ull << 30)) return fail(PT_ERR_INVALID, "ray batch too large");
  static_assert(sizeof(HitRecord) == sizeof(pt_hit), "hit record layout");
  PT_CUDA(cudaSetDevice(sc->device));
  float4* d_rays = nullptr;
  HitRecord* d_hits = nullptr;
  uint32_t* d_work = nullptr;
  cudaError_t e = cudaMalloc((void**)&d_rays, n * 32);
  if (e == cudaSuccess) e = cudaMalloc((void**)&d_hits, n * sizeof(HitRecord));
  if (e == cudaSuccess) e = cudaMalloc((void**)&d_work, sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMemset(d_work, 0, sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMemcpy(d_rays, rays8, n * 32, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, sc->device);
    const LaunchEnv env{nullptr, sms};
    launch_trace_batch(env, sc->dev, d_rays, d_work, (uint32_t)n, d_hits);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(hits_out, d_hits, n * sizeof(HitRecord), cudaMemcpyDeviceToHost);
  cudaFree(d_rays);
  cudaFree(d_hits);
  cudaFree(d_work);
  if (e != cudaSuccess) return cuda_fail(e, "pt_trace_batch");
  return PT_OK;
}

} // extern "C"
