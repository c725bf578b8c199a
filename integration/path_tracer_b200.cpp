// path_tracer_b200.cpp — the reference's `PathTracer` (src/lib/path_tracer.hpp:60-99) implemented
// on the C ABI of include/b200pt.h.  This is the translation unit a maintainer of
// LesleyLai/cuda-path-tracer drops in place of src/lib/path_tracer.cu (and ray_gen.cu,
// denoising/*.cu, accelerators/bvh.cpp): the class's public interface, and therefore every caller
// (src/cli/cli.cpp:86-105, src/interactive-app/app.cpp:130-157), stays as it is.
//
// It is compiled against the UNMODIFIED reference headers (integration/build.sh), so:
//   * the class's private reference buffers (Scene, Paths, cuda::Buffer members) stay in the
//     layout but remain empty; our two opaque handles live in a side table keyed by the object
//     (a maintainer would replace the private section by `pt_scene* scene_; pt_ctx* ctx_;`);
//   * SceneDescription's four containers are private without accessors (a `class` with its data
//     members first); `flatten()` reads them through the one-line change a maintainer would make
//     (`friend void flatten(...)`), spelled here as `class` -> `struct` around that one include,
//     after everything it includes has been seen.
// Nothing here touches CUDA directly: all device work happens behind libb200pt.so.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <unordered_map>
#include <variant>
#include <vector>

#include <functional>
#include <optional>

#include "camera.hpp"
#include "material.hpp"
#include "mesh.hpp"
#include "prelude.hpp"
#include "scene.hpp"
#include "sphere.hpp"
#include "transform.hpp"
#define class struct
#include "scene_description.hpp"
#undef class
#include "path_tracer.hpp"

#include "../include/b200pt.h"

namespace {

struct Handles {
  pt_scene* scene = nullptr;
  pt_ctx* ctx = nullptr;
  GPUMethod method = GPUMethod::streaming;
  int max_depth = 50; // static constexpr max_bounces (path_tracer.cu:27)
};
std::mutex g_mutex;
std::unordered_map<const PathTracer*, Handles>& table()
{
  static std::unordered_map<const PathTracer*, Handles> t;
  return t;
}
Handles& handles(const PathTracer* self)
{
  std::lock_guard<std::mutex> lock(g_mutex);
  return table()[self];
}

void check(int rc)
{
  if (rc != PT_OK) panic(pt_last_error()); // prelude.hpp: print and exit, like CUDA_CHECK
}

pt_camera to_pt(const Camera& c) // camera.hpp:17-23
{
  return pt_camera{{c.position.x, c.position.y, c.position.z},
                   {c.rotation.w, c.rotation.x, c.rotation.y, c.rotation.z},
                   c.vfov};
}

void put_mat(const glm::mat4& m, float* out) // glm::mat4 memory order = column-major
{
  for (int c = 0; c < 4; ++c)
    for (int r = 0; r < 4; ++r) out[c * 4 + r] = m[c][r];
}

// SceneDescription::build_scene() (scene_description.cpp:12-117) as a flattening into the plain
// arrays of pt_scene_desc.  Same semantics: the material table is in std::map (alphabetical)
// order, an object's material is looked up by name, only the FIRST mesh of the map is uploaded and
// every mesh object instances it (scene_description.cpp:95), spheres are numbered in object order.
struct FlatScene {
  std::vector<float> positions;
  std::vector<uint32_t> indices;
  std::vector<pt_object> objects;
  std::vector<pt_sphere> spheres;
  std::vector<pt_material> materials;
  pt_scene_desc desc{};
};

void flatten(const SceneDescription& s, FlatScene& f)
{
  std::map<std::string, uint32_t, std::less<>> material_index;
  for (const auto& [name, m] : s.material_map_) {
    pt_material pm{};
    switch (m.type) {
    case Material::Type::Diffuse:
      pm.type = PT_MAT_DIFFUSE;
      pm.albedo[0] = m.data.diffuse.albedo.x, pm.albedo[1] = m.data.diffuse.albedo.y, pm.albedo[2] = m.data.diffuse.albedo.z;
      break;
    case Material::Type::Metal:
      pm.type = PT_MAT_METAL;
      pm.albedo[0] = m.data.metal.albedo.x, pm.albedo[1] = m.data.metal.albedo.y, pm.albedo[2] = m.data.metal.albedo.z;
      pm.fuzz = m.data.metal.fuzz;
      break;
    case Material::Type::Dielectric:
      pm.type = PT_MAT_DIELECTRIC;
      pm.refraction_index = m.data.dielectric.refraction_index;
      break;
    }
    material_index.insert({name, (uint32_t)f.materials.size()});
    f.materials.push_back(pm);
  }
  for (size_t i = 0; i < s.objects_.size(); ++i) {
    const Object& o = s.objects_[i];
    pt_object po{};
    const auto it = material_index.find(s.objects_material_mapping_[i]);
    if (it == material_index.end()) panic("Cannot find material " + s.objects_material_mapping_[i]);
    po.material = it->second;
    put_mat(o.transform.m(), po.m);
    put_mat(o.transform.inverse_m(), po.inv);
    if (const Sphere* sp = std::get_if<Sphere>(&o.shape)) {
      po.type = PT_OBJ_SPHERE;
      po.prim_index = (uint32_t)f.spheres.size();
      f.spheres.push_back(pt_sphere{{sp->center.x, sp->center.y, sp->center.z}, sp->radius});
    } else {
      po.type = PT_OBJ_MESH;
      po.prim_index = 0;
    }
    f.objects.push_back(po);
  }
  if (!s.mesh_map_.empty()) {
    const Mesh& mesh = s.mesh_map_.begin()->second;
    f.positions.resize(mesh.positions.size() * 3);
    for (size_t i = 0; i < mesh.positions.size(); ++i) {
      f.positions[3 * i + 0] = mesh.positions[i].x;
      f.positions[3 * i + 1] = mesh.positions[i].y;
      f.positions[3 * i + 2] = mesh.positions[i].z;
    }
    f.indices = mesh.indices;
  }
  f.desc.positions = f.positions.data();
  f.desc.n_vertices = f.positions.size() / 3;
  f.desc.indices = f.indices.data();
  f.desc.n_indices = f.indices.size();
  f.desc.objects = f.objects.data();
  f.desc.n_objects = (uint32_t)f.objects.size();
  f.desc.spheres = f.spheres.data();
  f.desc.n_spheres = (uint32_t)f.spheres.size();
  f.desc.materials = f.materials.data();
  f.desc.n_materials = (uint32_t)f.materials.size();
}

void release(Handles& h)
{
  if (h.ctx) pt_ctx_destroy(h.ctx);
  if (h.scene) pt_scene_destroy(h.scene);
  h.ctx = nullptr;
  h.scene = nullptr;
}

} // namespace

PathTracer::PathTracer() = default;

// path_tracer.cu:559-564
void PathTracer::create_buffers(UResolution resolution, const SceneDescription& scene)
{
  Handles& h = handles(this);
  release(h);
  FlatScene flat;
  flatten(scene, flat);
  check(pt_scene_create(&flat.desc, /*device=*/0, &h.scene)); // cli.cpp:71: device 0
  pt_params p;
  pt_params_default(&p); // max_depth 50 == max_bounces
  p.max_depth = h.max_depth;
  // the reference reads current_gpu_method on every path_trace call; our context fixes the RNG
  // discipline at creation, so a later change re-creates it (path_trace below)
  p.rng_mode = current_gpu_method == GPUMethod::streaming ? PT_RNG_SLOT_RESEED : PT_RNG_PIXEL_STREAM;
  h.method = current_gpu_method;
  check(pt_ctx_create(h.scene, resolution.width, resolution.height, &p, /*stream=*/nullptr, &h.ctx));
  iteration_ = 0;
}

// path_tracer.cu:527-545
void PathTracer::resize_image(UResolution resolution)
{
  Handles& h = handles(this);
  check(pt_ctx_resize(h.ctx, resolution.width, resolution.height));
  iteration_ = 0;
}

// path_tracer.cu:522-525
void PathTracer::restart()
{
  check(pt_ctx_restart(handles(this).ctx));
  iteration_ = 0;
}

// path_tracer.cu:389-477: one sample per pixel per call, no-op once iteration() >= max_iterations
void PathTracer::path_trace(const Camera& camera, UResolution resolution)
{
  Handles& h = handles(this);
  if (h.method != current_gpu_method) { // the GUI's method toggle (gui.cpp:84-108)
    pt_params p;
    pt_params_default(&p);
    p.max_depth = h.max_depth;
    p.rng_mode = current_gpu_method == GPUMethod::streaming ? PT_RNG_SLOT_RESEED : PT_RNG_PIXEL_STREAM;
    pt_ctx* fresh = nullptr;
    check(pt_ctx_create(h.scene, resolution.width, resolution.height, &p, nullptr, &fresh));
    pt_ctx_destroy(h.ctx);
    h.ctx = fresh;
    h.method = current_gpu_method;
    iteration_ = 0;
  }
  check(pt_ctx_set_max_iterations(h.ctx, max_iterations));
  const pt_camera c = to_pt(camera);
  check(pt_path_trace(h.ctx, &c));
  iteration_ = pt_ctx_iteration(h.ctx);
}

// path_tracer.cu:479-485
void PathTracer::denoise(UResolution)
{
  pt_denoise_params d;
  pt_denoise_params_default(&d);
  d.filter_size = atrous_denoiser.filter_size;
  d.color_weight = atrous_denoiser.color_weight;
  d.normal_weight = atrous_denoiser.normal_weight;
  d.position_weight = atrous_denoiser.position_weight;
  check(pt_denoise(handles(this).ctx, &d));
}

// path_tracer.cu:487-520: tonemap into the caller's device (or managed) buffer, then synchronise
void PathTracer::send_to_preview(uchar4* dev_pbo, UResolution, DisplayBufferType type) const
{
  check(pt_resolve_rgba8(handles(this).ctx, static_cast<int>(type), dev_pbo, /*dst_is_device=*/1));
}

// Not part of the reference class: lets the test driver pick the reference's compile-time bounce
// limit (`max_bounces`, path_tracer.cu:27) and release the handles (the reference's destructor is
// implicit; a maintainer's version would own the handles as members).
extern "C" void b200_shim_set_max_depth(PathTracer* self, int max_depth) { handles(self).max_depth = max_depth; }
extern "C" void b200_shim_release(PathTracer* self)
{
  Handles& h = handles(self);
  release(h);
  std::lock_guard<std::mutex> lock(g_mutex);
  table().erase(self);
}
