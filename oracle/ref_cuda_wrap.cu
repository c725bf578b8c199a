// ref_cuda_wrap.cu — C-ABI wrapper around the REFERENCE's own CUDA path tracer.
// TEST INFRASTRUCTURE / BASELINE ONLY (built into oracle/_ref/libref_cuda.so by build_ref.sh).
//
// This translation unit #includes the reference's sources where they lie under
// /root/reference/src/lib (unity build: one TU; -rdc=true is still passed because the reference's
// `inline __constant__` camera, constant_memory.cuh:6, does not compile in whole-program mode and
// separable compilation is the reference's own setting, src/lib/CMakeLists.txt:70).  No reference source is copied
// into the repository.  path_tracer.cu alone is included from a build-time patched copy in a
// temporary directory (REF_PATCHED_DIR) with three sed edits, listed in build_ref.sh:
//   1. the streaming loop bound  `i < max_bounces`  reads the run-time g_ref_max_bounces
//      (reference: static constexpr 50, path_tracer.cu:27) and the megakernel loop reads the
//      __constant__ c_ref_max_bounces;
//   2. the traversal stack is StaticStack<unsigned, REF_STACK_SIZE> (reference: 24, unchecked);
//   3. a host-side ray counter  g_ref_ray_count += paths_count  before intersection_kernel.
// glm / fmt / spdlog come from oracle/ref_shim (absent third-party dependencies).
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <optional>
#include <queue>
#include <span>
#include <string>
#include <string_view>
#include <tuple>
#include <variant>
#include <vector>

#include <chrono>

#include <cuda.h>
#include <cuda_runtime_api.h>
#include <device_launch_parameters.h>
// NOTE: thrust/libcu++ headers must not be included before the reference's cuda_utils
// (they declare namespace cuda::std, which hijacks `std::` inside the reference's namespace cuda)

#include <fmt/format.h>
#include <glm/glm.hpp>
#include <spdlog/spdlog.h>

#include "../include/b200pt.h"

#ifndef REF_STACK_SIZE
#define REF_STACK_SIZE 64
#endif

static int g_ref_max_bounces = 50;
static unsigned long long g_ref_ray_count = 0;
__constant__ int c_ref_max_bounces = 50;

#define private public
#include "prelude.cpp"                       // /root/reference/src/lib/prelude.cpp
#include "cuda_utils/cuda_check.cpp"
#include "camera.cpp"
#include "accelerators/bvh.cpp"
#include "scene_description.cpp"
#include REF_PATCHED_PATH_TRACER            // patched copy of src/lib/path_tracer.cu
#include "ray_gen.cu"
#include "denoising/edge_avoiding_a_trous_denoiser.cu"
#undef private

namespace {

struct RefTracer {
  SceneDescription desc;
  PathTracer tracer;
  UResolution res{};
  cudaEvent_t e0{}, e1{};
};

glm::mat4 mat_from(const float* m)
{
  glm::mat4 r;
  for (int c = 0; c < 4; ++c) r[c] = glm::vec4(m[c * 4 + 0], m[c * 4 + 1], m[c * 4 + 2], m[c * 4 + 3]);
  return r;
}

Camera camera_from(const pt_camera* c)
{
  Camera cam;
  cam.position = glm::vec3(c->position[0], c->position[1], c->position[2]);
  cam.rotation = glm::quat(c->rotation[0], c->rotation[1], c->rotation[2], c->rotation[3]);
  cam.vfov = c->vfov;
  return cam;
}

std::string mat_name(uint32_t i)
{
  char buf[32];
  std::snprintf(buf, sizeof(buf), "m%06u", i); // std::map order == index order
  return buf;
}

__global__ void ref_trace_batch_kernel(const Ray* rays, unsigned n, AggregateView aggregate, pt_hit* out)
{
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Intersection rec;
  const bool hit = ray_scene_intersection_test(rays[i], aggregate, rec);
  pt_hit h;
  memset(&h, 0, sizeof(h));
  if (hit) {
    h.t = rec.t;
    h.point[0] = rec.point.x, h.point[1] = rec.point.y, h.point[2] = rec.point.z;
    h.normal[0] = rec.normal.x, h.normal[1] = rec.normal.y, h.normal[2] = rec.normal.z;
    h.material = static_cast<uint32_t>(rec.material_id);
    h.side = rec.side == HitFaceSide::front ? 0u : 1u;
    h.object = -2; // the reference's Intersection carries no object / primitive id
    h.prim = -2;
  } else {
    h.t = -1.0f;
    h.object = -1;
    h.prim = -1;
  }
  out[i] = h;
}

} // namespace

extern "C" {

__attribute__((visibility("default"))) void* ref_tracer_create(const pt_scene_desc* d, uint32_t w, uint32_t h,
                                                                int max_bounces, int megakernel)
{
  auto* t = new RefTracer();
  for (uint32_t i = 0; i < d->n_materials; ++i) {
    const pt_material& m = d->materials[i];
    const glm::vec3 albedo(m.albedo[0], m.albedo[1], m.albedo[2]);
    if (m.type == PT_MAT_DIFFUSE) t->desc.add_material(mat_name(i), Material{DiffuseMateral{albedo}});
    else if (m.type == PT_MAT_METAL) t->desc.add_material(mat_name(i), Material{MetalMaterial{albedo, m.fuzz}});
    else t->desc.add_material(mat_name(i), Material{DielectricMaterial{m.refraction_index}});
  }
  Mesh mesh;
  if (d->n_indices >= 3) {
    AABB box;
    for (uint64_t i = 0; i < d->n_vertices; ++i) {
      const glm::vec3 p(d->positions[3 * i], d->positions[3 * i + 1], d->positions[3 * i + 2]);
      mesh.positions.push_back(p);
      box = box.enclose(p); // Assimp aiProcess_GenBoundingBoxes
    }
    mesh.indices.assign(d->indices, d->indices + d->n_indices);
    mesh.aabb = box;
  } else {
    // sphere-only scene: build_scene() panics on an empty mesh (scene_description.cpp:95-100 ->
    // bvh.cpp:200-201), so an unreferenced one-triangle mesh keeps the unmodified builder happy
    mesh.positions = {glm::vec3(0, -1e6f, 0), glm::vec3(1e-3f, -1e6f, 0), glm::vec3(0, -1e6f, 1e-3f)};
    mesh.indices = {0, 1, 2};
    mesh.aabb = AABB{}.enclose(mesh.positions[0]).enclose(mesh.positions[1]).enclose(mesh.positions[2]);
  }
  const MeshRef mesh_ref = t->desc.add_mesh("mesh", std::move(mesh));
  for (uint32_t i = 0; i < d->n_objects; ++i) {
    const pt_object& o = d->objects[i];
    const Transform tf(mat_from(o.m), mat_from(o.inv));
    if (o.type == PT_OBJ_SPHERE) {
      const pt_sphere& s = d->spheres[o.prim_index];
      t->desc.add_object(Sphere{glm::vec3(s.center[0], s.center[1], s.center[2]), s.radius}, tf, mat_name(o.material));
    } else {
      t->desc.add_object(mesh_ref, tf, mat_name(o.material));
    }
  }
  t->res = UResolution{w, h};
  t->tracer.current_gpu_method = megakernel ? GPUMethod::megakernel : GPUMethod::streaming;
  t->tracer.max_iterations = 1 << 30;
  g_ref_max_bounces = max_bounces;
  cudaMemcpyToSymbol(c_ref_max_bounces, &max_bounces, sizeof(int));
  t->tracer.create_buffers(t->res, t->desc);
  cudaEventCreate(&t->e0);
  cudaEventCreate(&t->e1);
  cudaDeviceSynchronize();
  return t;
}

__attribute__((visibility("default"))) void ref_tracer_destroy(void* p)
{
  auto* t = static_cast<RefTracer*>(p);
  if (!t) return;
  cudaDeviceSynchronize();
  cudaEventDestroy(t->e0);
  cudaEventDestroy(t->e1);
  delete t;
}

__attribute__((visibility("default"))) void ref_tracer_restart(void* p) { static_cast<RefTracer*>(p)->tracer.restart(); }

// `for i < spp: path_trace` + cudaDeviceSynchronize (cli.cpp:96-100). Returns elapsed GPU ms.
__attribute__((visibility("default"))) float ref_tracer_render(void* p, const pt_camera* c, int n_iterations,
                                                               int max_bounces, unsigned long long* rays_out)
{
  auto* t = static_cast<RefTracer*>(p);
  const Camera cam = camera_from(c);
  g_ref_max_bounces = max_bounces;
  cudaMemcpyToSymbol(c_ref_max_bounces, &max_bounces, sizeof(int));
  g_ref_ray_count = 0;
  cudaEventRecord(t->e0, 0);
  for (int i = 0; i < n_iterations; ++i) t->tracer.path_trace(cam, t->res);
  cudaEventRecord(t->e1, 0);
  cudaDeviceSynchronize();
  float ms = 0.f;
  cudaEventElapsedTime(&ms, t->e0, t->e1);
  if (rays_out) *rays_out = g_ref_ray_count;
  return ms;
}

// kind: 1 colour, 2 normal, 3 depth, 0/4 path_trace_result_buffer_ (final / denoised)
__attribute__((visibility("default"))) int ref_tracer_download(void* p, int kind, float* out)
{
  auto* t = static_cast<RefTracer*>(p);
  const size_t n = static_cast<size_t>(t->res.width) * t->res.height;
  const void* src = nullptr;
  size_t bytes = n * 12;
  switch (kind) {
  case 1: src = t->tracer.dev_color_buffer_.data(); break;
  case 2: src = t->tracer.dev_normal_buffer_.data(); break;
  case 3: src = t->tracer.dev_depth_buffer_.data(); bytes = n * 4; break;
  default: src = t->tracer.path_trace_result_buffer_; break;
  }
  if (!src) return 1;
  return cudaMemcpy(out, src, bytes, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : 2;
}

__attribute__((visibility("default"))) int ref_tracer_upload_frame(void* p, const float* color3, const float* normal3,
                                                                   const float* depth1, const pt_camera* c)
{
  auto* t = static_cast<RefTracer*>(p);
  const size_t n = static_cast<size_t>(t->res.width) * t->res.height;
  cudaMemcpy(t->tracer.dev_color_buffer_.data(), color3, n * 12, cudaMemcpyHostToDevice);
  cudaMemcpy(t->tracer.dev_normal_buffer_.data(), normal3, n * 12, cudaMemcpyHostToDevice);
  cudaMemcpy(t->tracer.dev_depth_buffer_.data(), depth1, n * 4, cudaMemcpyHostToDevice);
  // the denoiser reads the camera the last path_trace left in __constant__ memory
  const auto gpu_camera = camera_from(c).to_gpu_camera(t->res);
  cudaMemcpyToSymbol(constant_memory::gpu_camera, &gpu_camera, sizeof(GPUCamera));
  return cudaDeviceSynchronize() == cudaSuccess ? 0 : 2;
}

// PathTracer::denoise (path_tracer.cu:479-485). Returns elapsed GPU ms.
__attribute__((visibility("default"))) float ref_tracer_denoise(void* p, int filter_size, float cw, float nw, float pw)
{
  auto* t = static_cast<RefTracer*>(p);
  t->tracer.atrous_denoiser.filter_size = filter_size;
  t->tracer.atrous_denoiser.color_weight = cw;
  t->tracer.atrous_denoiser.normal_weight = nw;
  t->tracer.atrous_denoiser.position_weight = pw;
  cudaEventRecord(t->e0, 0);
  t->tracer.denoise(t->res);
  cudaEventRecord(t->e1, 0);
  cudaDeviceSynchronize();
  float ms = 0.f;
  cudaEventElapsedTime(&ms, t->e0, t->e1);
  return ms;
}

// PathTracer::send_to_preview into a host RGBA8 image
__attribute__((visibility("default"))) int ref_tracer_preview(void* p, int kind, void* rgba_host)
{
  auto* t = static_cast<RefTracer*>(p);
  const size_t n = static_cast<size_t>(t->res.width) * t->res.height;
  uchar4* dev = nullptr;
  if (cudaMalloc(reinterpret_cast<void**>(&dev), n * 4) != cudaSuccess) return 2;
  t->tracer.send_to_preview(dev, t->res, static_cast<DisplayBufferType>(kind));
  const cudaError_t e = cudaMemcpy(rgba_host, dev, n * 4, cudaMemcpyDeviceToHost);
  cudaFree(dev);
  return e == cudaSuccess ? 0 : 2;
}

// ray_scene_intersection_test on a ray batch (the reference's intersection_kernel body)
__attribute__((visibility("default"))) int ref_trace_batch(void* p, const float* rays8, uint64_t n, pt_hit* out)
{
  auto* t = static_cast<RefTracer*>(p);
  if (n == 0) return 0;
  Ray* d_rays = nullptr;
  pt_hit* d_out = nullptr;
  static_assert(sizeof(Ray) == 32, "Ray layout");
  cudaMalloc(reinterpret_cast<void**>(&d_rays), n * sizeof(Ray));
  cudaMalloc(reinterpret_cast<void**>(&d_out), n * sizeof(pt_hit));
  cudaMemcpy(d_rays, rays8, n * sizeof(Ray), cudaMemcpyHostToDevice);
  const AggregateView view{t->tracer.dev_scene_.aggregate};
  ref_trace_batch_kernel<<<static_cast<unsigned>((n + 63) / 64), 64>>>(d_rays, static_cast<unsigned>(n), view, d_out);
  const cudaError_t e = cudaMemcpy(out, d_out, n * sizeof(pt_hit), cudaMemcpyDeviceToHost);
  cudaFree(d_rays);
  cudaFree(d_out);
  return e == cudaSuccess ? 0 : 2;
}

// Depth of the reference's own BVH for a mesh (root = depth 1).  ray_mesh_intersection_test pushes
// both children of every inner node it enters onto StaticStack<unsigned, 24> (path_tracer.cu:46-73)
// and never checks for overflow, so the stock stack is correct exactly when depth <= 23.
__attribute__((visibility("default"))) int ref_bvh_depth(const float* positions, uint64_t n_vertices,
                                                         const uint32_t* indices, uint64_t n_indices)
{
  Mesh mesh;
  for (uint64_t i = 0; i < n_vertices; ++i)
    mesh.positions.emplace_back(positions[3 * i], positions[3 * i + 1], positions[3 * i + 2]);
  mesh.indices.assign(indices, indices + n_indices);
  const auto bvh = bvh_from_mesh(mesh);
  std::vector<std::pair<uint32_t, int>> todo{{0u, 1}};
  int depth = 0;
  while (!todo.empty()) {
    const auto [node, d] = todo.back();
    todo.pop_back();
    depth = std::max(depth, d);
    if (!bvh[node].is_leaf()) {
      todo.push_back({bvh[node].first_child_or_primitive, d + 1});
      todo.push_back({bvh[node].first_child_or_primitive + 1, d + 1});
    }
  }
  return depth;
}

// what this library was compiled with (recorded in bench.py's reference line)
__attribute__((visibility("default"))) int ref_stack_size(void) { return REF_STACK_SIZE; }

// host-side scene build timing (the "Initialization" stage, cli.cpp:86-94): bvh_from_mesh only
__attribute__((visibility("default"))) double ref_bvh_build_seconds(const float* positions, uint64_t n_vertices,
                                                                    const uint32_t* indices, uint64_t n_indices,
                                                                    uint64_t* n_nodes)
{
  Mesh mesh;
  for (uint64_t i = 0; i < n_vertices; ++i)
    mesh.positions.emplace_back(positions[3 * i], positions[3 * i + 1], positions[3 * i + 2]);
  mesh.indices.assign(indices, indices + n_indices);
  const auto t0 = std::chrono::steady_clock::now();
  const auto bvh = bvh_from_mesh(mesh);
  const auto t1 = std::chrono::steady_clock::now();
  if (n_nodes) *n_nodes = bvh.size();
  return std::chrono::duration<double>(t1 - t0).count();
}

} // extern "C"
