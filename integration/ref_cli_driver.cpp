// ref_cli_driver.cpp — test driver for integration/path_tracer_b200.cpp: the body of the
// reference's execute_cli_version (src/cli/cli.cpp:62-115) with the reference's own types,
//   PathTracer path_tracer{}; create_buffers; max_iterations = spp; for (i < spp) path_trace;
//   send_to_preview(managed buffer),
// linked against the B200 implementation of `PathTracer` instead of the reference's
// path_tracer.cu.  The scene reaches it the way it reaches the reference: through
// SceneDescription's public add_material / add_mesh / add_object.  (The reference's own scene
// reader needs Assimp and nlohmann/json, which this image lacks, so the driver parses the scene
// file with pt_scene_file_read and replays it through that public API.)
//
//   ref_cli_driver <scene.json> <out.rgba> <megakernel|streaming> <max_depth> [spp] [filter_size]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime_api.h>

#include "path_tracer.hpp"
#include "scene_description.hpp"

#include "../include/b200pt.h"

extern "C" void b200_shim_set_max_depth(PathTracer* self, int max_depth);
extern "C" void b200_shim_release(PathTracer* self);

static glm::mat4 mat_from(const float* m)
{
  glm::mat4 r;
  for (int c = 0; c < 4; ++c) r[c] = glm::vec4(m[c * 4 + 0], m[c * 4 + 1], m[c * 4 + 2], m[c * 4 + 3]);
  return r;
}

int main(int argc, char** argv)
{
  if (argc < 5) {
    std::fprintf(stderr, "usage: ref_cli_driver <scene.json> <out.rgba> <megakernel|streaming> <max_depth> [spp] [filter]\n");
    return 2;
  }
  pt_scene_file* file = nullptr;
  pt_scene_desc d{};
  pt_scene_file_info info{};
  if (pt_scene_file_read(argv[1], &file, &d, &info) != PT_OK) {
    std::fprintf(stderr, "%s\n", pt_last_error());
    return 1;
  }
  SceneDescription scene_desc;
  scene_desc.filename = argv[1];
  char name[32];
  for (uint32_t i = 0; i < d.n_materials; ++i) {
    const pt_material& m = d.materials[i];
    std::snprintf(name, sizeof(name), "m%06u", i); // std::map order == index order
    const glm::vec3 albedo(m.albedo[0], m.albedo[1], m.albedo[2]);
    if (m.type == PT_MAT_DIFFUSE) scene_desc.add_material(name, Material{DiffuseMateral{albedo}});
    else if (m.type == PT_MAT_METAL) scene_desc.add_material(name, Material{MetalMaterial{albedo, m.fuzz}});
    else scene_desc.add_material(name, Material{DielectricMaterial{m.refraction_index}});
  }
  Mesh mesh;
  for (uint64_t i = 0; i < d.n_vertices; ++i)
    mesh.positions.emplace_back(d.positions[3 * i], d.positions[3 * i + 1], d.positions[3 * i + 2]);
  mesh.indices.assign(d.indices, d.indices + d.n_indices);
  const MeshRef mesh_ref = scene_desc.add_mesh("mesh", std::move(mesh));
  for (uint32_t i = 0; i < d.n_objects; ++i) {
    const pt_object& o = d.objects[i];
    std::snprintf(name, sizeof(name), "m%06u", o.material);
    const Transform tf(mat_from(o.m), mat_from(o.inv));
    if (o.type == PT_OBJ_SPHERE) {
      const pt_sphere& s = d.spheres[o.prim_index];
      scene_desc.add_object(Sphere{glm::vec3(s.center[0], s.center[1], s.center[2]), s.radius}, tf, name);
    } else {
      scene_desc.add_object(mesh_ref, tf, name);
    }
  }
  scene_desc.camera.position = glm::vec3(info.camera.position[0], info.camera.position[1], info.camera.position[2]);
  scene_desc.camera.rotation =
      glm::quat(info.camera.rotation[0], info.camera.rotation[1], info.camera.rotation[2], info.camera.rotation[3]);
  scene_desc.camera.vfov = info.camera.vfov;
  scene_desc.resolution = Resolution{info.width, info.height};
  scene_desc.spp = argc > 5 ? std::atoi(argv[5]) : info.spp;
  const int filter = argc > 6 ? std::atoi(argv[6]) : 0;

  // ---- from here on: cli.cpp:78-105, verbatim in structure
  const UResolution resolution = scene_desc.resolution.to_unsigned();
  const auto [width, height] = resolution;
  const int spp = scene_desc.spp;
  const auto& camera = scene_desc.camera;

  PathTracer path_tracer{};
  path_tracer.current_gpu_method = std::strcmp(argv[3], "megakernel") == 0 ? GPUMethod::megakernel : GPUMethod::streaming;
  b200_shim_set_max_depth(&path_tracer, std::atoi(argv[4]));
  path_tracer.create_buffers(resolution, scene_desc);
  cudaDeviceSynchronize();

  path_tracer.max_iterations = spp;
  for (int i = 0; i < spp + 2; ++i) { // two calls too many: no-ops once iteration() == max_iterations
    path_tracer.path_trace(camera, resolution);
  }
  cudaDeviceSynchronize();
  if (path_tracer.iteration() != spp) {
    std::fprintf(stderr, "iteration() = %d, expected %d\n", path_tracer.iteration(), spp);
    return 1;
  }
  if (filter > 0) {
    path_tracer.atrous_denoiser.filter_size = filter;
    path_tracer.denoise(resolution);
  }
  uchar4* buffer = nullptr;
  if (cudaMallocManaged(reinterpret_cast<void**>(&buffer), static_cast<size_t>(width) * height * 4) != cudaSuccess) return 1;
  path_tracer.send_to_preview(buffer, resolution);
  cudaDeviceSynchronize();

  FILE* f = std::fopen(argv[2], "wb");
  if (!f) return 1;
  std::fwrite(buffer, 4, static_cast<size_t>(width) * height, f);
  std::fclose(f);
  std::printf("Done path tracing %s! %ux%u spp %d iteration %d\n", scene_desc.filename.c_str(), width, height, spp,
              path_tracer.iteration());
  cudaFree(buffer);
  b200_shim_release(&path_tracer);
  pt_scene_file_free(file);
  return 0;
}
