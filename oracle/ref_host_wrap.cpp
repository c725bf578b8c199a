// ref_host_wrap.cpp — C-ABI wrapper around the REFERENCE's own host-callable code, compiled with
// g++ from the sources where they lie under /root/reference/src/lib.  TEST INFRASTRUCTURE ONLY:
// it exists to pin oracle/oracle.c (tests/test_oracle_pinning.py) against the real
// implementation of  bvh_from_mesh (accelerators/bvh.cpp:211-253),
// ray_triangle/sphere/aabb_intersection_test (intersections.cuh:7-103), inverse_transform_ray /
// transform_aabb (transform.hpp:50-88), AABB (aabb.hpp) and hash (hash.cuh:4-14).
// glm / spdlog / fmt come from oracle/ref_shim.
#include <chrono>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#define __host__
#define __device__

#include <glm/glm.hpp>

#include "../include/b200pt.h"

#include "aabb.hpp"
#include "accelerators/bvh.hpp"
#include "hash.cuh"
#include "intersections.cuh"
#include "mesh.hpp"
#include "transform.hpp"

#include "accelerators/bvh.cpp"
#include "prelude.cpp"

namespace {
Ray ray_from(const float* r) { return Ray{glm::vec3(r[0], r[1], r[2]), r[3], glm::vec3(r[4], r[5], r[6]), r[7]}; }
glm::mat4 mat_from(const float* m)
{
  glm::mat4 r;
  for (int c = 0; c < 4; ++c) r[c] = glm::vec4(m[c * 4 + 0], m[c * 4 + 1], m[c * 4 + 2], m[c * 4 + 3]);
  return r;
}
void fill(pt_hit* h, bool hit, const Intersection& rec, int prim)
{
  std::memset(h, 0, sizeof(*h));
  if (!hit) {
    h->t = -1.0f;
    h->object = -1;
    h->prim = -1;
    return;
  }
  h->t = rec.t;
  h->point[0] = rec.point.x, h->point[1] = rec.point.y, h->point[2] = rec.point.z;
  h->normal[0] = rec.normal.x, h->normal[1] = rec.normal.y, h->normal[2] = rec.normal.z;
  h->material = static_cast<uint32_t>(rec.material_id);
  h->side = rec.side == HitFaceSide::front ? 0u : 1u;
  h->prim = prim;
}
} // namespace

#define EXPORT extern "C" __attribute__((visibility("default")))

EXPORT uint32_t ref_hash(uint32_t a) { return hash(a); }

EXPORT uint64_t ref_bvh_from_mesh(const float* positions, uint64_t n_vertices, const uint32_t* indices,
                                  uint64_t n_indices, void* out_nodes, uint64_t capacity, double* seconds)
{
  Mesh mesh;
  for (uint64_t i = 0; i < n_vertices; ++i)
    mesh.positions.emplace_back(positions[3 * i], positions[3 * i + 1], positions[3 * i + 2]);
  mesh.indices.assign(indices, indices + n_indices);
  const auto t0 = std::chrono::steady_clock::now();
  const std::vector<BVHNode> bvh = bvh_from_mesh(mesh);
  const auto t1 = std::chrono::steady_clock::now();
  if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
  static_assert(sizeof(BVHNode) == 32, "BVHNode layout");
  if (out_nodes && bvh.size() <= capacity) std::memcpy(out_nodes, bvh.data(), bvh.size() * sizeof(BVHNode));
  return bvh.size();
}

EXPORT int ref_ray_triangle(const float* ray8, const float* p0, const float* p1, const float* p2, pt_hit* out)
{
  Intersection rec;
  const bool hit = ray_triangle_intersection_test(ray_from(ray8), glm::vec3(p0[0], p0[1], p0[2]),
                                                  glm::vec3(p1[0], p1[1], p1[2]), glm::vec3(p2[0], p2[1], p2[2]), rec);
  fill(out, hit, rec, 0);
  return hit;
}

EXPORT int ref_ray_sphere(const float* ray8, const float* c, float radius, pt_hit* out)
{
  Intersection rec;
  const bool hit = ray_sphere_intersection_test(ray_from(ray8), Sphere{glm::vec3(c[0], c[1], c[2]), radius}, rec);
  fill(out, hit, rec, -1);
  return hit;
}

EXPORT int ref_ray_aabb(const float* ray8, const float* mn, const float* mx)
{
  return ray_aabb_intersection_test(ray_from(ray8),
                                    AABB{glm::vec3(mn[0], mn[1], mn[2]), glm::vec3(mx[0], mx[1], mx[2])});
}

EXPORT void ref_inverse_transform_ray(const float* m, const float* inv, const float* ray8, float* out8)
{
  const Transform t(mat_from(m), mat_from(inv));
  const Ray r = inverse_transform_ray(t, ray_from(ray8));
  out8[0] = r.origin.x, out8[1] = r.origin.y, out8[2] = r.origin.z, out8[3] = r.t_min;
  out8[4] = r.direction.x, out8[5] = r.direction.y, out8[6] = r.direction.z, out8[7] = r.t_max;
}

EXPORT void ref_transform_aabb(const float* m, const float* inv, const float* mn, const float* mx, float* omn,
                               float* omx)
{
  const Transform t(mat_from(m), mat_from(inv));
  const AABB b = transform_aabb(t, AABB{glm::vec3(mn[0], mn[1], mn[2]), glm::vec3(mx[0], mx[1], mx[2])});
  omn[0] = b.min.x, omn[1] = b.min.y, omn[2] = b.min.z;
  omx[0] = b.max.x, omx[1] = b.max.y, omx[2] = b.max.z;
}

EXPORT void ref_aabb_props(const float* mn, const float* mx, const float* p, float* extent, int* max_extent,
                           float* surface_area, float* offset)
{
  const AABB b{glm::vec3(mn[0], mn[1], mn[2]), glm::vec3(mx[0], mx[1], mx[2])};
  const glm::vec3 e = b.extent();
  extent[0] = e.x, extent[1] = e.y, extent[2] = e.z;
  *max_extent = b.max_extent();
  *surface_area = b.surface_area();
  const glm::vec3 o = b.offset(glm::vec3(p[0], p[1], p[2]));
  offset[0] = o.x, offset[1] = o.y, offset[2] = o.z;
}
