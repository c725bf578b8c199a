// bvh_build.h — host BVH builder + flattener (replaces the reference's
// bvh_from_mesh, src/lib/accelerators/bvh.cpp:211-253).
#pragma once
#include <cstdint>
#include <memory>
#include <new>
#include <utility>
#include <vector>

namespace pt {

// One world-space triangle (instance already applied).
struct BuildTri {
  float v0[3], v1[3], v2[3];
  uint32_t prim;     // triangle index in the input index buffer / 3
  uint32_t object;   // scene object index
  uint32_t material; // material-table index of that object
};

// std::vector whose resize() leaves new elements uninitialised: the big output arrays are filled
// by parallel loops, a serial zero fill first would cost as much as filling them.
template <typename T> struct DefaultInitAlloc : std::allocator<T> {
  template <typename U> struct rebind {
    using other = DefaultInitAlloc<U>;
  };
  template <typename U> void construct(U* p) noexcept { ::new (static_cast<void*>(p)) U; }
  template <typename U, typename... A> void construct(U* p, A&&... a)
  {
    ::new (static_cast<void*>(p)) U(std::forward<A>(a)...);
  }
};
template <typename T> using RawVector = std::vector<T, DefaultInitAlloc<T>>;
// the baked world-space triangles: 480 MB at 10 M triangles, filled by a parallel loop
using BuildTris = RawVector<BuildTri>;

struct FlatBVH {
  RawVector<float> nodes; // 16 floats per node  (layout: common.cuh)
  RawVector<float> tris;  // 12 floats per triangle, leaf order
  uint32_t n_nodes = 0;
  uint32_t n_tris = 0; // including the trailing null triangle, if any
  uint32_t depth = 0;
  double sah_cost = 0.0;
  float root_lo[3] = {0, 0, 0}, root_hi[3] = {0, 0, 0}; // padded bounds of everything
  // compressed 8-wide tree over the same triangle array (layout: common.cuh), breadth-first
  RawVector<uint32_t> nodes8; // 20 words (80 B) per node
  uint32_t n_nodes8 = 0;
  uint32_t depth8 = 0;
};

// Binned-SAH (16 bins x 3 axes) top-down build, <= 4 triangles per leaf,
// OpenMP task-parallel; depth is bounded so the 64-entry traversal stack of
// the extend kernel can never overflow (the reference's 24-entry stack is
// unchecked, src/lib/static_stack.hpp:21-25).
//
// wide = true additionally collapses the binary tree into the compressed 8-wide tree
// (<= 3 triangles per leaf then, the triangle array is laid out wide-node by wide-node and the
// binary tree's leaves reference the same array).
void build_bvh(const BuildTris& tris, FlatBVH& out, bool wide = true);


// Sequential restatement of the device LBVH builder (lbvh.cu / lbvh.h) — CPU tests and the
// node-for-node check of the device result.  Returns false when the scene is too small or the
// tree too deep for the traversal stack (the caller then uses build_bvh).
bool build_lbvh_host(const BuildTris& tris, FlatBVH& out);

// Device LBVH builder (lbvh.cu).  h_tris: n world-space triangles in host memory.  With
// out.built == true, out.nodes / out.tris are device allocations the caller owns; out.built ==
// false (nothing allocated) means the builder declined (tiny scene, tree too deep) and the SAH
// builder should run.  Returns a cudaError_t as int (0 = success).
struct DeviceLBVH {
  bool built = false;
  float* nodes = nullptr; // 16 floats per node, device
  float* tris = nullptr;  // 12 floats per triangle, device
  uint32_t n_nodes = 0, n_tris = 0, depth = 0;
  float root_lo[3] = {0, 0, 0}, root_hi[3] = {0, 0, 0};
  float device_ms = 0.f;
};
int build_lbvh_device_c(const BuildTri* h_tris, uint32_t n, DeviceLBVH& out);

// Same, from the raw mesh: every mesh instance is baked to world space ON THE DEVICE (bit-identical
// to the host bake: same operation order, no fused multiply-adds), so neither the host bake nor
// the 48-byte-per-triangle upload happens.
struct MeshInstance {
  float m[16];                     // object -> world, column-major
  uint64_t first_tri, n_tri;       // triangles of its mesh in the index buffer
  uint64_t out_at;                 // first world triangle it produces
  uint32_t object, material;
};
int build_lbvh_device_mesh_c(const float* h_positions, uint64_t n_vertices, const uint32_t* h_indices,
                             uint64_t n_indices, const MeshInstance* inst, uint32_t n_inst, uint64_t n_world,
                             DeviceLBVH& out);

// Structural check of both trees against the triangle array: returns the number of violations
// (0 = every triangle is referenced exactly once by each tree, every child box contains its
// content, every reference is in range).
uint64_t validate_bvh(const FlatBVH& bvh);


// Host walk of a tree in the device kernels' visiting order; out = {inner-node visits, leaf visits,
// triangle tests, rays that hit, deepest stack}.  wide = 0: binary tree, 1: compressed 8-wide tree.
// Analysis tool only (tree-layout decisions without GPU time).
void trace_stats(const FlatBVH& bvh, const float* rays8, uint64_t n_rays, int wide, uint64_t out[5]);

} // namespace pt
