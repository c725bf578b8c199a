// lbvh.cu — device LBVH builder (Karras 2012) for sm_100a: the scene's world-space triangles go
// up once (48 B each) and the node / triangle arrays the traversal kernels read are produced in
// HBM, never touching the host again.  Opt-in (PT_BUILD=lbvh / cuda_pt --fast-build): a 10-M
// triangle scene builds in tens of milliseconds instead of a second of host SAH, at the price of a
// lower-quality tree.  The construction logic is lbvh.h, shared with the sequential host
// restatement (lbvh_host.cpp) the CPU tests check; this file adds what only exists on the device:
// the radix sort (CUB), the bottom-up fit with one atomic flag per inner node, and the scan that
// compacts the surviving inner nodes.
#include "bvh_build.h"
#include "lbvh.h"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cuda_runtime.h>

#include <float.h>

#include <algorithm>
#include <vector>

namespace pt {

using namespace lbvh;

namespace {

#define LB_THREADS 256

// float <-> int key with the same ordering (for atomicMin / atomicMax on floats)
__device__ __forceinline__ int f2key(float f)
{
  const int b = __float_as_int(f);
  return b >= 0 ? b : b ^ 0x7fffffff;
}
__device__ __forceinline__ float key2f(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff); }

__global__ void __launch_bounds__(LB_THREADS)
prim_kernel(const BuildTri* __restrict__ tris, int n, Box6* __restrict__ pbox, int* __restrict__ cbounds)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float c[3] = {0.f, 0.f, 0.f};
  const bool valid = i < n;
  if (valid) {
    const BuildTri t = tris[i];
    Box6 b;
    for (int a = 0; a < 3; ++a) {
      b.lo[a] = fminf(t.v0[a], fminf(t.v1[a], t.v2[a]));
      b.hi[a] = fmaxf(t.v0[a], fmaxf(t.v1[a], t.v2[a]));
      c[a] = 0.5f * (b.lo[a] + b.hi[a]);
    }
    pbox[i] = b;
  }
  // centroid bounds: warp reduction, then one atomic pair per warp and axis
  for (int a = 0; a < 3; ++a) {
    float lo = valid ? c[a] : FLT_MAX, hi = valid ? c[a] : -FLT_MAX;
    for (int off = 16; off > 0; off >>= 1) {
      lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, off));
      hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, off));
    }
    if ((threadIdx.x & 31) == 0) {
      atomicMin(cbounds + a, f2key(lo));
      atomicMax(cbounds + 3 + a, f2key(hi));
    }
  }
}

__global__ void __launch_bounds__(LB_THREADS)
code_kernel(const Box6* __restrict__ pbox, int n, const int* __restrict__ cbounds,
            uint64_t* __restrict__ codes, uint32_t* __restrict__ order)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // one scale for the three axes (a cubic grid), as in lbvh_host.cpp
  float lo[3], inv[3], ext = 0.f;
  for (int a = 0; a < 3; ++a) {
    lo[a] = key2f(cbounds[a]);
    ext = fmaxf(ext, key2f(cbounds[3 + a]) - lo[a]);
  }
  for (int a = 0; a < 3; ++a) inv[a] = ext > 0.f ? kMortonScale / ext : 0.f;
  const Box6 b = pbox[i];
  const float c[3] = {0.5f * (b.lo[0] + b.hi[0]), 0.5f * (b.lo[1] + b.hi[1]), 0.5f * (b.lo[2] + b.hi[2])};
  codes[i] = morton63(c, lo, inv);
  order[i] = (uint32_t)i;
}

// ids: inner node k -> k, primitive at sorted position p -> (n - 1) + p
__global__ void __launch_bounds__(LB_THREADS)
hierarchy_kernel(const uint64_t* __restrict__ codes, int n, int* __restrict__ first, int* __restrict__ last,
                 int* __restrict__ split, int* __restrict__ parent)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1) return;
  int f, l, s;
  node_range(codes, n, i, f, l, s);
  first[i] = f;
  last[i] = l;
  split[i] = s;
  parent[f == s ? (n - 1) + s : s] = i;
  parent[l == s + 1 ? (n - 1) + s + 1 : s + 1] = i;
  if (i == 0) parent[0] = -1;
}

// Bottom-up: one thread per primitive climbs; the second thread to reach an inner node (atomic
// flag) owns it, merges the children's boxes and keeps climbing.
__global__ void __launch_bounds__(LB_THREADS)
fit_kernel(int n, const int* __restrict__ first, const int* __restrict__ last, const int* __restrict__ split,
           const int* __restrict__ parent, const uint32_t* __restrict__ order,
           const Box6* __restrict__ pbox, Box6* nbox, int* depth, unsigned int* __restrict__ flag)
{
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  int cur = parent[(n - 1) + p];
  while (cur >= 0) {
    __threadfence();
    if (atomicAdd(flag + cur, 1u) == 0u) return; // the sibling subtree is not done yet
    const int s = split[cur];
    const bool lleaf = first[cur] == s, rleaf = last[cur] == s + 1;
    Box6 lb, rb;
    int ld = 0, rd = 0;
    if (lleaf) {
      lb = pbox[order[s]];
    } else {
      const volatile Box6* q = nbox + s;
      for (int a = 0; a < 3; ++a) lb.lo[a] = q->lo[a], lb.hi[a] = q->hi[a];
      ld = *((volatile int*)depth + s);
    }
    if (rleaf) {
      rb = pbox[order[s + 1]];
    } else {
      const volatile Box6* q = nbox + s + 1;
      for (int a = 0; a < 3; ++a) rb.lo[a] = q->lo[a], rb.hi[a] = q->hi[a];
      rd = *((volatile int*)depth + s + 1);
    }
    Box6 b;
    for (int a = 0; a < 3; ++a) {
      b.lo[a] = fminf(lb.lo[a], rb.lo[a]);
      b.hi[a] = fmaxf(lb.hi[a], rb.hi[a]);
    }
    nbox[cur] = b;
    depth[cur] = 1 + max(ld, rd);
    cur = parent[cur];
  }
}

__global__ void __launch_bounds__(LB_THREADS)
real_kernel(int n, const int* __restrict__ first, const int* __restrict__ last, uint32_t* __restrict__ real,
            int kLeafMax)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1) return;
  real[i] = last[i] - first[i] + 1 > kLeafMax ? 1u : 0u;
}

__global__ void __launch_bounds__(LB_THREADS)
emit_kernel(int n, const int* __restrict__ first, const int* __restrict__ last, const int* __restrict__ split,
            const uint32_t* __restrict__ order, const Box6* __restrict__ pbox, const Box6* __restrict__ nbox,
            const uint32_t* __restrict__ index, float* __restrict__ nodes, int kLeafMax)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1) return;
  if (last[i] - first[i] + 1 <= kLeafMax) return;
  float nd[16];
  nd[14] = nd[15] = 0.f;
  for (int c = 0; c < 2; ++c) {
    const int k = split[i] + c;
    const bool leaf = c == 0 ? first[i] == k : last[i] == k;
    if (leaf) {
      write_child(nd, c, pbox[order[k]], leaf_code((uint32_t)k, 1u));
    } else {
      const int size = last[k] - first[k] + 1;
      write_child(nd, c, nbox[k], size <= kLeafMax ? leaf_code((uint32_t)first[k], (uint32_t)size) : index[k]);
    }
  }
  float4* dst = reinterpret_cast<float4*>(nodes + (size_t)index[i] * 16);
  dst[0] = make_float4(nd[0], nd[1], nd[2], nd[3]);
  dst[1] = make_float4(nd[4], nd[5], nd[6], nd[7]);
  dst[2] = make_float4(nd[8], nd[9], nd[10], nd[11]);
  dst[3] = make_float4(nd[12], nd[13], nd[14], nd[15]);
}

__global__ void __launch_bounds__(LB_THREADS)
tris_kernel(const BuildTri* __restrict__ tris, int n, const uint32_t* __restrict__ order, float* __restrict__ out)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const BuildTri t = tris[order[i]];
  float4* dst = reinterpret_cast<float4*>(out + (size_t)i * 12);
  dst[0] = make_float4(t.v0[0], t.v0[1], t.v0[2], __uint_as_float(t.prim));
  dst[1] = make_float4(t.v1[0] - t.v0[0], t.v1[1] - t.v0[1], t.v1[2] - t.v0[2], __uint_as_float(t.object));
  dst[2] = make_float4(t.v2[0] - t.v0[0], t.v2[1] - t.v0[1], t.v2[2] - t.v0[2], __uint_as_float(t.material));
}


// transform_point (transform.hpp:37-42) with the host's operation order and NO fused
// multiply-adds, so that a scene baked here equals the host-baked one bit for bit
__device__ __forceinline__ void xform_point_exact(const float* m, const float* p, float* out)
{
  float v[4];
  for (int r = 0; r < 4; ++r)
    v[r] = __fadd_rn(__fadd_rn(__fmul_rn(m[0 * 4 + r], p[0]), __fmul_rn(m[1 * 4 + r], p[1])),
                     __fadd_rn(__fmul_rn(m[2 * 4 + r], p[2]), __fmul_rn(m[3 * 4 + r], 1.0f)));
  out[0] = __fdiv_rn(v[0], v[3]);
  out[1] = __fdiv_rn(v[1], v[3]);
  out[2] = __fdiv_rn(v[2], v[3]);
}

__global__ void __launch_bounds__(LB_THREADS)
bake_kernel(const float* __restrict__ pos, const uint32_t* __restrict__ idx, const MeshInstance inst,
            BuildTri* __restrict__ out)
{
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= inst.n_tri) return;
  const uint64_t g = inst.first_tri + t;
  BuildTri bt;
  xform_point_exact(inst.m, pos + 3 * (size_t)idx[3 * g + 0], bt.v0);
  xform_point_exact(inst.m, pos + 3 * (size_t)idx[3 * g + 1], bt.v1);
  xform_point_exact(inst.m, pos + 3 * (size_t)idx[3 * g + 2], bt.v2);
  bt.prim = (uint32_t)g;
  bt.object = inst.object;
  bt.material = inst.material;
  out[inst.out_at + t] = bt;
}

struct Scratch {
  std::vector<void*> ptrs;
  ~Scratch()
  {
    for (void* p : ptrs) cudaFree(p);
  }
  template <typename T> cudaError_t alloc(T** p, size_t count)
  {
    *p = nullptr;
    const cudaError_t e = cudaMalloc((void**)p, std::max<size_t>(1, count) * sizeof(T));
    if (e == cudaSuccess) ptrs.push_back(*p);
    return e;
  }
};

} // namespace

#define LB_TRY(call)                                                                              \
  do {                                                                                            \
    const cudaError_t e__ = (call);                                                               \
    if (e__ != cudaSuccess) return e__;                                                           \
  } while (0)

// Source of the world-space triangles: a host array that is uploaded, or the raw mesh + instance
// table, baked on the device.  On success with out.built == true, out.nodes / out.tris are device
// allocations the caller owns (cudaFree).  out.built == false (and no allocation) when the scene
// is too small or the tree would be too deep: use the SAH builder.
struct TriSource {
  const BuildTri* h_tris = nullptr;
  const float* h_positions = nullptr;
  uint64_t n_vertices = 0;
  const uint32_t* h_indices = nullptr;
  uint64_t n_indices = 0;
  const MeshInstance* inst = nullptr;
  uint32_t n_inst = 0;
};

static cudaError_t build_lbvh_device(const TriSource& src, uint32_t n_u, DeviceLBVH& out)
{
  out = DeviceLBVH{};
  const int n = (int)n_u;
  const int kLeafMax = leaf_max_setting();
  if (n <= 4 || n_u >= (1u << 28)) return cudaSuccess;
  // timing events released on every return path (early LB_TRY exits, the too-deep decline)
  struct Events {
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    ~Events()
    {
      if (e0) cudaEventDestroy(e0);
      if (e1) cudaEventDestroy(e1);
    }
  } ev;
  LB_TRY(cudaEventCreate(&ev.e0));
  LB_TRY(cudaEventCreate(&ev.e1));
  cudaEvent_t& e0 = ev.e0;
  cudaEvent_t& e1 = ev.e1;
  Scratch sc;
  BuildTri* d_tris;
  Box6 *pbox, *nbox;
  int *cbounds, *first, *last, *split, *parent, *depth;
  uint64_t *codes, *codes2;
  uint32_t *order, *order2, *real, *index;
  unsigned int* flag;
  LB_TRY(sc.alloc(&d_tris, (size_t)n));
  LB_TRY(sc.alloc(&pbox, (size_t)n));
  LB_TRY(sc.alloc(&nbox, (size_t)n));
  LB_TRY(sc.alloc(&cbounds, 6));
  LB_TRY(sc.alloc(&first, (size_t)n));
  LB_TRY(sc.alloc(&last, (size_t)n));
  LB_TRY(sc.alloc(&split, (size_t)n));
  LB_TRY(sc.alloc(&parent, (size_t)2 * n));
  LB_TRY(sc.alloc(&depth, (size_t)n));
  LB_TRY(sc.alloc(&codes, (size_t)n));
  LB_TRY(sc.alloc(&codes2, (size_t)n));
  LB_TRY(sc.alloc(&order, (size_t)n));
  LB_TRY(sc.alloc(&order2, (size_t)n));
  LB_TRY(sc.alloc(&real, (size_t)n));
  LB_TRY(sc.alloc(&index, (size_t)n));
  LB_TRY(sc.alloc(&flag, (size_t)n));
  if (src.h_tris) {
    LB_TRY(cudaMemcpy(d_tris, src.h_tris, (size_t)n * sizeof(BuildTri), cudaMemcpyHostToDevice));
  } else {
    float* d_pos;
    uint32_t* d_idx;
    LB_TRY(sc.alloc(&d_pos, (size_t)src.n_vertices * 3));
    LB_TRY(sc.alloc(&d_idx, (size_t)src.n_indices));
    LB_TRY(cudaMemcpy(d_pos, src.h_positions, (size_t)src.n_vertices * 12, cudaMemcpyHostToDevice));
    LB_TRY(cudaMemcpy(d_idx, src.h_indices, (size_t)src.n_indices * 4, cudaMemcpyHostToDevice));
    for (uint32_t k = 0; k < src.n_inst; ++k) {
      const MeshInstance& mi = src.inst[k];
      if (mi.n_tri == 0) continue;
      bake_kernel<<<(unsigned)((mi.n_tri + LB_THREADS - 1) / LB_THREADS), LB_THREADS>>>(d_pos, d_idx, mi, d_tris);
    }
  }
  LB_TRY(cudaEventRecord(e0));
  {
    // ordered-int keys (f2key) of +FLT_MAX and -FLT_MAX: identities of the min / max reductions
    const int keys[6] = {0x7f7fffff, 0x7f7fffff, 0x7f7fffff, (int)0x80800000u, (int)0x80800000u, (int)0x80800000u};
    LB_TRY(cudaMemcpy(cbounds, keys, sizeof(keys), cudaMemcpyHostToDevice));
  }
  LB_TRY(cudaMemset(flag, 0, (size_t)n * sizeof(unsigned int)));
  const int grid = (n + LB_THREADS - 1) / LB_THREADS;
  prim_kernel<<<grid, LB_THREADS>>>(d_tris, n, pbox, cbounds);
  code_kernel<<<grid, LB_THREADS>>>(pbox, n, cbounds, codes, order);
  {
    size_t tmp_bytes = 0;
    LB_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, codes, codes2, order, order2, n, 0, kMortonBits));
    void* tmp;
    LB_TRY(sc.alloc((char**)&tmp, tmp_bytes));
    LB_TRY(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, codes, codes2, order, order2, n, 0, kMortonBits));
  }
  hierarchy_kernel<<<grid, LB_THREADS>>>(codes2, n, first, last, split, parent);
  fit_kernel<<<grid, LB_THREADS>>>(n, first, last, split, parent, order2, pbox, nbox, depth, flag);
  real_kernel<<<grid, LB_THREADS>>>(n, first, last, real, kLeafMax);
  {
    size_t tmp_bytes = 0;
    LB_TRY(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, real, index, n - 1));
    void* tmp;
    LB_TRY(sc.alloc((char**)&tmp, tmp_bytes));
    LB_TRY(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, real, index, n - 1));
  }
  uint32_t last_index = 0, last_real = 0;
  int root_depth = 0;
  Box6 root_box;
  LB_TRY(cudaMemcpy(&last_index, index + (n - 2), 4, cudaMemcpyDeviceToHost));
  LB_TRY(cudaMemcpy(&last_real, real + (n - 2), 4, cudaMemcpyDeviceToHost));
  LB_TRY(cudaMemcpy(&root_depth, depth, 4, cudaMemcpyDeviceToHost));
  LB_TRY(cudaMemcpy(&root_box, nbox, sizeof(Box6), cudaMemcpyDeviceToHost));
  if (root_depth > 60) return cudaSuccess; // too deep for the 64-entry traversal stack
  const uint32_t n_real = last_index + last_real;
  float *nodes = nullptr, *tris_out = nullptr;
  LB_TRY(cudaMalloc((void**)&nodes, (size_t)std::max(1u, n_real) * 16 * sizeof(float)));
  cudaError_t e = cudaMalloc((void**)&tris_out, (size_t)n * 12 * sizeof(float));
  if (e != cudaSuccess) {
    cudaFree(nodes);
    return e;
  }
  emit_kernel<<<grid, LB_THREADS>>>(n, first, last, split, order2, pbox, nbox, index, nodes, kLeafMax);
  tris_kernel<<<grid, LB_THREADS>>>(d_tris, n, order2, tris_out);
  e = cudaEventRecord(e1);
  if (e == cudaSuccess) e = cudaEventSynchronize(e1);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) {
    cudaFree(nodes);
    cudaFree(tris_out);
    return e;
  }
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  out.built = true;
  out.nodes = nodes;
  out.tris = tris_out;
  out.n_nodes = n_real;
  out.n_tris = n_u;
  out.depth = (uint32_t)root_depth + 1u;
  out.device_ms = ms;
  for (int a = 0; a < 3; ++a) {
    const float lo = root_box.lo[a], hi = root_box.hi[a];
    out.root_lo[a] = pad_lo(pad_lo(lo, hi), hi);
    out.root_hi[a] = pad_hi(lo, pad_hi(lo, hi));
  }
  return cudaSuccess;
}

int build_lbvh_device_c(const BuildTri* h_tris, uint32_t n, DeviceLBVH& out)
{
  TriSource src;
  src.h_tris = h_tris;
  return (int)build_lbvh_device(src, n, out);
}

int build_lbvh_device_mesh_c(const float* h_positions, uint64_t n_vertices, const uint32_t* h_indices,
                             uint64_t n_indices, const MeshInstance* inst, uint32_t n_inst, uint64_t n_world,
                             DeviceLBVH& out)
{
  out = DeviceLBVH{};
  if (n_world >= (1ull << 28)) return 0;
  TriSource src;
  src.h_positions = h_positions;
  src.n_vertices = n_vertices;
  src.h_indices = h_indices;
  src.n_indices = n_indices;
  src.inst = inst;
  src.n_inst = n_inst;
  return (int)build_lbvh_device(src, (uint32_t)n_world, out);
}

} // namespace pt
