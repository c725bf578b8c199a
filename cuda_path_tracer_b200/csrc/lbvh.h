// lbvh.h — linear BVH construction (Karras, "Maximizing Parallelism in the Construction of BVHs,
// Octrees, and k-d Trees", HPG 2012) shared by the device builder (lbvh.cu) and its sequential
// host restatement (lbvh_host.cpp, CPU tests).  Replaces bvh_from_mesh
// (src/lib/accelerators/bvh.cpp:211-253) when a scene must be (re)built faster than the SAH
// builder can: the tree is result-equivalent (same closest hits), its quality is lower.
//
// Output is the same two-box 64-byte node array + leaf-ordered triangle array the SAH builder
// emits (common.cuh), so the traversal kernels do not know which builder ran:
//   - primitives sorted by the 63-bit Morton code of their centroid (ties broken by index; 30-bit
//     codes put ~20 triangles of the 10-M-triangle terrain into one cell and cost 4x in ray rate),
//   - Karras internal node i covers the sorted range [first_i, last_i], split after `split_i`,
//   - an internal node covering <= PT_LEAF_MAX primitives becomes a leaf (its range is contiguous),
//   - the other internal nodes keep their Karras order, compacted by an exclusive scan.
#pragma once
#include <stdint.h>
#include <stdlib.h>

#ifdef __CUDACC__
#define LB_HD __host__ __device__ __forceinline__
#else
#define LB_HD inline
#endif

namespace pt {
namespace lbvh {

// Largest range that becomes a leaf (PT_LBVH_LEAF overrides, 1..4).  The instruction-cost model
// (71 per inner visit + 119 per triangle test, scripts/tree_stats.py) prefers 2, the measured ray
// rate does not care (bunny_1m 4 599 vs 4 640, terrain 1 432 vs 1 447 Mrays/s for 2 vs 4) and 4
// keeps the node array 40 % smaller.
constexpr int kLeafMaxDefault = 4;
inline int leaf_max_setting()
{
  const char* v = getenv("PT_LBVH_LEAF");
  const int k = v ? atoi(v) : kLeafMaxDefault;
  return k < 1 ? 1 : (k > 4 ? 4 : k);
}

LB_HD int clz32(uint32_t x)
{
#ifdef __CUDA_ARCH__
  return __clz((int)x);
#else
  return x == 0u ? 32 : __builtin_clz(x);
#endif
}

LB_HD int clz64(uint64_t x)
{
#ifdef __CUDA_ARCH__
  return __clzll((long long)x);
#else
  return x == 0ull ? 64 : __builtin_clzll(x);
#endif
}

// spreads the low 21 bits of v so that two zero bits follow each of them
LB_HD uint64_t expand21(uint64_t v)
{
  v &= 0x1fffffull;
  v = (v | v << 32) & 0x1f00000000ffffull;
  v = (v | v << 16) & 0x1f0000ff0000ffull;
  v = (v | v << 8) & 0x100f00f00f00f00full;
  v = (v | v << 4) & 0x10c30c30c30c30c3ull;
  v = (v | v << 2) & 0x1249249249249249ull;
  return v;
}

// c = centroid, lo/inv = origin and 2^21 / extent of the centroid bounds (inv = 0 on a flat axis)
LB_HD uint64_t morton63(const float c[3], const float lo[3], const float inv[3])
{
  uint64_t q[3];
  for (int a = 0; a < 3; ++a) {
    float f = (c[a] - lo[a]) * inv[a];
    f = f < 0.f ? 0.f : (f > 2097151.f ? 2097151.f : f);
    q[a] = (uint64_t)f;
  }
  return (expand21(q[0]) << 2) | (expand21(q[1]) << 1) | expand21(q[2]);
}
constexpr float kMortonScale = 2097152.0f;
constexpr int kMortonBits = 63;

// length of the common prefix of sorted keys i and j (index appended, so keys are distinct);
// -1 when j is outside [0, n)
LB_HD int delta(const uint64_t* codes, int n, int i, int j)
{
  if (j < 0 || j >= n) return -1;
  const uint64_t a = codes[i], b = codes[j];
  if (a == b) return 64 + clz32((uint32_t)i ^ (uint32_t)j);
  return clz64(a ^ b);
}

// Internal node i of n sorted keys (0 <= i < n-1): the range it covers and where it splits.
// Children: left = split (a primitive if first == split, else internal node `split`),
//           right = split + 1 (a primitive if last == split + 1, else internal node split + 1).
LB_HD void node_range(const uint64_t* codes, int n, int i, int& first, int& last, int& split)
{
  const int d = delta(codes, n, i, i + 1) - delta(codes, n, i, i - 1) >= 0 ? 1 : -1;
  const int dmin = delta(codes, n, i, i - d);
  int lmax = 2;
  while (delta(codes, n, i, i + lmax * d) > dmin) lmax *= 2;
  int l = 0;
  for (int t = lmax / 2; t >= 1; t /= 2)
    if (delta(codes, n, i, i + (l + t) * d) > dmin) l += t;
  const int j = i + l * d;
  const int dnode = delta(codes, n, i, j);
  int s = 0;
  int t = l;
  do {
    t = (t + 1) >> 1;
    if (delta(codes, n, i, i + (s + t) * d) > dnode) s += t;
  } while (t > 1);
  split = i + s * d + (d < 0 ? d : 0);
  first = i < j ? i : j;
  last = i < j ? j : i;
}

LB_HD uint32_t leaf_code(uint32_t first, uint32_t count) { return ~((first << 3) | (count - 1u)); }

// same conservative padding as the SAH flattener (bvh_build.cpp pad_lo / pad_hi)
LB_HD float pad_lo(float lo, float hi)
{
  const float al = lo < 0.f ? -lo : lo, ah = hi < 0.f ? -hi : hi;
  const float m = al > ah ? al : ah;
  return lo - (m * 2.4e-7f + 1e-30f);
}
LB_HD float pad_hi(float lo, float hi)
{
  const float al = lo < 0.f ? -lo : lo, ah = hi < 0.f ? -hi : hi;
  const float m = al > ah ? al : ah;
  return hi + (m * 2.4e-7f + 1e-30f);
}

struct Box6 {
  float lo[3], hi[3];
};

// Writes child c (0/1) of a 16-float node: padded box + reference.
LB_HD void write_child(float* nd, int c, const Box6& b, uint32_t ref)
{
  const int o = c == 0 ? 0 : 4, z = c == 0 ? 8 : 10;
  nd[o + 0] = pad_lo(b.lo[0], b.hi[0]), nd[o + 1] = pad_hi(b.lo[0], b.hi[0]);
  nd[o + 2] = pad_lo(b.lo[1], b.hi[1]), nd[o + 3] = pad_hi(b.lo[1], b.hi[1]);
  nd[z + 0] = pad_lo(b.lo[2], b.hi[2]), nd[z + 1] = pad_hi(b.lo[2], b.hi[2]);
  union {
    uint32_t u;
    float f;
  } cv;
  cv.u = ref;
  nd[12 + c] = cv.f;
}

} // namespace lbvh
} // namespace pt
