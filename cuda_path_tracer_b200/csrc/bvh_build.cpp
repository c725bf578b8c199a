// bvh_build.cpp — binned-SAH BVH2 builder and 64-byte node flattener.
//
// Replaces the reference's single-threaded shared_ptr builder
// (src/lib/accelerators/bvh.cpp:74-253).  The tree is result-equivalent, not
// structure-equivalent: closest-hit answers are the same, the node layout is
// the two-box 64-byte node documented in common.cuh.
//
// Build: top-down, 16 bins x 3 axes evaluated in ONE pass per node, primitives kept
// physically partitioned (28-byte records move, so every pass is a sequential stream, not a
// gather through an index array), child bounds derived from the winning bins (no separate
// bounds pass), big nodes binned by parallel chunks, subtrees as OpenMP tasks.
#include "bvh_build.h"

#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

namespace pt {
namespace {

constexpr int kBins = 16;
constexpr int kSahDepthLimit = 32; // beyond this depth: balanced median splits
constexpr float kTravCost = 1.0f;
constexpr uint32_t kParallelBin = 1u << 18; // nodes at least this big are binned by chunks
constexpr uint32_t kTaskMin = 1u << 13;     // subtrees at least this big become tasks

struct Box {
  float lo[3], hi[3];
  void reset()
  {
    for (int a = 0; a < 3; ++a) {
      lo[a] = FLT_MAX;
      hi[a] = -FLT_MAX;
    }
  }
  void grow(const float* p)
  {
    for (int a = 0; a < 3; ++a) {
      lo[a] = std::min(lo[a], p[a]);
      hi[a] = std::max(hi[a], p[a]);
    }
  }
  void grow(const Box& b)
  {
    for (int a = 0; a < 3; ++a) {
      lo[a] = std::min(lo[a], b.lo[a]);
      hi[a] = std::max(hi[a], b.hi[a]);
    }
  }
  float area() const
  {
    const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    if (dx < 0.f) return 0.f;
    return 2.0f * (dx * dy + dx * dz + dy * dz);
  }
};

// one primitive: its box and the index of the triangle it came from (centroid = box centre)
struct Prim {
  Box b;
  uint32_t id;
  float c(int a) const { return 0.5f * (b.lo[a] + b.hi[a]); }
};

struct Bin {
  Box box; // bounds of the primitives
  uint32_t count;
  void merge(const Bin& o)
  {
    box.grow(o.box);
    count += o.count;
  }
};
// Bins are SPARSE: only the bins named in mask[axis] hold data.  Most nodes of a tree are small
// (a 10-M-triangle tree has 5 M nodes of a handful of primitives each), and clearing, merging and
// sweeping 48 bins per node cost more than binning the node's primitives.  Skipping empty bins
// changes no result: growing a box by an empty bin is a no-op, and a split candidate that ends
// on an empty bin repeats the previous candidate's cost, which the strict `<` never prefers.
struct BinSet {
  Bin bin[3][kBins];
  uint32_t mask[3];
  void reset() { mask[0] = mask[1] = mask[2] = 0u; }
  void merge(const BinSet& o)
  {
    for (int a = 0; a < 3; ++a) {
      for (uint32_t m = o.mask[a]; m; m &= m - 1u) {
        const int b = __builtin_ctz(m);
        if (mask[a] >> b & 1u)
          bin[a][b].merge(o.bin[a][b]);
        else
          bin[a][b] = o.bin[a][b];
      }
      mask[a] |= o.mask[a];
    }
  }
};

// trivially constructible on purpose: the node and primitive arrays are allocated without being
// cleared (a single-threaded clear of ~0.5 GB cost more than the parallel passes that fill them)
struct BuildNode {
  Box box;
  uint32_t left, right;  // children (inner) ...
  uint32_t first, count; // ... or primitive range (leaf, count > 0)
  uint32_t inner;        // inner nodes in this subtree, itself included (0 for a leaf)
  uint32_t height;       // edges to the deepest leaf below (0 for a leaf)
  float sah;             // sum over the subtree of area x (leaf ? count x C_isect : C_trav), not normalised
};

template <typename T> struct RawArray {
  T* p = nullptr;
  size_t n = 0;
  RawArray() = default;
  RawArray(const RawArray&) = delete;
  RawArray& operator=(const RawArray&) = delete;
  ~RawArray() { std::free(p); }
  void resize(size_t count)
  {
    std::free(p);
    p = static_cast<T*>(std::malloc(std::max<size_t>(1, count) * sizeof(T)));
    if (!p) throw std::bad_alloc();
    n = count;
  }
  T* data() { return p; }
  const T* data() const { return p; }
  T* begin() { return p; }
  T& operator[](size_t i) { return p[i]; }
  const T& operator[](size_t i) const { return p[i]; }
};

struct Builder {
  int kLeafMax = 4; // 3 when the compressed 8-wide tree is derived from this one
  // SAH cost of one triangle test relative to one inner-node visit.  1.0 is what every measured
  // number of this round was taken with; PT_SAH_ISECT / PT_SAH_LEAF are experiment knobs (the
  // traversal kernel spends 71 SASS instructions per inner visit and ~119 per triangle test).
  float kIsectCost = 1.0f;
  RawArray<Prim> prims;
  RawArray<BuildNode> nodes;
  // Node slots are handed out by RANGE, not by a shared counter: the subtree over `count`
  // primitives rooted at slot ni owns slots [ni, ni + 2 count - 1).  Subtrees big enough to be
  // tasks split their range (left child at ni + 1, right child at ni + 2 lc); a subtree that is
  // built sequentially fills the front of its range through a private bump counter, so the
  // touched memory stays compact.  No atomic, no cache line shared between the threads building
  // different subtrees, and the array position of every node is independent of scheduling.
  // Unused slots stay uninitialised and are never read (every consumer follows left/right).
  // (A chunk-parallel partition of the biggest nodes was measured as well: no gain at 10 M
  // triangles on 8 threads — the build is bound by per-primitive work, not by the serial top.)

  void make_leaf(uint32_t ni, uint32_t first, uint32_t count)
  {
    BuildNode& nd = nodes[ni];
    nd.first = first;
    nd.count = count;
    nd.left = nd.right = 0;
    nd.inner = 0;
    nd.height = 0;
    nd.sah = nd.box.area() * (float)count * kIsectCost;
  }

#if defined(__SSE2__)
  // Four-lane binning: one min/max pair per box instead of six scalar ones.  Lane 3 is a
  // don't-care (it carries hi[0] / the id bits and is masked before any arithmetic).
  struct VBins {
    struct VBin {
      __m128 blo, bhi;
    };
    VBin vb[3][kBins];
    uint32_t cnt[3][kBins];
    uint32_t mask[3];
  };
  struct BinXform {
    __m128 mask3, cblo, sc4, half, top;
    BinXform(const Box& cb, const float* scale)
        : mask3(_mm_castsi128_ps(_mm_set_epi32(0, -1, -1, -1))), cblo(_mm_set_ps(0.f, cb.lo[2], cb.lo[1], cb.lo[0])),
          sc4(_mm_set_ps(0.f, scale[2], scale[1], scale[0])), half(_mm_set1_ps(0.5f)),
          top(_mm_set1_ps((float)(kBins - 1)))
    {
    }
    // bin indices of the centroid on the three axes (lanes 0-2), clamped to [0, kBins)
    __m128i bins_of(const Prim& q, __m128& lo, __m128& hi) const
    {
      lo = _mm_and_ps(_mm_loadu_ps(q.b.lo), mask3);
      hi = _mm_and_ps(_mm_loadu_ps(q.b.hi), mask3);
      const __m128 c = _mm_mul_ps(half, _mm_add_ps(lo, hi));
      __m128 f = _mm_mul_ps(_mm_sub_ps(c, cblo), sc4);
      f = _mm_min_ps(_mm_max_ps(f, _mm_setzero_ps()), top);
      return _mm_cvttps_epi32(f);
    }
  };
  static void store_bins(const VBins& v, BinSet& out)
  {
    for (int a = 0; a < 3; ++a) {
      out.mask[a] = v.mask[a];
      for (uint32_t m = v.mask[a]; m; m &= m - 1u) {
        const int k = __builtin_ctz(m);
        alignas(16) float t[2][4];
        _mm_store_ps(t[0], v.vb[a][k].blo);
        _mm_store_ps(t[1], v.vb[a][k].bhi);
        Bin& bn = out.bin[a][k];
        for (int c = 0; c < 3; ++c) {
          bn.box.lo[c] = t[0][c];
          bn.box.hi[c] = t[1][c];
        }
        bn.count = v.cnt[a][k];
      }
    }
  }
  // Big nodes: every bin is initialised up front and the loop is branch-free (measured: 16 against
  // 36 TSC ticks per primitive for the first-touch loop on spatially coherent input).
  __attribute__((noinline)) static void bin_dense(const Prim* p, uint32_t n, const BinXform& x, const bool* use,
                                                  BinSet& out)
  {
    VBins v;
    const __m128 big = _mm_set1_ps(FLT_MAX), nbig = _mm_set1_ps(-FLT_MAX);
    for (int a = 0; a < 3; ++a)
      for (int k = 0; k < kBins; ++k) {
        v.vb[a][k].blo = big;
        v.vb[a][k].bhi = nbig;
        v.cnt[a][k] = 0;
      }
    for (uint32_t i = 0; i < n; ++i) {
      __m128 lo, hi;
      const __m128i bi = x.bins_of(p[i], lo, hi);
      const int k0 = _mm_cvtsi128_si32(bi);
      const int k1 = _mm_cvtsi128_si32(_mm_shuffle_epi32(bi, 0x55));
      const int k2 = _mm_cvtsi128_si32(_mm_shuffle_epi32(bi, 0xAA));
      v.vb[0][k0].blo = _mm_min_ps(v.vb[0][k0].blo, lo), v.vb[0][k0].bhi = _mm_max_ps(v.vb[0][k0].bhi, hi);
      v.vb[1][k1].blo = _mm_min_ps(v.vb[1][k1].blo, lo), v.vb[1][k1].bhi = _mm_max_ps(v.vb[1][k1].bhi, hi);
      v.vb[2][k2].blo = _mm_min_ps(v.vb[2][k2].blo, lo), v.vb[2][k2].bhi = _mm_max_ps(v.vb[2][k2].bhi, hi);
      v.cnt[0][k0]++, v.cnt[1][k1]++, v.cnt[2][k2]++;
    }
    for (int a = 0; a < 3; ++a) {
      v.mask[a] = 0u;
      if (!use[a]) continue; // a flat axis (scale 0) puts everything into bin 0: not a candidate
      for (int k = 0; k < kBins; ++k)
        if (v.cnt[a][k]) v.mask[a] |= 1u << k;
    }
    store_bins(v, out);
  }
  // Small nodes (most of the tree): bins are initialised on first touch, so the cost follows the
  // node's size, not the 48 bins.
  static void bin_sparse(const Prim* p, uint32_t n, const BinXform& x, const bool* use, BinSet& out)
  {
    VBins v;
    v.mask[0] = v.mask[1] = v.mask[2] = 0u;
    for (uint32_t i = 0; i < n; ++i) {
      __m128 lo, hi;
      alignas(16) int bi[4];
      _mm_store_si128(reinterpret_cast<__m128i*>(bi), x.bins_of(p[i], lo, hi));
      for (int a = 0; a < 3; ++a) {
        if (!use[a]) continue;
        const int k = bi[a];
        VBins::VBin& b = v.vb[a][k];
        if (v.mask[a] >> k & 1u) {
          b.blo = _mm_min_ps(b.blo, lo);
          b.bhi = _mm_max_ps(b.bhi, hi);
          v.cnt[a][k]++;
        } else {
          b.blo = lo, b.bhi = hi;
          v.cnt[a][k] = 1;
          v.mask[a] |= 1u << k;
        }
      }
    }
    store_bins(v, out);
  }
#endif

  // Bins `n` primitives on the three axes at once into `out` (this call OVERWRITES out; chunks are
  // combined with BinSet::merge).
  static void bin_range(const Prim* p, uint32_t n, const Box& cb, const float* scale, BinSet& out)
  {
    out.reset();
#if defined(__SSE2__)
    const BinXform x(cb, scale);
    const bool use[3] = {scale[0] != 0.f, scale[1] != 0.f, scale[2] != 0.f};
    if (n >= 32) // break-even of initialising all 48 bins against the first-touch branches
      bin_dense(p, n, x, use, out);
    else
      bin_sparse(p, n, x, use, out);
#else
    for (uint32_t i = 0; i < n; ++i) {
      const Prim& q = p[i];
      float c[3] = {q.c(0), q.c(1), q.c(2)};
      for (int a = 0; a < 3; ++a) {
        if (scale[a] == 0.f) continue;
        int b = (int)((c[a] - cb.lo[a]) * scale[a]);
        b = b < 0 ? 0 : (b >= kBins ? kBins - 1 : b);
        Bin& bn = out.bin[a][b];
        if (!(out.mask[a] >> b & 1u)) {
          bn.box.reset();
          bn.count = 0;
          out.mask[a] |= 1u << b;
        }
        bn.box.grow(q.b);
        bn.count++;
      }
    }
#endif
  }

  // bounds of a primitive range (median-split fallback only)
  void range_bounds(uint32_t first, uint32_t count, Box& nb, Box& cb) const
  {
    nb.reset();
    cb.reset();
    for (uint32_t i = first; i < first + count; ++i) {
      nb.grow(prims[i].b);
      const float c[3] = {prims[i].c(0), prims[i].c(1), prims[i].c(2)};
      cb.grow(c);
    }
  }

  // Partitions p[0, n) into the primitives whose centroid falls into bins <= bb of `axis`, then the
  // rest, and returns the size of the first part.  Element order is std::partition's (libstdc++,
  // bidirectional iterators: converge from both ends, swap), and every element is classified
  // exactly once — which is where the children's primitive and centroid bounds are accumulated, in
  // registers, instead of in the per-bin records of the binning pass.
  static uint32_t partition_by_bin(Prim* p, uint32_t n, int axis, float lo, float sc, int bb, Box& lbox,
                                   Box& lcb, Box& rbox, Box& rcb)
  {
#if defined(__SSE2__)
    const __m128 big = _mm_set1_ps(FLT_MAX), nbig = _mm_set1_ps(-FLT_MAX), half = _mm_set1_ps(0.5f);
    __m128 llo = big, lhi = nbig, lclo = big, lchi = nbig, rlo = big, rhi = nbig, rclo = big, rchi = nbig;
    auto goes_left = [&](const Prim& q) -> bool {
      const __m128 qlo = _mm_loadu_ps(q.b.lo), qhi = _mm_loadu_ps(q.b.hi); // lane 3: don't care
      const __m128 c = _mm_mul_ps(half, _mm_add_ps(qlo, qhi));
      int b = (int)((q.c(axis) - lo) * sc);
      b = b < 0 ? 0 : (b >= kBins ? kBins - 1 : b);
      if (b <= bb) {
        llo = _mm_min_ps(llo, qlo), lhi = _mm_max_ps(lhi, qhi);
        lclo = _mm_min_ps(lclo, c), lchi = _mm_max_ps(lchi, c);
        return true;
      }
      rlo = _mm_min_ps(rlo, qlo), rhi = _mm_max_ps(rhi, qhi);
      rclo = _mm_min_ps(rclo, c), rchi = _mm_max_ps(rchi, c);
      return false;
    };
#else
    lbox.reset(), lcb.reset(), rbox.reset(), rcb.reset();
    auto goes_left = [&](const Prim& q) -> bool {
      const float c[3] = {q.c(0), q.c(1), q.c(2)};
      int b = (int)((c[axis] - lo) * sc);
      b = b < 0 ? 0 : (b >= kBins ? kBins - 1 : b);
      (b <= bb ? lbox : rbox).grow(q.b);
      (b <= bb ? lcb : rcb).grow(c);
      return b <= bb;
    };
#endif
    Prim* f = p;
    Prim* l = p + n;
    for (;;) {
      for (;;) {
        if (f == l) goto done;
        if (!goes_left(*f)) break;
        ++f;
      }
      --l;
      for (;;) {
        if (f == l) goto done;
        if (goes_left(*l)) break;
        --l;
      }
      std::swap(*f, *l);
      ++f;
    }
  done:
#if defined(__SSE2__)
    {
      alignas(16) float t[8][4];
      _mm_store_ps(t[0], llo), _mm_store_ps(t[1], lhi), _mm_store_ps(t[2], lclo), _mm_store_ps(t[3], lchi);
      _mm_store_ps(t[4], rlo), _mm_store_ps(t[5], rhi), _mm_store_ps(t[6], rclo), _mm_store_ps(t[7], rchi);
      for (int x = 0; x < 3; ++x) {
        lbox.lo[x] = t[0][x], lbox.hi[x] = t[1][x], lcb.lo[x] = t[2][x], lcb.hi[x] = t[3][x];
        rbox.lo[x] = t[4][x], rbox.hi[x] = t[5][x], rcb.lo[x] = t[6][x], rcb.hi[x] = t[7][x];
      }
    }
#endif
    return (uint32_t)(f - p);
  }

  void build(uint32_t ni, uint32_t first, uint32_t count, int depth, const Box& nb, const Box& cb,
             uint32_t* bump = nullptr)
  {
    uint32_t local_next = ni + 1u;
    if (!bump && count < kTaskMin) bump = &local_next; // sequential from here down
    nodes[ni].box = nb;
    if (count == 1) {
      make_leaf(ni, first, count);
      return;
    }
    const float ext[3] = {cb.hi[0] - cb.lo[0], cb.hi[1] - cb.lo[1], cb.hi[2] - cb.lo[2]};
    const int longest = (ext[0] > ext[1] && ext[0] > ext[2]) ? 0 : (ext[1] > ext[2] ? 1 : 2);

    uint32_t mid = 0;
    bool have_split = false;
    Box lbox, lcb, rbox, rcb;
    if (depth < kSahDepthLimit && ext[longest] > 0.0f) {
      float scale[3];
      for (int a = 0; a < 3; ++a) scale[a] = ext[a] > 0.0f ? (float)kBins / ext[a] : 0.f;
      BinSet bins;
      if (count >= kParallelBin) {
        bins.reset();
        const int chunks = (int)std::min<uint32_t>(64, count / (kParallelBin / 4));
        std::vector<BinSet> part(chunks);
        const Prim* base = prims.data() + first;
#pragma omp taskloop default(shared) grainsize(1)
        for (int c = 0; c < chunks; ++c) {
          const uint32_t b0 = (uint32_t)((uint64_t)count * c / chunks);
          const uint32_t b1 = (uint32_t)((uint64_t)count * (c + 1) / chunks);
          bin_range(base + b0, b1 - b0, cb, scale, part[c]);
        }
        for (int c = 0; c < chunks; ++c) bins.merge(part[c]);
      } else {
        bin_range(prims.data() + first, count, cb, scale, bins);
      }

      // SAH sweep over the non-empty bins of every axis (see BinSet: same winner as a sweep over
      // all 16 bins)
      float best_cost = FLT_MAX;
      int best_axis = -1, best_bin = -1;
      for (int axis = 0; axis < 3; ++axis) {
        const uint32_t used = bins.mask[axis];
        if ((used & (used - 1u)) == 0u) continue; // everything in one bin: no candidate
        int idx[kBins];
        int k = 0;
        for (uint32_t m = used; m; m &= m - 1u) idx[k++] = __builtin_ctz(m);
        float right_area[kBins];
        uint32_t right_cnt[kBins];
        Box acc;
        acc.reset();
        uint32_t c = 0;
        for (int j = k - 1; j > 0; --j) {
          acc.grow(bins.bin[axis][idx[j]].box);
          c += bins.bin[axis][idx[j]].count;
          right_area[j] = acc.area();
          right_cnt[j] = c;
        }
        acc.reset();
        c = 0;
        for (int j = 0; j + 1 < k; ++j) {
          acc.grow(bins.bin[axis][idx[j]].box);
          c += bins.bin[axis][idx[j]].count;
          const float cost = acc.area() * (float)c + right_area[j + 1] * (float)right_cnt[j + 1];
          if (cost < best_cost) {
            best_cost = cost;
            best_axis = axis;
            best_bin = idx[j];
          }
        }
      }
      if (best_axis >= 0) {
        const float area = nb.area();
        const float split_cost = kTravCost + (area > 0.f ? best_cost / area : 0.f) * kIsectCost;
        const float leaf_cost = (float)count * kIsectCost;
        if (count <= (uint32_t)kLeafMax && leaf_cost <= split_cost) {
          make_leaf(ni, first, count);
          return;
        }
        mid = first + partition_by_bin(prims.data() + first, count, best_axis, cb.lo[best_axis],
                                       scale[best_axis], best_bin, lbox, lcb, rbox, rcb);
        have_split = mid > first && mid < first + count;
      }
    }
    if (!have_split) {
      if (count <= (uint32_t)kLeafMax) {
        make_leaf(ni, first, count);
        return;
      }
      // balanced median split (degenerate centroids or depth guard)
      mid = first + count / 2;
      std::nth_element(prims.begin() + first, prims.begin() + mid, prims.begin() + first + count,
                       [&](const Prim& a, const Prim& b) {
                         const float ca = a.c(longest), cbv = b.c(longest);
                         return ca < cbv || (ca == cbv && a.id < b.id);
                       });
      range_bounds(first, mid - first, lbox, lcb);
      range_bounds(mid, first + count - mid, rbox, rcb);
    }

    const uint32_t lc = mid - first, rc = first + count - mid;
    uint32_t l, r;
    if (bump) {
      l = (*bump)++;
      r = (*bump)++;
    } else {
      l = ni + 1u;
      r = ni + 2u * lc;
    }
    nodes[ni].left = l;
    nodes[ni].right = r;
    nodes[ni].count = 0;
    nodes[ni].first = first;
    if (count >= kTaskMin) {
#pragma omp task default(shared) firstprivate(l, first, lc, depth, lbox, lcb)
      build(l, first, lc, depth + 1, lbox, lcb);
#pragma omp task default(shared) firstprivate(r, mid, rc, depth, rbox, rcb)
      build(r, mid, rc, depth + 1, rbox, rcb);
#pragma omp taskwait
    } else {
      build(l, first, lc, depth + 1, lbox, lcb, bump);
      build(r, mid, rc, depth + 1, rbox, rcb, bump);
    }
    nodes[ni].inner = 1u + nodes[l].inner + nodes[r].inner;
    nodes[ni].height = 1u + std::max(nodes[l].height, nodes[r].height);
    nodes[ni].sah = nodes[ni].box.area() * kTravCost + (nodes[l].sah + nodes[r].sah);
  }
};

inline float pad_lo(float lo, float hi)
{
  const float m = std::max(std::fabs(lo), std::fabs(hi));
  return lo - (m * 2.4e-7f + 1e-30f);
}
inline float pad_hi(float lo, float hi)
{
  const float m = std::max(std::fabs(lo), std::fabs(hi));
  return hi + (m * 2.4e-7f + 1e-30f);
}

inline float u2f(uint32_t u)
{
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}

// ------------------------------------------------------------------ 8-wide collapse
// Compressed wide BVH in the spirit of Ylitie, Karras, Laine, "Efficient Incoherent Ray
// Traversal on GPUs Through Compressed Wide BVHs" (HPG 2017), derived from the binary tree:
// every wide node absorbs binary descendants (largest surface area first) until it has up to
// eight children; children are placed in octant-ordered slots; child boxes are quantised to
// 8 bits per plane on a per-node power-of-two grid, always outwards.
struct WideTmp {
  uint32_t child[8]; // binary node per slot, ~0u = empty
  uint32_t n_inner = 0, n_tris = 0;
};

struct WideBuilder {
  const Builder& B;
  std::vector<uint32_t> bnode; // binary node of every wide node (breadth-first)
  std::vector<uint32_t> new_first; // per binary leaf: position of its first triangle in the new order
  std::vector<uint32_t> order;     // new triangle position -> index into B.prims
  explicit WideBuilder(const Builder& b) : B(b) {}

  static Box padded(const Box& b)
  {
    Box r;
    for (int a = 0; a < 3; ++a) {
      r.lo[a] = pad_lo(b.lo[a], b.hi[a]);
      r.hi[a] = pad_hi(b.lo[a], b.hi[a]);
    }
    return r;
  }

  void gather(uint32_t b, WideTmp& w) const
  {
    uint32_t ch[8];
    int n = 0;
    const BuildNode& nd = B.nodes[b];
    if (nd.count != 0) {
      ch[n++] = b; // the whole tree is one leaf
    } else {
      ch[n++] = nd.left;
      ch[n++] = nd.right;
      while (n < 8) {
        int pick = -1;
        float best = -1.f;
        for (int i = 0; i < n; ++i) {
          const BuildNode& c = B.nodes[ch[i]];
          if (c.count != 0) continue;
          const float a = c.box.area();
          if (a > best) {
            best = a;
            pick = i;
          }
        }
        if (pick < 0) break;
        const BuildNode& c = B.nodes[ch[pick]];
        ch[pick] = c.left;
        ch[n++] = c.right;
      }
    }
    // octant-ordered slots: slot bit a set = the child lies towards +axis a of the node centre;
    // greedy assignment of the (child, slot) pair with the largest projection
    float cen[3];
    for (int a = 0; a < 3; ++a) cen[a] = 0.5f * (nd.box.lo[a] + nd.box.hi[a]);
    float cost[8][8];
    for (int i = 0; i < n; ++i) {
      const Box& cb = B.nodes[ch[i]].box;
      float d[3];
      for (int a = 0; a < 3; ++a) d[a] = 0.5f * (cb.lo[a] + cb.hi[a]) - cen[a];
      for (int s = 0; s < 8; ++s)
        cost[i][s] = ((s & 1) ? d[0] : -d[0]) + ((s & 2) ? d[1] : -d[1]) + ((s & 4) ? d[2] : -d[2]);
    }
    for (int s = 0; s < 8; ++s) w.child[s] = 0xffffffffu;
    bool used_c[8] = {false, false, false, false, false, false, false, false};
    for (int k = 0; k < n; ++k) {
      int bi = -1, bs = -1;
      float bc = -FLT_MAX;
      for (int i = 0; i < n; ++i) {
        if (used_c[i]) continue;
        for (int s = 0; s < 8; ++s) {
          if (w.child[s] != 0xffffffffu) continue;
          if (cost[i][s] > bc) {
            bc = cost[i][s];
            bi = i;
            bs = s;
          }
        }
      }
      used_c[bi] = true;
      w.child[bs] = ch[bi];
    }
    w.n_inner = 0;
    w.n_tris = 0;
    for (int s = 0; s < 8; ++s) {
      if (w.child[s] == 0xffffffffu) continue;
      const BuildNode& c = B.nodes[w.child[s]];
      if (c.count != 0) {
        w.n_tris += c.count;
      } else {
        w.n_inner++;
      }
    }
  }

  // 80-byte node, 20 words:
  //   0..2 p (grid origin)   3 = ex' | ey'<<8 | ez'<<16 | imask<<24   (e' = biased exponent - 8)
  //   4 child base   5 triangle base   6..7 meta[8]
  //   8..9 qlo.x[8]  10..11 qlo.y[8]  12..13 qlo.z[8]  14..15 qhi.x[8]  16..17 qhi.y[8]  18..19 qhi.z[8]
  void emit(uint32_t b, const WideTmp& w, uint32_t child_base, uint32_t tri_base, uint32_t* o)
  {
    const Box nb = padded(B.nodes[b].box);
    float p[3];
    int eb[3];
    double scale[3];
    Box cbx[8];
    for (int s = 0; s < 8; ++s)
      if (w.child[s] != 0xffffffffu) cbx[s] = padded(B.nodes[w.child[s]].box);
    for (int a = 0; a < 3; ++a) {
      const float ext0 = nb.hi[a] - nb.lo[a];
      p[a] = nb.lo[a] - (ext0 * 3e-5f + 1e-30f);
      const double ext = (double)nb.hi[a] - (double)p[a];
      int e = ext > 0.0 ? (int)std::ceil(std::log2(ext / 255.0)) : -110;
      e = std::max(e, -110);
      for (;; ++e) {
        const double sc = std::ldexp(1.0, e);
        bool ok = true;
        for (int s = 0; s < 8 && ok; ++s) {
          if (w.child[s] == 0xffffffffu) continue;
          if (std::ceil(((double)cbx[s].hi[a] - (double)p[a]) / sc + 1.0 / 64.0) > 255.0) ok = false;
        }
        if (ok) break;
      }
      eb[a] = e + 127;
      scale[a] = std::ldexp(1.0, e);
    }
    uint8_t q[6][8];
    uint8_t meta[8];
    uint32_t imask = 0, tri_off = 0;
    for (int s = 0; s < 8; ++s) {
      if (w.child[s] == 0xffffffffu) {
        for (int a = 0; a < 3; ++a) {
          q[a][s] = 255;
          q[3 + a][s] = 0;
        }
        meta[s] = 0;
        continue;
      }
      for (int a = 0; a < 3; ++a) {
        double lo = std::floor(((double)cbx[s].lo[a] - (double)p[a]) / scale[a] - 1.0 / 64.0);
        double hi = std::ceil(((double)cbx[s].hi[a] - (double)p[a]) / scale[a] + 1.0 / 64.0);
        lo = std::min(std::max(lo, 0.0), 255.0);
        hi = std::min(std::max(hi, 0.0), 255.0);
        q[a][s] = (uint8_t)lo;
        q[3 + a][s] = (uint8_t)hi;
      }
      const BuildNode& c = B.nodes[w.child[s]];
      if (c.count != 0) {
        meta[s] = (uint8_t)((((1u << c.count) - 1u) << 5) | tri_off); // unary count | offset
        new_first[w.child[s]] = tri_base + tri_off;
        for (uint32_t k = 0; k < c.count; ++k) order[tri_base + tri_off + k] = c.first + k;
        tri_off += c.count;
      } else {
        meta[s] = (uint8_t)((1u << 5) | (24u + (uint32_t)s));
        imask |= 1u << s;
      }
    }
    std::memcpy(o + 0, p, 12);
    o[3] = (uint32_t)(eb[0] - 8) | ((uint32_t)(eb[1] - 8) << 8) | ((uint32_t)(eb[2] - 8) << 16) | (imask << 24);
    o[4] = child_base;
    o[5] = tri_base;
    std::memcpy(o + 6, meta, 8);
    for (int k = 0; k < 6; ++k) std::memcpy(o + 8 + 2 * k, q[k], 8);
  }

  void run(uint32_t root, uint32_t n_prims, FlatBVH& out)
  {
    new_first.assign(B.nodes.n, 0u);
    order.assign(n_prims, 0u);
    bnode.clear();
    bnode.push_back(root);
    std::vector<WideTmp> tmp;
    std::vector<uint32_t> cbase, tbase;
    uint32_t level_begin = 0, next_tri = 0, depth = 0;
    while (level_begin < bnode.size()) {
      const uint32_t level_end = (uint32_t)bnode.size();
      const uint32_t nl = level_end - level_begin;
      tmp.resize(nl);
      cbase.resize(nl);
      tbase.resize(nl);
#pragma omp parallel for schedule(static) if (nl > 256)
      for (long long i = 0; i < (long long)nl; ++i) gather(bnode[level_begin + i], tmp[i]);
      uint32_t next_node = level_end;
      for (uint32_t i = 0; i < nl; ++i) {
        cbase[i] = next_node;
        tbase[i] = next_tri;
        next_node += tmp[i].n_inner;
        next_tri += tmp[i].n_tris;
      }
      bnode.resize(next_node);
      out.nodes8.resize((size_t)next_node * 20); // grows level by level
#pragma omp parallel for schedule(static) if (nl > 256)
      for (long long i = 0; i < (long long)nl; ++i) {
        uint32_t r = 0;
        for (int s = 0; s < 8; ++s) {
          const uint32_t c = tmp[i].child[s];
          if (c != 0xffffffffu && B.nodes[c].count == 0) bnode[cbase[i] + r++] = c;
        }
        emit(bnode[level_begin + i], tmp[i], cbase[i], tbase[i], &out.nodes8[(size_t)(level_begin + i) * 20]);
      }
      level_begin = level_end;
      ++depth;
    }
    out.n_nodes8 = (uint32_t)bnode.size();
    out.nodes8.resize((size_t)out.n_nodes8 * 20);
    out.depth8 = depth;
  }
};

} // namespace

void build_bvh(const BuildTris& tris, FlatBVH& out, bool wide)
{
  out = FlatBVH{};
  const size_t n = tris.size();
  if (n == 0) return;

  // PT_BUILD_TIMING=1 prints the phases to stderr
  const bool timing = std::getenv("PT_BUILD_TIMING") != nullptr;
  auto t_prev = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!timing) return;
    const auto now = std::chrono::steady_clock::now();
    std::fprintf(stderr, "build_bvh: %-22s %8.1f ms\n", what,
                 std::chrono::duration<double, std::milli>(now - t_prev).count());
    t_prev = now;
  };
  Builder B;
  B.kLeafMax = wide ? 3 : 4;
  if (!wide) {
    if (const char* v = std::getenv("PT_SAH_LEAF")) B.kLeafMax = std::min(8, std::max(1, std::atoi(v)));
    if (const char* v = std::getenv("PT_SAH_ISECT")) B.kIsectCost = std::min(16.f, std::max(0.0625f, (float)std::atof(v)));
  }
  B.prims.resize(n);
  B.nodes.resize(2 * n);
  Box root_box, root_cb;
  root_box.reset();
  root_cb.reset();
#pragma omp parallel
  {
    Box lb, lc;
    lb.reset();
    lc.reset();
#pragma omp for schedule(static) nowait
    for (long long i = 0; i < (long long)n; ++i) {
      Prim& p = B.prims[i];
      p.b.reset();
      p.b.grow(tris[i].v0);
      p.b.grow(tris[i].v1);
      p.b.grow(tris[i].v2);
      p.id = (uint32_t)i;
      lb.grow(p.b);
      const float c[3] = {p.c(0), p.c(1), p.c(2)};
      lc.grow(c);
    }
#pragma omp critical
    {
      root_box.grow(lb);
      root_cb.grow(lc);
    }
  }
  lap("primitive boxes");
  const uint32_t root = 0; // owns slots [0, 2n - 1)
#pragma omp parallel
  {
#pragma omp single
    B.build(root, 0, (uint32_t)n, 0, root_box, root_cb);
  }
  lap("binned SAH recursion");

  // ---- wide tree first: it decides the triangle order both trees share
  WideBuilder W(B);
  if (wide) W.run(root, (uint32_t)n, out);
  lap("wide collapse");

  // ---- flatten: inner nodes in DFS pre-order, triangles in leaf order
  const bool root_is_leaf = B.nodes[root].count != 0;
  for (int a = 0; a < 3; ++a) {
    const float lo = B.nodes[root].box.lo[a], hi = B.nodes[root].box.hi[a];
    out.root_lo[a] = pad_lo(pad_lo(lo, hi), hi);
    out.root_hi[a] = pad_hi(lo, pad_hi(lo, hi));
  }
  const bool need_null = root_is_leaf;
  out.n_tris = (uint32_t)n + (need_null ? 1u : 0u);
  out.tris.resize((size_t)out.n_tris * 12); // not cleared: every element is written below
#pragma omp parallel for schedule(static)
  for (long long i = 0; i < (long long)n; ++i) {
    const BuildTri& t = tris[B.prims[wide ? W.order[i] : (uint32_t)i].id];
    float* o = &out.tris[(size_t)i * 12];
    o[0] = t.v0[0], o[1] = t.v0[1], o[2] = t.v0[2], o[3] = u2f(t.prim);
    o[4] = t.v1[0] - t.v0[0], o[5] = t.v1[1] - t.v0[1], o[6] = t.v1[2] - t.v0[2];
    o[7] = u2f(t.object);
    o[8] = t.v2[0] - t.v0[0], o[9] = t.v2[1] - t.v0[1], o[10] = t.v2[2] - t.v0[2];
    o[11] = u2f(t.material);
  }
  if (need_null) {
    float* o = &out.tris[n * 12];
    for (int k = 0; k < 12; ++k) o[k] = 0.f;
    o[3] = u2f(0xffffffffu);
  }

  auto leaf_code = [](uint32_t first, uint32_t count) -> uint32_t {
    return ~((first << 3) | (count - 1u));
  };
  auto write_child = [&](float* nd, int c, const BuildNode& ch, uint32_t code) {
    const Box& b = ch.box;
    if (c == 0) {
      nd[0] = pad_lo(b.lo[0], b.hi[0]), nd[1] = pad_hi(b.lo[0], b.hi[0]);
      nd[2] = pad_lo(b.lo[1], b.hi[1]), nd[3] = pad_hi(b.lo[1], b.hi[1]);
      nd[8] = pad_lo(b.lo[2], b.hi[2]), nd[9] = pad_hi(b.lo[2], b.hi[2]);
      nd[12] = u2f(code);
    } else {
      nd[4] = pad_lo(b.lo[0], b.hi[0]), nd[5] = pad_hi(b.lo[0], b.hi[0]);
      nd[6] = pad_lo(b.lo[1], b.hi[1]), nd[7] = pad_hi(b.lo[1], b.hi[1]);
      nd[10] = pad_lo(b.lo[2], b.hi[2]), nd[11] = pad_hi(b.lo[2], b.hi[2]);
      nd[13] = u2f(code);
    }
  };

  if (root_is_leaf) {
    out.n_nodes = 1;
    out.nodes.assign(16, 0.f);
    write_child(out.nodes.data(), 0, B.nodes[root], leaf_code(0, (uint32_t)n));
    write_child(out.nodes.data(), 1, B.nodes[root], leaf_code((uint32_t)n, 1));
    out.depth = 1;
    return;
  }

  // DFS pre-order index of an inner node = its parent's + 1 (left child) or + 1 + the left
  // subtree's inner-node count (right child): known without a sequential walk, so the nodes are
  // emitted by parallel tasks
  const uint32_t n_inner = B.nodes[root].inner;
  out.n_nodes = n_inner;
  out.nodes.resize((size_t)n_inner * 16);
  struct Emit {
    const Builder& B;
    const WideBuilder& W;
    FlatBVH& out;
    bool wide;
    void run(uint32_t ni, uint32_t index) const
    {
      for (;;) {
        const BuildNode& nd = B.nodes[ni];
        const BuildNode& l = B.nodes[nd.left];
        const BuildNode& r = B.nodes[nd.right];
        const uint32_t li = index + 1u, ri = index + 1u + l.inner;
        const uint32_t lf = wide ? W.new_first[nd.left] : l.first, rf = wide ? W.new_first[nd.right] : r.first;
        float* o = &out.nodes[(size_t)index * 16];
        o[14] = o[15] = 0.f;
        write_child(o, 0, l, l.count ? leaf_code(lf, l.count) : li);
        write_child(o, 1, r, r.count ? leaf_code(rf, r.count) : ri);
        const bool go_l = l.count == 0, go_r = r.count == 0;
        if (go_l && go_r) {
          if (r.inner >= 2048u) {
            const uint32_t rn = nd.right;
#pragma omp task default(shared) firstprivate(rn, ri)
            run(rn, ri);
          } else {
            run(nd.right, ri);
          }
          ni = nd.left;
          index = li;
        } else if (go_l) {
          ni = nd.left;
          index = li;
        } else if (go_r) {
          ni = nd.right;
          index = ri;
        } else {
          return;
        }
      }
    }
    static uint32_t leaf_code(uint32_t first, uint32_t count) { return ~((first << 3) | (count - 1u)); }
    static void write_child(float* nd, int c, const BuildNode& ch, uint32_t code)
    {
      const Box& b = ch.box;
      const int o = c == 0 ? 0 : 4, z = c == 0 ? 8 : 10;
      nd[o + 0] = pad_lo(b.lo[0], b.hi[0]), nd[o + 1] = pad_hi(b.lo[0], b.hi[0]);
      nd[o + 2] = pad_lo(b.lo[1], b.hi[1]), nd[o + 3] = pad_hi(b.lo[1], b.hi[1]);
      nd[z + 0] = pad_lo(b.lo[2], b.hi[2]), nd[z + 1] = pad_hi(b.lo[2], b.hi[2]);
      nd[12 + c] = u2f(code);
    }
  };
  const Emit emit{B, W, out, wide};
#pragma omp parallel
  {
#pragma omp single
    emit.run(root, 0u);
  }
  const float root_area = B.nodes[root].box.area();
  out.sah_cost = root_area > 0.f ? (double)B.nodes[root].sah / root_area : 0.0;
  out.depth = B.nodes[root].height + 1;
  lap("flatten");
}


// ------------------------------------------------------------------- validation
namespace {
struct DBox {
  double lo[3], hi[3];
  void reset()
  {
    for (int a = 0; a < 3; ++a) lo[a] = 1e300, hi[a] = -1e300;
  }
  void grow(const DBox& b)
  {
    for (int a = 0; a < 3; ++a) lo[a] = std::min(lo[a], b.lo[a]), hi[a] = std::max(hi[a], b.hi[a]);
  }
  bool inside(const DBox& outer) const
  {
    for (int a = 0; a < 3; ++a)
      if (lo[a] < outer.lo[a] || hi[a] > outer.hi[a]) return false;
    return true;
  }
};

struct Validator {
  const FlatBVH& b;
  std::vector<uint8_t> ref2, ref8;
  uint64_t bad = 0;
  explicit Validator(const FlatBVH& f) : b(f), ref2(f.n_tris, 0), ref8(f.n_tris, 0) {}

  DBox tri_box(uint32_t t) const
  {
    const float* o = &b.tris[(size_t)t * 12];
    DBox r;
    r.reset();
    for (int k = 0; k < 3; ++k) {
      for (int a = 0; a < 3; ++a) {
        // the vertices as the kernel sees them: v0, v0 + e1, v0 + e2
        const double v = k == 0 ? (double)o[a] : (double)o[a] + (double)o[4 * k + a];
        r.lo[a] = std::min(r.lo[a], v);
        r.hi[a] = std::max(r.hi[a], v);
      }
    }
    return r;
  }

  // binary tree: content box of a child reference, checked against the box stored in the parent
  DBox walk2(int32_t ref, int depth)
  {
    DBox content;
    content.reset();
    if (depth > 64) {
      ++bad;
      return content;
    }
    if (ref < 0) {
      const uint32_t code = (uint32_t)~ref;
      const uint32_t first = code >> 3, count = (code & 7u) + 1u;
      for (uint32_t k = 0; k < count; ++k) {
        if (first + k >= b.n_tris) {
          ++bad;
          continue;
        }
        if (ref2[first + k]++) ++bad;
        content.grow(tri_box(first + k));
      }
      return content;
    }
    if ((uint32_t)ref >= b.n_nodes) {
      ++bad;
      return content;
    }
    const float* nd = &b.nodes[(size_t)ref * 16];
    for (int c = 0; c < 2; ++c) {
      DBox stored;
      const int o = c * 4, z = 8 + c * 2;
      stored.lo[0] = nd[o + 0], stored.hi[0] = nd[o + 1];
      stored.lo[1] = nd[o + 2], stored.hi[1] = nd[o + 3];
      stored.lo[2] = nd[z + 0], stored.hi[2] = nd[z + 1];
      int32_t child;
      std::memcpy(&child, &nd[12 + c], 4);
      const DBox cc = walk2(child, depth + 1);
      if (cc.lo[0] <= cc.hi[0] && !cc.inside(stored)) ++bad;
      content.grow(cc);
    }
    return content;
  }

  DBox walk8(uint32_t node, uint32_t depth)
  {
    DBox content;
    content.reset();
    if (node >= b.n_nodes8 || depth > b.depth8) {
      ++bad;
      return content;
    }
    const uint32_t* w = &b.nodes8[(size_t)node * 20];
    float p[3];
    std::memcpy(p, w, 12);
    double scale[3];
    for (int a = 0; a < 3; ++a) scale[a] = std::ldexp(1.0, (int)((w[3] >> (8 * a)) & 0xffu) + 8 - 127);
    const uint32_t imask = w[3] >> 24;
    const uint8_t* meta = reinterpret_cast<const uint8_t*>(w + 6);
    const uint8_t* q = reinterpret_cast<const uint8_t*>(w + 8);
    for (int s = 0; s < 8; ++s) {
      const uint32_t m = meta[s];
      if (m == 0) {
        if (imask & (1u << s)) ++bad;
        continue;
      }
      DBox stored;
      for (int a = 0; a < 3; ++a) {
        stored.lo[a] = (double)p[a] + q[8 * a + s] * scale[a];
        stored.hi[a] = (double)p[a] + q[8 * (3 + a) + s] * scale[a];
      }
      DBox cc;
      cc.reset();
      if ((m & 31u) >= 24u) {
        if ((m & 31u) != 24u + (uint32_t)s || (m >> 5) != 1u || !(imask & (1u << s))) ++bad;
        uint32_t rel = 0;
        for (int k = 0; k < s; ++k) rel += (imask >> k) & 1u;
        cc = walk8(w[4] + rel, depth + 1);
      } else {
        if (imask & (1u << s)) ++bad;
        const uint32_t un = m >> 5, count = un == 1 ? 1 : (un == 3 ? 2 : (un == 7 ? 3 : 0));
        if (count == 0) ++bad;
        for (uint32_t k = 0; k < count; ++k) {
          const uint32_t t = w[5] + (m & 31u) + k;
          if (t >= b.n_tris || (m & 31u) + k >= 24u) {
            ++bad;
            continue;
          }
          if (ref8[t]++) ++bad;
          cc.grow(tri_box(t));
        }
      }
      if (cc.lo[0] <= cc.hi[0] && !cc.inside(stored)) ++bad;
      content.grow(cc);
    }
    return content;
  }
};
} // namespace

uint64_t validate_bvh(const FlatBVH& bvh)
{
  if (bvh.n_tris == 0) return 0;
  Validator v(bvh);
  const uint32_t real_tris = (uint32_t)(bvh.tris.size() / 12);
  (void)real_tris;
  v.walk2(0, 0);
  // the binary root-leaf form carries one null triangle nothing else references
  uint32_t null_tri = 0xffffffffu;
  if (bvh.n_nodes == 1) {
    uint32_t last;
    std::memcpy(&last, &bvh.tris[(size_t)(bvh.n_tris - 1) * 12 + 3], 4);
    if (last == 0xffffffffu) null_tri = bvh.n_tris - 1;
  }
  for (uint32_t t = 0; t < bvh.n_tris; ++t)
    if (v.ref2[t] != 1) ++v.bad;
  if (bvh.n_nodes8 != 0) {
    v.walk8(0, 0);
    for (uint32_t t = 0; t < bvh.n_tris; ++t)
      if (v.ref8[t] != (t == null_tri ? 0 : 1)) ++v.bad;
  }
  return v.bad;
}


// ------------------------------------------------------------- traversal statistics
// Host walk of either tree in the device kernels' visiting order (binary: nearer child first,
// farther child stacked; wide: octant-ordered node groups), counting what a ray costs.  A tool for
// choosing between tree layouts without GPU time: it does not compute images and nothing in the
// product path calls it.
namespace {
inline float safe_inv_h(float x)
{
  const float e = 8.271806125530277e-25f;
  return 1.0f / (std::fabs(x) > e ? x : std::copysign(e, x));
}
struct RayH {
  float o[3], d[3], id[3], tmin, tbest;
};
inline bool tri_hit(const float* t, RayH& r)
{
  const float* v0 = t;
  const float* e1 = t + 4;
  const float* e2 = t + 8;
  const float hx = r.d[1] * e2[2] - e2[1] * r.d[2], hy = r.d[2] * e2[0] - e2[2] * r.d[0],
              hz = r.d[0] * e2[1] - e2[0] * r.d[1];
  const float a = e1[0] * hx + e1[1] * hy + e1[2] * hz;
  if (a > -1e-7f && a < 1e-7f) return false;
  const float f = 1.0f / a;
  const float sx = r.o[0] - v0[0], sy = r.o[1] - v0[1], sz = r.o[2] - v0[2];
  const float u = f * (sx * hx + sy * hy + sz * hz);
  if (u < 0.f || u > 1.f) return false;
  const float qx = sy * e1[2] - e1[1] * sz, qy = sz * e1[0] - e1[2] * sx, qz = sx * e1[1] - e1[0] * sy;
  const float v = f * (r.d[0] * qx + r.d[1] * qy + r.d[2] * qz);
  if (v < 0.f || u + v > 1.f) return false;
  const float tt = f * (e2[0] * qx + e2[1] * qy + e2[2] * qz);
  if (tt < r.tmin || tt > r.tbest) return false;
  r.tbest = tt;
  return true;
}
inline bool slab(const float lo[3], const float hi[3], const RayH& r, float& tnear)
{
  float t0 = r.tmin, t1 = r.tbest;
  for (int a = 0; a < 3; ++a) {
    const float x0 = (lo[a] - r.o[a]) * r.id[a], x1 = (hi[a] - r.o[a]) * r.id[a];
    t0 = std::max(t0, std::min(x0, x1));
    t1 = std::min(t1, std::max(x0, x1));
  }
  tnear = t0;
  return t1 * 1.0000004f >= t0;
}
} // namespace

void trace_stats(const FlatBVH& b, const float* rays8, uint64_t n_rays, int wide, uint64_t out[5])
{
  uint64_t inner = 0, leaves = 0, tests = 0, hits = 0, max_stack = 0;
  if (b.n_tris != 0) {
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : inner, leaves, tests, hits) reduction(max : max_stack)
    for (long long i = 0; i < (long long)n_rays; ++i) {
      RayH r;
      const float* q = rays8 + 8 * i;
      for (int a = 0; a < 3; ++a) r.o[a] = q[a], r.d[a] = q[4 + a], r.id[a] = safe_inv_h(q[4 + a]);
      r.tmin = q[3];
      r.tbest = q[7];
      bool any = false;
      if (!wide) {
        int32_t stack[128];
        int sp = 0;
        int32_t node = 0;
        for (;;) {
          if (node >= 0) {
            ++inner;
            const float* nd = &b.nodes[(size_t)node * 16];
            const float l0[3] = {nd[0], nd[2], nd[8]}, h0[3] = {nd[1], nd[3], nd[9]};
            const float l1[3] = {nd[4], nd[6], nd[10]}, h1[3] = {nd[5], nd[7], nd[11]};
            float n0, n1;
            const bool t0 = slab(l0, h0, r, n0), t1 = slab(l1, h1, r, n1);
            int32_t c0, c1;
            std::memcpy(&c0, &nd[12], 4);
            std::memcpy(&c1, &nd[13], 4);
            if (!t0 && !t1) {
              if (sp == 0) break;
              node = stack[--sp];
            } else {
              const bool swap = t1 && (!t0 || n1 < n0);
              node = swap ? c1 : c0;
              if (t0 && t1 && sp < 128) stack[sp++] = swap ? c0 : c1;
              max_stack = std::max<uint64_t>(max_stack, (uint64_t)sp);
            }
          } else {
            ++leaves;
            const uint32_t code = (uint32_t)~node, first = code >> 3, count = (code & 7u) + 1u;
            for (uint32_t k = 0; k < count; ++k) {
              ++tests;
              if (tri_hit(&b.tris[(size_t)(first + k) * 12], r)) any = true;
            }
            if (sp == 0) break;
            node = stack[--sp];
          }
        }
      } else if (wide == 2) {
        // virtual 4-wide tree: a node = a binary node fused with its inner children (every other
        // level collapsed); children tested together, visited nearest first
        int32_t stack[256];
        int sp = 0;
        int32_t node = 0;
        for (;;) {
          if (node >= 0) {
            ++inner;
            int32_t ch[4];
            float lo[4][3], hi[4][3];
            int nc = 0;
            const float* nd = &b.nodes[(size_t)node * 16];
            for (int c = 0; c < 2; ++c) {
              int32_t ref;
              std::memcpy(&ref, &nd[12 + c], 4);
              if (ref >= 0) { // inner child: take its two children instead
                const float* gd = &b.nodes[(size_t)ref * 16];
                for (int g = 0; g < 2; ++g) {
                  std::memcpy(&ch[nc], &gd[12 + g], 4);
                  const int o = g * 4, z = 8 + g * 2;
                  lo[nc][0] = gd[o], hi[nc][0] = gd[o + 1], lo[nc][1] = gd[o + 2], hi[nc][1] = gd[o + 3];
                  lo[nc][2] = gd[z], hi[nc][2] = gd[z + 1];
                  ++nc;
                }
              } else {
                ch[nc] = ref;
                const int o = c * 4, z = 8 + c * 2;
                lo[nc][0] = nd[o], hi[nc][0] = nd[o + 1], lo[nc][1] = nd[o + 2], hi[nc][1] = nd[o + 3];
                lo[nc][2] = nd[z], hi[nc][2] = nd[z + 1];
                ++nc;
              }
            }
            float tn[4];
            int order4[4], nh = 0;
            for (int c = 0; c < nc; ++c)
              if (slab(lo[c], hi[c], r, tn[c])) order4[nh++] = c;
            std::sort(order4, order4 + nh, [&](int x, int y) { return tn[x] > tn[y]; }); // far first
            if (nh == 0) {
              if (sp == 0) break;
              node = stack[--sp];
            } else {
              for (int k = 0; k + 1 < nh && sp < 256; ++k) stack[sp++] = ch[order4[k]];
              node = ch[order4[nh - 1]];
              max_stack = std::max<uint64_t>(max_stack, (uint64_t)sp);
            }
          } else {
            ++leaves;
            const uint32_t code = (uint32_t)~node, first = code >> 3, count = (code & 7u) + 1u;
            for (uint32_t k = 0; k < count; ++k) {
              ++tests;
              if (tri_hit(&b.tris[(size_t)(first + k) * 12], r)) any = true;
            }
            if (sp == 0) break;
            node = stack[--sp];
          }
        }
      } else if (b.n_nodes8 != 0) {
        struct G {
          uint32_t base, bits;
        } stack[64];
        int sp = 0;
        const uint32_t octm = (r.id[0] < 0.f ? 0u : 1u) | (r.id[1] < 0.f ? 0u : 2u) | (r.id[2] < 0.f ? 0u : 4u);
        G ng{0u, 0x80000000u};
        for (;;) {
          // pop the highest-priority child of the current node group
          const int bit = 31 - __builtin_clz(ng.bits);
          const uint32_t rest = ng.bits & ~(1u << bit);
          const uint32_t slot = (uint32_t)(bit - 24) ^ octm;
          const uint32_t node = ng.base + (uint32_t)__builtin_popcount(ng.bits & 0xffu & ((1u << slot) - 1u));
          const G rest_g{ng.base, rest};
          ++inner;
          const uint32_t* w = &b.nodes8[(size_t)node * 20];
          float p[3];
          std::memcpy(p, w, 12);
          float scale[3];
          for (int a = 0; a < 3; ++a) scale[a] = std::ldexp(1.0f, (int)((w[3] >> (8 * a)) & 0xffu) + 8 - 127);
          const uint32_t imask = w[3] >> 24;
          const uint8_t* meta = reinterpret_cast<const uint8_t*>(w + 6);
          const uint8_t* qq = reinterpret_cast<const uint8_t*>(w + 8);
          uint32_t hm = 0;
          for (int s = 0; s < 8; ++s) {
            if (meta[s] == 0) continue;
            float lo[3], hi[3], tn;
            for (int a = 0; a < 3; ++a) {
              lo[a] = p[a] + qq[8 * a + s] * scale[a];
              hi[a] = p[a] + qq[8 * (3 + a) + s] * scale[a];
            }
            if (!slab(lo, hi, r, tn)) continue;
            uint32_t bi = meta[s] & 31u;
            if (bi >= 24u) bi ^= octm;
            hm |= (uint32_t)(meta[s] >> 5) << bi;
          }
          uint32_t tri_bits = hm & 0x00ffffffu;
          if (tri_bits) ++leaves;
          while (tri_bits) {
            const int k = __builtin_ctz(tri_bits);
            tri_bits &= tri_bits - 1;
            ++tests;
            if (tri_hit(&b.tris[(size_t)(w[5] + (uint32_t)k) * 12], r)) any = true;
          }
          const G ng2{w[4], (hm & 0xff000000u) | imask};
          if (ng2.bits & 0xff000000u) {
            if ((rest & 0xff000000u) && sp < 64) stack[sp++] = rest_g;
            ng = ng2;
          } else if (rest & 0xff000000u) {
            ng = rest_g;
          } else {
            if (sp == 0) break;
            ng = stack[--sp];
          }
          max_stack = std::max<uint64_t>(max_stack, (uint64_t)sp);
        }
      }
      if (any) ++hits;
    }
  }
  out[0] = inner, out[1] = leaves, out[2] = tests, out[3] = hits, out[4] = max_stack;
}

} // namespace pt
