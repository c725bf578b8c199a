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
#include <cstring>

namespace pt {
namespace {

constexpr int kBins = 16;
constexpr int kLeafMax = 4;
constexpr int kSahDepthLimit = 32; // beyond this depth: balanced median splits
constexpr float kTravCost = 1.0f;
constexpr float kIsectCost = 1.0f;
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
  Box box;  // bounds of the primitives
  Box cbox; // bounds of their centroids
  uint32_t count;
  void reset()
  {
    box.reset();
    cbox.reset();
    count = 0;
  }
  void merge(const Bin& o)
  {
    box.grow(o.box);
    cbox.grow(o.cbox);
    count += o.count;
  }
};
struct BinSet {
  Bin bin[3][kBins];
  void reset()
  {
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < kBins; ++b) bin[a][b].reset();
  }
};

struct BuildNode {
  Box box;
  uint32_t left = 0, right = 0;  // children (inner) ...
  uint32_t first = 0, count = 0; // ... or primitive range (leaf, count > 0)
};

struct Builder {
  std::vector<Prim> prims;
  std::vector<BuildNode> nodes;
  std::atomic<uint32_t> n_nodes{0};
  std::atomic<uint32_t> max_depth{0};

  uint32_t alloc_node() { return n_nodes.fetch_add(1, std::memory_order_relaxed); }

  void make_leaf(uint32_t ni, uint32_t first, uint32_t count, int depth)
  {
    nodes[ni].first = first;
    nodes[ni].count = count;
    uint32_t d = (uint32_t)depth, cur = max_depth.load(std::memory_order_relaxed);
    while (d > cur && !max_depth.compare_exchange_weak(cur, d)) {}
  }

  static void bin_range(const Prim* p, uint32_t n, const Box& cb, const float* scale, BinSet& out)
  {
    for (uint32_t i = 0; i < n; ++i) {
      const Prim& q = p[i];
      float c[3] = {q.c(0), q.c(1), q.c(2)};
      for (int a = 0; a < 3; ++a) {
        if (scale[a] == 0.f) continue;
        int b = (int)((c[a] - cb.lo[a]) * scale[a]);
        b = b < 0 ? 0 : (b >= kBins ? kBins - 1 : b);
        Bin& bn = out.bin[a][b];
        bn.box.grow(q.b);
        bn.cbox.grow(c);
        bn.count++;
      }
    }
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

  void build(uint32_t ni, uint32_t first, uint32_t count, int depth, const Box& nb, const Box& cb)
  {
    nodes[ni].box = nb;
    if (count == 1) {
      make_leaf(ni, first, count, depth);
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
      bins.reset();
      if (count >= kParallelBin) {
        const int chunks = (int)std::min<uint32_t>(64, count / (kParallelBin / 4));
        std::vector<BinSet> part(chunks);
        const Prim* base = prims.data() + first;
#pragma omp taskloop default(shared) grainsize(1)
        for (int c = 0; c < chunks; ++c) {
          const uint32_t b0 = (uint32_t)((uint64_t)count * c / chunks);
          const uint32_t b1 = (uint32_t)((uint64_t)count * (c + 1) / chunks);
          part[c].reset();
          bin_range(base + b0, b1 - b0, cb, scale, part[c]);
        }
        for (int c = 0; c < chunks; ++c)
          for (int a = 0; a < 3; ++a)
            for (int b = 0; b < kBins; ++b) bins.bin[a][b].merge(part[c].bin[a][b]);
      } else {
        bin_range(prims.data() + first, count, cb, scale, bins);
      }

      float best_cost = FLT_MAX;
      int best_axis = -1, best_bin = -1;
      for (int axis = 0; axis < 3; ++axis) {
        if (scale[axis] == 0.f) continue;
        float right_area[kBins];
        uint32_t right_cnt[kBins];
        Box acc;
        acc.reset();
        uint32_t c = 0;
        for (int b = kBins - 1; b > 0; --b) {
          acc.grow(bins.bin[axis][b].box);
          c += bins.bin[axis][b].count;
          right_area[b] = acc.area();
          right_cnt[b] = c;
        }
        acc.reset();
        c = 0;
        for (int b = 0; b < kBins - 1; ++b) {
          acc.grow(bins.bin[axis][b].box);
          c += bins.bin[axis][b].count;
          if (c == 0 || right_cnt[b + 1] == 0) continue;
          const float cost = acc.area() * (float)c + right_area[b + 1] * (float)right_cnt[b + 1];
          if (cost < best_cost) {
            best_cost = cost;
            best_axis = axis;
            best_bin = b;
          }
        }
      }
      if (best_axis >= 0) {
        const float area = nb.area();
        const float split_cost = kTravCost + (area > 0.f ? best_cost / area : 0.f) * kIsectCost;
        const float leaf_cost = (float)count * kIsectCost;
        if (count <= (uint32_t)kLeafMax && leaf_cost <= split_cost) {
          make_leaf(ni, first, count, depth);
          return;
        }
        const float sc = scale[best_axis];
        const float lo = cb.lo[best_axis];
        const int ax = best_axis, bb = best_bin;
        Prim* b0 = prims.data() + first;
        Prim* m = std::partition(b0, b0 + count, [=](const Prim& q) {
          int b = (int)((q.c(ax) - lo) * sc);
          b = b < 0 ? 0 : (b >= kBins ? kBins - 1 : b);
          return b <= bb;
        });
        mid = (uint32_t)(m - prims.data());
        have_split = mid > first && mid < first + count;
        if (have_split) {
          // child bounds come from the bins: no extra pass over the primitives
          lbox.reset(), lcb.reset(), rbox.reset(), rcb.reset();
          for (int b = 0; b < kBins; ++b) {
            const Bin& bn = bins.bin[best_axis][b];
            if (bn.count == 0) continue;
            (b <= best_bin ? lbox : rbox).grow(bn.box);
            (b <= best_bin ? lcb : rcb).grow(bn.cbox);
          }
        }
      }
    }
    if (!have_split) {
      if (count <= (uint32_t)kLeafMax) {
        make_leaf(ni, first, count, depth);
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

    const uint32_t l = alloc_node();
    const uint32_t r = alloc_node();
    nodes[ni].left = l;
    nodes[ni].right = r;
    nodes[ni].count = 0;
    const uint32_t lc = mid - first, rc = first + count - mid;
    if (count >= kTaskMin) {
#pragma omp task default(shared) firstprivate(l, first, lc, depth, lbox, lcb)
      build(l, first, lc, depth + 1, lbox, lcb);
#pragma omp task default(shared) firstprivate(r, mid, rc, depth, rbox, rcb)
      build(r, mid, rc, depth + 1, rbox, rcb);
#pragma omp taskwait
    } else {
      build(l, first, lc, depth + 1, lbox, lcb);
      build(r, mid, rc, depth + 1, rbox, rcb);
    }
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

} // namespace

void build_bvh(const std::vector<BuildTri>& tris, FlatBVH& out)
{
  out = FlatBVH{};
  const size_t n = tris.size();
  if (n == 0) return;

  Builder B;
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
  const uint32_t root = B.alloc_node();
#pragma omp parallel
  {
#pragma omp single
    B.build(root, 0, (uint32_t)n, 0, root_box, root_cb);
  }

  // ---- flatten: inner nodes in DFS pre-order, triangles in leaf order
  const bool root_is_leaf = B.nodes[root].count != 0;
  for (int a = 0; a < 3; ++a) {
    const float lo = B.nodes[root].box.lo[a], hi = B.nodes[root].box.hi[a];
    out.root_lo[a] = pad_lo(pad_lo(lo, hi), hi);
    out.root_hi[a] = pad_hi(lo, pad_hi(lo, hi));
  }
  const bool need_null = root_is_leaf;
  out.n_tris = (uint32_t)n + (need_null ? 1u : 0u);
  out.tris.resize((size_t)out.n_tris * 12);
#pragma omp parallel for schedule(static)
  for (long long i = 0; i < (long long)n; ++i) {
    const BuildTri& t = tris[B.prims[i].id];
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

  // count inner nodes and assign pre-order indices iteratively
  const uint32_t total = B.n_nodes.load();
  std::vector<uint32_t> inner_index(total, 0xffffffffu);
  std::vector<uint32_t> stack;
  stack.reserve(128);
  uint32_t n_inner = 0;
  stack.push_back(root);
  while (!stack.empty()) {
    const uint32_t ni = stack.back();
    stack.pop_back();
    const BuildNode& nd = B.nodes[ni];
    if (nd.count != 0) continue;
    inner_index[ni] = n_inner++;
    stack.push_back(nd.right);
    stack.push_back(nd.left);
  }
  out.n_nodes = n_inner;
  out.nodes.assign((size_t)n_inner * 16, 0.f);
  double sah = 0.0;
  const double root_area = B.nodes[root].box.area();
#pragma omp parallel for schedule(static) reduction(+ : sah)
  for (long long ni = 0; ni < (long long)total; ++ni) {
    const BuildNode& nd = B.nodes[ni];
    if (nd.count != 0) {
      if (root_area > 0) sah += (double)nd.box.area() / root_area * nd.count * kIsectCost;
      continue;
    }
    if (inner_index[ni] == 0xffffffffu) continue;
    if (root_area > 0) sah += (double)nd.box.area() / root_area * kTravCost;
    float* o = &out.nodes[(size_t)inner_index[ni] * 16];
    const BuildNode& l = B.nodes[nd.left];
    const BuildNode& r = B.nodes[nd.right];
    write_child(o, 0, l, l.count ? leaf_code(l.first, l.count) : inner_index[nd.left]);
    write_child(o, 1, r, r.count ? leaf_code(r.first, r.count) : inner_index[nd.right]);
  }
  out.sah_cost = sah;
  out.depth = B.max_depth.load() + 1;
}

} // namespace pt
