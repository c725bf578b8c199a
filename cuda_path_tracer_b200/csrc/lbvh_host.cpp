// lbvh_host.cpp — sequential host restatement of the device LBVH builder (lbvh.cu): the same
// lbvh.h functions, plain loops instead of kernels, std::stable_sort instead of the radix sort.
// It exists so that the construction logic is checked on CPU (tests/test_host_bvh.py) and so
// that the device result can be compared node for node on the GPU box; the product path builds
// on the device.
#include "bvh_build.h"
#include "lbvh.h"

#include <algorithm>
#include <cfloat>
#include <cstring>
#include <numeric>

namespace pt {

using namespace lbvh;

static inline float u2f(uint32_t u)
{
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}

bool build_lbvh_host(const BuildTris& tris, FlatBVH& out)
{
  out = FlatBVH{};
  const int n = (int)tris.size();
  const int kLeafMax = leaf_max_setting();
  if (n <= 4) return false; // tiny scenes stay with the SAH builder (root-leaf form)

  // primitive boxes, centroid bounds
  std::vector<Box6> pbox(n);
  float clo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, chi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (int i = 0; i < n; ++i) {
    const BuildTri& t = tris[i];
    for (int a = 0; a < 3; ++a) {
      pbox[i].lo[a] = std::min(t.v0[a], std::min(t.v1[a], t.v2[a]));
      pbox[i].hi[a] = std::max(t.v0[a], std::max(t.v1[a], t.v2[a]));
      const float c = 0.5f * (pbox[i].lo[a] + pbox[i].hi[a]);
      clo[a] = std::min(clo[a], c);
      chi[a] = std::max(chi[a], c);
    }
  }
  // ONE scale for the three axes (a cubic grid): per-axis scales would make a flat axis split
  // as often as the long ones and slice a terrain into iso-height bands (measured: 4x fewer rays/s)
  float inv[3];
  const float ext = std::max(chi[0] - clo[0], std::max(chi[1] - clo[1], chi[2] - clo[2]));
  for (int a = 0; a < 3; ++a) inv[a] = ext > 0.f ? kMortonScale / ext : 0.f;

  // Morton codes, stable sort by code (== radix sort of (code, index) pairs)
  std::vector<uint64_t> code(n);
  std::vector<uint32_t> order(n);
  for (int i = 0; i < n; ++i) {
    const float c[3] = {0.5f * (pbox[i].lo[0] + pbox[i].hi[0]), 0.5f * (pbox[i].lo[1] + pbox[i].hi[1]),
                        0.5f * (pbox[i].lo[2] + pbox[i].hi[2])};
    code[i] = morton63(c, clo, inv);
  }
  std::iota(order.begin(), order.end(), 0u);
  std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return code[a] < code[b]; });
  std::vector<uint64_t> sorted(n);
  for (int i = 0; i < n; ++i) sorted[i] = code[order[i]];

  // hierarchy
  const int ni = n - 1;
  std::vector<int> first(ni), last(ni), split(ni);
  for (int i = 0; i < ni; ++i) node_range(sorted.data(), n, i, first[i], last[i], split[i]);

  // bottom-up boxes and depths (post-order walk from the root)
  std::vector<Box6> nbox(ni);
  std::vector<int> depth(ni, 0);
  {
    std::vector<std::pair<int, int>> stack; // (node, phase)
    stack.push_back({0, 0});
    while (!stack.empty()) {
      auto [i, phase] = stack.back();
      stack.pop_back();
      const bool lleaf = first[i] == split[i], rleaf = last[i] == split[i] + 1;
      if (phase == 0) {
        stack.push_back({i, 1});
        if (!lleaf) stack.push_back({split[i], 0});
        if (!rleaf) stack.push_back({split[i] + 1, 0});
      } else {
        const Box6& lb = lleaf ? pbox[order[split[i]]] : nbox[split[i]];
        const Box6& rb = rleaf ? pbox[order[split[i] + 1]] : nbox[split[i] + 1];
        for (int a = 0; a < 3; ++a) {
          nbox[i].lo[a] = std::min(lb.lo[a], rb.lo[a]);
          nbox[i].hi[a] = std::max(lb.hi[a], rb.hi[a]);
        }
        depth[i] = 1 + std::max(lleaf ? 0 : depth[split[i]], rleaf ? 0 : depth[split[i] + 1]);
      }
    }
  }
  if (depth[0] > 60) return false; // the traversal stack holds 64 entries

  // compact the real inner nodes (range larger than a leaf)
  std::vector<uint32_t> index(ni);
  uint32_t n_real = 0;
  for (int i = 0; i < ni; ++i) {
    index[i] = n_real;
    if (last[i] - first[i] + 1 > kLeafMax) ++n_real;
  }
  out.n_nodes = n_real;
  out.nodes.resize((size_t)n_real * 16);
  for (int i = 0; i < ni; ++i) {
    if (last[i] - first[i] + 1 <= kLeafMax) continue;
    float* nd = &out.nodes[(size_t)index[i] * 16];
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
  }

  // triangles in sorted order
  out.n_tris = (uint32_t)n;
  out.tris.resize((size_t)n * 12);
  for (int i = 0; i < n; ++i) {
    const BuildTri& t = tris[order[i]];
    float* o = &out.tris[(size_t)i * 12];
    o[0] = t.v0[0], o[1] = t.v0[1], o[2] = t.v0[2], o[3] = u2f(t.prim);
    o[4] = t.v1[0] - t.v0[0], o[5] = t.v1[1] - t.v0[1], o[6] = t.v1[2] - t.v0[2], o[7] = u2f(t.object);
    o[8] = t.v2[0] - t.v0[0], o[9] = t.v2[1] - t.v0[1], o[10] = t.v2[2] - t.v0[2], o[11] = u2f(t.material);
  }
  for (int a = 0; a < 3; ++a) {
    const float lo = nbox[0].lo[a], hi = nbox[0].hi[a];
    out.root_lo[a] = pad_lo(pad_lo(lo, hi), hi);
    out.root_hi[a] = pad_hi(lo, pad_hi(lo, hi));
  }
  out.depth = (uint32_t)depth[0] + 1;
  return true;
}

} // namespace pt
