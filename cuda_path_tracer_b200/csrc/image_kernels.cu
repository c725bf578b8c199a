// image_kernels.cu — per-pixel kernels: stable compaction (streaming-compat mode),
// accumulate, resolve/tonemap, frame import/export, A-Trous denoiser.
//   accumulate  final_gather                                 (reference path_tracer.cu:203-219, 317-330)
//   resolve     preview kernels / linear_to_gamma            (path_tracer.cu:221-225, 334-385)
//   atrous      denoising_kernel                             (denoising/...denoiser.cu:24-86)
#include "kernels.h"

#include <algorithm>
#include <stdlib.h>

#include <float.h>

namespace pt {

PT_D float4 ldg4(const float4* p) { return __ldg(p); }
PT_D float4 mk4(f3 v, float w) { return make_float4(v.x, v.y, v.z, w); }

// ========================================================= stable compaction
// PT_RNG_SLOT_RESEED only: reproduces thrust::stable_partition's order of the
// live paths (path_tracer.cu:454-457) on 4-byte ids instead of 65-byte records.
#define SC_THREADS 256
#define SC_ITEMS 2048 // per block

__global__ void __launch_bounds__(SC_THREADS)
sc_count_kernel(const uint8_t* __restrict__ flags, const uint32_t* __restrict__ n_ptr,
                uint32_t n_first, uint32_t* __restrict__ block_sums)
{
  const uint32_t n = n_ptr ? *n_ptr : n_first;
  const uint32_t begin = blockIdx.x * SC_ITEMS;
  uint32_t c = 0;
  for (uint32_t i = begin + threadIdx.x; i < begin + SC_ITEMS && i < n; i += SC_THREADS)
    c += flags[i];
  __shared__ uint32_t warp_sums[SC_THREADS / 32];
  for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t s = 0;
    for (int w = 0; w < SC_THREADS / 32; ++w) s += warp_sums[w];
    block_sums[blockIdx.x] = s;
  }
}

__global__ void __launch_bounds__(1024)
sc_scan_kernel(uint32_t* __restrict__ block_sums, uint32_t n_blocks,
               uint32_t* __restrict__ total_out)
{
  // single block exclusive scan, chunks of 1024
  __shared__ uint32_t sh[1024];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < n_blocks; base += 1024) {
    const uint32_t i = base + threadIdx.x;
    const uint32_t v = i < n_blocks ? block_sums[i] : 0u;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (uint32_t o = 1; o < 1024; o <<= 1) {
      uint32_t t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0u;
      __syncthreads();
      sh[threadIdx.x] += t;
      __syncthreads();
    }
    const uint32_t incl = sh[threadIdx.x];
    if (i < n_blocks) block_sums[i] = carry + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total_out = carry;
}

__global__ void __launch_bounds__(SC_THREADS)
sc_scatter_kernel(const uint8_t* __restrict__ flags, const uint32_t* __restrict__ n_ptr,
                  uint32_t n_first, const uint32_t* __restrict__ block_offsets,
                  const uint32_t* __restrict__ queue, uint32_t* __restrict__ next_queue)
{
  const uint32_t n = n_ptr ? *n_ptr : n_first;
  const uint32_t begin = blockIdx.x * SC_ITEMS;
  __shared__ uint32_t warp_sums[SC_THREADS / 32];
  __shared__ uint32_t running;
  if (threadIdx.x == 0) running = block_offsets[blockIdx.x];
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  for (uint32_t chunk = begin; chunk < begin + SC_ITEMS; chunk += SC_THREADS) {
    const uint32_t i = chunk + threadIdx.x;
    const bool f = i < n && flags[i] != 0;
    const uint32_t mask = __ballot_sync(0xffffffffu, f);
    if (lane == 0) warp_sums[warp] = __popc(mask);
    __syncthreads();
    uint32_t off = running;
    for (uint32_t w = 0; w < warp; ++w) off += warp_sums[w];
    if (f) next_queue[off + __popc(mask & ((1u << lane) - 1u))] = queue ? queue[i] : i;
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t s = 0;
      for (int w = 0; w < SC_THREADS / 32; ++w) s += warp_sums[w];
      running += s;
    }
    __syncthreads();
  }
}

// =============================================================== accumulate
__global__ void __launch_bounds__(256)
accumulate_kernel(const PathState ps, const PassParams pp, float4* __restrict__ sum_color,
                  float4* __restrict__ sum_gbuf, uint32_t* __restrict__ counters,
                  uint32_t n_first, uint32_t max_depth,
                  unsigned long long* __restrict__ total_rays)
{
  const uint32_t p = pp.pixel_begin + blockIdx.x * blockDim.x + threadIdx.x;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    // rays = live paths entering extend, summed over bounces (counters[0] is host-known)
    // (bounce-synchronous mode only: chain_kernel counts its own rays)
    if (pp.rng_mode == 1u) {
      unsigned long long r = (unsigned long long)(pp.pixel_end - pp.pixel_begin) * pp.samples;
      for (uint32_t b = 1; b < max_depth; ++b) r += counters[b];
      atomicAdd(total_rays + 0, r);
    }
    // rays that entered the BVH traversal queue (tcounters follow counters and work cursors)
    const uint32_t* tc = counters + 2 * (max_depth + 2);
    unsigned long long tr = 0;
    for (uint32_t b = 0; b < max_depth; ++b) tr += tc[b];
    atomicAdd(total_rays + 1, tr);
  }
  if (p >= pp.pixel_end) return;
  float4 c = sum_color[p];
  float4 g = sum_gbuf[p];
  for (uint32_t s = 0; s < pp.samples; ++s) {
    const size_t pid = (size_t)s * pp.band_pixels + (p - pp.pixel_begin);
    const float4 t = ps.thr[pid];
    const float4 gb = ps.gbuf[pid];
    c.x += t.x;
    c.y += t.y;
    c.z += t.z;
    g.x += gb.x;
    g.y += gb.y;
    g.z += gb.z;
    g.w += gb.w;
  }
  c.w += (float)pp.samples;
  sum_color[p] = c;
  sum_gbuf[p] = g;
}

// ================================================================== resolve
PT_D unsigned char to_255(float v)
{
  // static_cast<unsigned char>(glm::clamp(v, 0.f, 1.f) * 255.99f)
  return (unsigned char)(fminf(fmaxf(v, 0.0f), 1.0f) * 255.99f);
}

PT_D f3 fetch_kind(int kind, const float4* sum_color, const float4* sum_gbuf,
                   const float4* final_rgb, bool final_is_mean, uint32_t p, float& depth)
{
  const float4 c = sum_color[p];
  const float inv_dummy = c.w; // sample count
  depth = 0.f;
  if (kind == 0 && final_is_mean) { // FINAL after denoise
    const float4 f = final_rgb[p];
    return mk3(f.x, f.y, f.z);
  }
  if (kind == 4) {
    const float4 f = final_rgb[p];
    return mk3(f.x, f.y, f.z);
  }
  if (kind == 0 || kind == 1) return mk3(c.x / inv_dummy, c.y / inv_dummy, c.z / inv_dummy);
  const float4 g = sum_gbuf[p];
  depth = g.w / inv_dummy;
  return mk3(g.x / inv_dummy, g.y / inv_dummy, g.z / inv_dummy);
}

__global__ void __launch_bounds__(256)
resolve_kernel(int kind, const float4* __restrict__ sum_color, const float4* __restrict__ sum_gbuf,
               const float4* __restrict__ final_rgb, bool final_is_mean, uint32_t pixels,
               uchar4* __restrict__ out)
{
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= pixels) return;
  float depth;
  f3 c = fetch_kind(kind, sum_color, sum_gbuf, final_rgb, final_is_mean, p, depth);
  unsigned char alpha = 255;
  if (kind == 2) c = c * 0.5f + mk3(0.5f, 0.5f, 0.5f); // neg1_1_to_0_1
  if (kind == 3) {                                     // preview_depth_kernel: 1/depth, alpha 1
    c = mk3(1.0f / depth, 1.0f / depth, 1.0f / depth);
    alpha = 1;
  }
  const float g = 1.0f / 2.2f;
  c = mk3(powf(c.x, g), powf(c.y, g), powf(c.z, g)); // linear_to_gamma
  out[p] = make_uchar4(to_255(c.x), to_255(c.y), to_255(c.z), alpha);
}

__global__ void __launch_bounds__(256)
export_kernel(int kind, const float4* __restrict__ sum_color, const float4* __restrict__ sum_gbuf,
              const float4* __restrict__ final_rgb, bool final_is_mean, uint32_t pixels,
              float* __restrict__ out)
{
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= pixels) return;
  float depth;
  const f3 c = fetch_kind(kind, sum_color, sum_gbuf, final_rgb, final_is_mean, p, depth);
  if (kind == 3) {
    out[p] = depth;
  } else {
    out[3 * (size_t)p + 0] = c.x;
    out[3 * (size_t)p + 1] = c.y;
    out[3 * (size_t)p + 2] = c.z;
  }
}

__global__ void __launch_bounds__(256)
import_kernel(const float* __restrict__ color3, const float* __restrict__ normal3,
              const float* __restrict__ depth1, uint32_t pixels, float4* __restrict__ sum_color,
              float4* __restrict__ sum_gbuf)
{
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= pixels) return;
  sum_color[p] = make_float4(color3[3 * (size_t)p], color3[3 * (size_t)p + 1],
                             color3[3 * (size_t)p + 2], 1.0f);
  sum_gbuf[p] = make_float4(normal3[3 * (size_t)p], normal3[3 * (size_t)p + 1],
                            normal3[3 * (size_t)p + 2], depth1[p]);
}

// ================================================================== denoise
// Pre-pass: per-pixel means and the world position the reference rebuilds per
// tap as generate_ray(camera, x+0.5, y+0.5)(depth) (denoiser.cu:44-45,71-72).
__global__ void __launch_bounds__(256)
denoise_prepare_kernel(const DevCamera cam, const float4* __restrict__ sum_color,
                       const float4* __restrict__ sum_gbuf, float4* __restrict__ color0,
                       float4* __restrict__ normal_depth, float4* __restrict__ position,
                       uint32_t row_lo)
{
  const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t y = row_lo + blockIdx.y;
  if (x >= cam.width) return;
  const uint32_t p = y * cam.width + x;
  const float4 c = sum_color[p];
  const float4 g = sum_gbuf[p];
  const float n = c.w;
  const float depth = g.w / n;
  color0[p] = make_float4(c.x / n, c.y / n, c.z / n, 0.f);
  normal_depth[p] = make_float4(g.x / n, g.y / n, g.z / n, depth);
  f3 o, d;
  camera_ray(cam, (float)x + 0.5f, (float)y + 0.5f, o, d);
  position[p] = mk4(o + d * depth, 0.f);
}

// One a-trous iteration (denoiser.cu:24-86), register-tiled: a thread produces ATR_R outputs
// that are `step` rows apart, (x, y0 + r*step), so the 5 x (ATR_R + 4) distinct taps are
// loaded once and shared (40 float4-triples for 4 outputs instead of 100); a warp covers 32
// consecutive x, so every tap row is one coalesced 512-byte request.
//
// Arithmetic follows the reference with two re-associations that stay far inside the 1e-4
// parity bound: the three edge-stopping weights min(exp(-a),1)*min(exp(-b),1)*min(exp(-c),1)
// (a, b, c >= 0, so the clamps never bind) are evaluated as one exp2(-(a+b+c) * log2 e), and
// colour*weight*kernel is grouped as colour*(weight*kernel).  The tap weight is
// kernel[min(|dx|,|dy|)] with kernel = {3/8, 1/4, 1/16} as in the reference (not the separable
// B3 product).  Taps are clamped to [0,W]x[0,H] inclusive like the reference: u == W aliases
// pixel (0, v+1) while its position is still rebuilt from the ray through (W+0.5, v+0.5);
// reads that would fall past the end of the buffer (undefined in the reference) use the last
// row / last pixel instead.  clamp_fix selects the sane W-1/H-1 clamp.
// Taps clamped to u == W or v == H (clamp_fix = 0): u == W aliases pixel (0, v+1) while its
// position is still rebuilt from the ray through (W+0.5, v+0.5); reads past the end of the buffer
// (undefined in the reference) use the last row / last pixel.
// sm_100 packed FP32 (FFMA2 / FADD2 / FMUL2): two independent lanes per instruction.  The filter is
// bound by instruction issue (ncu: 69 % issue-active, FMA pipe 54 %), so the arithmetic of two
// outputs that share a tap is issued as one packed stream.
typedef unsigned long long f2x;
PT_D f2x pk2(float a, float b)
{
  f2x r;
  asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
PT_D void upk2(f2x v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
PT_D f2x sub2(f2x a, f2x b)
{
  f2x r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
PT_D f2x add2(f2x a, f2x b)
{
  f2x r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
PT_D f2x mul2(f2x a, f2x b)
{
  f2x r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
PT_D f2x fma2(f2x a, f2x b, f2x c)
{
  f2x r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
PT_D float kernel_weight(int dx, int dy)
{
  const int adx = dx < 0 ? -dx : dx, ady = dy < 0 ? -dy : dy;
  const int m = adx < ady ? adx : ady;
  return m == 0 ? 3.f / 8.f : (m == 1 ? 1.f / 4.f : 1.f / 16.f);
}

struct Tap9 {
  f3 c, n, p;
};
__device__ __noinline__ Tap9 edge_tap(const DevCamera& cam, const float4* __restrict__ color_in,
                                      const float4* __restrict__ normal_depth, int u, int v, int W, int H)
{
  Tap9 t;
  int qi = u + v * W;
  if (qi >= W * H) qi = min(u, W - 1) + (H - 1) * W;
  const float4 c = ldg4(color_in + qi), n = ldg4(normal_depth + qi);
  t.c = mk3(c.x, c.y, c.z);
  t.n = mk3(n.x, n.y, n.z);
  f3 o, d;
  camera_ray(cam, (float)u + 0.5f, (float)v + 0.5f, o, d);
  t.p = o + d * n.w;
  return t;
}

// ATR_R (outputs per thread, `step` rows apart) is a template parameter: 2 for the edge strips,
// PT_ATR_R (default below) for the interior.
#ifndef ATR_PACK_DEFAULT
#define ATR_PACK_DEFAULT 0
#endif
#ifndef ATR_R_DEFAULT
#define ATR_R_DEFAULT 3 // measured (1080p, 5 iterations): 0.539 / 0.505 / 0.504 ms at 2 / 3 / 4
#endif

// One launch per iteration covers three regions of the output window, each a list of CTAs:
//   region 0  the interior: outputs none of whose taps reaches the clamp values u == W or v == H
//             (everything with clamp_fix).  EDGE = false: the edge path is compiled out of this
//             body — a third of the instructions and no call in the unrolled loop;
//   region 1/2 the right and bottom strips (2 * step wide / high), EDGE = true.
// A CTA runs exactly one of the bodies, so the interior never fetches the edge code, and the strips
// overlap with the interior instead of trailing it as two latency-bound launches (measured:
// 13 us each per iteration).
struct AtrousRegion {
  int x_lo, x_hi, row_lo, row_hi, n_groups, grid_x, n_ctas;
};
struct AtrousRegions {
  AtrousRegion r[3];
};

template <bool EDGE, int ATR_R, bool PACK>
PT_D void atrous_body(const DevCamera& cam, const DenoiseParams& dp, const float4* __restrict__ color_in,
                      const float4* __restrict__ normal_depth, const float4* __restrict__ position,
                      float4* __restrict__ color_out, int step, const AtrousRegion& rg, int cta)
{
  const int n_groups = rg.n_groups, row_lo = rg.row_lo, row_hi = rg.row_hi, x_lo = rg.x_lo, x_hi = rg.x_hi;
  const int bx = cta % rg.grid_x, by = cta / rg.grid_x;
  // rows [row_lo, row_hi) are written (the whole frame, or one GPU's band plus its halo); taps
  // are clamped at the FRAME edges either way, so a band computes what the full frame would
  const int W = (int)cam.width, H = (int)cam.height;
  const int x = x_lo + bx * 32 + (threadIdx.x & 31);
  const int g = by * 4 + (threadIdx.x >> 5); // (phase, group of ATR_R dilated rows)
  const int ph = g % step, k = g / step;
  if (x >= x_hi || k >= n_groups) return;
  const int q0 = row_lo / step + k * ATR_R;

  const float log2e = 1.4426950408889634f;
  const float kc = log2e / dp.c_phi;
  const float kn = log2e / (dp.n_phi * (float)(step * step));
  const float kp = log2e / dp.p_phi;
  const int umax = !EDGE || dp.clamp_fix ? W - 1 : W;
  const int vmax = !EDGE || dp.clamp_fix ? H - 1 : H;

  f3 cv[ATR_R], nv[ATR_R], pv[ATR_R], sum[ATR_R];
  float cum[ATR_R];
  bool valid[ATR_R];
#pragma unroll
  for (int r = 0; r < ATR_R; ++r) {
    const int y = (q0 + r) * step + ph;
    valid[r] = y < row_hi && y >= row_lo;
    const int p = min(y, H - 1) * W + x;
    const float4 c = ldg4(color_in + p), n = ldg4(normal_depth + p), q = ldg4(position + p);
    cv[r] = mk3(c.x, c.y, c.z);
    nv[r] = mk3(n.x, n.y, n.z);
    pv[r] = mk3(q.x, q.y, q.z);
    sum[r] = mk3(0.f, 0.f, 0.f);
    cum[r] = 0.f;
  }

  // fully unrolled (ATR_R + 4 tap rows): dy = j - r below must be a compile-time constant
#pragma unroll
  for (int j = -2; j < ATR_R + 2; ++j) {
    const int v = min(max((q0 + j) * step + ph, 0), vmax);
#pragma unroll
    for (int dx = -2; dx <= 2; ++dx) {
      const int u = min(max(x + dx * step, 0), umax);
      int qi = u + v * W;
      f3 ct, nt, pt3;
      if (!EDGE || (u < W && v < H)) {
        const float4 c = ldg4(color_in + qi), n = ldg4(normal_depth + qi), q = ldg4(position + qi);
        ct = mk3(c.x, c.y, c.z);
        nt = mk3(n.x, n.y, n.z);
        pt3 = mk3(q.x, q.y, q.z);
      } else {
        // the reference's right/bottom edge quirks: rare, kept out of line so that the unrolled
        // tap loop stays small (inlined 30 times it was most of an 11 000-instruction kernel)
        const Tap9 t = edge_tap(cam, color_in, normal_depth, u, v, W, H);
        ct = t.c, nt = t.n, pt3 = t.p;
      }
      if (PACK) {
        // outputs in aligned pairs (r, r+1): where this tap row serves both, one packed stream
        const f2x tcx = pk2(ct.x, ct.x), tcy = pk2(ct.y, ct.y), tcz = pk2(ct.z, ct.z);
        const f2x tnx = pk2(nt.x, nt.x), tny = pk2(nt.y, nt.y), tnz = pk2(nt.z, nt.z);
        const f2x tpx = pk2(pt3.x, pt3.x), tpy = pk2(pt3.y, pt3.y), tpz = pk2(pt3.z, pt3.z);
#pragma unroll
        for (int r = 0; r < ATR_R; r += 2) {
          const int dy0 = j - r, dy1 = j - (r + 1);
          const bool v0 = dy0 >= -2 && dy0 <= 2;                    // compile-time after unrolling
          const bool v1 = r + 1 < ATR_R && dy1 >= -2 && dy1 <= 2;
          if (v0 && v1) {
            const f2x dcx = sub2(pk2(cv[r].x, cv[r + 1].x), tcx), dcy = sub2(pk2(cv[r].y, cv[r + 1].y), tcy),
                      dcz = sub2(pk2(cv[r].z, cv[r + 1].z), tcz);
            const f2x dnx = sub2(pk2(nv[r].x, nv[r + 1].x), tnx), dny = sub2(pk2(nv[r].y, nv[r + 1].y), tny),
                      dnz = sub2(pk2(nv[r].z, nv[r + 1].z), tnz);
            const f2x dpx = sub2(pk2(pv[r].x, pv[r + 1].x), tpx), dpy = sub2(pk2(pv[r].y, pv[r + 1].y), tpy),
                      dpz = sub2(pk2(pv[r].z, pv[r + 1].z), tpz);
            const f2x cc = fma2(dcz, dcz, fma2(dcy, dcy, mul2(dcx, dcx)));
            const f2x nn = fma2(dnz, dnz, fma2(dny, dny, mul2(dnx, dnx)));
            const f2x pp2 = fma2(dpz, dpz, fma2(dpy, dpy, mul2(dpx, dpx)));
            const f2x e2 = fma2(pp2, pk2(kp, kp), fma2(nn, pk2(kn, kn), mul2(cc, pk2(kc, kc))));
            float e0, e1, x0, x1;
            upk2(e2, e0, e1);
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(x0) : "f"(-e0));
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(x1) : "f"(-e1));
            const f2x w2 = mul2(pk2(x0, x1), pk2(kernel_weight(dx, dy0), kernel_weight(dx, dy1)));
            f2x sx = pk2(sum[r].x, sum[r + 1].x), sy = pk2(sum[r].y, sum[r + 1].y), sz = pk2(sum[r].z, sum[r + 1].z);
            sx = fma2(tcx, w2, sx);
            sy = fma2(tcy, w2, sy);
            sz = fma2(tcz, w2, sz);
            const f2x cm = add2(pk2(cum[r], cum[r + 1]), w2);
            upk2(sx, sum[r].x, sum[r + 1].x);
            upk2(sy, sum[r].y, sum[r + 1].y);
            upk2(sz, sum[r].z, sum[r + 1].z);
            upk2(cm, cum[r], cum[r + 1]);
          } else if (v0 || v1) {
            const int rr = v0 ? r : r + 1, dy = v0 ? dy0 : dy1;
            const float kw = kernel_weight(dx, dy);
            const f3 dc = cv[rr] - ct, dn = nv[rr] - nt, dq = pv[rr] - pt3;
            const float e = dot3(dc, dc) * kc + dot3(dn, dn) * kn + dot3(dq, dq) * kp;
            float ex;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(-e));
            const float w = ex * kw;
            sum[rr] = sum[rr] + ct * w;
            cum[rr] += w;
          }
        }
      } else {
#pragma unroll
        for (int r = 0; r < ATR_R; ++r) {
          const int dy = j - r;
          if (dy < -2 || dy > 2) continue; // compile-time after unrolling
          const float kw = kernel_weight(dx, dy);
          const f3 dc = cv[r] - ct, dn = nv[r] - nt, dq = pv[r] - pt3;
          const float e = dot3(dc, dc) * kc + dot3(dn, dn) * kn + dot3(dq, dq) * kp;
          float ex;
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(-e)); // 2 ulp; no range fix-up code
          const float w = ex * kw;
          sum[r] = sum[r] + ct * w;
          cum[r] += w;
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < ATR_R; ++r) {
    if (!valid[r]) continue;
    const int y = (q0 + r) * step + ph;
    color_out[y * W + x] = make_float4(sum[r].x / cum[r], sum[r].y / cum[r], sum[r].z / cum[r], 0.f);
  }
}

template <int R_INTERIOR, bool PACK>
__global__ void __launch_bounds__(128)
atrous_kernel(const DevCamera cam, const DenoiseParams dp, const float4* __restrict__ color_in,
              const float4* __restrict__ normal_depth, const float4* __restrict__ position,
              float4* __restrict__ color_out, int step, const AtrousRegions regions)
{
  int cta = (int)blockIdx.x;
  if (cta < regions.r[0].n_ctas) {
    atrous_body<false, R_INTERIOR, PACK>(cam, dp, color_in, normal_depth, position, color_out, step, regions.r[0], cta);
    return;
  }
  cta -= regions.r[0].n_ctas;
  if (cta < regions.r[1].n_ctas) {
    atrous_body<true, 2, false>(cam, dp, color_in, normal_depth, position, color_out, step, regions.r[1], cta);
    return;
  }
  cta -= regions.r[1].n_ctas;
  atrous_body<true, 2, false>(cam, dp, color_in, normal_depth, position, color_out, step, regions.r[2], cta);
}

// ================================================================ launchers
static inline uint32_t cdiv(uint32_t a, uint32_t b) { return (a + b - 1) / b; }

void launch_stable_compact(const LaunchEnv& env, const PassBuffers& pb, const PassParams& pp,
                           int q, uint32_t bounce)
{
  const uint32_t n_blocks = cdiv(pb.capacity, SC_ITEMS);
  const uint32_t* n_ptr = bounce == 0 ? nullptr : pb.counters + bounce;
  const uint32_t n_first = pp.pixels; // slot == pixel index at bounce 0
  sc_count_kernel<<<n_blocks, SC_THREADS, 0, env.stream>>>(pb.flags, n_ptr, n_first,
                                                           pb.block_sums);
  sc_scan_kernel<<<1, 1024, 0, env.stream>>>(pb.block_sums, n_blocks, pb.counters + bounce + 1);
  sc_scatter_kernel<<<n_blocks, SC_THREADS, 0, env.stream>>>(
      pb.flags, n_ptr, n_first, pb.block_sums, bounce == 0 ? nullptr : pb.queue[q],
      pb.queue[q ^ 1]);
}

void launch_accumulate(const LaunchEnv& env, const PassBuffers& pb, const PassParams& pp,
                       float4* sum_color, float4* sum_gbuf, uint32_t max_depth)
{
  accumulate_kernel<<<cdiv(pp.pixel_end - pp.pixel_begin, 256), 256, 0, env.stream>>>(
      pb.ps, pp, sum_color, sum_gbuf, pb.counters, 0u, max_depth, pb.total_rays);
}

void launch_resolve_rgba8(const LaunchEnv& env, int kind, const float4* sum_color,
                          const float4* sum_gbuf, const float4* final_rgb, bool final_is_mean,
                          uint32_t pixels, uchar4* out)
{
  resolve_kernel<<<cdiv(pixels, 256), 256, 0, env.stream>>>(kind, sum_color, sum_gbuf, final_rgb,
                                                            final_is_mean, pixels, out);
}

void launch_export_f32(const LaunchEnv& env, int kind, const float4* sum_color,
                       const float4* sum_gbuf, const float4* final_rgb, bool final_is_mean,
                       uint32_t pixels, float* out)
{
  export_kernel<<<cdiv(pixels, 256), 256, 0, env.stream>>>(kind, sum_color, sum_gbuf, final_rgb,
                                                           final_is_mean, pixels, out);
}

void launch_import_frame(const LaunchEnv& env, const float* color3, const float* normal3,
                         const float* depth1, uint32_t pixels, float4* sum_color,
                         float4* sum_gbuf)
{
  import_kernel<<<cdiv(pixels, 256), 256, 0, env.stream>>>(color3, normal3, depth1, pixels,
                                                           sum_color, sum_gbuf);
}

void launch_denoise_prepare(const LaunchEnv& env, const DevCamera& cam, const float4* sum_color,
                            const float4* sum_gbuf, float4* color0, float4* normal_depth,
                            float4* position, uint32_t row_lo, uint32_t row_hi)
{
  dim3 grid(cdiv(cam.width, 256), row_hi - row_lo);
  denoise_prepare_kernel<<<grid, 256, 0, env.stream>>>(cam, sum_color, sum_gbuf, color0,
                                                       normal_depth, position, row_lo);
}

void launch_atrous(const LaunchEnv& env, const DevCamera& cam, const DenoiseParams& dp,
                   const float4* color_in, const float4* normal_depth, const float4* position,
                   float4* color_out, int step_width, uint32_t row_lo, uint32_t row_hi)
{
  // rows are grouped per phase (y mod step): ATR_R outputs of one thread are `step` apart;
  // dilated row index q = y / step runs over the window [row_lo, row_hi)
  const int W = (int)cam.width, H = (int)cam.height, step = step_width;
  static const int interior_r = [] {
    const char* v = getenv("PT_ATR_R");
    const int r = v ? atoi(v) : ATR_R_DEFAULT;
    return r == 3 || r == 4 ? r : 2;
  }();
  auto region = [&](bool edge, int y0, int y1, int x0, int x1) {
    AtrousRegion rg{};
    if (y0 >= y1 || x0 >= x1) return rg; // empty: n_ctas = 0
    const uint32_t R = edge ? 2u : (uint32_t)interior_r;
    const uint32_t q_lo = (uint32_t)y0 / (uint32_t)step;
    const uint32_t nq = cdiv((uint32_t)y1, (uint32_t)step) - q_lo;
    const uint32_t n_groups = cdiv(nq, R);
    rg.x_lo = x0, rg.x_hi = x1, rg.row_lo = y0, rg.row_hi = y1;
    rg.n_groups = (int)n_groups;
    rg.grid_x = (int)cdiv((uint32_t)(x1 - x0), 32);
    rg.n_ctas = rg.grid_x * (int)cdiv(n_groups * (uint32_t)step, 4);
    return rg;
  };
  // outputs whose taps stay inside [0, W-1] x [0, H-1]: x + 2 step <= W - 1, y + 2 step <= H - 1
  const int xin = dp.clamp_fix ? W : std::max(0, W - 2 * step);
  const int yin = dp.clamp_fix ? H : std::max(0, H - 2 * step);
  const int r0 = (int)row_lo, r1 = (int)row_hi;
  AtrousRegions rs;
  rs.r[0] = region(false, r0, std::min(r1, yin), 0, xin);  // interior
  rs.r[1] = region(true, r0, r1, xin, W);                  // right strip
  rs.r[2] = region(true, std::max(r0, yin), r1, 0, xin);   // bottom strip
  const int total = rs.r[0].n_ctas + rs.r[1].n_ctas + rs.r[2].n_ctas;
  if (total == 0) return;
  static const bool pack = [] {
    const char* v = getenv("PT_ATR_PACK");
    return v ? atoi(v) != 0 : ATR_PACK_DEFAULT != 0;
  }();
#define PT_ATROUS(RR, P)                                                                           \
  atrous_kernel<RR, P><<<total, 128, 0, env.stream>>>(cam, dp, color_in, normal_depth, position, color_out, step, rs)
  if (interior_r == 4) {
    if (pack) PT_ATROUS(4, true); else PT_ATROUS(4, false);
  } else if (interior_r == 3) {
    if (pack) PT_ATROUS(3, true); else PT_ATROUS(3, false);
  } else {
    if (pack) PT_ATROUS(2, true); else PT_ATROUS(2, false);
  }
#undef PT_ATROUS
}

} // namespace pt
