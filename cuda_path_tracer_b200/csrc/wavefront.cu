// wavefront.cu — hand-written sm_100a kernels of the path-tracing wavefront.
//
// Default scheduler (PT_RNG_PIXEL_STREAM, every path carries its own RNG stream):
//
//   chain_kernel<true>   raygen (ray_gen.cu:11-61) + classification of the primary rays + every
//                        bounce that needs no BVH, shaded in registers: rebuilds the hit (sphere /
//                        triangle / miss), evaluate_material + sky + G-buffer (path_tracer.cu:29-34,
//                        138-201, 292-315), scatters, classifies the new ray (sphere tests,
//                        path_tracer.cu:87-98, and a stack-free walk of the top of the BVH).  A path
//                        whose next ray must enter the BVH is PARKED: its 68-byte state is appended
//                        to the next list with one warp-aggregated atomic (replaces
//                        thrust::stable_partition, path_tracer.cu:433-457).
//   traverse_kernel      BVH traversal ONLY, over the parked rays.  Persistent warps, lane-level
//                        refill with one atomic per warp, culled + ordered walk of 64-byte two-box
//                        nodes over world-space baked triangles, while-while scheduling.  Replaces
//                        the mesh branch of ray_scene_intersection_test (path_tracer.cu:36-76).
//   chain_kernel<false>  the same chain loop for the paths whose ray was just traversed.
//
// Streaming-compat scheduler (PT_RNG_SLOT_RESEED, bounce-synchronous like the reference):
// raygen_kernel, traverse_kernel over an id queue, shade_kernel, stable scan compaction
// (image_kernels.cu).
//
// Why traversal is split from everything else: ncu showed the fused extend+shade kernel
// issue-bound with 9-10 of 32 lanes active (profiles/r1_bounce_v2_*): in open scenes most rays
// never enter the BVH, and their sphere tests + shading ran in the partially filled retire/refill
// phases of a warp whose other lanes were long-running mesh rays.  Now everything that is not
// traversal runs at (near) full warp occupancy, and only rays that touch the mesh pay for the
// persistent machinery.  Opt-in: traverse8_kernel over the compressed 8-wide tree (PT_BVH=8).
#include "kernels.h"

#include <float.h>
#include <stdio.h>
#include <stdlib.h>

#include <mutex>

namespace pt {

// =================================================================== helpers
PT_D f3 xyz(const float4& v) { return mk3(v.x, v.y, v.z); }
PT_D float4 mk4(f3 v, float w) { return make_float4(v.x, v.y, v.z, w); }
PT_D float4 ldg4(const float4* p) { return __ldg(p); }
// Path-state streams are written once and read once, gigabytes per pass: with `cs` set they use
// the evict-first cache policy (ld/st.global.cs) so that they do not displace the scene in L1/L2.
PT_D float4 ld_state(const float4* p, uint32_t cs) { return cs ? __ldcs(p) : *p; }
PT_D uint4 ld_state(const uint4* p, uint32_t cs) { return cs ? __ldcs(p) : *p; }
PT_D uint32_t ld_state(const uint32_t* p, uint32_t cs) { return cs ? __ldcs(p) : *p; }
PT_D void st_state(float4* p, float4 v, uint32_t cs)
{
  if (cs)
    __stcs(p, v);
  else
    *p = v;
}
PT_D void st_state(uint4* p, uint4 v, uint32_t cs)
{
  if (cs)
    __stcs(p, v);
  else
    *p = v;
}
PT_D void st_state(uint32_t* p, uint32_t v, uint32_t cs)
{
  if (cs)
    __stcs(p, v);
  else
    *p = v;
}

struct Hit {
  float t;
  f3 p;
  f3 n;
  uint32_t material;
  uint32_t side;
  int32_t object;
  int32_t prim;
};

// aux code word (see PathState::aux)
#define AUX_TRI 0x80000000u     // low bits = triangle slot; the traversal found it
#define AUX_PENDING 0x40000000u // spheres that follow the first mesh object are still untested
#define AUX_VALUE 0x3fffffffu   // sphere index + 1, or 0 = nothing hit so far

// ------------------------------------------------------------- sphere test
// ray_object_intersection_test, sphere branch (path_tracer.cu:87-98) with
// inverse_transform_ray (transform.hpp:50-58) and ray_sphere_intersection_test
// (intersections.cuh:7-41).  The quirks are kept: the object-space direction is
// re-normalised while t_min/t_max stay world-space; the reported t is the
// world-space distance; the normal is transformed by the inverse transpose and
// not re-normalised.
// PRE = false: the caller knows the sphere is hit (resolve_hit rebuilding an accepted hit).
template <bool PRE = true>
PT_D bool sphere_test(const DevSphere* __restrict__ sp, f3 o, f3 d, float tmin, float tmax,
                      Hit& h)
{
  // Conservative pre-reject in world space (similarity transforms only).  The exact test below
  // costs two IEEE divisions and a square root before its first early-out; most rays either pass
  // a sphere by far or leave it behind.  A ray is dropped here only when the exact test must
  // reject it too: (a) its LINE misses the bounding sphere by more than 64e-6 of the magnitudes
  // the discriminant is formed from (the exact test's own rounding is ~1e-6 of them; the radius is
  // widened by 2e-6 of the coordinate magnitudes for the error of the world->object transform),
  // or (b) the origin lies outside by the same margin and the ray points away, so both roots are
  // negative and fail t >= t_min > 0.  NaN/inf operands fail both comparisons and fall through.
  if (PRE && sp->pre_ok) {
    const f3 ocw = o - mk3(sp->wx, sp->wy, sp->wz);
    const float dd2 = dot3(d, d), bq = dot3(ocw, d), oc2 = dot3(ocw, ocw);
    const float re = fmaf(fabsf(o.x) + fabsf(o.y) + fabsf(o.z) + sp->wl1, 2e-6f, sp->wr);
    const float cq = oc2 - re * re;
    const float b2 = bq * bq, m2 = dd2 * oc2;
    if (b2 - dd2 * cq < -64e-6f * (b2 + m2)) return false;
    if (cq > 64e-6f * oc2 && bq > 0.0f && b2 > 1e-8f * m2) return false;
  }
  const float* I = sp->inv;
  // transform_point(inverse): affine, w == 1 exactly
  f3 oo;
  oo.x = (I[0] * o.x + I[1] * o.y) + (I[2] * o.z + I[3]);
  oo.y = (I[4] * o.x + I[5] * o.y) + (I[6] * o.z + I[7]);
  oo.z = (I[8] * o.x + I[9] * o.y) + (I[10] * o.z + I[11]);
  f3 dd;
  dd.x = (I[0] * d.x + I[1] * d.y) + (I[2] * d.z);
  dd.y = (I[4] * d.x + I[5] * d.y) + (I[6] * d.z);
  dd.z = (I[8] * d.x + I[9] * d.y) + (I[10] * d.z);
  dd = normalize3(dd);

  const f3 center = mk3(sp->cx, sp->cy, sp->cz);
  const float radius = sp->radius;
  const f3 oc = oo - center;
  const float a = dot3(dd, dd);
  const float b = 2.0f * dot3(dd, oc);
  const float c = dot3(oc, oc) - radius * radius;
  const float disc = b * b - 4.0f * a * c;
  if (disc < 0.0f) return false;
  const float sq = sqrtf(disc);
  const float t1 = (-b - sq) / (2.0f * a);
  const float t2 = (-b + sq) / (2.0f * a);
  float t;
  if (t1 >= tmin && t1 <= tmax) {
    t = t1;
  } else if (t2 >= tmin && t2 <= tmax) {
    t = t2;
  } else {
    return false;
  }
  const f3 po = oo + dd * t;
  const f3 outward = (po - center) / radius;
  const bool front = dot3(dd, outward) < 0.0f;
  const f3 no = front ? outward : -outward;

  const float* M = sp->m;
  f3 pw;
  pw.x = (M[0] * po.x + M[1] * po.y) + (M[2] * po.z + M[3]);
  pw.y = (M[4] * po.x + M[5] * po.y) + (M[6] * po.z + M[7]);
  pw.z = (M[8] * po.x + M[9] * po.y) + (M[10] * po.z + M[11]);
  // transpose(inverse) * normal  ==  columns of the row-major inverse
  f3 nw;
  nw.x = (I[0] * no.x + I[4] * no.y) + (I[8] * no.z);
  nw.y = (I[1] * no.x + I[5] * no.y) + (I[9] * no.z);
  nw.z = (I[2] * no.x + I[6] * no.y) + (I[10] * no.z);

  h.t = length3(pw - o); // glm::distance(ray.origin, point)
  h.p = pw;
  h.n = nw;
  h.side = front ? 0u : 1u;
  h.material = sp->material;
  h.object = sp->object;
  h.prim = -1;
  return true;
}

PT_D float safe_inv(float x)
{
  const float ooeps = 8.271806125530277e-25f; // 2^-80
  return 1.0f / (fabsf(x) > ooeps ? x : copysignf(ooeps, x));
}

// ----------------------------------------------------------- classification
// The part of ray_scene_intersection_test (path_tracer.cu:110-128) that does not need the
// BVH stack: the spheres preceding the first mesh object, then a stack-free walk of the top of
// the BVH.  A ray whose walk dies out is "simple": the remaining spheres are tested right away
// and its intersection is final.  Otherwise it is queued for traverse_kernel (AUX_PENDING set)
// together with the node the walk stopped at.
// Slab tests of both children of an inner node against the segment [tmin, tbest]
// (the same arithmetic as trav_inner).
PT_D void node_test(const float4* __restrict__ nodes, int node, f3 id, f3 od, float tmin, float tbest, bool& t0,
                    bool& t1, int& c0, int& c1, float& c0min, float& c1min)
{
  const float4* np = nodes + (size_t)node * 4;
  const float4 n0 = ldg4(np + 0);
  const float4 n1 = ldg4(np + 1);
  const float4 n2 = ldg4(np + 2);
  const float4 n3 = ldg4(np + 3);
  const float c0lox = n0.x * id.x - od.x, c0hix = n0.y * id.x - od.x;
  const float c0loy = n0.z * id.y - od.y, c0hiy = n0.w * id.y - od.y;
  const float c0loz = n2.x * id.z - od.z, c0hiz = n2.y * id.z - od.z;
  const float c1lox = n1.x * id.x - od.x, c1hix = n1.y * id.x - od.x;
  const float c1loy = n1.z * id.y - od.y, c1hiy = n1.w * id.y - od.y;
  const float c1loz = n2.z * id.z - od.z, c1hiz = n2.w * id.z - od.z;
  c0min = fmaxf(fmaxf(fminf(c0lox, c0hix), fminf(c0loy, c0hiy)), fmaxf(fminf(c0loz, c0hiz), tmin));
  const float c0max =
      fminf(fminf(fmaxf(c0lox, c0hix), fmaxf(c0loy, c0hiy)), fminf(fmaxf(c0loz, c0hiz), tbest));
  c1min = fmaxf(fmaxf(fminf(c1lox, c1hix), fminf(c1loy, c1hiy)), fmaxf(fminf(c1loz, c1hiz), tmin));
  const float c1max =
      fminf(fminf(fmaxf(c1lox, c1hix), fmaxf(c1loy, c1hiy)), fminf(fmaxf(c1loz, c1hiz), tbest));
  // robust slab comparison (Ize 2013): widen the far side by 2 ulp
  t0 = c0max * 1.0000004f >= c0min;
  t1 = c1max * 1.0000004f >= c1min;
  c0 = __float_as_int(n3.x);
  c1 = __float_as_int(n3.y);
}

// 64-byte node fetch: four 128-bit loads, or (L256) two 256-bit loads (LDG.E.256: one L1
// data-pipe pass per 32-byte sector instead of two).
template <bool L256>
PT_D void load_node(const float4* __restrict__ np, float4& n0, float4& n1, float4& n2, float4& n3)
{
  if (L256) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(n0.x), "=f"(n0.y), "=f"(n0.z), "=f"(n0.w), "=f"(n1.x), "=f"(n1.y), "=f"(n1.z), "=f"(n1.w)
                 : "l"(np));
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(n2.x), "=f"(n2.y), "=f"(n2.z), "=f"(n2.w), "=f"(n3.x), "=f"(n3.y), "=f"(n3.z), "=f"(n3.w)
                 : "l"(np + 2));
  } else {
    n0 = ldg4(np + 0);
    n1 = ldg4(np + 1);
    n2 = ldg4(np + 2);
    n3 = ldg4(np + 3);
  }
}

// Closest sphere of one group (spheres[lo, hi)): linear scan like the reference's object loop,
// or — root >= 0 — an ordered walk of the group's tree.  A subtree is skipped when the ray's line
// misses its box or enters it farther than the closest accepted hit (entry parameter x |d| =
// distance, the unit sphere_test reports; 1e-5 slack).  Equivalent to the scan for rigidly placed
// spheres except for the winner of an exact tie.
// ST is a template parameter of everything that tests spheres: the kernels of scenes without a
// sphere tree (every bundled scene) are compiled without the walk (measured: with the walk
// compiled in, the chain kernels of the headline scene went from 71-80 to 78-80 registers plus
// 96 B of local stack and lost 2 %).
template <bool ST>
PT_D void spheres_closest(const DevScene& sc, int root, uint32_t lo, uint32_t hi, f3 o, f3 d, float tmin,
                          float& tbest, uint32_t& code, Hit& hs)
{
  if (!ST || root < 0) {
    for (uint32_t i = lo; i < hi; ++i) {
      if (sphere_test(sc.spheres + i, o, d, tmin, tbest, hs)) {
        code = i + 1u;
        tbest = hs.t;
      }
    }
    return;
  }
  const f3 id = mk3(safe_inv(d.x), safe_inv(d.y), safe_inv(d.z));
  const f3 od = mk3(o.x * id.x, o.y * id.y, o.z * id.z);
  const float len = length3(d);
  int stack[24]; // median-split tree over <= 2^24 spheres, 4 per leaf: at most 22 levels
  int sp = 0, node = root;
  for (;;) {
    if (node >= 0) {
      bool t0, t1;
      int c0, c1;
      float m0, m1;
      node_test(sc.sph_nodes, node, id, od, 0.0f, FLT_MAX, t0, t1, c0, c1, m0, m1);
      const float far = tbest * 1.00001f;
      t0 = t0 && m0 * len <= far;
      t1 = t1 && m1 * len <= far;
      if (t0 && t1) {
        const bool swap = m1 < m0;
        PT_CHECK(sp < 24, "sphere tree stack overflow");
        stack[sp++] = swap ? c0 : c1;
        node = swap ? c1 : c0;
        continue;
      }
      if (t0 || t1) {
        node = t0 ? c0 : c1;
        continue;
      }
    } else {
      const uint32_t leaf = (uint32_t)(~node);
      const uint32_t first = leaf >> 3, count = (leaf & 7u) + 1u;
      for (uint32_t k = 0; k < count; ++k) {
        if (sphere_test(sc.spheres + first + k, o, d, tmin, tbest, hs)) {
          code = first + k + 1u;
          tbest = hs.t;
        }
      }
    }
    if (sp == 0) return;
    node = stack[--sp];
  }
}

// Length of the stack-free walk below.  Re-measured with the final traversal kernel (2 / 4 / 6 / 8
// nodes): bunny 11 213 / 11 204 / 11 320 / 11 252 Mrays/s, 150 k triangles 5 295 / 5 212 / 5 204 /
// 5 233, terrain 2 727 / 2 686 / 2 672 / 2 638 (profiles/r2_ab_prefix_walk_chain_bounds.log): a
// shorter walk moves work from the chain kernels into traversal, which wins or loses by scene.
#ifndef PREFIX_MAX
#define PREFIX_MAX 6
#endif
// `hs` receives the full Intersection of the sphere named by `code` (valid when the ray is simple
// and code != 0): a caller that shades the ray at once (chain_kernel) need not rebuild it.
template <bool ST = false>
PT_D bool classify(const DevScene& sc, f3 o, f3 d, float tmin, float tmax, float& tbest,
                   uint32_t& code, int& start, Hit& hs)
{
  tbest = tmax;
  code = 0u;
  start = 0;
  spheres_closest<ST>(sc, sc.sph_root_before, 0u, sc.n_spheres_before, o, d, tmin, tbest, code, hs);
  bool complex_ray = false;
  if (sc.n_tris != 0u) {
    // Walk the hot top of the tree while at most one child box is hit: such a prefix needs no
    // stack, runs here at full warp occupancy, and lets a ray that only grazes the outer boxes
    // never reach the traversal queue.  Traversal starts where the walk stopped.
    const f3 id = mk3(safe_inv(d.x), safe_inv(d.y), safe_inv(d.z));
    const f3 od = mk3(o.x * id.x, o.y * id.y, o.z * id.z);
    int node = 0;
    complex_ray = true;
#pragma unroll 1
    for (int it = 0; it < PREFIX_MAX; ++it) {
      bool t0, t1;
      int c0, c1;
      float m0, m1;
      node_test(sc.nodes, node, id, od, tmin, tbest, t0, t1, c0, c1, m0, m1);
      if (!t0 && !t1) {
        complex_ray = false;
        break;
      }
      if (t0 && t1) break;
      node = t0 ? c0 : c1;
      if (node < 0) break; // a leaf: traversal starts (and ends) there
    }
    start = node;
  }
  if (complex_ray) {
    code |= AUX_PENDING;
  } else {
    spheres_closest<ST>(sc, sc.sph_root_after, sc.n_spheres_before, sc.n_spheres, o, d, tmin, tbest, code, hs);
  }
  return complex_ray;
}
template <bool ST = false>
PT_D bool classify(const DevScene& sc, f3 o, f3 d, float tmin, float tmax, float& tbest,
                   uint32_t& code, int& start)
{
  Hit hs;
  return classify<ST>(sc, o, d, tmin, tmax, tbest, code, start, hs);
}

// Rebuilds the Intersection (intersection.hpp:8-14) from the 8-byte aux word.
template <bool ST = false>
PT_D bool resolve_hit(const DevScene& sc, f3 o, f3 d, float tmin, float t_aux, uint32_t code, Hit& h)
{
  bool hit = false;
  float tbest = t_aux;
  if (code & AUX_TRI) {
    PT_CHECK((code & AUX_VALUE) < sc.n_tris, "hit record names a triangle outside the array");
    const float4* tp = sc.tris + (size_t)(code & AUX_VALUE) * 3;
    const float4 t0 = ldg4(tp + 0);
    const float4 t1 = ldg4(tp + 1);
    const float4 t2 = ldg4(tp + 2);
    const f3 outward = normalize3(cross3(xyz(t1), xyz(t2))); // triangle_normal (intersections.cuh:43-47)
    const bool front = dot3(d, outward) < 0.0f;
    h.t = t_aux;
    h.p = o + d * t_aux; // ray(t)
    h.n = front ? outward : -outward;
    h.side = front ? 0u : 1u;
    h.prim = __float_as_int(t0.w);
    h.object = __float_as_int(t1.w);
    h.material = (uint32_t)__float_as_int(t2.w);
    hit = true;
  } else if (code & AUX_VALUE) {
    PT_CHECK((code & AUX_VALUE) - 1u < sc.n_spheres, "hit record names a sphere outside the table");
    // same root as when it was accepted: any t_max >= the accepted t selects it again
    hit = sphere_test<false>(sc.spheres + ((code & AUX_VALUE) - 1u), o, d, tmin, FLT_MAX, h);
  }
  if (code & (AUX_TRI | AUX_PENDING)) {
    uint32_t after = 0u;
    spheres_closest<ST>(sc, sc.sph_root_after, sc.n_spheres_before, sc.n_spheres, o, d, tmin, tbest, after, h);
    if (after != 0u) {
      hit = true;
      // h holds the LAST accepted sphere = the closest one; when the walk found none closer than
      // the triangle, h was not touched
    }
  }
  return hit;
}

// ---------------------------------------------------------------- traversal
// Resumable per-lane state machine (init / step) so that a persistent warp can retire
// finished lanes and refill them while the others keep walking.  Result-equivalent to the
// reference's un-culled traversal (a culled subtree cannot hold a nearer accepted hit);
// exact-tie winners may differ (reference: last tested wins).
struct Trav {
  f3 o, d;
  float tmin, tbest;
  float idx, idy, idz, odx, ody, odz;
  int node; // current node, PT_SENTINEL when done
  int sp;
  int best; // best triangle slot or -1
  // quantised nodes: PRMT selectors that drop the NEAR plane of an axis pair (lo | hi << 16) into
  // the float 2^23 + q — the low half for a ray that travels up the axis, the high half otherwise;
  // the far plane's selector is sel ^ 0x22.  The slab test then needs no per-axis min / max.
  uint32_t selx, sely, selz;
};

// (A hybrid stack — the first 8 or 16 entries in shared memory laid out [entry][thread], one
// conflict-free wavefront per warp access, the rest in local memory — was measured and rejected:
// traverse 15.27 -> 15.82 / 16.10 ms on the bunny frame, profiles/README.md.)
// QN: the inner nodes are read from DevScene::qnodes.  A plane arrives as the float 2^23 + q (its
// 16 bits dropped into the mantissa of 0x4B000000 by one PRMT), so with
//   idx = cell / d   and   odx = 2^23 * idx - (org - o) / d
// the slab distance (org + q * cell - o) / d is the SAME fused multiply-add f * idx - odx as for a
// 64-byte node: de-quantisation costs no arithmetic.  Rounding odx loses at most half a cell (the
// grid keeps a whole one in reserve); what is proportional to the distance is covered by the
// widened comparison below, as for the exact planes.
template <bool QN = false>
PT_D void trav_init(const DevScene& sc, Trav& T, f3 o, f3 d, float tmin, float tbest, int start, int* stack)
{
  T.o = o;
  T.d = d;
  T.tmin = tmin;
  T.tbest = tbest;
  T.best = -1;
  const float ix = safe_inv(d.x), iy = safe_inv(d.y), iz = safe_inv(d.z);
  if (QN) {
    T.idx = sc.q_cell[0] * ix;
    T.idy = sc.q_cell[1] * iy;
    T.idz = sc.q_cell[2] * iz;
    T.odx = fmaf(8388608.0f, T.idx, -((sc.q_org[0] - o.x) * ix));
    T.ody = fmaf(8388608.0f, T.idy, -((sc.q_org[1] - o.y) * iy));
    T.odz = fmaf(8388608.0f, T.idz, -((sc.q_org[2] - o.z) * iz));
    T.selx = T.idx >= 0.0f ? 0x7610u : 0x7632u;
    T.sely = T.idy >= 0.0f ? 0x7610u : 0x7632u;
    T.selz = T.idz >= 0.0f ? 0x7610u : 0x7632u;
  } else {
    T.idx = ix;
    T.idy = iy;
    T.idz = iz;
    T.odx = o.x * T.idx;
    T.ody = o.y * T.idy;
    T.odz = o.z * T.idz;
  }
  stack[0] = PT_SENTINEL;
  T.sp = 1;
  T.node = start;
}

// Inner-node visit: two slab tests against the children's boxes stored in the node, ordered
// descent (nearer child first), farther child pushed.
#ifndef PT_QN_SIGNSEL
#define PT_QN_SIGNSEL 1
#endif
PT_D float q_pick(uint32_t w, uint32_t sel) { return __uint_as_float(__byte_perm(w, 0x4B000000u, sel)); }

template <bool L256, bool QN = false>
PT_D void trav_inner(const DevScene& sc, Trav& T, int* stack)
{
  PT_CHECK((uint32_t)T.node < sc.n_nodes, "inner node index outside the tree");
  float c0min, c0max, c1min, c1max;
  int c0, c1;
  if (QN) {
    uint4 a, b;
    const uint4* qp = sc.qnodes + (size_t)T.node * 2;
    if (L256) {
      asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
                   : "l"(qp));
    } else {
      a = __ldg(qp);
      b = __ldg(qp + 1);
    }
#if PT_QN_SIGNSEL
    // f * idx - odx is monotonic in f with the sign of idx, so the plane the ray meets first is
    // known per axis from the ray alone: 4 min/max per child instead of 10, same values
    const uint32_t fx = T.selx ^ 0x22u, fy = T.sely ^ 0x22u, fz = T.selz ^ 0x22u;
    const float n0x = q_pick(a.x, T.selx) * T.idx - T.odx, f0x = q_pick(a.x, fx) * T.idx - T.odx;
    const float n0y = q_pick(a.y, T.sely) * T.idy - T.ody, f0y = q_pick(a.y, fy) * T.idy - T.ody;
    const float n0z = q_pick(a.z, T.selz) * T.idz - T.odz, f0z = q_pick(a.z, fz) * T.idz - T.odz;
    const float n1x = q_pick(a.w, T.selx) * T.idx - T.odx, f1x = q_pick(a.w, fx) * T.idx - T.odx;
    const float n1y = q_pick(b.x, T.sely) * T.idy - T.ody, f1y = q_pick(b.x, fy) * T.idy - T.ody;
    const float n1z = q_pick(b.y, T.selz) * T.idz - T.odz, f1z = q_pick(b.y, fz) * T.idz - T.odz;
    c0min = fmaxf(fmaxf(n0x, n0y), fmaxf(n0z, T.tmin));
    c0max = fminf(fminf(f0x, f0y), fminf(f0z, T.tbest));
    c1min = fmaxf(fmaxf(n1x, n1y), fmaxf(n1z, T.tmin));
    c1max = fminf(fminf(f1x, f1y), fminf(f1z, T.tbest));
#else
    const float c0lox = q_pick(a.x, 0x7610u) * T.idx - T.odx, c0hix = q_pick(a.x, 0x7632u) * T.idx - T.odx;
    const float c0loy = q_pick(a.y, 0x7610u) * T.idy - T.ody, c0hiy = q_pick(a.y, 0x7632u) * T.idy - T.ody;
    const float c0loz = q_pick(a.z, 0x7610u) * T.idz - T.odz, c0hiz = q_pick(a.z, 0x7632u) * T.idz - T.odz;
    const float c1lox = q_pick(a.w, 0x7610u) * T.idx - T.odx, c1hix = q_pick(a.w, 0x7632u) * T.idx - T.odx;
    const float c1loy = q_pick(b.x, 0x7610u) * T.idy - T.ody, c1hiy = q_pick(b.x, 0x7632u) * T.idy - T.ody;
    const float c1loz = q_pick(b.y, 0x7610u) * T.idz - T.odz, c1hiz = q_pick(b.y, 0x7632u) * T.idz - T.odz;
    c0min = fmaxf(fmaxf(fminf(c0lox, c0hix), fminf(c0loy, c0hiy)), fmaxf(fminf(c0loz, c0hiz), T.tmin));
    c0max = fminf(fminf(fmaxf(c0lox, c0hix), fmaxf(c0loy, c0hiy)), fminf(fmaxf(c0loz, c0hiz), T.tbest));
    c1min = fmaxf(fmaxf(fminf(c1lox, c1hix), fminf(c1loy, c1hiy)), fmaxf(fminf(c1loz, c1hiz), T.tmin));
    c1max = fminf(fminf(fmaxf(c1lox, c1hix), fmaxf(c1loy, c1hiy)), fminf(fmaxf(c1loz, c1hiz), T.tbest));
#endif
    c0 = (int)b.z;
    c1 = (int)b.w;
  } else {
    float4 n0, n1, n2, n3;
    load_node<L256>(sc.nodes + (size_t)T.node * 4, n0, n1, n2, n3);
    const float c0lox = n0.x * T.idx - T.odx, c0hix = n0.y * T.idx - T.odx;
    const float c0loy = n0.z * T.idy - T.ody, c0hiy = n0.w * T.idy - T.ody;
    const float c0loz = n2.x * T.idz - T.odz, c0hiz = n2.y * T.idz - T.odz;
    const float c1lox = n1.x * T.idx - T.odx, c1hix = n1.y * T.idx - T.odx;
    const float c1loy = n1.z * T.idy - T.ody, c1hiy = n1.w * T.idy - T.ody;
    const float c1loz = n2.z * T.idz - T.odz, c1hiz = n2.w * T.idz - T.odz;
    c0min = fmaxf(fmaxf(fminf(c0lox, c0hix), fminf(c0loy, c0hiy)), fmaxf(fminf(c0loz, c0hiz), T.tmin));
    c0max = fminf(fminf(fmaxf(c0lox, c0hix), fmaxf(c0loy, c0hiy)), fminf(fmaxf(c0loz, c0hiz), T.tbest));
    c1min = fmaxf(fmaxf(fminf(c1lox, c1hix), fminf(c1loy, c1hiy)), fmaxf(fminf(c1loz, c1hiz), T.tmin));
    c1max = fminf(fminf(fmaxf(c1lox, c1hix), fmaxf(c1loy, c1hiy)), fminf(fmaxf(c1loz, c1hiz), T.tbest));
    c0 = __float_as_int(n3.x);
    c1 = __float_as_int(n3.y);
  }
  // robust slab comparison (Ize 2013): widen the far side by 2 ulp
  const bool trav0 = c0max * 1.0000004f >= c0min;
  const bool trav1 = c1max * 1.0000004f >= c1min;
  if (!trav0 && !trav1) {
    PT_CHECK(T.sp > 0, "traversal stack underflow");
    T.node = stack[--T.sp];
  } else {
    const bool swap = trav1 && (!trav0 || c1min < c0min);
    T.node = swap ? c1 : c0;
    if (trav0 && trav1) {
      PT_CHECK(T.sp < PT_STACK, "traversal stack overflow");
      stack[T.sp++] = swap ? c0 : c1;
    }
  }
}

// Leaf visit: <= 4 Moller-Trumbore tests with the reference's operation order for a, f, u, v, t
// (intersections.cuh:49-85), evaluated branch-free so the lanes of a warp stay converged; the
// accept conditions are the reference's (EPSILON 1e-7 on the determinant, u,v in [0,1],
// u+v <= 1, closed t range).
PT_D void trav_leaf(const DevScene& sc, Trav& T, int* stack)
{
  const uint32_t code = (uint32_t)(~T.node);
  const uint32_t first = code >> 3;
  const uint32_t count = (code & 7u) + 1u;
  PT_CHECK(first + count <= sc.n_tris, "leaf outside the triangle array");
  for (uint32_t k = 0; k < count; ++k) {
    const float4* tp = sc.tris + (size_t)(first + k) * 3;
    const float4 t0 = ldg4(tp + 0);
    const float4 t1 = ldg4(tp + 1);
    const float4 t2 = ldg4(tp + 2);
    const f3 e1 = xyz(t1), e2 = xyz(t2);
    const f3 hh = cross3(T.d, e2);
    const float a = dot3(e1, hh);
    const float f = 1.0f / a;
    const f3 s = T.o - xyz(t0);
    const float u = f * dot3(s, hh);
    const f3 q = cross3(s, e1);
    const float v = f * dot3(T.d, q);
    const float t = f * dot3(e2, q);
    const bool ok = !(a > -0.0000001f && a < 0.0000001f) && !(u < 0.0f || u > 1.0f) &&
                    !(v < 0.0f || u + v > 1.0f) && (t >= T.tmin && t <= T.tbest);
    if (ok) {
      T.tbest = t;
      T.best = (int)(first + k);
    }
  }
  PT_CHECK(T.sp > 0, "traversal stack underflow");
  T.node = stack[--T.sp];
}

// ------------------------------------------------------ bounce-0 index map
// Work item -> (sample, pixel): one warp covers an 8x4 pixel tile so primary
// rays of a warp are coherent.  Returns false for padding lanes.
PT_D bool first_item(const PassParams& pp, uint32_t idx, uint32_t& pid, uint32_t& pixel,
                     uint32_t& s, uint32_t& x, uint32_t& y)
{
  uint32_t tx, ty;
  const uint32_t lane = idx & 31u;
  if (pp.order == 0u) {
    // sample-major: all tiles of sample 0, then sample 1, ...
    const uint32_t per_sample = pp.tiles_x * pp.tiles_y * 32u;
    s = fd_div(idx, pp.fd_per_sample);
    const uint32_t tile = (idx - s * per_sample) >> 5;
    ty = fd_div(tile, pp.fd_tiles_x), tx = tile - ty * pp.tiles_x;
  } else {
    // tile-major: the pass's samples of one tile are consecutive warps, so the paths in flight at
    // any moment (and the parked lists they append to) cover a compact patch of the image
    // instead of a full-width band — a smaller scene working set for L1/L2.  Order 2 also walks
    // the tiles in 8x8-tile blocks.  The path id, hence every result, is unchanged.
    const uint32_t group = idx >> 5;
    uint32_t tile = fd_div(group, pp.fd_samples);
    s = group - tile * pp.samples;
    if (pp.order == 1u) {
      ty = fd_div(tile, pp.fd_tiles_x), tx = tile - ty * pp.tiles_x;
    } else {
      const uint32_t sb = tile >> 6, in = tile & 63u;
      const uint32_t sby = fd_div(sb, pp.fd_sbx), sbxi = sb - sby * pp.sbx;
      tx = sbxi * 8u + (in & 7u);
      ty = sby * 8u + (in >> 3);
      if (tx >= pp.tiles_x || ty >= pp.tiles_y) return false;
    }
  }
  x = tx * 8u + (lane & 7u);
  y = (ty + pp.tile_y0) * 4u + (lane >> 3);
  pixel = y * pp.cam.width + x;
  pid = s * pp.band_pixels + (pixel - pp.pixel_begin); // relative to the band: a lane's buffers hold its band only
  return x < pp.cam.width && y < pp.cam.height && s < pp.samples;
}
PT_D bool first_item(const PassParams& pp, uint32_t idx, uint32_t& pid, uint32_t& pixel, uint32_t& s)
{
  uint32_t x, y;
  return first_item(pp, idx, pid, pixel, s, x, y);
}

// Appends `pid` of every lane with `push` set: one atomic per warp (ballot + popc prefix).
PT_D void warp_append(bool push, uint32_t pid, uint32_t* __restrict__ queue,
                      uint32_t* __restrict__ count, uint32_t lane)
{
  const uint32_t mask = __ballot_sync(0xffffffffu, push);
  if (mask == 0u) return;
  const int leader = __ffs(mask) - 1;
  uint32_t base = 0;
  if ((int)lane == leader) base = atomicAdd(count, (uint32_t)__popc(mask));
  base = __shfl_sync(0xffffffffu, base, leader);
  if (push) queue[base + (uint32_t)__popc(mask & ((1u << lane) - 1u))] = pid;
}

// =================================================================== raygen
#define FULL_THREADS 256

__global__ void __launch_bounds__(FULL_THREADS)
raygen_kernel(const DevScene sc, const PathState ps, const PassParams pp, uint32_t n_items,
              uint32_t* __restrict__ tq, uint32_t* __restrict__ tq_count)
{
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t stride = gridDim.x * blockDim.x;
  const uint32_t n_round = (n_items + 31u) & ~31u;
  for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n_round; idx += stride) {
    uint32_t pid = 0, pixel, s;
    const bool valid = idx < n_items && first_item(pp, idx, pid, pixel, s);
    bool complex_ray = false;
    if (valid) {
      // raygen_kernel (ray_gen.cu:11-32): seed, jitter (x then y), pinhole ray
      uint32_t rng = minstd_seed(wang_hash(wang_hash(pixel) ^ (pp.first_iteration + s)));
      const uint32_t y = pixel / pp.cam.width, x = pixel - y * pp.cam.width;
      const float fx = (float)x + minstd_uniform(rng);
      const float fy = (float)y + minstd_uniform(rng);
      f3 o, d;
      camera_ray(pp.cam, fx, fy, o, d);
      float tbest;
      uint32_t code;
      int start;
      complex_ray = classify(sc, o, d, 1e-4f, FLT_MAX, tbest, code, start);
      ps.ray[2 * (size_t)pid] = mk4(o, 1e-4f);
      ps.ray[2 * (size_t)pid + 1] = mk4(d, FLT_MAX);
      ps.thr[pid] = make_float4(1.0f, 1.0f, 1.0f, __uint_as_float(rng));
      ps.aux[pid] = make_uint4(__float_as_uint(tbest), code, (uint32_t)start, 0u);
    }
    warp_append(complex_ray, pid, tq, tq_count, lane);
  }
}

PT_D uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Stages `bytes` (a multiple of 16) from global to shared memory with ONE TMA bulk copy issued by
// thread 0; every thread of the CTA waits on the mbarrier the copy completes on.
PT_D void stage_prefix(void* smem_dst, const void* gsrc, uint32_t bytes, unsigned long long* s_bar)
{
  const uint32_t bar = smem_u32(s_bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(bar)
        : "memory");
  }
  uint32_t ready = 0;
  while (!ready) {
    asm volatile(
        "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
        : "=r"(ready)
        : "r"(bar)
        : "memory");
  }
}

// ================================================================= traverse
// Persistent warps (grid = SMs x resident CTAs).  Each lane owns one ray at a time.  Work is
// fetched from a device-side cursor with ONE atomic per refill for all idle lanes of the warp;
// lanes that finish early are retired (8-byte result) and refilled as soon as fewer than
// EXT_REFILL lanes are still walking, so the warp stays populated although per-ray traversal
// lengths vary by orders of magnitude.
#define EXT_THREADS 128
#ifndef EXT_MIN_BLOCKS
#define EXT_MIN_BLOCKS 8
#endif
#define EXT_REFILL 16
#define EXT_INNER_MIN 8
#ifndef PT_DEFAULT_FINISH
#define PT_DEFAULT_FINISH 2
#endif
#ifndef PT_DEFAULT_FINISH_RAYS
#define PT_DEFAULT_FINISH_RAYS 98304
#endif
#ifndef PT_DEFAULT_ORDER
#define PT_DEFAULT_ORDER 2 // measured: bunny +0.5 %, bunny_1m +1.5 %, terrain +5.5 % over order 0
#endif

#ifndef PT_INNER_STEPS
// Inner nodes a lane visits per pair of warp votes in traverse_kernel.  The kernel is bound by
// instruction issue and by the length of one lane's dependent chain, not by the L1 data pipe
// (halving the node bytes bought 2.5 %), and the two ballots, the population count and the branches
// of the while-while loop are ~20 of the ~80 instructions of a visit.  Measured (bunny, Mrays/s):
// 1 / 2 / 3 / 4 steps = 10 301 / 10 600 / 10 784 / 10 593; same order on 150 k and 2.6 M triangles.
#define PT_INNER_STEPS 3
#endif

enum { SRC_QUEUE = 1, SRC_BATCH = 2 };

// (Staging the hottest nodes of the binary tree in shared memory — area-ordered prefix, 80-byte
// stride against bank conflicts, up to 200 KB per SM — was measured and rejected: 15.6-17.1 ms
// against 15.7 ms for plain L1-cached loads on the bunny frame, profiles/README.md.)
//
// stream_state: the parked ray state is fetched with the evict-first policy (ld_state).
// MINB: resident CTAs per SM the register allocation is bounded for (8 -> 64 registers, 50 % of
// the warp slots; 10 -> 48; 12 -> 40); L256: node fetch with two 256-bit loads.  Both are
// run-time choices between instantiations (PT_TRAV="minb,l256"), measured in profiles/README.md.
// QN: the inner nodes are read in their 32-byte quantised form (DevScene::qnodes; trav_inner).
template <int SRC, int MINB, bool L256, bool ST, bool QN = false>
__global__ void __launch_bounds__(EXT_THREADS, MINB)
traverse_kernel(const DevScene sc, const PathState ps, const uint32_t* __restrict__ tq,
                const uint32_t* __restrict__ n_ptr, uint32_t n_host, uint32_t* __restrict__ work,
                const float4* __restrict__ batch_rays, HitRecord* __restrict__ batch_out,
                int refill_min, int inner_min, int stream_state, const BinLists bins)
{
  const uint32_t n = SRC == SRC_QUEUE ? *n_ptr : n_host;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t lt_mask = (1u << lane) - 1u;
  int stack[PT_STACK];
  Trav T;
  T.node = PT_SENTINEL;
  bool has = false;
  bool exhausted = false; // warp-uniform
  uint32_t pid = 0, code = 0;

  for (;;) {
    // ---- refill idle lanes: one atomic per warp
    const uint32_t need = __ballot_sync(0xffffffffu, !has);
    if (need != 0u && !exhausted) {
      const uint32_t cnt = (uint32_t)__popc(need);
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(work, cnt);
      base = __shfl_sync(0xffffffffu, base, 0);
      if (base + cnt >= n) exhausted = true;
      if (!has) {
        const uint32_t idx = base + (uint32_t)__popc(need & lt_mask);
        if (idx < n) {
          if (SRC == SRC_QUEUE) {
            pid = tq ? tq[idx] : idx; // no list: the state is compacted, slot == index
            const float4 ro = ld_state(ps.ray + 2 * (size_t)pid, (uint32_t)stream_state);
            const float4 rd = ld_state(ps.ray + 2 * (size_t)pid + 1, (uint32_t)stream_state);
            const uint4 ax = ld_state(ps.aux + pid, (uint32_t)stream_state);
            trav_init<QN>(sc, T, xyz(ro), xyz(rd), ro.w, __uint_as_float(ax.x), (int)ax.z, stack);
            has = true;
          } else {
            pid = idx;
            const float4 ro = batch_rays[2 * (size_t)idx], rd = batch_rays[2 * (size_t)idx + 1];
            float tbest;
            int start;
            const bool complex_ray = classify<ST>(sc, xyz(ro), xyz(rd), ro.w, rd.w, tbest, code, start);
            trav_init<QN>(sc, T, xyz(ro), xyz(rd), ro.w, tbest, start, stack);
            if (!complex_ray) T.node = PT_SENTINEL;
            has = true;
          }
        }
      }
    }
    if (__ballot_sync(0xffffffffu, has) == 0u) {
      if (exhausted) break;
      continue;
    }
    // ---- walk until too few lanes are still busy (all of them once the queue is drained).
    // while-while scheduling: lanes descend inner nodes together until (almost) all of them
    // stand on a leaf, then the leaves are intersected together — inner and leaf code never
    // interleave inside a warp, which is what kept 10 of 32 lanes busy in the if/else form.
    const int threshold = exhausted ? 1 : refill_min;
    for (;;) {
      const bool busy = has && T.node != PT_SENTINEL;
      if (__popc(__ballot_sync(0xffffffffu, busy)) < threshold) break;
      for (;;) {
        const bool inner = has && (uint32_t)T.node < (uint32_t)PT_SENTINEL;
        const uint32_t m_inner = __ballot_sync(0xffffffffu, inner);
        const uint32_t m_leaf = __ballot_sync(0xffffffffu, has && T.node < 0);
        if (m_inner == 0u) break;
        if (__popc(m_inner) < inner_min && m_leaf != 0u) break;
        if (inner) {
          trav_inner<L256, QN>(sc, T, stack);
          // further nodes before the next pair of votes (lanes that reached a leaf sit them out)
#pragma unroll
          for (int s = 1; s < PT_INNER_STEPS; ++s)
            if ((uint32_t)T.node < (uint32_t)PT_SENTINEL) trav_inner<L256, QN>(sc, T, stack);
        }
      }
      if (has && T.node < 0) trav_leaf(sc, T, stack);
    }
    // ---- retire finished lanes
    if (SRC == SRC_QUEUE && bins.list != nullptr) {
      // ray binning: one atomic per (warp, bin) for the lanes that retire now
      const bool retiring = has && T.node == PT_SENTINEL;
      uint32_t bin = PT_BINS; // not retiring
      if (retiring) {
        bin = 0u;
        if (T.best >= 0) {
          const uint32_t mat = (uint32_t)__float_as_int(ldg4(sc.tris + (size_t)T.best * 3 + 2).w);
          bin = 1u + (uint32_t)sc.materials[mat].type;
        }
      }
      const uint32_t peers = __match_any_sync(0xffffffffu, bin);
      if (retiring) {
        const int leader = __ffs(peers) - 1;
        uint32_t base = 0;
        if ((int)lane == leader) base = atomicAdd(bins.count + bin, (uint32_t)__popc(peers));
        base = __shfl_sync(peers, base, leader);
        PT_CHECK(base + (uint32_t)__popc(peers & lt_mask) < bins.cap && pid < bins.cap, "bin list slot out of range");
        bins.list[(size_t)bin * bins.cap + base + (uint32_t)__popc(peers & lt_mask)] = pid;
      }
    }
    if (has && T.node == PT_SENTINEL) {
      if (SRC == SRC_QUEUE) {
        if (T.best >= 0)
          *reinterpret_cast<uint2*>(ps.aux + pid) =
              make_uint2(__float_as_uint(T.tbest), AUX_TRI | (uint32_t)T.best);
      } else {
        // parity hook (pt_trace_batch): full Intersection + primitive ids
        if (T.best >= 0) code = AUX_TRI | (uint32_t)T.best;
        Hit h;
        HitRecord r;
        if (resolve_hit<ST>(sc, T.o, T.d, T.tmin, T.tbest, code, h)) {
          r.t = h.t;
          r.px = h.p.x, r.py = h.p.y, r.pz = h.p.z;
          r.nx = h.n.x, r.ny = h.n.y, r.nz = h.n.z;
          r.material = h.material, r.side = h.side;
          r.object = h.object, r.prim = h.prim;
        } else {
          r.t = -1.0f;
          r.px = r.py = r.pz = r.nx = r.ny = r.nz = 0.f;
          r.material = 0, r.side = 0, r.object = -1, r.prim = -1;
        }
        r.pad = 0;
        batch_out[pid] = r;
      }
      has = false;
    }
  }
}


// ================================================================ traverse8
// Traversal of the compressed 8-wide tree (node layout: common.cuh).  Per node visit a lane
// fetches 80 bytes and tests eight quantised child boxes, instead of 64 bytes and two boxes:
// about a third of the dependent fetch -> test -> fetch steps of the binary walk, and half its
// L1 wavefronts (ncu: the binary traverse_kernel keeps l1tex__data_pipe_lsu_wavefronts at 83 %
// of peak).  The top of the tree (breadth-first prefix) is staged into shared memory once per
// CTA with one TMA bulk copy; deeper nodes and the triangles come through L1/L2.
//
// Per-lane state follows Ylitie et al. 2017: a node group (first child, hit bits of the inner
// children in octant order | imask) and a triangle group (first triangle, hit bits); the stack
// holds node groups only.  Scheduling is the while-while form of traverse_kernel: lanes take
// node steps together, then triangle steps together, and a persistent warp refills lanes that
// ran out of work.
PT_D void tri_test(const DevScene& sc, Trav& T, uint32_t slot)
{
  const float4* tp = sc.tris + (size_t)slot * 3;
  const float4 t0 = ldg4(tp + 0);
  const float4 t1 = ldg4(tp + 1);
  const float4 t2 = ldg4(tp + 2);
  const f3 e1 = xyz(t1), e2 = xyz(t2);
  const f3 hh = cross3(T.d, e2);
  const float a = dot3(e1, hh);
  const float f = 1.0f / a;
  const f3 s = T.o - xyz(t0);
  const float u = f * dot3(s, hh);
  const f3 q = cross3(s, e1);
  const float v = f * dot3(T.d, q);
  const float t = f * dot3(e2, q);
  const bool ok = !(a > -0.0000001f && a < 0.0000001f) && !(u < 0.0f || u > 1.0f) &&
                  !(v < 0.0f || u + v > 1.0f) && (t >= T.tmin && t <= T.tbest);
  if (ok) {
    T.tbest = t;
    T.best = (int)slot;
  }
}

// One quantised plane: byte k of `w` dropped into the mantissa of 2^23 (bits 8..15) by a single
// PRMT, so that t = m * A + B needs no integer->float conversion.  A already carries the
// 1/256 of the byte position, B the -2^23 * A of the magic offset.  `magic` (0x4B000000) is kept
// in a register on purpose: PRMT takes one immediate, and it has to be the selector.
template <int K> PT_D float qplane(uint32_t w, uint32_t magic, float A, float B)
{
  return fmaf(__uint_as_float(__byte_perm(w, magic, 0x7404u | (K << 4))), A, B);
}

// bit4/cnt4: per-byte hit-bit position and unary triangle count (1 for an inner child) of four slots
template <int K>
PT_D void child8_test(uint32_t nx, uint32_t ny, uint32_t nz, uint32_t fx, uint32_t fy, uint32_t fz,
                      uint32_t bit4, uint32_t cnt4, uint32_t magic, float Ax, float Ay, float Az,
                      float Bx, float By, float Bz, float tmin, float tbest, uint32_t& hitmask)
{
  const float tnx = qplane<K>(nx, magic, Ax, Bx), tny = qplane<K>(ny, magic, Ay, By),
              tnz = qplane<K>(nz, magic, Az, Bz);
  const float tfx = qplane<K>(fx, magic, Ax, Bx), tfy = qplane<K>(fy, magic, Ay, By),
              tfz = qplane<K>(fz, magic, Az, Bz);
  const float cmin = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, tmin));
  const float cmax = fminf(fminf(tfx, tfy), fminf(tfz, tbest));
  const uint32_t bits = __byte_perm(cnt4, 0u, 0x4440u | K) << __byte_perm(bit4, 0u, 0x4440u | K);
  // robust slab comparison (Ize 2013): widen the far side by 2 ulp
  if (cmax * 1.0000004f >= cmin) hitmask |= bits;
}

PT_D void node8_test(const uint4* __restrict__ np, const Trav& T, uint32_t octm, uint32_t magic,
                     uint2& ng, uint2& tg)
{
  const uint4 w0 = np[0], w1 = np[1], w2 = np[2], w3 = np[3], w4 = np[4];
  const float Ax = __uint_as_float((w0.w & 0xffu) << 23) * T.idx;
  const float Ay = __uint_as_float((w0.w & 0xff00u) << 15) * T.idy;
  const float Az = __uint_as_float((w0.w & 0xff0000u) << 7) * T.idz;
  const float Bx = fmaf(-8388608.0f, Ax, (__uint_as_float(w0.x) - T.o.x) * T.idx);
  const float By = fmaf(-8388608.0f, Ay, (__uint_as_float(w0.y) - T.o.y) * T.idy);
  const float Bz = fmaf(-8388608.0f, Az, (__uint_as_float(w0.z) - T.o.z) * T.idz);
  // entry planes are the low planes for a positive direction, the high planes otherwise
  const bool sx = T.idx < 0.0f, sy = T.idy < 0.0f, sz = T.idz < 0.0f;
  const uint32_t nx0 = sx ? w3.z : w2.x, nx1 = sx ? w3.w : w2.y;
  const uint32_t fx0 = sx ? w2.x : w3.z, fx1 = sx ? w2.y : w3.w;
  const uint32_t ny0 = sy ? w4.x : w2.z, ny1 = sy ? w4.y : w2.w;
  const uint32_t fy0 = sy ? w2.z : w4.x, fy1 = sy ? w2.w : w4.y;
  const uint32_t nz0 = sz ? w4.z : w3.x, nz1 = sz ? w4.w : w3.y;
  const uint32_t fz0 = sz ? w3.x : w4.z, fz1 = sz ? w3.y : w4.w;
  // hit-bit positions, four slots at a time: low five bits of meta, inner slots (>= 24) XORed
  // with the ray's octant so that the nearest inner child owns the highest bit
  const uint32_t oct4 = octm * 0x01010101u;
  const uint32_t i0 = w1.z & 0x1f1f1f1fu, i1 = w1.w & 0x1f1f1f1fu;
  const uint32_t in0 = ((i0 + 0x08080808u) >> 5) & 0x01010101u, in1 = ((i1 + 0x08080808u) >> 5) & 0x01010101u;
  const uint32_t bit0 = i0 ^ (oct4 & (in0 * 7u)), bit1 = i1 ^ (oct4 & (in1 * 7u));
  const uint32_t cnt0 = (w1.z >> 5) & 0x07070707u, cnt1 = (w1.w >> 5) & 0x07070707u;
  uint32_t hm = 0u;
  child8_test<0>(nx0, ny0, nz0, fx0, fy0, fz0, bit0, cnt0, magic, Ax, Ay, Az, Bx, By, Bz, T.tmin, T.tbest, hm);
  child8_test<1>(nx0, ny0, nz0, fx0, fy0, fz0, bit0, cnt0, magic, Ax, Ay, Az, Bx, By, Bz, T.tmin, T.tbest, hm);
  child8_test<2>(nx0, ny0, nz0, fx0, fy0, fz0, bit0, cnt0, magic, Ax, Ay, Az, Bx, By, Bz, T.tmin, T.tbest, hm);
  child8_test<3>(nx0, ny0, nz0, fx0, fy0, fz0, bit0, cnt0, magic, Ax, Ay, Az, Bx, By, Bz, T.tmin, T.tbest, hm);
  child8_test<0>(nx1, ny1, nz1, fx1, fy1, fz1, bit1, cnt1, magic, Ax, Ay, Az, Bx, By, Bz, T.tmin, T.tbest, hm);
  child8_test<1>(nx1, ny1, nz1, fx1, fy1, fz1, bit1, cnt1, magic, Ax, Ay, Az, Bx, By, Bz, T.tmin, T.tbest, hm);
  child8_test<2>(nx1, ny1, nz1, fx1, fy1, fz1, bit1, cnt1, magic, Ax, Ay, Az, Bx, By, Bz, T.tmin, T.tbest, hm);
  child8_test<3>(nx1, ny1, nz1, fx1, fy1, fz1, bit1, cnt1, magic, Ax, Ay, Az, Bx, By, Bz, T.tmin, T.tbest, hm);
  ng = make_uint2(w1.x, (hm & 0xff000000u) | (w0.w >> 24));
  tg = make_uint2(w1.y, hm & 0x00ffffffu);
}

template <int SRC, int THREADS, int BLOCKS>
__global__ void __launch_bounds__(THREADS, BLOCKS)
traverse8_kernel(const DevScene sc, const PathState ps, const uint32_t* __restrict__ tq,
                 const uint32_t* __restrict__ n_ptr, uint32_t n_host, uint32_t* __restrict__ work,
                 const float4* __restrict__ batch_rays, HitRecord* __restrict__ batch_out,
                 int refill_min, int node_min, int tri_min, uint32_t n_staged,
                 uint32_t magic /* 0x4B000000, a parameter so that ptxas keeps it in a register */)
{
  extern __shared__ __align__(128) uint4 s_nodes[];
  __shared__ __align__(8) unsigned long long s_bar;
  const uint32_t n = SRC == SRC_QUEUE ? *n_ptr : n_host;
  if (n == 0u) return;
  // ---- stage the breadth-first prefix of the tree: one TMA bulk copy, completion on an mbarrier
  if (n_staged != 0u) stage_prefix(s_nodes, sc.nodes8, n_staged * 80u, &s_bar);

  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t lt_mask = (1u << lane) - 1u;
  uint2 stack[PT_STACK8];
  Trav T;
  T.best = -1;
  uint2 ng = make_uint2(0u, 0u), tg = make_uint2(0u, 0u);
  int sp = 0;
  uint32_t octm = 0;
  bool has = false, done = false;
  bool exhausted = false; // warp-uniform
  uint32_t pid = 0, code = 0;

  for (;;) {
    // ---- refill idle lanes: one atomic per warp
    const uint32_t need = __ballot_sync(0xffffffffu, !has);
    if (need != 0u && !exhausted) {
      const uint32_t cnt = (uint32_t)__popc(need);
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(work, cnt);
      base = __shfl_sync(0xffffffffu, base, 0);
      if (base + cnt >= n) exhausted = true;
      if (!has) {
        const uint32_t idx = base + (uint32_t)__popc(need & lt_mask);
        if (idx < n) {
          float4 ro, rd;
          float tbest;
          bool complex_ray = true;
          if (SRC == SRC_QUEUE) {
            pid = tq ? tq[idx] : idx; // no list: the state is compacted, slot == index
            ro = ps.ray[2 * (size_t)pid];
            rd = ps.ray[2 * (size_t)pid + 1];
            tbest = __uint_as_float(ps.aux[pid].x);
          } else {
            pid = idx;
            ro = batch_rays[2 * (size_t)idx], rd = batch_rays[2 * (size_t)idx + 1];
            int start;
            complex_ray = classify(sc, xyz(ro), xyz(rd), ro.w, rd.w, tbest, code, start);
          }
          T.o = xyz(ro);
          T.d = xyz(rd);
          T.tmin = ro.w;
          T.tbest = tbest;
          T.best = -1;
          T.idx = safe_inv(rd.x);
          T.idy = safe_inv(rd.y);
          T.idz = safe_inv(rd.z);
          octm = (T.idx < 0.0f ? 0u : 1u) | (T.idy < 0.0f ? 0u : 2u) | (T.idz < 0.0f ? 0u : 4u);
          ng = make_uint2(0u, 0x80000000u); // "child 0 of nothing" = the root
          tg = make_uint2(0u, 0u);
          sp = 0;
          done = !complex_ray;
          has = true;
        }
      }
    }
    if (__ballot_sync(0xffffffffu, has) == 0u) {
      if (exhausted) break;
      continue;
    }
    const int threshold = exhausted ? 1 : refill_min;
    for (;;) {
      if (__popc(__ballot_sync(0xffffffffu, has && !done)) < threshold) break;
      // ---- node steps
      for (;;) {
        const bool act = has && !done;
        const bool nw = act && tg.y == 0u;
        const uint32_t m_node = __ballot_sync(0xffffffffu, nw);
        const uint32_t m_tri = __ballot_sync(0xffffffffu, act && tg.y != 0u);
        if (m_node == 0u) break;
        if (__popc(m_node) < node_min && m_tri != 0u) break;
        if (nw) {
          const uint32_t hits = ng.y;
          const int bit = 31 - __clz((int)hits);
          const uint32_t rest = hits & ~(1u << bit);
          const uint32_t slot = (uint32_t)(bit - 24) ^ octm;
          const uint32_t node = ng.x + (uint32_t)__popc(hits & 0xffu & ((1u << slot) - 1u));
          const uint4* np = node < n_staged ? s_nodes + (size_t)node * 5 : sc.nodes8 + (size_t)node * 5;
          const uint2 rest_g = make_uint2(ng.x, rest);
          uint2 ng2;
          node8_test(np, T, octm, magic, ng2, tg);
          if (ng2.y & 0xff000000u) {
            if (rest & 0xff000000u) stack[sp++] = rest_g;
            ng = ng2;
          } else {
            ng = rest_g;
          }
          if (tg.y == 0u && (ng.y & 0xff000000u) == 0u) {
            if (sp == 0) {
              done = true;
            } else {
              ng = stack[--sp];
            }
          }
        }
      }
      // ---- triangle steps (the first one is unconditional: with few lanes on either side the
      // two phases must not keep deferring to each other)
      for (bool first_step = true;; first_step = false) {
        const bool act = has && !done;
        const bool tw = act && tg.y != 0u;
        const uint32_t m_tri = __ballot_sync(0xffffffffu, tw);
        const uint32_t m_node = __ballot_sync(0xffffffffu, act && tg.y == 0u);
        if (m_tri == 0u) break;
        if (!first_step && __popc(m_tri) < tri_min && m_node != 0u) break;
        if (tw) {
          const uint32_t b = (uint32_t)__ffs((int)tg.y) - 1u;
          tg.y &= tg.y - 1u;
          tri_test(sc, T, tg.x + b);
          if (tg.y == 0u && (ng.y & 0xff000000u) == 0u) {
            if (sp == 0) {
              done = true;
            } else {
              ng = stack[--sp];
            }
          }
        }
      }
    }
    // ---- retire finished lanes
    if (has && done) {
      if (SRC == SRC_QUEUE) {
        if (T.best >= 0)
          *reinterpret_cast<uint2*>(ps.aux + pid) =
              make_uint2(__float_as_uint(T.tbest), AUX_TRI | (uint32_t)T.best);
      } else {
        if (T.best >= 0) code = AUX_TRI | (uint32_t)T.best;
        Hit h;
        HitRecord r;
        if (resolve_hit(sc, T.o, T.d, T.tmin, T.tbest, code, h)) {
          r.t = h.t;
          r.px = h.p.x, r.py = h.p.y, r.pz = h.p.z;
          r.nx = h.n.x, r.ny = h.n.y, r.nz = h.n.z;
          r.material = h.material, r.side = h.side;
          r.object = h.object, r.prim = h.prim;
        } else {
          r.t = -1.0f;
          r.px = r.py = r.pz = r.nx = r.ny = r.nz = 0.f;
          r.material = 0, r.side = 0, r.object = -1, r.prim = -1;
        }
        r.pad = 0;
        batch_out[pid] = r;
      }
      has = false;
    }
  }
}

// ==================================================================== shade
// random_in_unit_sphere (distributions.cuh:6-19): uniform ON the sphere.
PT_D f3 random_on_sphere(uint32_t& rng)
{
  const float phi = (2.0f * 3.14159265358979323846264338327950288f) * minstd_uniform(rng);
  const float cos_theta = 2.0f * minstd_uniform(rng) - 1.0f;
  const float sin_theta = sqrtf(1.0f - cos_theta * cos_theta);
  float sp, cp; // one shared range reduction; same values as sinf(phi), cosf(phi)
  sincosf(phi, &sp, &cp);
  return mk3(cp * sin_theta, sp * sin_theta, cos_theta);
}

PT_D float sign1(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }

// get_background_color (path_tracer.cu:29-34)
PT_D f3 sky_color(f3 d)
{
  const f3 unit = normalize3(d);
  const float t = 0.5f * (unit.y + 1.0f);
  return mk3(0.5f, 0.7f, 1.0f) * (1.0f - t) + mk3(1.0f, 1.0f, 1.0f) * t;
}

// evaluate_material (path_tracer.cu:138-201): scatters the ray in place and updates the
// throughput.  Draw order and count per material are the reference's (lambert/metal: 2,
// dielectric: 1 and only when refraction is possible).
PT_D void scatter(const DevMaterial& mat, const Hit& h, f3& o, f3& d, float& tmin, f3& color,
                  uint32_t& rng)
{
  const f3 n = h.n;
  f3 origin = h.p - (1e-4f * sign1(dot3(d, n))) * n;
  f3 dir;
  if (mat.type == 0) {
    f3 sd = normalize3(n + random_on_sphere(rng));
    if (fabsf(sd.x) < 1e-8f && fabsf(sd.y) < 1e-8f && fabsf(sd.z) < 1e-8f) sd = n;
    dir = sd;
    color = color * mk3(mat.r, mat.g, mat.b);
  } else if (mat.type == 1) {
    const f3 reflected = d - n * dot3(n, d) * 2.0f; // glm::reflect
    dir = reflected + mat.param * random_on_sphere(rng);
    if (dot3(dir, n) > 0.0f) {
      color = color * mk3(mat.r, mat.g, mat.b);
    } else {
      color = mk3(0.f, 0.f, 0.f);
    }
  } else {
    const float ior = mat.param;
    const float ratio = h.side == 0u ? (1.0f / ior) : ior;
    const f3 unit = normalize3(d);
    const float cos_theta = fminf(dot3(-unit, n), 1.0f);
    const float sin_theta = sqrtf(1.0f - cos_theta * cos_theta);
    const bool cannot_refract = ratio * sin_theta > 1.0f;
    bool reflect_it = cannot_refract;
    if (!cannot_refract) {
      // Schlick (path_tracer.cu:130-136); the draw happens only on this branch
      float r0 = (1.0f - ratio) / (1.0f + ratio);
      r0 = r0 * r0;
      const float refl = r0 + (1.0f - r0) * powf(1.0f - cos_theta, 5.0f);
      reflect_it = refl > minstd_uniform(rng);
    }
    if (reflect_it) {
      dir = unit - n * dot3(n, unit) * 2.0f;
    } else {
      // glm::refract
      const float dv = dot3(n, unit);
      const float k = 1.0f - ratio * ratio * (1.0f - dv * dv);
      dir = k >= 0.0f ? (ratio * unit - (ratio * dv + sqrtf(k)) * n) : mk3(0.f, 0.f, 0.f);
    }
    origin = h.p;
    tmin = 1e-5f;
  }
  o = origin;
  d = dir;
}

template <bool FIRST>
__global__ void __launch_bounds__(FULL_THREADS)
shade_kernel(const DevScene sc, const PathState ps, const PassParams pp,
             const uint32_t* __restrict__ queue, const uint32_t* __restrict__ n_ptr,
             uint32_t n_first, uint32_t* __restrict__ next_queue,
             uint32_t* __restrict__ next_count, uint32_t* __restrict__ tq,
             uint32_t* __restrict__ tq_count, uint8_t* __restrict__ flags, uint32_t bounce,
             uint32_t last_bounce)
{
  const uint32_t n = FIRST ? n_first : *n_ptr;
  const uint32_t n_round = (n + 31u) & ~31u;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n_round; idx += stride) {
    bool alive = false, complex_ray = false;
    uint32_t pid = 0, slot = idx;
    bool valid = idx < n;
    if (valid) {
      if (FIRST) {
        uint32_t s;
        valid = first_item(pp, idx, pid, slot, s); // slot = pixel index at bounce 0
      } else {
        pid = queue[idx];
      }
    }
    if (valid) {
      const float4 ro = ps.ray[2 * (size_t)pid];
      const float4 rd = ps.ray[2 * (size_t)pid + 1];
      const float4 th = ps.thr[pid];
      const uint4 ax = ps.aux[pid];
      f3 o = xyz(ro), d = xyz(rd);
      float tmin = ro.w;
      f3 color = xyz(th);
      uint32_t rng = __float_as_uint(th.w);
      if (pp.rng_mode == 1u) {
        // reference streaming mode: re-seed from the compacted slot index and
        // discard(bounce) (path_tracer.cu:300-301)
        rng = minstd_discard(minstd_seed(wang_hash(wang_hash(slot) ^ pp.first_iteration)), bounce);
      }
      Hit h;
      if (!resolve_hit(sc, o, d, tmin, __uint_as_float(ax.x), ax.y, h)) {
        color = color * sky_color(d);
        if (FIRST) ps.gbuf[pid] = make_float4(-d.x, -d.y, -d.z, 1e6f);
      } else {
        if (FIRST) ps.gbuf[pid] = make_float4(h.n.x, h.n.y, h.n.z, h.t);
        const DevMaterial mat = sc.materials[h.material];
        scatter(mat, h, o, d, tmin, color, rng);
        alive = last_bounce == 0u;
        if (alive) {
          float tbest;
          uint32_t code;
          int start;
          complex_ray = classify(sc, o, d, tmin, FLT_MAX, tbest, code, start);
          ps.ray[2 * (size_t)pid] = mk4(o, tmin);
          ps.ray[2 * (size_t)pid + 1] = mk4(d, FLT_MAX);
          ps.aux[pid] = make_uint4(__float_as_uint(tbest), code, (uint32_t)start, 0u);
        }
      }
      ps.thr[pid] = mk4(color, __uint_as_float(rng));
      if (pp.rng_mode == 1u && last_bounce == 0u) flags[slot] = alive ? 1 : 0;
    }
    if (last_bounce == 0u) {
      if (pp.rng_mode != 1u) warp_append(alive, pid, next_queue, next_count, lane);
      warp_append(complex_ray, pid, tq, tq_count, lane);
    }
  }
}

// ==================================================================== chain
// PT_RNG_PIXEL_STREAM scheduler.  Paths carry their own RNG stream and depth, so nothing forces
// all paths to advance bounce by bounce: a thread keeps shading its path IN REGISTERS for as
// long as the next ray is "simple" (classification proves it never enters the mesh BVH: it hits
// a sphere or the sky) and only parks it — ray 32 B, throughput 16 B, aux 16 B, path id 4 B —
// when a ray needs the traversal kernel.  In open scenes most rays are simple (59 % in the
// bunny scene), so most bounces never touch HBM, and the queue that does go through
// traverse_kernel is homogeneous.  The counted quantity is unchanged: one ray per intersection
// resolved (== paths entering the reference's intersection_kernel, path_tracer.cu:428).
//   FIRST: items are the tile-ordered primary samples (raygen fused);
//   else : items are the paths whose ray traverse_kernel just finished.
// (A persistent variant in which a lane that parks or terminates its path immediately takes the
// next item — lane refill as in traverse_kernel — was measured and rejected: lane occupancy
// rose but chain<true> went 4.65 -> 7.0 ms and chain<false> 6.05 -> 6.6 ms on the bunny frame;
// these kernels are bound by the latency of their path-state loads/stores, not by issue slots.)
//
// Parked paths are COMPACTED: the state a parked path needs to continue (ray 32 B, throughput +
// RNG 16 B, aux 16 B, path id 4 B) is written to slot = its position in the next iteration's
// list, so the append, the traversal's ray fetch / result store and the next chain launch's
// reload are contiguous streams — no path-id gather anywhere between the two ends of a path.
// (Measured: regrouping the list by direction octant, which scatters these accesses, slowed the
// next chain launch by 30-100 %.)  Two park buffers alternate between iterations.
PT_D uint32_t warp_slot(bool push, uint32_t* __restrict__ count, uint32_t lane)
{
  const uint32_t mask = __ballot_sync(0xffffffffu, push);
  if (mask == 0u) return 0u;
  const int leader = __ffs(mask) - 1;
  uint32_t base = 0;
  if ((int)lane == leader) base = atomicAdd(count, (uint32_t)__popc(mask));
  base = __shfl_sync(0xffffffffu, base, leader);
  return base + (uint32_t)__popc(mask & ((1u << lane) - 1u));
}

// TMA (re-entry launches only, opt-in PT_CHAIN_TMA=1|2): the parked state a launch consumes is
// four contiguous streams, so it can be staged into shared memory with cp.async.bulk completing on
// an mbarrier, double-buffered, while the previous tile is shaded.  Two granularities were built
// and measured (profiles/README.md, round 2): 1 = one 256-record tile (17 KB) per CTA and stage,
// with a CTA barrier before a stage is refilled; 2 = one 32-record tile (2 176 B) per WARP and
// stage, no CTA barrier.  Both take the stream reads off the long-scoreboard list (4.0 -> 2.4
// stalled warps per issue) and both are SLOWER than plain loads (chain<false> +3..+10 %): the
// kernel's time is in the dependent scene fetches after the state arrives, and the staging adds a
// synchronisation per tile.  Default off.
#define CHAIN_TILE FULL_THREADS
#define CHAIN_STAGE_BYTES (CHAIN_TILE * 68u)
#define CHAIN_WSTAGE_BYTES (32u * 68u)
template <uint32_t RECORDS>
PT_D void chain_stage_issue(const ParkBuf& in, size_t first, unsigned char* stage, unsigned long long* bar)
{
  const uint32_t b = smem_u32(bar);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // generic reads of this stage are done
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(RECORDS * 68u) : "memory");
  const uint32_t dst = smem_u32(stage);
#define PT_BULK(off, src, bytes)                                                                   \
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"( \
                   dst + (off)),                                                                   \
               "l"(src), "r"(bytes), "r"(b)                                                        \
               : "memory")
  PT_BULK(0u, in.ray + 2 * first, RECORDS * 32u);
  PT_BULK(RECORDS * 32u, in.thr + first, RECORDS * 16u);
  PT_BULK(RECORDS * 48u, in.aux + first, RECORDS * 16u);
  PT_BULK(RECORDS * 64u, in.pid + first, RECORDS * 4u);
#undef PT_BULK
}
PT_D void mbar_wait(unsigned long long* bar, uint32_t parity)
{
  const uint32_t b = smem_u32(bar);
  uint32_t ready = 0;
  while (!ready) {
    asm volatile(
        "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
        : "=r"(ready)
        : "r"(b), "r"(parity)
        : "memory");
  }
}

// Three CTAs of 256 threads per SM, stated: left to itself nvcc also stops at 80 registers but keeps
// 64 B of stack instead of 48 (re-entry launches 5.30 -> 5.14 ms on the bunny, 15.5 -> 14.7 on the
// terrain).  The kernel is latency bound, yet more warps do not pay for fewer registers: 4 CTAs
// (64 registers) 5.89 / 25.5 ms, 5 CTAs (48) 6.84 / 29.9 ms; nor do more registers for fewer warps:
// 2 CTAs (86 registers) 6.54 / 18.6 ms (profiles/r2_ab_chain_register_bounds.log).
#ifndef CHAIN_MIN_BLOCKS
#define CHAIN_MIN_BLOCKS 3
#endif
template <bool FIRST, int TMA, bool ST>
__global__ void __launch_bounds__(FULL_THREADS, CHAIN_MIN_BLOCKS)
chain_kernel(const DevScene sc, const PathState ps, const PassParams pp, const ParkBuf in,
             const uint32_t* __restrict__ n_ptr, uint32_t n_first, const ParkBuf out,
             uint32_t* __restrict__ out_count, uint32_t max_depth,
             unsigned long long* __restrict__ total_rays, const BinLists bins)
{
  const uint32_t n = FIRST ? n_first : *n_ptr;
  // ray binning: the items are the bins' lists back to back (their counts sum to n)
  uint32_t bin_end[PT_BINS] = {0u, 0u, 0u, 0u};
  if (!FIRST && bins.list != nullptr) {
    uint32_t acc = 0;
#pragma unroll
    for (int b = 0; b < PT_BINS; ++b) {
      acc += bins.count[b];
      bin_end[b] = acc;
    }
  }
  const uint32_t n_round = (n + 31u) & ~31u;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t stride = gridDim.x * blockDim.x;
  const uint32_t cs = pp.stream_state;
  uint32_t rays_local = 0;
  extern __shared__ __align__(128) unsigned char s_stage[]; // TMA: 2 stages per CTA, or per warp
  __shared__ __align__(8) unsigned long long s_full[2 * (FULL_THREADS / 32)];
  const uint32_t n_tiles = n_round == 0u ? 0u : (n_round + CHAIN_TILE - 1u) / CHAIN_TILE;
  const uint32_t warp = threadIdx.x >> 5;
  if (TMA == 1) {
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_full[0])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_full[1])));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      if (blockIdx.x < n_tiles)
        chain_stage_issue<CHAIN_TILE>(in, (size_t)blockIdx.x * CHAIN_TILE, s_stage, &s_full[0]);
      if (blockIdx.x + gridDim.x < n_tiles)
        chain_stage_issue<CHAIN_TILE>(in, (size_t)(blockIdx.x + gridDim.x) * CHAIN_TILE, s_stage + CHAIN_STAGE_BYTES,
                                      &s_full[1]);
    }
  } else if (TMA == 2) {
    if (lane == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_full[2 * warp])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_full[2 * warp + 1])));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      const uint32_t first = blockIdx.x * blockDim.x + warp * 32u;
      unsigned char* ws = s_stage + warp * 2u * CHAIN_WSTAGE_BYTES;
      if (first < n_round) chain_stage_issue<32u>(in, first, ws, &s_full[2 * warp]);
      if (first + stride < n_round)
        chain_stage_issue<32u>(in, (size_t)first + stride, ws + CHAIN_WSTAGE_BYTES, &s_full[2 * warp + 1]);
    }
    __syncwarp();
  }
  // TMA 1: whole CTAs iterate together (the barrier below), padding lanes are simply not valid
  const uint32_t n_loop = TMA == 1 ? n_tiles * CHAIN_TILE : n_round;
  uint32_t k_iter = 0;
  for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n_loop; idx += stride, ++k_iter) {
    uint32_t pid = 0, pixel, s, px, py;
    bool valid = idx < n;
    float4 t_ro, t_rd, t_th;
    uint4 t_ax;
    uint32_t t_pid = 0;
    if (TMA == 1) {
      const uint32_t st = k_iter & 1u;
      mbar_wait(&s_full[st], (k_iter >> 1) & 1u);
      const unsigned char* base = s_stage + st * CHAIN_STAGE_BYTES;
      const float4* sr = reinterpret_cast<const float4*>(base);
      t_ro = sr[2 * threadIdx.x];
      t_rd = sr[2 * threadIdx.x + 1];
      t_th = reinterpret_cast<const float4*>(base + CHAIN_TILE * 32u)[threadIdx.x];
      t_ax = reinterpret_cast<const uint4*>(base + CHAIN_TILE * 48u)[threadIdx.x];
      t_pid = reinterpret_cast<const uint32_t*>(base + CHAIN_TILE * 64u)[threadIdx.x];
      __syncthreads(); // the stage has been read by everyone: refill it with the tile after next
      if (threadIdx.x == 0) {
        const uint32_t next = blockIdx.x + (k_iter + 2u) * gridDim.x;
        if (next < n_tiles)
          chain_stage_issue<CHAIN_TILE>(in, (size_t)next * CHAIN_TILE, s_stage + st * CHAIN_STAGE_BYTES, &s_full[st]);
      }
    } else if (TMA == 2) {
      const uint32_t st = k_iter & 1u;
      mbar_wait(&s_full[2 * warp + st], (k_iter >> 1) & 1u);
      unsigned char* base = s_stage + (warp * 2u + st) * CHAIN_WSTAGE_BYTES;
      const float4* sr = reinterpret_cast<const float4*>(base);
      t_ro = sr[2 * lane];
      t_rd = sr[2 * lane + 1];
      t_th = reinterpret_cast<const float4*>(base + 32u * 32u)[lane];
      t_ax = reinterpret_cast<const uint4*>(base + 32u * 48u)[lane];
      t_pid = reinterpret_cast<const uint32_t*>(base + 32u * 64u)[lane];
      __syncwarp(); // the warp's stage has been read: refill it with its tile after next
      if (lane == 0) {
        const size_t next = (size_t)idx + 2u * (size_t)stride; // lane 0's idx = the warp's first item
        if (next < n_round) chain_stage_issue<32u>(in, next, base, &s_full[2 * warp + st]);
      }
    }
    if (FIRST && valid) valid = first_item(pp, idx, pid, pixel, s, px, py);
    bool park = false; // the path leaves this kernel with a ray that needs the BVH
    f3 o = mk3(0.f, 0.f, 0.f), d = o, color = o;
    float tmin = 0.f, tbest = 0.f;
    uint32_t rng = 0u, code = 0u, depth = 0u;
    int start = 0;
    if (valid) {
      bool need_traversal;
      Hit h;
      bool have_hit = false; // h already holds the Intersection of `code` (a sphere found by classify)
      if (FIRST) {
        // raygen_kernel (ray_gen.cu:11-32): seed, jitter (x then y), pinhole ray
        rng = minstd_seed(wang_hash(wang_hash(pixel) ^ (pp.first_iteration + s)));
        const float fx = (float)px + minstd_uniform(rng);
        const float fy = (float)py + minstd_uniform(rng);
        camera_ray(pp.cam, fx, fy, o, d);
        tmin = 1e-4f;
        color = mk3(1.0f, 1.0f, 1.0f);
        depth = 0;
        need_traversal = classify<ST>(sc, o, d, tmin, FLT_MAX, tbest, code, start, h);
        have_hit = true;
      } else {
        uint32_t slot = idx;
        if (TMA == 0 && bins.list != nullptr) {
          int b = 0;
#pragma unroll
          for (int k = 0; k < PT_BINS - 1; ++k) b += idx >= bin_end[k] ? 1 : 0;
          slot = bins.list[(size_t)b * bins.cap + (idx - (b ? bin_end[b - 1] : 0u))];
          PT_CHECK(slot < bins.cap, "bin list entry out of range");
        }
        const float4 ro = TMA ? t_ro : ld_state(in.ray + 2 * (size_t)slot, cs);
        const float4 rd = TMA ? t_rd : ld_state(in.ray + 2 * (size_t)slot + 1, cs);
        const float4 th = TMA ? t_th : ld_state(in.thr + slot, cs);
        const uint4 ax = TMA ? t_ax : ld_state(in.aux + slot, cs);
        pid = TMA ? t_pid : ld_state(in.pid + slot, cs);
        o = xyz(ro), d = xyz(rd);
        tmin = ro.w;
        color = xyz(th);
        rng = __float_as_uint(th.w);
        tbest = __uint_as_float(ax.x);
        code = ax.y;
        depth = ax.w;
        need_traversal = false; // traverse_kernel has refined aux already
      }
      for (;;) {
        if (need_traversal) {
          park = true;
          break;
        }
        ++rays_local;
        // a simple ray's sphere hit comes straight from classification (same arithmetic, same
        // root as resolve_hit would pick); traversed rays rebuild theirs from the aux word
        const bool hit = have_hit ? code != 0u : resolve_hit<ST>(sc, o, d, tmin, tbest, code, h);
        if (depth == 0u) {
          st_state(ps.gbuf + pid,
                   hit ? make_float4(h.n.x, h.n.y, h.n.z, h.t) : make_float4(-d.x, -d.y, -d.z, 1e6f), cs);
        }
        if (!hit) {
          color = color * sky_color(d);
          break;
        }
        const DevMaterial mat = sc.materials[h.material];
        scatter(mat, h, o, d, tmin, color, rng);
        if (++depth == max_depth) break; // survivors contribute their throughput (path_tracer.cu:252-265)
        need_traversal = classify<ST>(sc, o, d, tmin, FLT_MAX, tbest, code, start, h);
        have_hit = true;
      }
      if (!park) st_state(ps.thr + pid, mk4(color, __uint_as_float(rng)), cs); // the path's contribution
    }
    const uint32_t slot = warp_slot(park, out_count, lane);
    if (park) {
      PT_CHECK(slot < pp.capacity && pid < pp.capacity, "parked-state slot out of range");
      st_state(out.ray + 2 * (size_t)slot, mk4(o, tmin), cs);
      st_state(out.ray + 2 * (size_t)slot + 1, mk4(d, FLT_MAX), cs);
      st_state(out.thr + slot, mk4(color, __uint_as_float(rng)), cs);
      st_state(out.aux + slot, make_uint4(__float_as_uint(tbest), code, (uint32_t)start, depth), cs);
      st_state(out.pid + slot, pid, cs);
    }
  }
  // one 64-bit atomic per warp for the ray counter
  for (int off = 16; off > 0; off >>= 1) rays_local += __shfl_down_sync(0xffffffffu, rays_local, off);
  if (lane == 0 && rays_local != 0u) atomicAdd(total_rays, (unsigned long long)rays_local);
}

// =================================================================== finish
// The late bounces of a pass in one launch, once few rays are left.  A pass is a chain of dependent launches
// (traverse, chain per bounce); the late ones hold a few thousand rays and still cost ~50 us of
// traversal + ~12 us of shading each, because one ray's dependent walk sets the floor (frame launch
// list, profiles/README.md).  From the first bounce at which the previous pass of the context had
// at most PT_FINISH_RAYS rays parked (scaled to this pass's size; render_pass), the paths still
// parked are finished here: one lane per path, traversal inlined where the wavefront would park — the
// chain loop of chain_kernel with trav_inner/trav_leaf in place of the hand-over.  Divergent, but
// over few paths (forced on millions of rays it costs 32 % on the 10 M-triangle terrain: hence the
// prediction); results are bit-identical (every path carries its own RNG stream and the same
// arithmetic runs in the same order per path).
template <bool L256, bool ST, bool QN = false>
__global__ void __launch_bounds__(EXT_THREADS)
finish_kernel(const DevScene sc, const PathState ps, const ParkBuf in, const uint32_t* __restrict__ n_ptr,
              uint32_t max_depth, unsigned long long* __restrict__ total_rays)
{
  const uint32_t n = *n_ptr;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t stride = gridDim.x * blockDim.x;
  uint32_t rays_local = 0, extra_traversed = 0;
  int stack[PT_STACK];
  for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += stride) {
    const float4 ro = in.ray[2 * (size_t)idx], rd = in.ray[2 * (size_t)idx + 1];
    const float4 th = in.thr[idx];
    const uint4 ax = in.aux[idx];
    const uint32_t pid = in.pid[idx];
    f3 o = xyz(ro), d = xyz(rd), color = xyz(th);
    float tmin = ro.w, tbest = __uint_as_float(ax.x);
    uint32_t rng = __float_as_uint(th.w), code = ax.y, depth = ax.w;
    int start = (int)ax.z;
    bool need_traversal = true, have_hit = false, first = true;
    Hit h;
    for (;;) {
      if (need_traversal) {
        Trav T;
        trav_init<QN>(sc, T, o, d, tmin, tbest, start, stack);
        while (T.node != PT_SENTINEL) {
          if (T.node >= 0)
            trav_inner<L256, QN>(sc, T, stack);
          else
            trav_leaf(sc, T, stack);
        }
        if (T.best >= 0) { // what traverse_kernel writes back into aux
          tbest = T.tbest;
          code = AUX_TRI | (uint32_t)T.best;
        }
        if (!first) ++extra_traversed; // (the first one is counted by the parked list's length)
        first = false;
        need_traversal = false;
        have_hit = false;
      }
      ++rays_local;
      const bool hit = have_hit ? code != 0u : resolve_hit<ST>(sc, o, d, tmin, tbest, code, h);
      if (depth == 0u) ps.gbuf[pid] = hit ? make_float4(h.n.x, h.n.y, h.n.z, h.t) : make_float4(-d.x, -d.y, -d.z, 1e6f);
      if (!hit) {
        color = color * sky_color(d);
        break;
      }
      const DevMaterial mat = sc.materials[h.material];
      scatter(mat, h, o, d, tmin, color, rng);
      if (++depth == max_depth) break;
      need_traversal = classify<ST>(sc, o, d, tmin, FLT_MAX, tbest, code, start, h);
      have_hit = true;
    }
    ps.thr[pid] = mk4(color, __uint_as_float(rng));
  }
  for (int off = 16; off > 0; off >>= 1) {
    rays_local += __shfl_down_sync(0xffffffffu, rays_local, off);
    extra_traversed += __shfl_down_sync(0xffffffffu, extra_traversed, off);
  }
  if (lane == 0 && rays_local != 0u) atomicAdd(total_rays, (unsigned long long)rays_local);
  if (lane == 0 && extra_traversed != 0u) atomicAdd(total_rays + 1, (unsigned long long)extra_traversed);
}

// ================================================================ launchers
static inline uint32_t cdiv(uint32_t a, uint32_t b) { return (a + b - 1) / b; }

// Experiment switches, read from the environment ONCE per process (not per launch).
struct Tunables {
  int refill, inner_min, stream_state, node_min, tri_min, chain_grid, chain_grid0;
  int trav_minb, trav_l256; // traverse_kernel instantiation (PT_TRAV="minb,l256")
  int order;                // bounce-0 item order (PT_ORDER)
  int chain_tma;            // TMA-staged parked state in the re-entry chain launches (PT_CHAIN_TMA)
  int finish_after;         // first wavefront bounce finish_kernel may replace (PT_FINISH, 0 = never)
  int finish_max_rays;      // ... once at most this many rays are predicted to be parked (PT_FINISH_RAYS)
};
static int env_int(const char* name, int dflt)
{
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}
static const Tunables& tunables()
{
  static const Tunables t = [] {
    Tunables t{};
    t.refill = env_int("PT_REFILL", 0); // 0 = choose by scene size (launch_t2v)
    t.inner_min = env_int("PT_INNER_MIN", EXT_INNER_MIN);
    t.stream_state = env_int("PT_STREAM_STATE", 0);
    t.node_min = env_int("PT_NODE_MIN", 12);
    t.tri_min = env_int("PT_TRI_MIN", 4);
    t.chain_grid = env_int("PT_CHAIN_GRID", 6);
    t.chain_grid0 = env_int("PT_CHAIN_GRID0", 24);
    t.trav_minb = 0; // 0 = choose by scene size (launch_t2)
    t.trav_l256 = 0;
    if (const char* v = getenv("PT_TRAV")) sscanf(v, "%d,%d", &t.trav_minb, &t.trav_l256);
    t.order = env_int("PT_ORDER", PT_DEFAULT_ORDER);
    t.chain_tma = env_int("PT_CHAIN_TMA", 0);
    t.finish_after = env_int("PT_FINISH", PT_DEFAULT_FINISH);
    t.finish_max_rays = env_int("PT_FINISH_RAYS", PT_DEFAULT_FINISH_RAYS);
    return t;
  }();
  return t;
}
int tunable_order() { return tunables().order; }
int tunable_finish_after() { return tunables().finish_after; }
int tunable_finish_rays() { return tunables().finish_max_rays; }

void launch_finish(const LaunchEnv& env, const DevScene& sc, const PassBuffers& pb, uint32_t iter, uint32_t max_depth)
{
  // consumes the list chain iteration `iter` parked (not yet traversed)
  const bool sphere_trees = sc.sph_root_before >= 0 || sc.sph_root_after >= 0;
  const bool l256 = (size_t)sc.n_nodes * 64 <= (2ull << 20);
  const uint32_t grid = (uint32_t)env.sms * 4u;
#define PT_FIN(L, S, Q)                                                                            \
  finish_kernel<L, S, Q><<<grid, EXT_THREADS, 0, env.stream>>>(sc, pb.ps, pb.park[iter & 1], pb.tcounters + iter, \
                                                                max_depth, pb.total_rays)
  // (the same node fetch as traverse_kernel, so that a path's walk — hence the winner of an exact
  // tie — does not depend on which of the two kernels finishes it)
  if (sphere_trees) {
    if (l256) PT_FIN(true, true, false); else PT_FIN(false, true, false);
  } else if (sc.qnodes != nullptr) {
    PT_FIN(true, false, true);
  } else {
    if (l256) PT_FIN(true, false, false); else PT_FIN(false, false, false);
  }
#undef PT_FIN
}
int tunable_stream_state() { return tunables().stream_state; }

void launch_raygen(const LaunchEnv& env, const DevScene& sc, const PassBuffers& pb,
                   const PassParams& pp, uint32_t n_items)
{
  const uint32_t grid = min((uint32_t)env.sms * 8u, cdiv(n_items, FULL_THREADS));
  raygen_kernel<<<grid, FULL_THREADS, 0, env.stream>>>(sc, pb.ps, pp, n_items, pb.tq,
                                                       pb.tcounters + 0);
}

// ---- traversal launch shape: CTA size, CTAs per SM and the shared-memory staging budget per CTA
// are run-time choices for the wide kernel (PT_T8 = "threads,blocks_per_sm,smem_kb"); every shape
// is a separate instantiation.  Grids are persistent: SMs x resident CTAs.
struct TShape {
  int threads, blocks, smem_kb;
};
static TShape shape_from_env(const char* name, TShape d)
{
  if (const char* v = getenv(name)) sscanf(v, "%d,%d,%d", &d.threads, &d.blocks, &d.smem_kb);
  return d;
}
static TShape t8_shape()
{
  static TShape sh = shape_from_env("PT_T8", TShape{128, 8, 0});
  return sh;
}

static const int kMaxDynSmem = 226 * 1024; // 227 KB per CTA minus the static barrier and alignment

// per device and kernel: opt in to the big dynamic shared-memory window, size the persistent grid
template <typename K>
static uint32_t persistent_grid(K kern, const LaunchEnv& env, int threads, size_t smem, int* cache_nb,
                                size_t* cache_smem)
{
  // (one lock for all kernels: contexts of a multi-GPU group launch from one host thread each)
  static std::mutex lock;
  std::lock_guard<std::mutex> guard(lock);
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (cache_nb[dev] == 0 || cache_smem[dev] != smem) {
    // (only kernels that stage into dynamic shared memory: with static shared memory on top, the
    // full window would exceed the per-CTA limit and the call would fail with "invalid argument")
    if (smem > 0) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem);
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, threads, smem) != cudaSuccess || nb <= 0) nb = 1;
    cache_nb[dev] = nb;
    cache_smem[dev] = smem;
  }
  return (uint32_t)(env.sms * cache_nb[dev]);
}

// Lanes still walking below which a warp goes back to refill.  Measured with two lanes: 12 / 16 /
// 20 / 24 give 10 229 / 10 230 / 10 120 / 9 905 Mrays/s on the bunny and 2 627 / 2 596 / - / 2 331
// on the 10 M-triangle terrain (walks there are long and uneven: refilling earlier keeps more of
// the misses in flight).
static int refill_for(const DevScene& sc)
{
  const int r = tunables().refill;
  if (r > 0) return r;
  const size_t scene_bytes = (size_t)sc.n_nodes * 64 + (size_t)sc.n_tris * 48;
  return scene_bytes > (512ull << 20) ? 12 : EXT_REFILL;
}

template <int SRC, int MINB, bool L256, bool QN = false>
static void launch_t2v(const LaunchEnv& env, const DevScene& sc, const PathState& ps, const uint32_t* tq,
                       const uint32_t* n_ptr, uint32_t n_host, uint32_t* work, const float4* rays,
                       HitRecord* out, uint32_t max_grid, const BinLists& bins)
{
  auto kern = traverse_kernel<SRC, MINB, L256, false, QN>;
  static int nb[64] = {0};
  static size_t sm[64] = {0};
  const Tunables& t = tunables();
  const uint32_t grid = min(persistent_grid(kern, env, EXT_THREADS, 0, nb, sm), max_grid);
  kern<<<grid, EXT_THREADS, 0, env.stream>>>(sc, ps, tq, n_ptr, n_host, work, rays, out, refill_for(sc),
                                             t.inner_min, t.stream_state, bins);
}

template <int SRC>
static void launch_t2(const LaunchEnv& env, const DevScene& sc, const PathState& ps, const uint32_t* tq,
                      const uint32_t* n_ptr, uint32_t n_host, uint32_t* work, const float4* rays,
                      HitRecord* out, uint32_t max_grid, const BinLists& bins)
{
  const Tunables& t = tunables();
  if (SRC == SRC_BATCH && (sc.sph_root_before >= 0 || sc.sph_root_after >= 0)) {
    // the parity hook of a scene with sphere trees: classification + hit rebuild walk them
    auto kern = traverse_kernel<SRC, EXT_MIN_BLOCKS, false, true>;
    static int nb[64] = {0};
    static size_t sm[64] = {0};
    const uint32_t grid = min(persistent_grid(kern, env, EXT_THREADS, 0, nb, sm), max_grid);
    kern<<<grid, EXT_THREADS, 0, env.stream>>>(sc, ps, tq, n_ptr, n_host, work, rays, out, refill_for(sc), t.inner_min,
                                               t.stream_state, bins);
    return;
  }
  int minb = t.trav_minb, l256 = t.trav_l256;
  if (minb == 0) {
    // The exact-node instantiations (trees the quantiser is not used for).  Measured
    // (profiles/README.md, round 2): a tree that lives in L1/L2 likes the 256-bit node fetch — fewer
    // instructions (bunny: traverse -3 %); a scene far beyond the
    // 126 MB L2 likes more resident warps to hide its misses: 10 CTAs per SM at 48 registers
    // (10 M-triangle terrain, two lanes: 2 476 -> 2 596 Mrays/s; 12 CTAs at 40 registers spill
    // into the same L1 data pipe and only matched it in one-lane passes); in between (2.6 M
    // triangles) neither moves the result by more than 0.5 %.
    // (ld.global.nc.L2::256B / ::128B prefetch-size hints on the triangle and node loads and
    // cudaLimitMaxL2FetchGranularity = 32 / 128 were measured on all three scenes: no effect
    // beyond 0.1 %, profiles/README.md; removed.)
    const size_t node_bytes = (size_t)sc.n_nodes * 64, scene_bytes = node_bytes + (size_t)sc.n_tris * 48;
    minb = scene_bytes > (512ull << 20) ? 10 : EXT_MIN_BLOCKS;
    l256 = node_bytes <= (2ull << 20) ? 1 : 0;
  }
  // Quantised nodes (scene_upload builds them for host-built trees that are not 'big').  Their
  // node data takes 8 registers instead of 16, so one more CTA fits per SM: 9 (56 registers) —
  // measured with the final node test: bunny 11 300 (9) / 11 168 (10) Mrays/s, 150 k triangles
  // 5 164 / 5 099, 2.6 M triangles 6 656 / 6 617 (profiles/r2_ab_traverse_sign_selected_planes.log).
  // A quantised node is one 32-byte sector: always ONE 256-bit load, whatever the tree's size
  // (2.6 M triangles: 6 647 -> 6 789 Mrays/s against two 128-bit loads).
  if (sc.qnodes != nullptr) {
    const int qb = t.trav_minb != 0 ? t.trav_minb : 9;
    if (qb >= 10) return launch_t2v<SRC, 10, true, true>(env, sc, ps, tq, n_ptr, n_host, work, rays, out, max_grid, bins);
    if (qb <= 8) return launch_t2v<SRC, 8, true, true>(env, sc, ps, tq, n_ptr, n_host, work, rays, out, max_grid, bins);
    return launch_t2v<SRC, 9, true, true>(env, sc, ps, tq, n_ptr, n_host, work, rays, out, max_grid, bins);
  }
#define PT_T2_CASE(B, L)                                                                           \
  if (minb == B && l256 == L)                                                                      \
    return launch_t2v<SRC, B, L != 0>(env, sc, ps, tq, n_ptr, n_host, work, rays, out, max_grid, bins);
  PT_T2_CASE(8, 0)
  PT_T2_CASE(8, 1)
  PT_T2_CASE(10, 0)
  PT_T2_CASE(10, 1)
  PT_T2_CASE(12, 0)
  PT_T2_CASE(12, 1)
#undef PT_T2_CASE
  launch_t2v<SRC, EXT_MIN_BLOCKS, false>(env, sc, ps, tq, n_ptr, n_host, work, rays, out, max_grid, bins);
}

template <int SRC, int THREADS, int BLOCKS>
static void launch_t8(const LaunchEnv& env, const DevScene& sc, const PathState& ps, const uint32_t* tq,
                      const uint32_t* n_ptr, uint32_t n_host, uint32_t* work, const float4* rays,
                      HitRecord* out, int smem_kb, uint32_t max_grid)
{
  auto kern = traverse8_kernel<SRC, THREADS, BLOCKS>;
  const uint32_t n_staged = min(sc.n_nodes8, (uint32_t)(min(smem_kb * 1024, kMaxDynSmem) / 80));
  const size_t smem = (size_t)n_staged * 80;
  static int nb[64] = {0};
  static size_t sm[64] = {0};
  const uint32_t grid = min(persistent_grid(kern, env, THREADS, smem, nb, sm), max_grid);
  kern<<<grid, THREADS, smem, env.stream>>>(sc, ps, tq, n_ptr, n_host, work, rays, out,
                                            refill_for(sc), tunables().node_min,
                                            tunables().tri_min, n_staged, 0x4B000000u);
}

template <int SRC>
static void launch_traverse_shape(const LaunchEnv& env, const DevScene& sc, const PathState& ps,
                                  const uint32_t* tq, const uint32_t* n_ptr, uint32_t n_host,
                                  uint32_t* work, const float4* rays, HitRecord* out, bool batch,
                                  const BinLists& bins = BinLists{nullptr, nullptr, 0u})
{
  if (sc.n_nodes8 == 0u) {
    launch_t2<SRC>(env, sc, ps, tq, n_ptr, n_host, work, rays, out,
                   batch ? cdiv(n_host, (uint32_t)EXT_THREADS) : 0xffffffffu, bins);
    return;
  }
  const TShape sh = t8_shape();
  const uint32_t max_grid = batch ? cdiv(n_host, (uint32_t)sh.threads) : 0xffffffffu;
#define PT_SHAPE_CASE(T, B)                                                                        \
  if (sh.threads == T && sh.blocks == B)                                                           \
    return launch_t8<SRC, T, B>(env, sc, ps, tq, n_ptr, n_host, work, rays, out, sh.smem_kb, max_grid);
  PT_SHAPE_CASE(128, 8)
  PT_SHAPE_CASE(256, 4)
  PT_SHAPE_CASE(256, 2)
  PT_SHAPE_CASE(512, 2)
  PT_SHAPE_CASE(512, 1)
  PT_SHAPE_CASE(1024, 1)
#undef PT_SHAPE_CASE
  launch_t8<SRC, 128, 8>(env, sc, ps, tq, n_ptr, n_host, work, rays, out, sh.smem_kb, max_grid);
}

void launch_traverse(const LaunchEnv& env, const DevScene& sc, const PassBuffers& pb,
                     const uint32_t* tq, uint32_t bounce)
{
  launch_traverse_shape<SRC_QUEUE>(env, sc, pb.ps, tq, pb.tcounters + bounce, 0u, pb.work + bounce,
                                   nullptr, nullptr, false);
}

// the bin lists traverse iteration `iter` fills and chain iteration `iter + 1` consumes
static BinLists bins_of(const PassBuffers& pb, uint32_t iter)
{
  if (!pb.bin_list) return BinLists{nullptr, nullptr, 0u};
  return BinLists{pb.bin_list, pb.bin_counts + (size_t)iter * PT_BINS, pb.capacity};
}

void launch_traverse_parked(const LaunchEnv& env, const DevScene& sc, const PassBuffers& pb, uint32_t iter)
{
  // the compacted state of iteration `iter` is its own work list
  PathState view{};
  view.ray = pb.park[iter & 1].ray;
  view.aux = pb.park[iter & 1].aux;
  launch_traverse_shape<SRC_QUEUE>(env, sc, view, nullptr, pb.tcounters + iter, 0u, pb.work + iter,
                                   nullptr, nullptr, false, bins_of(pb, iter));
}

void launch_chain(const LaunchEnv& env, const DevScene& sc, const PassBuffers& pb,
                  const PassParams& pp, uint32_t iter, uint32_t n_items_first, uint32_t max_depth)
{
  // CTAs per SM of the grid-stride loops (3 are resident at 77-80 registers).  Measured on the
  // bunny frame: the primary launch likes a fine grid (its per-item cost varies with how many
  // in-register bounces follow: 6 -> 4.03 ms, 24 -> 3.73 ms), the re-entry launches do not
  // (6 -> 5.09 ms, 24 -> 5.17 ms).
  const bool sphere_trees = sc.sph_root_before >= 0 || sc.sph_root_after >= 0;
  const uint32_t grid = (uint32_t)env.sms * (uint32_t)tunables().chain_grid;
  const uint32_t grid_first = (uint32_t)env.sms * (uint32_t)tunables().chain_grid0;
  if (iter == 0) {
    const uint32_t g0 = min(grid_first, cdiv(n_items_first, FULL_THREADS));
    if (sphere_trees)
      chain_kernel<true, 0, true><<<g0, FULL_THREADS, 0, env.stream>>>(sc, pb.ps, pp, ParkBuf{}, nullptr, n_items_first,
                                                                     pb.park[0], pb.tcounters + 0, max_depth,
                                                                     pb.total_rays, BinLists{nullptr, nullptr, 0u});
    else
      chain_kernel<true, 0, false><<<g0, FULL_THREADS, 0, env.stream>>>(sc, pb.ps, pp, ParkBuf{}, nullptr, n_items_first,
                                                                      pb.park[0], pb.tcounters + 0, max_depth,
                                                                      pb.total_rays, BinLists{nullptr, nullptr, 0u});
  } else {
    // consumes the traversed state of iteration iter-1, parks into the buffer of iteration iter
    const BinLists bins = bins_of(pb, iter - 1);
    const int tma = bins.list == nullptr && !sphere_trees ? tunables().chain_tma : 0;
    if (sphere_trees) {
      chain_kernel<false, 0, true><<<grid, FULL_THREADS, 0, env.stream>>>(
          sc, pb.ps, pp, pb.park[(iter - 1) & 1], pb.tcounters + (iter - 1), 0u, pb.park[iter & 1],
          pb.tcounters + iter, max_depth, pb.total_rays, bins);
    } else if (tma == 1) {
      // (34 KB of dynamic shared memory: below the 48 KB that needs no opt-in)
      chain_kernel<false, 1, false><<<grid, FULL_THREADS, 2 * CHAIN_STAGE_BYTES, env.stream>>>(
          sc, pb.ps, pp, pb.park[(iter - 1) & 1], pb.tcounters + (iter - 1), 0u, pb.park[iter & 1],
          pb.tcounters + iter, max_depth, pb.total_rays, bins);
    } else if (tma == 2) {
      chain_kernel<false, 2, false><<<grid, FULL_THREADS, 2 * CHAIN_STAGE_BYTES, env.stream>>>(
          sc, pb.ps, pp, pb.park[(iter - 1) & 1], pb.tcounters + (iter - 1), 0u, pb.park[iter & 1],
          pb.tcounters + iter, max_depth, pb.total_rays, bins);
    } else {
      chain_kernel<false, 0, false><<<grid, FULL_THREADS, 0, env.stream>>>(
          sc, pb.ps, pp, pb.park[(iter - 1) & 1], pb.tcounters + (iter - 1), 0u, pb.park[iter & 1],
          pb.tcounters + iter, max_depth, pb.total_rays, bins);
    }
  }
}

void launch_shade(const LaunchEnv& env, const DevScene& sc, const PassBuffers& pb,
                  const PassParams& pp, int q, uint32_t bounce, uint32_t n_items_first,
                  bool last_bounce)
{
  const uint32_t grid = (uint32_t)env.sms * 8u;
  if (bounce == 0) {
    shade_kernel<true><<<min(grid, cdiv(n_items_first, FULL_THREADS)), FULL_THREADS, 0, env.stream>>>(
        sc, pb.ps, pp, nullptr, nullptr, n_items_first, pb.queue[q ^ 1], pb.counters + 1, pb.tq,
        pb.tcounters + 1, pb.flags, 0u, last_bounce ? 1u : 0u);
  } else {
    shade_kernel<false><<<grid, FULL_THREADS, 0, env.stream>>>(
        sc, pb.ps, pp, pb.queue[q], pb.counters + bounce, 0u, pb.queue[q ^ 1],
        pb.counters + bounce + 1, pb.tq, pb.tcounters + bounce + 1, pb.flags, bounce,
        last_bounce ? 1u : 0u);
  }
}

void launch_trace_batch(const LaunchEnv& env, const DevScene& sc, const float4* rays, uint32_t* work,
                        uint32_t n, HitRecord* out)
{
  // the parity hook runs the SAME classification + persistent traversal code as the renderer
  launch_traverse_shape<SRC_BATCH>(env, sc, PathState{}, nullptr, nullptr, n, work, rays, out, true);
}

} // namespace pt
