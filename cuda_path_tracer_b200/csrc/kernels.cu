// kernels.cu — hand-written sm_100a kernels of the wavefront path tracer.
//
//   extend      closest hit:  ray_scene_intersection_test   (reference path_tracer.cu:36-128,
//               intersections.cuh:7-103) re-designed as a culled, ordered while-while
//               traversal of a 64-byte two-box BVH node over world-space baked triangles,
//               persistent warps fetching 32-ray batches from a device-side queue.
//   shade       evaluate_material + sky + G-buffer           (path_tracer.cu:138-201, 29-34,
//               292-315) with warp-aggregated compaction of 4-byte path ids.
//   accumulate  final_gather                                 (path_tracer.cu:203-219, 317-330)
//   resolve     preview kernels / linear_to_gamma            (path_tracer.cu:221-225, 334-385)
//   atrous      denoising_kernel                             (denoising/...denoiser.cu:24-86)
#include "kernels.h"

#include <float.h>

namespace pt {

// =================================================================== helpers
PT_D f3 xyz(const float4& v) { return mk3(v.x, v.y, v.z); }
PT_D float4 mk4(f3 v, float w) { return make_float4(v.x, v.y, v.z, w); }
PT_D float4 ldg4(const float4* p) { return __ldg(p); }

struct Hit {
  float t;
  f3 p;
  f3 n;
  uint32_t material;
  uint32_t side;
  int32_t object;
  int32_t prim;
};

// ------------------------------------------------------------- sphere test
// ray_object_intersection_test, sphere branch (path_tracer.cu:87-98) with
// inverse_transform_ray (transform.hpp:50-58) and ray_sphere_intersection_test
// (intersections.cuh:7-41).  The quirks are kept: the object-space direction is
// re-normalised while t_min/t_max stay world-space; the reported t is the
// world-space distance; the normal is transformed by the inverse transpose and
// not re-normalised.
PT_D bool sphere_test(const DevSphere* __restrict__ sp, f3 o, f3 d, float tmin, float tmax,
                      Hit& h)
{
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

// ---------------------------------------------------------------- traversal
// Closest hit over spheres + the world-space BVH, result-equivalent to the reference's
// un-culled traversal (a culled subtree cannot hold a nearer accepted hit); exact-tie
// winners may differ (reference: last tested wins).  The traversal is written as a
// resumable per-lane state machine (init / step / finish) so that a persistent warp can
// retire finished lanes and refill them with new rays while the others keep walking.
struct Trav {
  f3 o, d;
  float tmin, tbest;
  float idx, idy, idz, odx, ody, odz;
  int node; // current node, PT_SENTINEL when done
  int sp;
  int best;   // best triangle slot or -1
  bool hit;   // a sphere tested before the mesh was hit (record in `h`)
  Hit h;
};

PT_D void trav_init(const DevScene& sc, Trav& T, f3 o, f3 d, float tmin, float tmax, int* stack)
{
  T.o = o;
  T.d = d;
  T.tmin = tmin;
  T.tbest = tmax;
  T.hit = false;
  T.best = -1;
  for (uint32_t i = 0; i < sc.n_spheres_before; ++i) {
    if (sphere_test(sc.spheres + i, o, d, tmin, T.tbest, T.h)) {
      T.hit = true;
      T.tbest = T.h.t;
    }
  }
  const float ooeps = 8.271806125530277e-25f; // 2^-80
  T.idx = 1.0f / (fabsf(d.x) > ooeps ? d.x : copysignf(ooeps, d.x));
  T.idy = 1.0f / (fabsf(d.y) > ooeps ? d.y : copysignf(ooeps, d.y));
  T.idz = 1.0f / (fabsf(d.z) > ooeps ? d.z : copysignf(ooeps, d.z));
  T.odx = o.x * T.idx;
  T.ody = o.y * T.idy;
  T.odz = o.z * T.idz;
  stack[0] = PT_SENTINEL;
  T.sp = 1;
  T.node = sc.n_tris != 0 ? 0 : PT_SENTINEL;
}

// One node visit: an inner node (two slab tests, ordered descent, far child pushed) or a
// leaf (<= 4 Moller-Trumbore tests in the reference's operation order, intersections.cuh:49-85).
PT_D void trav_step(const DevScene& sc, Trav& T, int* stack)
{
  if (T.node >= 0) {
    const float4* np = sc.nodes + (size_t)T.node * 4;
    const float4 n0 = ldg4(np + 0);
    const float4 n1 = ldg4(np + 1);
    const float4 n2 = ldg4(np + 2);
    const float4 n3 = ldg4(np + 3);
    const float c0lox = n0.x * T.idx - T.odx, c0hix = n0.y * T.idx - T.odx;
    const float c0loy = n0.z * T.idy - T.ody, c0hiy = n0.w * T.idy - T.ody;
    const float c0loz = n2.x * T.idz - T.odz, c0hiz = n2.y * T.idz - T.odz;
    const float c1lox = n1.x * T.idx - T.odx, c1hix = n1.y * T.idx - T.odx;
    const float c1loy = n1.z * T.idy - T.ody, c1hiy = n1.w * T.idy - T.ody;
    const float c1loz = n2.z * T.idz - T.odz, c1hiz = n2.w * T.idz - T.odz;
    const float c0min =
        fmaxf(fmaxf(fminf(c0lox, c0hix), fminf(c0loy, c0hiy)), fmaxf(fminf(c0loz, c0hiz), T.tmin));
    const float c0max =
        fminf(fminf(fmaxf(c0lox, c0hix), fmaxf(c0loy, c0hiy)), fminf(fmaxf(c0loz, c0hiz), T.tbest));
    const float c1min =
        fmaxf(fmaxf(fminf(c1lox, c1hix), fminf(c1loy, c1hiy)), fmaxf(fminf(c1loz, c1hiz), T.tmin));
    const float c1max =
        fminf(fminf(fmaxf(c1lox, c1hix), fmaxf(c1loy, c1hiy)), fminf(fmaxf(c1loz, c1hiz), T.tbest));
    // robust slab comparison (Ize 2013): widen the far side by 2 ulp
    const bool trav0 = c0max * 1.0000004f >= c0min;
    const bool trav1 = c1max * 1.0000004f >= c1min;
    const int c0 = __float_as_int(n3.x);
    const int c1 = __float_as_int(n3.y);
    if (!trav0 && !trav1) {
      T.node = stack[--T.sp];
    } else {
      T.node = trav0 ? c0 : c1;
      if (trav0 && trav1) {
        int other = c1;
        if (c1min < c0min) {
          other = c0;
          T.node = c1;
        }
        stack[T.sp++] = other;
      }
    }
  } else {
    const uint32_t code = (uint32_t)(~T.node);
    const uint32_t first = code >> 3;
    const uint32_t count = (code & 7u) + 1u;
    for (uint32_t k = 0; k < count; ++k) {
      const float4* tp = sc.tris + (size_t)(first + k) * 3;
      const float4 t0 = ldg4(tp + 0);
      const float4 t1 = ldg4(tp + 1);
      const float4 t2 = ldg4(tp + 2);
      const f3 e1 = xyz(t1), e2 = xyz(t2);
      const f3 hh = cross3(T.d, e2);
      const float a = dot3(e1, hh);
      if (a > -0.0000001f && a < 0.0000001f) continue;
      const float f = 1.0f / a;
      const f3 s = T.o - xyz(t0);
      const float u = f * dot3(s, hh);
      if (u < 0.0f || u > 1.0f) continue;
      const f3 q = cross3(s, e1);
      const float v = f * dot3(T.d, q);
      if (v < 0.0f || u + v > 1.0f) continue;
      const float t = f * dot3(e2, q);
      if (!(t >= T.tmin && t <= T.tbest)) continue;
      T.tbest = t;
      T.best = (int)(first + k);
    }
    T.node = stack[--T.sp];
  }
}

// Builds the Intersection of the winning primitive (triangle_normal + face side,
// intersections.cuh:43-85) and tests the spheres that follow the first mesh object.
PT_D bool trav_finish(const DevScene& sc, Trav& T)
{
  if (T.best >= 0) {
    const float4* tp = sc.tris + (size_t)T.best * 3;
    const float4 t0 = ldg4(tp + 0);
    const float4 t1 = ldg4(tp + 1);
    const float4 t2 = ldg4(tp + 2);
    const f3 outward = normalize3(cross3(xyz(t1), xyz(t2)));
    const bool front = dot3(T.d, outward) < 0.0f;
    T.h.t = T.tbest;
    T.h.p = T.o + T.d * T.tbest; // ray(t)
    T.h.n = front ? outward : -outward;
    T.h.side = front ? 0u : 1u;
    T.h.prim = __float_as_int(t0.w);
    T.h.object = __float_as_int(t1.w);
    T.h.material = (uint32_t)__float_as_int(t2.w);
    T.hit = true;
  }
  for (uint32_t i = sc.n_spheres_before; i < sc.n_spheres; ++i) {
    if (sphere_test(sc.spheres + i, T.o, T.d, T.tmin, T.tbest, T.h)) {
      T.hit = true;
      T.tbest = T.h.t;
    }
  }
  return T.hit;
}

// ------------------------------------------------------ bounce-0 index map
// Work item -> (sample, pixel): one warp covers an 8x4 pixel tile so primary
// rays of a warp are coherent.  Returns false for padding lanes.
PT_D bool first_item(const PassParams& pp, uint32_t idx, uint32_t& pid, uint32_t& x, uint32_t& y,
                     uint32_t& s)
{
  const uint32_t per_sample = pp.tiles_x * pp.tiles_y * 32u;
  s = idx / per_sample;
  const uint32_t r = idx - s * per_sample;
  const uint32_t tile = r >> 5, lane = r & 31u;
  const uint32_t ty = tile / pp.tiles_x, tx = tile - ty * pp.tiles_x;
  x = tx * 8u + (lane & 7u);
  y = ty * 4u + (lane >> 3);
  pid = s * pp.pixels + y * pp.cam.width + x;
  return x < pp.cam.width && y < pp.cam.height && s < pp.samples;
}

// =================================================================== extend
// Persistent warps (grid = SMs x resident CTAs).  Each lane owns one ray at a time.  Work is
// fetched from a device-side cursor with ONE atomic per refill for all idle lanes of the warp;
// lanes that finish early are retired and refilled as soon as fewer than EXT_REFILL lanes are
// still walking, so the warp stays populated although per-ray traversal lengths vary by orders
// of magnitude (sky rays leave after the root; silhouette rays visit dozens of nodes).
#define EXT_THREADS 128
#define EXT_MIN_BLOCKS 8
#define EXT_REFILL 22

enum { SRC_FIRST = 0, SRC_QUEUE = 1, SRC_BATCH = 2 };

template <int SRC>
__global__ void __launch_bounds__(EXT_THREADS, EXT_MIN_BLOCKS)
extend_kernel(const DevScene sc, const PathState ps, const PassParams pp,
              const uint32_t* __restrict__ queue, const uint32_t* __restrict__ n_ptr,
              uint32_t n_host, uint32_t* __restrict__ work, const float4* __restrict__ batch_rays,
              HitRecord* __restrict__ batch_out)
{
  const uint32_t n = SRC == SRC_QUEUE ? *n_ptr : n_host;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t lt_mask = (1u << lane) - 1u;
  int stack[PT_STACK];
  Trav T;
  T.node = PT_SENTINEL;
  bool has = false;
  bool exhausted = false; // warp-uniform
  uint32_t pid = 0;

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
          f3 o, d;
          float tmin, tmax;
          bool valid = true;
          if (SRC == SRC_FIRST) {
            uint32_t x, y, s;
            valid = first_item(pp, idx, pid, x, y, s);
            if (valid) {
              // raygen_kernel (ray_gen.cu:11-32)
              const uint32_t pixel = y * pp.cam.width + x;
              uint32_t rng = minstd_seed(wang_hash(wang_hash(pixel) ^ (pp.first_iteration + s)));
              const float fx = (float)x + minstd_uniform(rng);
              const float fy = (float)y + minstd_uniform(rng);
              camera_ray(pp.cam, fx, fy, o, d);
              tmin = 1e-4f;
              tmax = FLT_MAX;
              ps.ray_o[pid] = mk4(o, tmin);
              ps.ray_d[pid] = mk4(d, tmax);
              ps.thr[pid] = make_float4(1.0f, 1.0f, 1.0f, __uint_as_float(rng));
            }
          } else if (SRC == SRC_QUEUE) {
            pid = queue[idx];
            const float4 ro = ps.ray_o[pid];
            const float4 rd = ps.ray_d[pid];
            o = xyz(ro), d = xyz(rd);
            tmin = ro.w, tmax = rd.w;
          } else {
            pid = idx;
            const float4 ro = batch_rays[2 * (size_t)idx], rd = batch_rays[2 * (size_t)idx + 1];
            o = xyz(ro), d = xyz(rd);
            tmin = ro.w, tmax = rd.w;
          }
          if (valid) {
            trav_init(sc, T, o, d, tmin, tmax, stack);
            has = true;
          }
        }
      }
    }
    if (__ballot_sync(0xffffffffu, has) == 0u) {
      if (exhausted) break;
      continue;
    }
    // ---- walk until too few lanes are still busy (all of them once the queue is drained)
    const int threshold = exhausted ? 1 : EXT_REFILL;
    for (;;) {
      const bool busy = has && T.node != PT_SENTINEL;
      if (__popc(__ballot_sync(0xffffffffu, busy)) < threshold) break;
      if (busy) trav_step(sc, T, stack);
    }
    // ---- retire finished lanes
    if (has && T.node == PT_SENTINEL) {
      const bool hit = trav_finish(sc, T);
      if (SRC == SRC_BATCH) {
        HitRecord r;
        if (hit) {
          r.t = T.h.t;
          r.px = T.h.p.x, r.py = T.h.p.y, r.pz = T.h.p.z;
          r.nx = T.h.n.x, r.ny = T.h.n.y, r.nz = T.h.n.z;
          r.material = T.h.material, r.side = T.h.side;
          r.object = T.h.object, r.prim = T.h.prim;
        } else {
          r.t = -1.0f;
          r.px = r.py = r.pz = r.nx = r.ny = r.nz = 0.f;
          r.material = 0, r.side = 0, r.object = -1, r.prim = -1;
        }
        r.pad = 0;
        batch_out[pid] = r;
      } else if (hit) {
        ps.hit_a[pid] = make_float4(T.h.t, T.h.p.x, T.h.p.y, T.h.p.z);
        ps.hit_b[pid] = mk4(T.h.n, __uint_as_float(T.h.material | (T.h.side << 31)));
      } else {
        ps.hit_a[pid] = make_float4(-1.0f, 0.f, 0.f, 0.f);
      }
      has = false;
    }
  }
}


// ==================================================================== shade
#define SHD_THREADS 256

// random_in_unit_sphere (distributions.cuh:6-19): uniform ON the sphere.
PT_D f3 random_on_sphere(uint32_t& rng)
{
  const float phi = (2.0f * 3.14159265358979323846264338327950288f) * minstd_uniform(rng);
  const float cos_theta = 2.0f * minstd_uniform(rng) - 1.0f;
  const float sin_theta = sqrtf(1.0f - cos_theta * cos_theta);
  return mk3(cosf(phi) * sin_theta, sinf(phi) * sin_theta, cos_theta);
}

PT_D float sign1(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }

template <bool FIRST>
__global__ void __launch_bounds__(SHD_THREADS)
shade_kernel(const DevScene sc, const PathState ps, const PassParams pp,
             const uint32_t* __restrict__ queue, const uint32_t* __restrict__ n_ptr,
             uint32_t n_first, uint32_t* __restrict__ next_queue,
             uint32_t* __restrict__ next_count, uint8_t* __restrict__ flags, uint32_t bounce,
             uint32_t last_bounce)
{
  const uint32_t n = FIRST ? n_first : *n_ptr;
  const uint32_t n_round = (n + 31u) & ~31u;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n_round; idx += stride) {
    bool alive = false;
    uint32_t pid = 0;
    bool valid = idx < n;
    if (valid) {
      if (FIRST) {
        uint32_t x, y, s;
        valid = first_item(pp, idx, pid, x, y, s);
      } else {
        pid = queue[idx];
      }
    }
    if (valid) {
      const float4 ha = ps.hit_a[pid];
      float4 th = ps.thr[pid];
      const float4 rd = ps.ray_d[pid];
      const f3 d = xyz(rd);
      f3 color = xyz(th);
      uint32_t rng = __float_as_uint(th.w);
      if (pp.rng_mode == 1u) {
        // reference streaming mode: re-seed from the compacted slot index and
        // discard(bounce) (path_tracer.cu:300-301)
        const uint32_t slot = FIRST ? pid : idx;
        rng = minstd_seed(wang_hash(wang_hash(slot) ^ pp.first_iteration));
        rng = minstd_discard(rng, bounce);
      }
      if (ha.x < 0.0f) {
        // miss: get_background_color (path_tracer.cu:29-34)
        const f3 unit = normalize3(d);
        const float t = 0.5f * (unit.y + 1.0f);
        const f3 sky = mk3(0.5f, 0.7f, 1.0f) * (1.0f - t) + mk3(1.0f, 1.0f, 1.0f) * t;
        color = color * sky;
        ps.thr[pid] = mk4(color, __uint_as_float(rng));
        if (FIRST) ps.gbuf[pid] = make_float4(-d.x, -d.y, -d.z, 1e6f);
      } else {
        const float4 hb = ps.hit_b[pid];
        const f3 n = xyz(hb);
        const uint32_t meta = __float_as_uint(hb.w);
        const DevMaterial mat = sc.materials[meta & 0x7fffffffu];
        const f3 p = mk3(ha.y, ha.z, ha.w);
        if (FIRST) ps.gbuf[pid] = make_float4(n.x, n.y, n.z, ha.x);

        // evaluate_material (path_tracer.cu:138-201)
        f3 origin = p - (1e-4f * sign1(dot3(d, n))) * n;
        f3 dir;
        float tmin = ps.ray_o[pid].w;
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
          const float ratio = (meta >> 31) == 0u ? (1.0f / ior) : ior;
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
          origin = p;
          tmin = 1e-5f;
        }
        ps.ray_o[pid] = mk4(origin, tmin);
        ps.ray_d[pid] = mk4(dir, mat.type == 2 ? FLT_MAX : rd.w);
        ps.thr[pid] = mk4(color, __uint_as_float(rng));
        alive = last_bounce == 0u;
      }
    }
    if (pp.rng_mode == 1u) {
      // stable compaction happens in a separate pass keyed on the slot index
      if (idx < n && !last_bounce) {
        if (FIRST) {
          if (valid) flags[pid] = alive ? 1 : 0;
        } else {
          flags[idx] = alive ? 1 : 0;
        }
      }
    } else {
      // warp-aggregated compaction: one atomic per warp, ids stay warp-ordered
      const uint32_t mask = __ballot_sync(0xffffffffu, alive);
      if (mask) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(next_count, (uint32_t)__popc(mask));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (alive) next_queue[base + __popc(mask & ((1u << lane) - 1u))] = pid;
      }
    }
  }
}

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
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p == 0) {
    // rays = live paths entering extend, summed over bounces (counters[0] is host-known)
    unsigned long long r = (unsigned long long)pp.pixels * pp.samples;
    for (uint32_t b = 1; b < max_depth; ++b) r += counters[b];
    *total_rays += r;
  }
  if (p >= pp.pixels) return;
  float4 c = sum_color[p];
  float4 g = sum_gbuf[p];
  for (uint32_t s = 0; s < pp.samples; ++s) {
    const float4 t = ps.thr[(size_t)s * pp.pixels + p];
    const float4 gb = ps.gbuf[(size_t)s * pp.pixels + p];
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
                       float4* __restrict__ normal_depth, float4* __restrict__ position)
{
  const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t y = blockIdx.y;
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

// One a-trous iteration, reference arithmetic (denoiser.cu:24-86).  The tap
// weight is kernel[min(|dx|,|dy|)] with kernel = {3/8, 1/4, 1/16}; the three
// edge-stopping weights use exp() clamped to 1.  Taps are clamped to [0,W]x[0,H]
// inclusive like the reference: u == W aliases pixel (0, v+1) while its
// position is still rebuilt from the ray through (W+0.5, v+0.5); reads that
// would fall past the end of the buffer (undefined in the reference) use the
// last row / last pixel instead.  clamp_fix selects the sane W-1/H-1 clamp.
__global__ void __launch_bounds__(256)
atrous_kernel(const DevCamera cam, const DenoiseParams dp, const float4* __restrict__ color_in,
              const float4* __restrict__ normal_depth, const float4* __restrict__ position,
              float4* __restrict__ color_out, int step)
{
  const int x = blockIdx.x * 32 + (threadIdx.x & 31);
  const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
  const int W = (int)cam.width, H = (int)cam.height;
  if (x >= W || y >= H) return;
  const int p = y * W + x;
  const float4 cv = ldg4(color_in + p);
  const float4 nv = ldg4(normal_depth + p);
  const float4 pv = ldg4(position + p);
  const float kern[3] = {3.f / 8.f, 1.f / 4.f, 1.f / 16.f};
  const float step2 = (float)(step * step);
  const int umax = dp.clamp_fix ? W - 1 : W;
  const int vmax = dp.clamp_fix ? H - 1 : H;
  f3 sum = mk3(0.f, 0.f, 0.f);
  float cum_w = 0.f;
#pragma unroll
  for (int dy = -2; dy <= 2; ++dy) {
#pragma unroll
    for (int dx = -2; dx <= 2; ++dx) {
      const int u = min(max(x + dx * step, 0), umax);
      const int v = min(max(y + dy * step, 0), vmax);
      int q = u + v * W;
      f3 ptmp;
      float4 ct, nt;
      if (u < W && v < H) {
        ct = ldg4(color_in + q);
        nt = ldg4(normal_depth + q);
        const float4 pt4 = ldg4(position + q);
        ptmp = mk3(pt4.x, pt4.y, pt4.z);
      } else {
        if (q >= W * H) q = min(u, W - 1) + (H - 1) * W;
        ct = ldg4(color_in + q);
        nt = ldg4(normal_depth + q);
        f3 o, d;
        camera_ray(cam, (float)u + 0.5f, (float)v + 0.5f, o, d);
        ptmp = o + d * nt.w;
      }
      f3 t = mk3(cv.x - ct.x, cv.y - ct.y, cv.z - ct.z);
      float dist2 = dot3(t, t);
      const float c_w = fminf(expf(-dist2 / dp.c_phi), 1.0f);
      t = mk3(nv.x - nt.x, nv.y - nt.y, nv.z - nt.z);
      dist2 = fmaxf(dot3(t, t) / step2, 0.0f);
      const float n_w = fminf(expf(-dist2 / dp.n_phi), 1.0f);
      t = mk3(pv.x - ptmp.x, pv.y - ptmp.y, pv.z - ptmp.z);
      dist2 = dot3(t, t);
      const float p_w = fminf(expf(-dist2 / dp.p_phi), 1.0f);
      const float weight = c_w * n_w * p_w;
      const int ki = min(abs(dx), abs(dy));
      sum = sum + mk3(ct.x, ct.y, ct.z) * weight * kern[ki];
      cum_w += weight * kern[ki];
    }
  }
  color_out[p] = make_float4(sum.x / cum_w, sum.y / cum_w, sum.z / cum_w, 0.f);
}

// ================================================================ launchers
static inline uint32_t cdiv(uint32_t a, uint32_t b) { return (a + b - 1) / b; }

// persistent grid: SMs x the number of CTAs the kernel can keep resident per SM
template <int SRC> static uint32_t extend_grid(const LaunchEnv& env)
{
  static int per_sm = 0;
  if (per_sm == 0) {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, extend_kernel<SRC>, EXT_THREADS, 0) !=
            cudaSuccess ||
        nb <= 0)
      nb = EXT_MIN_BLOCKS;
    per_sm = nb;
  }
  return (uint32_t)(env.sms * per_sm);
}

void launch_extend_first(const LaunchEnv& env, const DevScene& sc, const PassBuffers& pb,
                         const PassParams& pp, uint32_t n_items)
{
  const uint32_t grid = min(extend_grid<SRC_FIRST>(env), cdiv(n_items, EXT_THREADS));
  extend_kernel<SRC_FIRST><<<grid, EXT_THREADS, 0, env.stream>>>(
      sc, pb.ps, pp, nullptr, nullptr, n_items, pb.work + 0, nullptr, nullptr);
}

void launch_extend(const LaunchEnv& env, const DevScene& sc, const PassBuffers& pb,
                   const PassParams& pp, int q, uint32_t bounce)
{
  const uint32_t grid = extend_grid<SRC_QUEUE>(env);
  extend_kernel<SRC_QUEUE><<<grid, EXT_THREADS, 0, env.stream>>>(
      sc, pb.ps, pp, pb.queue[q], pb.counters + bounce, 0u, pb.work + bounce, nullptr, nullptr);
}

void launch_shade(const LaunchEnv& env, const DevScene& sc, const PassBuffers& pb,
                  const PassParams& pp, int q, uint32_t bounce, uint32_t n_items_first,
                  bool last_bounce)
{
  const uint32_t grid = env.sms * 4;
  if (bounce == 0) {
    shade_kernel<true><<<grid, SHD_THREADS, 0, env.stream>>>(
        sc, pb.ps, pp, nullptr, nullptr, n_items_first, pb.queue[q ^ 1], pb.counters + 1,
        pb.flags, 0u, last_bounce ? 1u : 0u);
  } else {
    shade_kernel<false><<<grid, SHD_THREADS, 0, env.stream>>>(
        sc, pb.ps, pp, pb.queue[q], pb.counters + bounce, 0u, pb.queue[q ^ 1],
        pb.counters + bounce + 1, pb.flags, bounce, last_bounce ? 1u : 0u);
  }
}

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
  accumulate_kernel<<<cdiv(pp.pixels, 256), 256, 0, env.stream>>>(
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
                            float4* position)
{
  dim3 grid(cdiv(cam.width, 256), cam.height);
  denoise_prepare_kernel<<<grid, 256, 0, env.stream>>>(cam, sum_color, sum_gbuf, color0,
                                                       normal_depth, position);
}

void launch_atrous(const LaunchEnv& env, const DevCamera& cam, const DenoiseParams& dp,
                   const float4* color_in, const float4* normal_depth, const float4* position,
                   float4* color_out, int step_width)
{
  dim3 grid(cdiv(cam.width, 32), cdiv(cam.height, 8));
  atrous_kernel<<<grid, 256, 0, env.stream>>>(cam, dp, color_in, normal_depth, position,
                                              color_out, step_width);
}

void launch_trace_batch(const LaunchEnv& env, const DevScene& sc, const float4* rays, uint32_t* work,
                        uint32_t n, HitRecord* out)
{
  // the parity hook runs the SAME persistent traversal kernel as the renderer
  const uint32_t grid = min(extend_grid<SRC_BATCH>(env), cdiv(n, EXT_THREADS));
  extend_kernel<SRC_BATCH><<<grid, EXT_THREADS, 0, env.stream>>>(
      sc, PathState{}, PassParams{}, nullptr, nullptr, n, work, rays, out);
}

} // namespace pt
