/*
 * oracle.c — CPU restatement of LesleyLai/cuda-path-tracer's hot path.
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  Plain C11 + OpenMP; float arithmetic
 * written in the reference's operation order (glm 0.9.9.8 semantics restated by
 * hand — glm is an un-vendored Conan dependency, conanfile.txt:6).
 *
 * Build: make -C oracle   ->  oracle/liboracle.so
 */
#include "oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ vec3 */
typedef struct { float x, y, z; } v3;
static inline v3 V(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 vadd(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 vsub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 vneg(v3 a) { return V(-a.x, -a.y, -a.z); }
static inline v3 vmul(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 vscale(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }
static inline v3 vdivs(v3 a, float s) { return V(a.x / s, a.y / s, a.z / s); }
static inline v3 vdiv(v3 a, v3 b) { return V(a.x / b.x, a.y / b.y, a.z / b.z); }
/* glm::dot(vec3): tmp = a*b; tmp.x + tmp.y + tmp.z */
static inline float vdot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline v3 vcross(v3 a, v3 b)
{
  return V(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
/* glm::normalize: v * inversesqrt(dot(v, v)) */
static inline v3 vnormalize(v3 a) { return vscale(a, 1.0f / sqrtf(vdot(a, a))); }
static inline float vlength(v3 a) { return sqrtf(vdot(a, a)); }
/* glm::min(x,y) = (y < x) ? y : x ;  glm::max(x,y) = (x < y) ? y : x  (NaN behaviour kept) */
static inline float gmin(float x, float y) { return (y < x) ? y : x; }
static inline float gmax(float x, float y) { return (x < y) ? y : x; }
static inline v3 vmin(v3 a, v3 b) { return V(gmin(a.x, b.x), gmin(a.y, b.y), gmin(a.z, b.z)); }
static inline v3 vmax(v3 a, v3 b) { return V(gmax(a.x, b.x), gmax(a.y, b.y), gmax(a.z, b.z)); }
static inline float comp(v3 a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

/* glm mat4 (column-major) * vec4: (m0*v0 + m1*v1) + (m2*v2 + m3*v3) */
static inline void m4v4(const float* m, float x, float y, float z, float w, float out[4])
{
  for (int r = 0; r < 4; ++r)
    out[r] = (m[0 + r] * x + m[4 + r] * y) + (m[8 + r] * z + m[12 + r] * w);
}

/* ------------------------------------------------------------ hash + RNG */
/* src/lib/hash.cuh:4-14 */
uint32_t orc_hash(uint32_t a)
{
  a = (a + 0x7ed55d16u) + (a << 12);
  a = (a ^ 0xc761c23cu) ^ (a >> 19);
  a = (a + 0x165667b1u) + (a << 5);
  a = (a + 0xd3a2646cu) ^ (a << 9);
  a = (a + 0xfd7046c5u) + (a << 3);
  a = (a ^ 0xb55a4f09u) ^ (a >> 16);
  return a;
}

/* thrust::minstd_rand = linear_congruential_engine<uint32, 48271, 0, 2147483647>
 * (thrust/random/linear_congruential_engine.h; seed(): detail/linear_congruential_engine.inl:45-55) */
uint32_t orc_rng_seed(uint32_t s)
{
  uint32_t x = s % 2147483647u;
  return x == 0u ? 1u : x;
}
static inline uint32_t rng_next(uint32_t* st)
{
  *st = (uint32_t)(((uint64_t)(*st) * 48271ull) % 2147483647ull);
  return *st;
}
/* thrust::uniform_real_distribution<float>(0,1) (detail/uniform_real_distribution.inl:62-79):
 * float(urng() - min) / (1.f + float(max - min)), min = 1, max = 2147483646 */
float orc_rng_uniform(uint32_t* st)
{
  float result = (float)(rng_next(st) - 1u);
  result /= (1.0f + (float)(2147483646u - 1u));
  return (result * (1.0f - 0.0f)) + 0.0f;
}
void orc_rng_discard(uint32_t* st, uint64_t n)
{
  for (uint64_t i = 0; i < n; ++i) rng_next(st);
}

/* --------------------------------------------------------------- AABB */
typedef struct { v3 min, max; } aabb_t;
/* src/lib/aabb.hpp:15-17 default: min = FLT_MAX, max = -FLT_MAX */
static inline aabb_t aabb_empty(void)
{
  aabb_t b = {{FLT_MAX, FLT_MAX, FLT_MAX}, {-FLT_MAX, -FLT_MAX, -FLT_MAX}};
  return b;
}
static inline int aabb_is_empty(aabb_t b) /* aabb.hpp:24-29 */
{
  return b.min.x > b.max.x || b.min.y > b.max.y || b.min.z > b.max.z;
}
static inline aabb_t aabb_enclose_pt(aabb_t b, v3 p) /* aabb.hpp:31-34 */
{
  aabb_t r = {vmin(b.min, p), vmax(b.max, p)};
  return r;
}
static inline aabb_t aabb_union(aabb_t a, aabb_t b) /* aabb.hpp:53-56 */
{
  aabb_t r = {vmin(a.min, b.min), vmax(a.max, b.max)};
  return r;
}
static inline v3 aabb_center(aabb_t b) { return vdivs(vadd(b.min, b.max), 2.0f); } /* :19-22 */
static inline int aabb_max_extent(aabb_t b) /* aabb.hpp:47-51 */
{
  v3 e = vsub(b.max, b.min);
  return (e.x > e.y && e.x > e.z) ? 0 : (e.y > e.z) ? 1 : 2;
}
static inline float aabb_surface_area(aabb_t b) /* aabb.hpp:58-64 */
{
  v3 d = vsub(b.max, b.min);
  return 2.0f * (d.x * d.y + d.x * d.z + d.y * d.z);
}
static inline v3 aabb_offset(aabb_t b, v3 p) /* aabb.hpp:69-76 */
{
  v3 o = vsub(p, b.min);
  if (b.max.x > b.min.x) o.x /= b.max.x - b.min.x;
  if (b.max.y > b.min.y) o.y /= b.max.y - b.min.y;
  if (b.max.z > b.min.z) o.z /= b.max.z - b.min.z;
  return o;
}

void orc_aabb_props(const float mn[3], const float mx[3], const float p[3], float extent[3],
                    int* max_extent, float* surface_area, float offset[3])
{
  aabb_t b = {V(mn[0], mn[1], mn[2]), V(mx[0], mx[1], mx[2])};
  v3 e = vsub(b.max, b.min);
  extent[0] = e.x, extent[1] = e.y, extent[2] = e.z;
  *max_extent = aabb_max_extent(b);
  *surface_area = aabb_surface_area(b);
  v3 o = aabb_offset(b, V(p[0], p[1], p[2]));
  offset[0] = o.x, offset[1] = o.y, offset[2] = o.z;
}

/* ---------------------------------------------------------- ray + hits */
typedef struct { v3 origin; float t_min; v3 direction; float t_max; } ray_t; /* ray.hpp:8-20 */
typedef struct { /* intersection.hpp:8-14 (+ ids for the parity tests) */
  float t; v3 point; v3 normal; uint32_t material_id; uint32_t side; int32_t object; int32_t prim;
} isect_t;

static inline ray_t ray_from8(const float* r)
{
  ray_t q = {V(r[0], r[1], r[2]), r[3], V(r[4], r[5], r[6]), r[7]};
  return q;
}
static inline v3 ray_at(ray_t r, float t) { return vadd(r.origin, vscale(r.direction, t)); }

/* Transform (transform.hpp:9-35): m and inverse_m, both column-major */
typedef struct { float m[16], inv[16]; } xform_t;

/* transform.hpp:37-42 */
static inline v3 transform_point_m(const float* m, v3 p)
{
  float v[4];
  m4v4(m, p.x, p.y, p.z, 1.0f, v);
  return V(v[0] / v[3], v[1] / v[3], v[2] / v[3]);
}
/* transform.hpp:44-48 */
static inline v3 transform_vector_m(const float* m, v3 p)
{
  float v[4];
  m4v4(m, p.x, p.y, p.z, 0.0f, v);
  return V(v[0], v[1], v[2]);
}
/* transform.hpp:50-58 */
static inline ray_t inverse_transform_ray(const xform_t* t, ray_t r)
{
  ray_t o;
  o.origin = transform_point_m(t->inv, r.origin);
  o.direction = vnormalize(transform_vector_m(t->inv, r.direction));
  o.t_min = r.t_min;
  o.t_max = r.t_max;
  return o;
}
/* transform.hpp:60-66: transpose(inverse_m) * vec4(normal, 0) */
static inline v3 transform_normal(const xform_t* t, v3 n)
{
  float tr[16];
  for (int c = 0; c < 4; ++c)
    for (int r = 0; r < 4; ++r) tr[c * 4 + r] = t->inv[r * 4 + c];
  float v[4];
  m4v4(tr, n.x, n.y, n.z, 0.0f, v);
  return V(v[0], v[1], v[2]);
}
/* transform.hpp:69-88 */
static aabb_t transform_aabb(const xform_t* t, aabb_t b)
{
  if (aabb_is_empty(b)) return b;
  v3 pts[8];
  pts[0].x = pts[1].x = pts[2].x = pts[3].x = b.min.x;
  pts[4].x = pts[5].x = pts[6].x = pts[7].x = b.max.x;
  pts[0].y = pts[1].y = pts[4].y = pts[5].y = b.min.y;
  pts[2].y = pts[3].y = pts[6].y = pts[7].y = b.max.y;
  pts[0].z = pts[2].z = pts[4].z = pts[6].z = b.min.z;
  pts[1].z = pts[3].z = pts[5].z = pts[7].z = b.max.z;
  v3 p0 = transform_point_m(t->m, pts[0]);
  aabb_t nb = {p0, p0};
  for (int i = 1; i < 8; ++i) nb = aabb_enclose_pt(nb, transform_point_m(t->m, pts[i]));
  return nb;
}

void orc_inverse_transform_ray(const float m[16], const float inv[16], const float ray8[8],
                               float out8[8])
{
  xform_t t;
  memcpy(t.m, m, 64);
  memcpy(t.inv, inv, 64);
  ray_t r = inverse_transform_ray(&t, ray_from8(ray8));
  out8[0] = r.origin.x, out8[1] = r.origin.y, out8[2] = r.origin.z, out8[3] = r.t_min;
  out8[4] = r.direction.x, out8[5] = r.direction.y, out8[6] = r.direction.z, out8[7] = r.t_max;
}

/* glm::inverse(mat4) — cofactor formulation (glm/detail/func_matrix.inl compute_inverse<4,4>) */
void orc_mat4_inverse(const float m[16], float out[16])
{
#define M(c, r) m[(c)*4 + (r)]
  float c00 = M(2, 2) * M(3, 3) - M(3, 2) * M(2, 3), c02 = M(1, 2) * M(3, 3) - M(3, 2) * M(1, 3);
  float c03 = M(1, 2) * M(2, 3) - M(2, 2) * M(1, 3), c04 = M(2, 1) * M(3, 3) - M(3, 1) * M(2, 3);
  float c06 = M(1, 1) * M(3, 3) - M(3, 1) * M(1, 3), c07 = M(1, 1) * M(2, 3) - M(2, 1) * M(1, 3);
  float c08 = M(2, 1) * M(3, 2) - M(3, 1) * M(2, 2), c10 = M(1, 1) * M(3, 2) - M(3, 1) * M(1, 2);
  float c11 = M(1, 1) * M(2, 2) - M(2, 1) * M(1, 2), c12 = M(2, 0) * M(3, 3) - M(3, 0) * M(2, 3);
  float c14 = M(1, 0) * M(3, 3) - M(3, 0) * M(1, 3), c15 = M(1, 0) * M(2, 3) - M(2, 0) * M(1, 3);
  float c16 = M(2, 0) * M(3, 2) - M(3, 0) * M(2, 2), c18 = M(1, 0) * M(3, 2) - M(3, 0) * M(1, 2);
  float c19 = M(1, 0) * M(2, 2) - M(2, 0) * M(1, 2), c20 = M(2, 0) * M(3, 1) - M(3, 0) * M(2, 1);
  float c22 = M(1, 0) * M(3, 1) - M(3, 0) * M(1, 1), c23 = M(1, 0) * M(2, 1) - M(2, 0) * M(1, 1);
  float f0[4] = {c00, c00, c02, c03}, f1[4] = {c04, c04, c06, c07}, f2[4] = {c08, c08, c10, c11};
  float f3[4] = {c12, c12, c14, c15}, f4[4] = {c16, c16, c18, c19}, f5[4] = {c20, c20, c22, c23};
  float v0[4] = {M(1, 0), M(0, 0), M(0, 0), M(0, 0)}, v1[4] = {M(1, 1), M(0, 1), M(0, 1), M(0, 1)};
  float v2[4] = {M(1, 2), M(0, 2), M(0, 2), M(0, 2)}, v3_[4] = {M(1, 3), M(0, 3), M(0, 3), M(0, 3)};
  float sa[4] = {1, -1, 1, -1}, sb[4] = {-1, 1, -1, 1};
  for (int i = 0; i < 4; ++i) {
    out[0 + i] = (v1[i] * f0[i] - v2[i] * f1[i] + v3_[i] * f2[i]) * sa[i];
    out[4 + i] = (v0[i] * f0[i] - v2[i] * f3[i] + v3_[i] * f4[i]) * sb[i];
    out[8 + i] = (v0[i] * f1[i] - v1[i] * f3[i] + v3_[i] * f5[i]) * sa[i];
    out[12 + i] = (v0[i] * f2[i] - v1[i] * f4[i] + v2[i] * f5[i]) * sb[i];
  }
  float d = (M(0, 0) * out[0] + M(0, 1) * out[4]) + (M(0, 2) * out[8] + M(0, 3) * out[12]);
#undef M
  float ood = 1.0f / d;
  for (int i = 0; i < 16; ++i) out[i] *= ood;
}

/* ------------------------------------------------------- primitive tests */
/* intersections.cuh:7-41 */
static int ray_sphere_intersection_test(ray_t ray, v3 center, float radius, isect_t* rec)
{
  v3 oc = vsub(ray.origin, center);
  float a = vdot(ray.direction, ray.direction);
  float b = 2 * vdot(ray.direction, oc);
  float c = vdot(oc, oc) - radius * radius;
  float disc = b * b - 4 * a * c;
  if (disc < 0) return 0;
  float sq = sqrtf(disc);
  float t1 = (-b - sq) / (2 * a);
  float t2 = (-b + sq) / (2 * a);
  float t;
  if (t1 >= ray.t_min && t1 <= ray.t_max) t = t1;
  else if (t2 >= ray.t_min && t2 <= ray.t_max) t = t2;
  else return 0;
  rec->t = t;
  rec->point = ray_at(ray, t);
  v3 outward = vdivs(vsub(rec->point, center), radius);
  rec->side = vdot(ray.direction, outward) < 0 ? 0u : 1u;
  rec->normal = rec->side == 0u ? outward : vneg(outward);
  return 1;
}

/* intersections.cuh:43-85 */
static int ray_triangle_intersection_test(ray_t ray, v3 p0, v3 p1, v3 p2, isect_t* rec)
{
  const float EPSILON = 0.0000001f;
  v3 edge1 = vsub(p1, p0), edge2 = vsub(p2, p0);
  v3 h = vcross(ray.direction, edge2);
  float a = vdot(edge1, h);
  if (a > -EPSILON && a < EPSILON) return 0;
  float f = 1.0f / a;
  v3 s = vsub(ray.origin, p0);
  float u = f * vdot(s, h);
  if (u < 0.0 || u > 1.0) return 0;
  v3 q = vcross(s, edge1);
  float v = f * vdot(ray.direction, q);
  if (v < 0.0 || u + v > 1.0) return 0;
  float t = f * vdot(edge2, q);
  if ((t < ray.t_min) || (t > ray.t_max)) return 0;
  rec->t = t;
  rec->point = ray_at(ray, t);
  v3 outward = vnormalize(vcross(vsub(p1, p0), vsub(p2, p0))); /* triangle_normal :43-47 */
  rec->side = vdot(ray.direction, outward) < 0 ? 0u : 1u;
  rec->normal = rec->side == 0u ? outward : vneg(outward);
  rec->material_id = 1;
  return 1;
}

/* intersections.cuh:87-103 — an infinite-LINE slab test (no t range, no direction sign) */
static int ray_aabb_intersection_test(ray_t ray, aabb_t b)
{
  if (aabb_is_empty(b)) return 0;
  v3 t_min = vdiv(vsub(b.min, ray.origin), ray.direction);
  v3 t_max = vdiv(vsub(b.max, ray.origin), ray.direction);
  v3 real_min = vmin(t_min, t_max);
  v3 real_max = vmax(t_min, t_max);
  /* std::min(a,b) = (b<a)?b:a ; std::max(a,b) = (a<b)?b:a */
  float minmax = gmin(gmin(real_max.x, real_max.y), real_max.z);
  float maxmin = gmax(gmax(real_min.x, real_min.y), real_min.z);
  return minmax >= maxmin;
}

static void rec_to_hit(const isect_t* r, int hit, pt_hit* h)
{
  memset(h, 0, sizeof(*h));
  if (!hit) {
    h->t = -1.0f;
    h->object = -1;
    h->prim = -1;
    return;
  }
  h->t = r->t;
  h->point[0] = r->point.x, h->point[1] = r->point.y, h->point[2] = r->point.z;
  h->normal[0] = r->normal.x, h->normal[1] = r->normal.y, h->normal[2] = r->normal.z;
  h->material = r->material_id;
  h->side = r->side;
  h->object = r->object;
  h->prim = r->prim;
}

int orc_ray_triangle(const float ray8[8], const float p0[3], const float p1[3], const float p2[3],
                     pt_hit* out)
{
  isect_t r;
  memset(&r, 0, sizeof(r));
  r.prim = 0;
  int hit = ray_triangle_intersection_test(ray_from8(ray8), V(p0[0], p0[1], p0[2]),
                                           V(p1[0], p1[1], p1[2]), V(p2[0], p2[1], p2[2]), &r);
  rec_to_hit(&r, hit, out);
  return hit;
}
int orc_ray_sphere(const float ray8[8], const float c[3], float radius, pt_hit* out)
{
  isect_t r;
  memset(&r, 0, sizeof(r));
  r.prim = -1;
  int hit = ray_sphere_intersection_test(ray_from8(ray8), V(c[0], c[1], c[2]), radius, &r);
  rec_to_hit(&r, hit, out);
  return hit;
}
int orc_ray_aabb(const float ray8[8], const float mn[3], const float mx[3])
{
  aabb_t b = {V(mn[0], mn[1], mn[2]), V(mx[0], mx[1], mx[2])};
  return ray_aabb_intersection_test(ray_from8(ray8), b);
}

/* ----------------------------------------------------------- BVH build */
/* accelerators/bvh.cpp:74-253 restated with index arrays instead of shared_ptr nodes.
 * std::ranges::partition / nth_element only fix WHICH leaves go left/right (for distinct
 * keys), not their order, and every leaf range is split down to one triangle, so the
 * breadth-first flattened tree is the reference's tree. */
typedef struct bnode {
  aabb_t aabb;
  int left, right; /* -1 for leaves */
  uint32_t tri_index_begin;
} bnode;
typedef struct {
  bnode* nodes;
  int n_nodes;
  aabb_t* leaf_aabb; /* per triangle */
} bbuild;

static int bb_new_leaf(bbuild* B, uint32_t tri)
{
  int i = B->n_nodes++;
  B->nodes[i].aabb = B->leaf_aabb[tri];
  B->nodes[i].left = B->nodes[i].right = -1;
  B->nodes[i].tri_index_begin = tri * 3;
  return i;
}
static int bb_new_inner(bbuild* B, int l, int r)
{
  int i = B->n_nodes++;
  B->nodes[i].aabb = aabb_union(B->nodes[l].aabb, B->nodes[r].aabb);
  B->nodes[i].left = l;
  B->nodes[i].right = r;
  B->nodes[i].tri_index_begin = 0;
  return i;
}

static int bb_build(bbuild* B, uint32_t* leaves, size_t n);

static size_t bucket_of(const bbuild* B, aabb_t cb, int axis, uint32_t tri)
{
  /* bvh.cpp:128-134 */
  size_t b = (size_t)(int)(12.0f * comp(aabb_offset(cb, aabb_center(B->leaf_aabb[tri])), axis));
  if (b == 12) b = 11;
  return b;
}

static int bb_split_sah(bbuild* B, uint32_t* leaves, size_t n, aabb_t cb, int axis)
{
  /* bvh.cpp:115-182 */
  enum { NB = 12 };
  int count[NB];
  aabb_t bounds[NB];
  for (int i = 0; i < NB; ++i) {
    count[i] = 0;
    bounds[i] = aabb_empty();
  }
  aabb_t bound = aabb_empty();
  for (size_t i = 0; i < n; ++i) {
    size_t b = bucket_of(B, cb, axis, leaves[i]);
    count[b]++;
    bounds[b] = aabb_union(bounds[b], B->leaf_aabb[leaves[i]]);
    bound = aabb_union(bound, B->leaf_aabb[leaves[i]]);
  }
  float cost[NB - 1];
  for (int i = 0; i < NB - 1; ++i) {
    aabb_t b0 = aabb_empty(), b1 = aabb_empty();
    int c0 = 0, c1 = 0;
    for (int j = 0; j <= i; ++j) {
      b0 = aabb_union(b0, bounds[j]);
      c0 += count[j];
    }
    for (int j = i + 1; j < NB; ++j) {
      b1 = aabb_union(b1, bounds[j]);
      c1 += count[j];
    }
    cost[i] = .125f + ((float)c0 * aabb_surface_area(b0) + (float)c1 * aabb_surface_area(b1)) /
                          aabb_surface_area(bound);
  }
  float min_cost = cost[0];
  size_t min_bucket = 0;
  for (size_t i = 1; i < NB - 1; ++i)
    if (cost[i] < min_cost) {
      min_cost = cost[i];
      min_bucket = i;
    }
  /* std::ranges::partition (bvh.cpp:170-175).  The order it leaves inside each half decides
     which of two equal-centroid leaves becomes the left sibling further down, so libstdc++'s
     bidirectional algorithm (bits/ranges_algo.h, the classic two-pointer Hoare scheme) is
     restated: advance `first` over elements satisfying the predicate, retreat `tail` over
     elements failing it, swap, repeat. */
  size_t first = 0, tail = n;
  for (;;) {
    for (;;) {
      if (first == tail) goto partitioned;
      if (bucket_of(B, cb, axis, leaves[first]) <= min_bucket) ++first;
      else break;
    }
    --tail;
    for (;;) {
      if (first == tail) goto partitioned;
      if (!(bucket_of(B, cb, axis, leaves[tail]) <= min_bucket)) --tail;
      else break;
    }
    uint32_t sw = leaves[first];
    leaves[first] = leaves[tail];
    leaves[tail] = sw;
    ++first;
  }
partitioned:;
  const size_t nl = first, nr = n - first;
  if (nl == 0 || nr == 0) return -1; /* reference: panic("Shouldn't happen!") bvh.cpp:84-85 */
  int l = bb_build(B, leaves, nl);
  int r = bb_build(B, leaves + nl, nr);
  if (l < 0 || r < 0) return -1;
  return bb_new_inner(B, l, r);
}

static int bb_build(bbuild* B, uint32_t* leaves, size_t n)
{
  /* bvh.cpp:74-113 */
  aabb_t cb = aabb_empty();
  for (size_t i = 0; i < n; ++i) cb = aabb_enclose_pt(cb, aabb_center(B->leaf_aabb[leaves[i]]));
  int axis = aabb_max_extent(cb);
  if (n == 0) return -1;
  if (n == 1) return bb_new_leaf(B, leaves[0]);
  if (n == 2) {
    uint32_t l = leaves[0], r = leaves[1];
    if (comp(aabb_center(B->leaf_aabb[l]), axis) > comp(aabb_center(B->leaf_aabb[r]), axis)) {
      uint32_t t = l;
      l = r;
      r = t;
    }
    int li = bb_new_leaf(B, l), ri = bb_new_leaf(B, r);
    return bb_new_inner(B, li, ri);
  }
  if (n <= 4) {
    /* std::ranges::nth_element at n/2 keyed on the centroid coordinate (bvh.cpp:96-101).
       libstdc++'s introselect for these sizes (bits/stl_algo.h): while more than 3 elements
       remain, one median-of-three pivot partition round; then insertion sort of what is left.
       Restated so that equal keys end up in libstdc++'s order. */
#define KEY(i) comp(aabb_center(B->leaf_aabb[leaves[i]]), axis)
#define SWAPL(a, b) do { uint32_t t__ = leaves[a]; leaves[a] = leaves[b]; leaves[b] = t__; } while (0)
    size_t lo = 0, hi = n;
    const size_t nth = n / 2;
    while (hi - lo > 3) {
      /* __move_median_to_first(lo, lo+1, mid, hi-1) */
      const size_t a = lo + 1, b = lo + (hi - lo) / 2, c = hi - 1;
      if (KEY(a) < KEY(b)) {
        if (KEY(b) < KEY(c)) SWAPL(lo, b);
        else if (KEY(a) < KEY(c)) SWAPL(lo, c);
        else SWAPL(lo, a);
      } else if (KEY(a) < KEY(c)) SWAPL(lo, a);
      else if (KEY(b) < KEY(c)) SWAPL(lo, c);
      else SWAPL(lo, b);
      /* __unguarded_partition(lo+1, hi, pivot = lo) */
      size_t f = lo + 1, l = hi;
      for (;;) {
        while (KEY(f) < KEY(lo)) ++f;
        --l;
        while (KEY(lo) < KEY(l)) --l;
        if (!(f < l)) break;
        SWAPL(f, l);
        ++f;
      }
      if (f <= nth) lo = f; else hi = f;
    }
    for (size_t i = lo + 1; i < hi; ++i) { /* __insertion_sort(lo, hi) */
      uint32_t k = leaves[i];
      float kv = comp(aabb_center(B->leaf_aabb[k]), axis);
      size_t j = i;
      while (j > lo && kv < comp(aabb_center(B->leaf_aabb[leaves[j - 1]]), axis)) {
        leaves[j] = leaves[j - 1];
        --j;
      }
      leaves[j] = k;
    }
#undef KEY
#undef SWAPL
    size_t half = n / 2;
    int l = bb_build(B, leaves, half);
    int r = bb_build(B, leaves + half, n - half);
    if (l < 0 || r < 0) return -1;
    return bb_new_inner(B, l, r);
  }
  return bb_split_sah(B, leaves, n, cb, axis);
}

/* bvh_from_mesh (bvh.cpp:211-253): breadth-first flatten, siblings adjacent */
static orc_bvh_node* build_reference_bvh(const float* positions, const uint32_t* indices,
                                         uint64_t n_indices, uint32_t* n_out)
{
  const uint64_t n_tri = n_indices / 3;
  *n_out = 0;
  if (n_tri == 0) return NULL;
  bbuild B;
  B.nodes = (bnode*)malloc(sizeof(bnode) * (2 * n_tri));
  B.n_nodes = 0;
  B.leaf_aabb = (aabb_t*)malloc(sizeof(aabb_t) * n_tri);
  uint32_t* leaves = (uint32_t*)malloc(sizeof(uint32_t) * n_tri);
  for (uint64_t t = 0; t < n_tri; ++t) {
    aabb_t b = aabb_empty(); /* bvh.cpp:188-197 */
    for (int k = 0; k < 3; ++k) {
      const float* p = positions + 3 * (size_t)indices[3 * t + k];
      b = aabb_enclose_pt(b, V(p[0], p[1], p[2]));
    }
    B.leaf_aabb[t] = b;
    leaves[t] = (uint32_t)t;
  }
  int root = n_tri == 1 ? bb_new_leaf(&B, 0) : bb_build(&B, leaves, n_tri);
  orc_bvh_node* out = NULL;
  if (root >= 0) {
    const uint32_t total = (uint32_t)(2 * n_tri - 1);
    out = (orc_bvh_node*)calloc(total, sizeof(orc_bvh_node));
    int* queue = (int*)malloc(sizeof(int) * total);
    uint32_t* lin = (uint32_t*)malloc(sizeof(uint32_t) * total);
    uint32_t qh = 0, qt = 0, size = 0;
#define PUSH(ni)                                                                                 \
  do {                                                                                           \
    const bnode* nd__ = &B.nodes[ni];                                                            \
    queue[qt] = (ni);                                                                            \
    lin[qt++] = size;                                                                            \
    out[size].min[0] = nd__->aabb.min.x, out[size].min[1] = nd__->aabb.min.y;                    \
    out[size].min[2] = nd__->aabb.min.z, out[size].max[0] = nd__->aabb.max.x;                    \
    out[size].max[1] = nd__->aabb.max.y, out[size].max[2] = nd__->aabb.max.z;                    \
    out[size].first_child_or_primitive = nd__->left < 0 ? nd__->tri_index_begin : 0;             \
    out[size].primitive_count = nd__->left < 0 ? 1 : 0;                                          \
    ++size;                                                                                      \
  } while (0)
    PUSH(root);
    while (qh < qt) {
      int ni = queue[qh];
      uint32_t li = lin[qh];
      ++qh;
      if (B.nodes[ni].left >= 0) {
        out[li].first_child_or_primitive = size;
        PUSH(B.nodes[ni].left);
        PUSH(B.nodes[ni].right);
      }
    }
#undef PUSH
    *n_out = size;
    free(queue);
    free(lin);
  }
  free(B.nodes);
  free(B.leaf_aabb);
  free(leaves);
  return out;
}

/* ----------------------------------------------------------------- scene */
typedef struct { /* GPUObject, scene.hpp:14-20 */
  int type;
  uint32_t index;
  xform_t transform;
  aabb_t aabb;
} gpu_object;

struct orc_scene {
  uint32_t n_objects;
  gpu_object* objects;
  uint32_t* object_material_indices;
  uint32_t n_spheres;
  pt_sphere* spheres;
  uint64_t n_vertices, n_indices;
  float* positions;
  uint32_t* indices;
  orc_bvh_node* bvh;
  uint32_t bvh_size;
  uint32_t n_materials;
  pt_material* materials;
};

/* SceneDescription::build_scene (scene_description.cpp:12-117) */
orc_scene* orc_scene_create(const pt_scene_desc* d)
{
  orc_scene* s = (orc_scene*)calloc(1, sizeof(orc_scene));
  s->n_objects = d->n_objects;
  s->objects = (gpu_object*)calloc(d->n_objects ? d->n_objects : 1, sizeof(gpu_object));
  s->object_material_indices = (uint32_t*)calloc(d->n_objects ? d->n_objects : 1, 4);
  s->n_spheres = d->n_spheres;
  s->spheres = (pt_sphere*)malloc(sizeof(pt_sphere) * (d->n_spheres ? d->n_spheres : 1));
  if (d->n_spheres) memcpy(s->spheres, d->spheres, sizeof(pt_sphere) * d->n_spheres);
  s->n_vertices = d->n_vertices;
  s->n_indices = d->n_indices;
  s->positions = (float*)malloc(12 * (d->n_vertices ? d->n_vertices : 1));
  if (d->n_vertices) memcpy(s->positions, d->positions, 12 * d->n_vertices);
  s->indices = (uint32_t*)malloc(4 * (d->n_indices ? d->n_indices : 1));
  if (d->n_indices) memcpy(s->indices, d->indices, 4 * d->n_indices);
  s->n_materials = d->n_materials;
  s->materials = (pt_material*)malloc(sizeof(pt_material) * (d->n_materials ? d->n_materials : 1));
  if (d->n_materials) memcpy(s->materials, d->materials, sizeof(pt_material) * d->n_materials);

  /* mesh.aabb: Assimp aiProcess_GenBoundingBoxes == min/max over the vertices */
  aabb_t mesh_aabb = aabb_empty();
  for (uint64_t i = 0; i < d->n_vertices; ++i)
    mesh_aabb = aabb_enclose_pt(mesh_aabb, V(d->positions[3 * i], d->positions[3 * i + 1], d->positions[3 * i + 2]));

  for (uint32_t i = 0; i < d->n_objects; ++i) {
    const pt_object* o = &d->objects[i];
    gpu_object* g = &s->objects[i];
    g->type = o->type;
    memcpy(g->transform.m, o->m, 64);
    memcpy(g->transform.inv, o->inv, 64);
    s->object_material_indices[i] = o->material;
    if (o->type == PT_OBJ_SPHERE) { /* scene_description.cpp:23-39 */
      g->index = o->prim_index;
      const pt_sphere* sp = &d->spheres[o->prim_index];
      v3 c = transform_point_m(o->m, V(sp->center[0], sp->center[1], sp->center[2]));
      float r = vlength(transform_vector_m(o->m, V(1.0f, 0.0f, 0.0f))) * sp->radius;
      g->aabb.min = vsub(c, V(r, r, r));
      g->aabb.max = vadd(c, V(r, r, r));
    } else { /* :40-44 */
      g->index = 0;
      g->aabb = transform_aabb(&g->transform, mesh_aabb);
    }
  }
  s->bvh = build_reference_bvh(s->positions, s->indices, s->n_indices, &s->bvh_size);
  return s;
}

void orc_scene_destroy(orc_scene* s)
{
  if (!s) return;
  free(s->objects);
  free(s->object_material_indices);
  free(s->spheres);
  free(s->positions);
  free(s->indices);
  free(s->bvh);
  free(s->materials);
  free(s);
}
uint32_t orc_scene_bvh_size(const orc_scene* s) { return s->bvh_size; }
const orc_bvh_node* orc_scene_bvh(const orc_scene* s) { return s->bvh; }
void orc_scene_object_aabb(const orc_scene* s, uint32_t i, float mn[3], float mx[3])
{
  mn[0] = s->objects[i].aabb.min.x, mn[1] = s->objects[i].aabb.min.y, mn[2] = s->objects[i].aabb.min.z;
  mx[0] = s->objects[i].aabb.max.x, mx[1] = s->objects[i].aabb.max.y, mx[2] = s->objects[i].aabb.max.z;
}

/* ------------------------------------------------------------- traversal */
static inline v3 pos_at(const orc_scene* s, uint32_t idx)
{
  return V(s->positions[3 * (size_t)idx], s->positions[3 * (size_t)idx + 1], s->positions[3 * (size_t)idx + 2]);
}

/* ray_mesh_intersection_test (path_tracer.cu:36-76).  The reference's 24-entry
 * StaticStack is unchecked; a growable stack is used here (no overflow UB). */
static int ray_mesh_intersection_test(const orc_scene* s, ray_t ray, const xform_t* tf,
                                      isect_t* rec, int mode)
{
  int hit = 0;
  if (s->n_indices == 0) return 0;
  if (mode == 1) { /* brute force: every triangle in index order */
    for (uint32_t i = 0; i + 2 < s->n_indices; i += 3) {
      v3 p0 = transform_point_m(tf->m, pos_at(s, s->indices[i]));
      v3 p1 = transform_point_m(tf->m, pos_at(s, s->indices[i + 1]));
      v3 p2 = transform_point_m(tf->m, pos_at(s, s->indices[i + 2]));
      if (ray_triangle_intersection_test(ray, p0, p1, p2, rec)) {
        hit = 1;
        rec->prim = (int32_t)(i / 3);
        ray.t_max = rec->t;
      }
    }
    return hit;
  }
  const ray_t tray = inverse_transform_ray(tf, ray);
  uint32_t stack_small[256];
  uint32_t* stack = stack_small;
  size_t cap = 256, sp = 0;
  stack[sp++] = 0;
  while (sp != 0) {
    const uint32_t ni = stack[--sp];
    const orc_bvh_node* node = &s->bvh[ni];
    if (node->primitive_count != 0) {
      const uint32_t i = node->first_child_or_primitive;
      v3 p0 = transform_point_m(tf->m, pos_at(s, s->indices[i]));
      v3 p1 = transform_point_m(tf->m, pos_at(s, s->indices[i + 1]));
      v3 p2 = transform_point_m(tf->m, pos_at(s, s->indices[i + 2]));
      if (ray_triangle_intersection_test(ray, p0, p1, p2, rec)) {
        hit = 1;
        rec->prim = (int32_t)(i / 3);
        ray.t_max = rec->t;
      }
    } else {
      aabb_t b = {V(node->min[0], node->min[1], node->min[2]), V(node->max[0], node->max[1], node->max[2])};
      if (ray_aabb_intersection_test(tray, b)) {
        if (sp + 2 > cap) {
          uint32_t* ns = (uint32_t*)malloc(cap * 2 * sizeof(uint32_t));
          memcpy(ns, stack, sp * sizeof(uint32_t));
          if (stack != stack_small) free(stack);
          stack = ns;
          cap *= 2;
        }
        stack[sp++] = node->first_child_or_primitive + 1;
        stack[sp++] = node->first_child_or_primitive;
      }
    }
  }
  if (stack != stack_small) free(stack);
  return hit;
}

/* ray_object_intersection_test (path_tracer.cu:78-108) */
static int ray_object_intersection_test(const orc_scene* s, ray_t ray, const gpu_object* obj,
                                        isect_t* rec, int mode)
{
  if (!ray_aabb_intersection_test(ray, obj->aabb)) return 0;
  int hit = 0;
  if (obj->type == PT_OBJ_SPHERE) {
    const ray_t tr = inverse_transform_ray(&obj->transform, ray);
    const pt_sphere* sp = &s->spheres[obj->index];
    hit = ray_sphere_intersection_test(tr, V(sp->center[0], sp->center[1], sp->center[2]), sp->radius, rec);
    if (hit) {
      rec->point = transform_point_m(obj->transform.m, rec->point);
      rec->t = vlength(vsub(rec->point, ray.origin)); /* glm::distance(ray.origin, point) */
      rec->normal = transform_normal(&obj->transform, rec->normal);
      rec->prim = -1;
    }
  } else {
    hit = ray_mesh_intersection_test(s, ray, &obj->transform, rec, mode);
  }
  return hit;
}

/* ray_scene_intersection_test (path_tracer.cu:110-128) */
static int ray_scene_intersection_test(const orc_scene* s, ray_t ray, isect_t* rec, int mode)
{
  int hit = 0;
  for (uint32_t i = 0; i < s->n_objects; ++i) {
    if (ray_object_intersection_test(s, ray, &s->objects[i], rec, mode)) {
      hit = 1;
      rec->material_id = s->object_material_indices[i];
      rec->object = (int32_t)i;
      ray.t_max = rec->t;
    }
  }
  return hit;
}

void orc_trace_batch(const orc_scene* s, const float* rays8, uint64_t n, pt_hit* out, int mode)
{
#pragma omp parallel for schedule(dynamic, 64)
  for (long long i = 0; i < (long long)n; ++i) {
    isect_t rec;
    memset(&rec, 0, sizeof(rec));
    int hit = ray_scene_intersection_test(s, ray_from8(rays8 + 8 * i), &rec, mode);
    rec_to_hit(&rec, hit, &out[i]);
  }
}

/* ---------------------------------------------------------------- camera */
typedef struct { float m[16]; float vfov; uint32_t width, height; } gpu_camera; /* camera.hpp:10-15 */

/* Camera::to_gpu_camera (camera.cpp:5-13): translate(identity, position) * mat4_cast(rotation) */
static gpu_camera to_gpu_camera(const pt_camera* c, uint32_t w, uint32_t h)
{
  gpu_camera g;
  const float qw = c->rotation[0], qx = c->rotation[1], qy = c->rotation[2], qz = c->rotation[3];
  float qxx = qx * qx, qyy = qy * qy, qzz = qz * qz, qxz = qx * qz, qxy = qx * qy, qyz = qy * qz;
  float qwx = qw * qx, qwy = qw * qy, qwz = qw * qz;
  float R[16] = {0};
  R[0] = 1.0f - 2.0f * (qyy + qzz), R[1] = 2.0f * (qxy + qwz), R[2] = 2.0f * (qxz - qwy);
  R[4] = 2.0f * (qxy - qwz), R[5] = 1.0f - 2.0f * (qxx + qzz), R[6] = 2.0f * (qyz + qwx);
  R[8] = 2.0f * (qxz + qwy), R[9] = 2.0f * (qyz - qwx), R[10] = 1.0f - 2.0f * (qxx + qyy);
  R[15] = 1.0f;
  float T[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, c->position[0], c->position[1], c->position[2], 1};
  for (int j = 0; j < 4; ++j)
    for (int r = 0; r < 4; ++r) {
      float acc = T[0 + r] * R[j * 4 + 0];
      acc = acc + T[4 + r] * R[j * 4 + 1];
      acc = acc + T[8 + r] * R[j * 4 + 2];
      acc = acc + T[12 + r] * R[j * 4 + 3];
      g.m[j * 4 + r] = acc;
    }
  g.vfov = c->vfov;
  g.width = w;
  g.height = h;
  return g;
}

/* generate_ray (ray_gen.cu:34-61) */
static ray_t generate_ray(const gpu_camera* cam, float x, float y)
{
  const float aspect_ratio = (float)cam->width / (float)cam->height;
  const float viewport_height = 2.0f * tanf(cam->vfov / 2);
  const float viewport_width = aspect_ratio * viewport_height;
  const float focal_length = 1.0f;
  const v3 origin = V(0, 0, 0);
  const v3 horizontal = V(viewport_width, 0, 0);
  const v3 vertical = V(0, viewport_height, 0);
  const v3 llc = vsub(vsub(vsub(origin, vdivs(horizontal, 2.f)), vdivs(vertical, 2.f)), V(0, 0, focal_length));
  const float u = x / (float)(cam->width - 1);
  const float v = ((float)cam->height - y) / (float)(cam->height - 1);
  const v3 direction = vsub(vadd(vadd(llc, vscale(horizontal, u)), vscale(vertical, v)), origin);
  float wo[4], wd[4];
  m4v4(cam->m, origin.x, origin.y, origin.z, 1.0f, wo);
  m4v4(cam->m, direction.x, direction.y, direction.z, 0.0f, wd);
  ray_t r;
  r.origin = V(wo[0], wo[1], wo[2]);
  r.t_min = 1e-4f;
  r.direction = vnormalize(V(wd[0], wd[1], wd[2]));
  r.t_max = FLT_MAX;
  return r;
}

void orc_generate_ray(const pt_camera* cam, uint32_t w, uint32_t h, float x, float y, float out8[8])
{
  gpu_camera g = to_gpu_camera(cam, w, h);
  ray_t r = generate_ray(&g, x, y);
  out8[0] = r.origin.x, out8[1] = r.origin.y, out8[2] = r.origin.z, out8[3] = r.t_min;
  out8[4] = r.direction.x, out8[5] = r.direction.y, out8[6] = r.direction.z, out8[7] = r.t_max;
}

/* --------------------------------------------------------------- shading */
/* get_background_color (path_tracer.cu:29-34); glm::lerp(x,y,a) = x*(1-a) + y*a */
static v3 get_background_color(ray_t r)
{
  v3 unit = vnormalize(r.direction);
  float t = 0.5f * (unit.y + 1.0f);
  return vadd(vscale(V(0.5f, 0.7f, 1.0f), 1.0f - t), vscale(V(1.0f, 1.0f, 1.0f), t));
}

/* random_in_unit_sphere (distributions.cuh:6-19) */
static v3 random_in_unit_sphere(uint32_t* rng)
{
  const float phi = 2.f * 3.14159265358979323846264338327950288f * orc_rng_uniform(rng);
  const float cos_theta = 2.f * orc_rng_uniform(rng) - 1.f;
  const float sin_theta = sqrtf(1 - cos_theta * cos_theta);
  return V(cosf(phi) * sin_theta, sinf(phi) * sin_theta, cos_theta);
}

/* reflectance (path_tracer.cu:130-136) */
static float reflectance(float cosine, float ref_idx)
{
  float r0 = (1 - ref_idx) / (1 + ref_idx);
  r0 = r0 * r0;
  return r0 + (1 - r0) * powf((1 - cosine), 5);
}
static inline float gsign(float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f); }

/* evaluate_material (path_tracer.cu:138-201) */
static void evaluate_material(ray_t* ray, const isect_t* is, uint32_t* rng, v3* color,
                              const pt_material* materials)
{
  ray->origin = vsub(is->point, vscale(is->normal, 1e-4f * gsign(vdot(ray->direction, is->normal))));
  const pt_material* mat = &materials[is->material_id];
  if (mat->type == PT_MAT_DIFFUSE) {
    v3 sd = vnormalize(vadd(is->normal, random_in_unit_sphere(rng)));
    if (fabs(sd.x) < 1e-8 && fabs(sd.y) < 1e-8 && fabs(sd.z) < 1e-8) sd = is->normal;
    ray->direction = sd;
    *color = vmul(*color, V(mat->albedo[0], mat->albedo[1], mat->albedo[2]));
  } else if (mat->type == PT_MAT_METAL) {
    /* glm::reflect(I, N) = I - N * dot(N, I) * 2 */
    v3 reflected = vsub(ray->direction, vscale(vscale(is->normal, vdot(is->normal, ray->direction)), 2.0f));
    v3 sd = vadd(reflected, vscale(random_in_unit_sphere(rng), mat->fuzz));
    ray->direction = sd;
    if (vdot(sd, is->normal) > 0) *color = vmul(*color, V(mat->albedo[0], mat->albedo[1], mat->albedo[2]));
    else *color = V(0, 0, 0);
  } else {
    const float ratio = is->side == 0u ? (1.0f / mat->refraction_index) : mat->refraction_index;
    v3 unit = vnormalize(ray->direction);
    float cos_theta = fminf(vdot(vneg(unit), is->normal), 1.0f);
    float sin_theta = sqrtf(1.0f - cos_theta * cos_theta);
    int cannot_refract = ratio * sin_theta > 1.0;
    v3 direction;
    if (cannot_refract || reflectance(cos_theta, ratio) > orc_rng_uniform(rng)) {
      direction = vsub(unit, vscale(vscale(is->normal, vdot(is->normal, unit)), 2.0f));
    } else {
      /* glm::refract */
      float dv = vdot(is->normal, unit);
      float k = 1.0f - ratio * ratio * (1.0f - dv * dv);
      direction = k >= 0.0f ? vsub(vscale(unit, ratio), vscale(is->normal, ratio * dv + sqrtf(k))) : V(0, 0, 0);
    }
    ray->origin = is->point;
    ray->t_min = 1e-5f;
    ray->direction = direction;
    ray->t_max = FLT_MAX;
  }
}

/* final_gather (path_tracer.cu:203-219): running mean */
static inline float temporal(float oldv, float newv, int iteration)
{
  float sc = (float)(iteration + 1);
  return iteration == 0 ? newv : (oldv * (sc - 1) + newv) / sc;
}
static void final_gather(int iteration, v3 c, v3 n, float d, float* color3, float* normal3, float* depth1)
{
  color3[0] = temporal(color3[0], c.x, iteration);
  color3[1] = temporal(color3[1], c.y, iteration);
  color3[2] = temporal(color3[2], c.z, iteration);
  normal3[0] = temporal(normal3[0], n.x, iteration);
  normal3[1] = temporal(normal3[1], n.y, iteration);
  normal3[2] = temporal(normal3[2], n.z, iteration);
  *depth1 = temporal(*depth1, d, iteration);
}

/* path_tracing_mega_kernel (path_tracer.cu:227-269) */
void orc_render_megakernel(const orc_scene* s, const pt_camera* cam, uint32_t w, uint32_t h,
                           int first_iteration, int n_iterations, int max_bounces, float* color3,
                           float* normal3, float* depth1, uint64_t* rays_out)
{
  const gpu_camera gc = to_gpu_camera(cam, w, h);
  uint64_t rays = 0;
  for (int it = first_iteration; it < first_iteration + n_iterations; ++it) {
#pragma omp parallel for schedule(dynamic, 64) reduction(+ : rays)
    for (long long index = 0; index < (long long)w * h; ++index) {
      const uint32_t x = (uint32_t)(index % w), y = (uint32_t)(index / w);
      uint32_t rng = orc_rng_seed(orc_hash(orc_hash((uint32_t)index) ^ (uint32_t)it));
      const float fx = (float)x + orc_rng_uniform(&rng);
      const float fy = (float)y + orc_rng_uniform(&rng);
      ray_t ray = generate_ray(&gc, fx, fy);
      v3 color = V(1.0f, 1.0f, 1.0f);
      v3 normal = vneg(ray.direction);
      float depth = 1e6f;
      for (int i = 0; i < max_bounces; ++i) {
        isect_t is;
        memset(&is, 0, sizeof(is));
        ++rays;
        if (!ray_scene_intersection_test(s, ray, &is, 0)) {
          color = vmul(color, get_background_color(ray));
          break;
        }
        if (i == 0) {
          normal = is.normal;
          depth = is.t;
        }
        evaluate_material(&ray, &is, &rng, &color, s->materials);
      }
      final_gather(it, color, normal, depth, color3 + 3 * index, normal3 + 3 * index, depth1 + index);
    }
  }
  if (rays_out) *rays_out = rays;
}

/* PathTracer::path_trace, streaming branch (path_tracer.cu:413-470) with raygen_kernel
 * (ray_gen.cu:11-32), intersection_kernel (:271-290), material_kernel (:292-315),
 * thrust::stable_partition (:454-457) and final_gathering_kernel (:317-330). */
void orc_render_streaming(const orc_scene* s, const pt_camera* cam, uint32_t w, uint32_t h,
                          int first_iteration, int n_iterations, int max_bounces, float* color3,
                          float* normal3, float* depth1, uint64_t* rays_out)
{
  const gpu_camera gc = to_gpu_camera(cam, w, h);
  const size_t P = (size_t)w * h;
  ray_t* rays = (ray_t*)malloc(sizeof(ray_t) * P);
  int* pixel = (int*)malloc(sizeof(int) * P);
  v3* color = (v3*)malloc(sizeof(v3) * P);
  v3* normal = (v3*)malloc(sizeof(v3) * P);
  float* depth = (float*)malloc(sizeof(float) * P);
  uint8_t* bounces_left = (uint8_t*)malloc(P);
  isect_t* isect = (isect_t*)calloc(P, sizeof(isect_t));
  /* partition scratch */
  ray_t* rays2 = (ray_t*)malloc(sizeof(ray_t) * P);
  int* pixel2 = (int*)malloc(sizeof(int) * P);
  v3* color2 = (v3*)malloc(sizeof(v3) * P);
  v3* normal2 = (v3*)malloc(sizeof(v3) * P);
  float* depth2 = (float*)malloc(sizeof(float) * P);
  uint8_t* bl2 = (uint8_t*)malloc(P);
  uint64_t total_rays = 0;

  for (int it = first_iteration; it < first_iteration + n_iterations; ++it) {
#pragma omp parallel for schedule(static)
    for (long long index = 0; index < (long long)P; ++index) {
      const uint32_t x = (uint32_t)(index % w), y = (uint32_t)(index / w);
      uint32_t rng = orc_rng_seed(orc_hash(orc_hash((uint32_t)index) ^ (uint32_t)it));
      const float fx = (float)x + orc_rng_uniform(&rng);
      const float fy = (float)y + orc_rng_uniform(&rng);
      ray_t ray = generate_ray(&gc, fx, fy);
      color[index] = V(1.0f, 1.0f, 1.0f);
      depth[index] = 1e6f;
      normal[index] = vneg(ray.direction);
      bounces_left[index] = 50;
      rays[index] = ray;
      pixel[index] = (int)index;
    }
    size_t paths_count = P;
    for (int i = 0; i < max_bounces && paths_count > 0; ++i) {
      total_rays += paths_count;
#pragma omp parallel for schedule(dynamic, 64)
      for (long long index = 0; index < (long long)paths_count; ++index) {
        isect_t is = isect[index];
        if (ray_scene_intersection_test(s, rays[index], &is, 0)) {
          isect[index] = is;
        } else {
          isect[index].t = -1.0f;
          bounces_left[index] = 0;
        }
      }
#pragma omp parallel for schedule(static)
      for (long long index = 0; index < (long long)paths_count; ++index) {
        uint32_t rng = orc_rng_seed(orc_hash(orc_hash((uint32_t)index) ^ (uint32_t)it));
        orc_rng_discard(&rng, (uint64_t)i);
        const isect_t is = isect[index];
        if (is.t < 0) {
          color[index] = vmul(color[index], get_background_color(rays[index]));
          continue;
        }
        if (i == 0) {
          depth[index] = is.t;
          normal[index] = is.normal;
        }
        evaluate_material(&rays[index], &is, &rng, &color[index], s->materials);
      }
      /* stable partition of [0, paths_count) by bounces_left > 0 */
      size_t nl = 0, nr = 0;
      for (size_t k = 0; k < paths_count; ++k) {
        if (bounces_left[k] > 0) {
          rays[nl] = rays[k], pixel[nl] = pixel[k], color[nl] = color[k];
          normal[nl] = normal[k], depth[nl] = depth[k], bounces_left[nl] = bounces_left[k];
          ++nl;
        } else {
          rays2[nr] = rays[k], pixel2[nr] = pixel[k], color2[nr] = color[k];
          normal2[nr] = normal[k], depth2[nr] = depth[k], bl2[nr] = bounces_left[k];
          ++nr;
        }
      }
      for (size_t k = 0; k < nr; ++k) {
        rays[nl + k] = rays2[k], pixel[nl + k] = pixel2[k], color[nl + k] = color2[k];
        normal[nl + k] = normal2[k], depth[nl + k] = depth2[k], bounces_left[nl + k] = bl2[k];
      }
      paths_count = nl;
    }
    for (size_t index = 0; index < P; ++index) {
      const int p = pixel[index];
      final_gather(it, color[index], normal[index], depth[index], color3 + 3 * (size_t)p,
                   normal3 + 3 * (size_t)p, depth1 + p);
    }
  }
  if (rays_out) *rays_out = total_rays;
  free(rays), free(pixel), free(color), free(normal), free(depth), free(bounces_left), free(isect);
  free(rays2), free(pixel2), free(color2), free(normal2), free(depth2), free(bl2);
}

/* ---------------------------------------------------------------- denoise */
/* denoising_kernel + EdgeAvoidingATrousDenoiser::denoise (denoiser.cu:24-116) */
void orc_denoise(uint32_t w, uint32_t h, const pt_camera* cam, const float* color3,
                 const float* normal3, const float* depth1, int filter_size, float c_phi,
                 float n_phi, float p_phi, int clamp_fix, float* out3, uint8_t* tainted)
{
  const gpu_camera gc = to_gpu_camera(cam, w, h);
  const size_t P = (size_t)w * h;
  float* bufA = (float*)malloc(P * 12);
  float* bufB = (float*)malloc(P * 12);
  uint8_t* taintA = (uint8_t*)calloc(P, 1);
  uint8_t* taintB = (uint8_t*)calloc(P, 1);
  const float* in = color3;
  const uint8_t* tin = taintA; /* all zero */
  float* outb = bufA;
  uint8_t* tout = taintB;
  static const float kernel[3] = {3.f / 8.f, 1.f / 4.f, 1.f / 16.f};
  const int W = (int)w, H = (int)h;
  for (int step = 1; step <= filter_size; step *= 2) {
#pragma omp parallel for schedule(static)
    for (long long index = 0; index < (long long)P; ++index) {
      const int x = (int)(index % w), y = (int)(index / w);
      const v3 cval = V(in[3 * index], in[3 * index + 1], in[3 * index + 2]);
      const v3 nval = V(normal3[3 * index], normal3[3 * index + 1], normal3[3 * index + 2]);
      const ray_t ray = generate_ray(&gc, x + 0.5f, y + 0.5f);
      const v3 pval = ray_at(ray, depth1[index]);
      v3 sum = V(0, 0, 0);
      float cum_w = 0.0f;
      uint8_t taint = 0;
      for (int dy = -2; dy <= 2; ++dy) {
        for (int dx = -2; dx <= 2; ++dx) {
          /* std::clamp to [0, width] / [0, height] INCLUSIVE (denoiser.cu:51-54) */
          const int umax = clamp_fix ? W - 1 : W, vmax = clamp_fix ? H - 1 : H;
          int u = x + dx * step, v = y + dy * step;
          u = u < 0 ? 0 : (u > umax ? umax : u);
          v = v < 0 ? 0 : (v > vmax ? vmax : v);
          long long ti = (long long)u + (long long)v * W;
          if (ti >= (long long)P) {
            /* the reference reads past the end of the buffers here (undefined);
               the product substitutes the last row, and the pixel is marked */
            taint = 1;
            ti = (long long)(u < W - 1 ? u : W - 1) + (long long)(H - 1) * W;
          }
          if (tin[ti]) taint = 1;
          const v3 ctemp = V(in[3 * ti], in[3 * ti + 1], in[3 * ti + 2]);
          v3 t = vsub(cval, ctemp);
          float dist2 = vdot(t, t);
          const float c_w = fminf(expf(-dist2 / c_phi), 1.0f);
          const v3 ntemp = V(normal3[3 * ti], normal3[3 * ti + 1], normal3[3 * ti + 2]);
          t = vsub(nval, ntemp);
          dist2 = fmaxf(vdot(t, t) / (float)(step * step), 0.0f);
          const float n_w = fminf(expf(-dist2 / n_phi), 1.0f);
          const ray_t tr = generate_ray(&gc, u + 0.5f, v + 0.5f);
          const v3 ptmp = ray_at(tr, depth1[ti]);
          t = vsub(pval, ptmp);
          dist2 = vdot(t, t);
          const float p_w = fminf(expf(-dist2 / p_phi), 1.0f);
          const float weight = c_w * n_w * p_w;
          const int ki = abs(dx) < abs(dy) ? abs(dx) : abs(dy);
          sum = vadd(sum, vscale(vscale(ctemp, weight), kernel[ki]));
          cum_w += weight * kernel[ki];
        }
      }
      outb[3 * index] = sum.x / cum_w;
      outb[3 * index + 1] = sum.y / cum_w;
      outb[3 * index + 2] = sum.z / cum_w;
      tout[index] = taint;
    }
    in = outb;
    tin = tout;
    outb = (outb == bufA) ? bufB : bufA;
    tout = (tout == taintA) ? taintB : taintA;
  }
  memcpy(out3, in, P * 12);
  if (tainted) memcpy(tainted, tin, P);
  free(bufA), free(bufB), free(taintA), free(taintB);
}

/* --------------------------------------------------------------- tonemap */
/* preview_kernel / preview_depth_kernel / linear_to_gamma (path_tracer.cu:221-225, 334-385) */
static inline uint8_t to255(float v)
{
  float c = gmin(gmax(v, 0.f), 1.f); /* glm::clamp = min(max(x, lo), hi) */
  return (uint8_t)(c * 255.99f);
}
void orc_tonemap(int kind, uint32_t n, const float* src, uint8_t* rgba)
{
  const float g = 1.f / 2.2f;
  for (uint32_t i = 0; i < n; ++i) {
    v3 c;
    uint8_t a = 255;
    if (kind == 3) {
      float d = 1 / src[i];
      c = V(d, d, d);
      a = 1;
    } else {
      c = V(src[3 * (size_t)i], src[3 * (size_t)i + 1], src[3 * (size_t)i + 2]);
      if (kind == 2) c = vadd(vscale(c, 0.5f), V(0.5f, 0.5f, 0.5f));
    }
    c = V(powf(c.x, g), powf(c.y, g), powf(c.z, g));
    rgba[4 * (size_t)i + 0] = to255(c.x);
    rgba[4 * (size_t)i + 1] = to255(c.y);
    rgba[4 * (size_t)i + 2] = to255(c.z);
    rgba[4 * (size_t)i + 3] = a;
  }
}

int orc_num_threads(void)
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
