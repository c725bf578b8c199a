// common.cuh — device-side data model of the B200 wavefront path tracer.
//
// Behavioural spec followed here (reference = LesleyLai/cuda-path-tracer):
//   RNG         src/lib/hash.cuh:4-14 + thrust minstd_rand / uniform_real_distribution
//   camera      src/lib/ray_gen.cu:34-61, src/lib/camera.cpp:5-13
//   Ray         src/lib/ray.hpp:8-20  (origin, t_min, direction, t_max == two float4)
// Nothing in this file is shared with the reference's sources; the layouts are
// float4-packed SoA so that every path-state access is one 16-byte load/store.
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#define PT_HD __host__ __device__ __forceinline__
#define PT_D __device__ __forceinline__

// Debug build (python -m cuda_path_tracer_b200.build --debug -> libb200pt_debug.so, selected with
// B200PT_LIB): every index into the traversal stack, the tree, the triangle array, the parked-state
// buffers and the bin lists is checked on the device; a violation prints what and where and traps
// (the launch fails, the parity suite turns red).  compute-sanitizer is not available on the GPU
// pool, this is its stand-in.  The release build compiles the checks away.
#ifdef PT_BOUNDS_CHECK
#include <stdio.h>
#define PT_CHECK(cond, what)                                                                       \
  do {                                                                                             \
    if (!(cond)) {                                                                                 \
      printf("b200pt bounds violation: %s (%s:%d, block %d thread %d)\n", what, __FILE__, __LINE__, \
             (int)blockIdx.x, (int)threadIdx.x);                                                   \
      __trap();                                                                                    \
    }                                                                                              \
  } while (0)
#else
#define PT_CHECK(cond, what) ((void)0)
#endif

namespace pt {

// ---------------------------------------------------------------- float3 math
struct f3 {
  float x, y, z;
};
PT_HD f3 mk3(float x, float y, float z) { return f3{x, y, z}; }
PT_HD f3 operator+(f3 a, f3 b) { return f3{a.x + b.x, a.y + b.y, a.z + b.z}; }
PT_HD f3 operator-(f3 a, f3 b) { return f3{a.x - b.x, a.y - b.y, a.z - b.z}; }
PT_HD f3 operator-(f3 a) { return f3{-a.x, -a.y, -a.z}; }
PT_HD f3 operator*(f3 a, f3 b) { return f3{a.x * b.x, a.y * b.y, a.z * b.z}; }
PT_HD f3 operator*(f3 a, float s) { return f3{a.x * s, a.y * s, a.z * s}; }
PT_HD f3 operator*(float s, f3 a) { return f3{a.x * s, a.y * s, a.z * s}; }
PT_HD f3 operator/(f3 a, float s) { return f3{a.x / s, a.y / s, a.z / s}; }
// glm::dot(vec3): (x*x' + y*y') + z*z'
PT_HD float dot3(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
PT_HD f3 cross3(f3 a, f3 b)
{
  return f3{a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y};
}
// glm::normalize(v) = v * inversesqrt(dot(v,v)), inversesqrt(x) = 1/sqrt(x)
PT_HD f3 normalize3(f3 v) { return v * (1.0f / sqrtf(dot3(v, v))); }
PT_HD float length3(f3 v) { return sqrtf(dot3(v, v)); }

// ---------------------------------------------------------------------- RNG
// Wang-style integer hash, the seed function of the reference (hash.cuh:4-14).
PT_HD uint32_t wang_hash(uint32_t a)
{
  a = (a + 0x7ed55d16u) + (a << 12);
  a = (a ^ 0xc761c23cu) ^ (a >> 19);
  a = (a + 0x165667b1u) + (a << 5);
  a = (a + 0xd3a2646cu) ^ (a << 9);
  a = (a + 0xfd7046c5u) + (a << 3);
  a = (a ^ 0xb55a4f09u) ^ (a >> 16);
  return a;
}

// thrust::default_random_engine == minstd_rand: x <- 48271 x mod (2^31 - 1).
// seed(s): x = s mod m, 0 -> 1 (linear_congruential_engine.inl:45-55).
PT_HD uint32_t minstd_seed(uint32_t s)
{
  uint32_t x = s % 2147483647u;
  return x == 0u ? 1u : x;
}
PT_HD uint32_t minstd_next(uint32_t x)
{
  // (48271 x) mod (2^31 - 1) without a 64-bit division: 2^31 == 1 (mod M), so the product
  // p = hi * 2^31 + lo reduces to hi + lo, minus M once if needed.  Exact.
  const uint64_t p = (uint64_t)x * 48271ull;
  uint32_t r = (uint32_t)(p & 0x7fffffffull) + (uint32_t)(p >> 31);
  return r >= 2147483647u ? r - 2147483647u : r;
}
// thrust::uniform_real_distribution<float>(0,1): float(x - min) / (1 + float(max - min))
// with min = 1, max = 2^31 - 2  ->  float(x - 1) / 2147483648.f
PT_HD float minstd_uniform(uint32_t& x)
{
  x = minstd_next(x);
  return (float)(x - 1u) / 2147483648.0f;
}
// x * a^n mod m by squaring == engine.discard(n)
PT_HD uint32_t minstd_discard(uint32_t x, uint32_t n)
{
  uint64_t a = 48271ull, r = x;
  while (n) {
    if (n & 1u) r = (r * a) % 2147483647ull;
    a = (a * a) % 2147483647ull;
    n >>= 1;
  }
  return (uint32_t)r;
}

// ------------------------------------------------------------------- camera
// Precomputed form of GPUCamera (camera.hpp:10-15) as a kernel parameter; one
// per launch, no __constant__ state shared between contexts.
struct DevCamera {
  float m[12];      // camera_matrix columns 0..2 (xyz each), then column 3 xyz = origin
  float vp_w, vp_h; // viewport_width, viewport_height (focal length 1)
  float fw_m1;      // float(W-1)
  float fh_m1;      // float(H-1)
  float fh;         // float(H)
  uint32_t width, height;
};

// generate_ray (ray_gen.cu:34-61). x,y are continuous pixel coordinates.
PT_HD void camera_ray(const DevCamera& c, float x, float y, f3& o, f3& d)
{
  const float u = x / c.fw_m1;
  const float v = (c.fh - y) / c.fh_m1;
  // lower_left_corner + u*horizontal + v*vertical - origin (zero terms are exact)
  const float lx = -(c.vp_w / 2.0f) + u * c.vp_w;
  const float ly = -(c.vp_h / 2.0f) + v * c.vp_h;
  const float lz = -1.0f;
  o = mk3(c.m[9], c.m[10], c.m[11]);
  // glm mat4*vec4 association: (m0*v0 + m1*v1) + (m2*v2 + m3*v3), v3 = 0
  f3 w;
  w.x = (c.m[0] * lx + c.m[3] * ly) + (c.m[6] * lz);
  w.y = (c.m[1] * lx + c.m[4] * ly) + (c.m[7] * lz);
  w.z = (c.m[2] * lx + c.m[5] * ly) + (c.m[8] * lz);
  d = normalize3(w);
}

// ------------------------------------------------------------ device scene
// World-space, instance-baked triangle: 48 B = three float4; leaves are
// contiguous runs of up to PT_LEAF_MAX triangles.
//   t0 = (v0.xyz, bits(triangle index in the input index buffer / 3))
//   t1 = (e1.xyz, bits(object index))       e1 = v1 - v0
//   t2 = (e2.xyz, bits(material index))     e2 = v2 - v0
//
// BVH2 node, 64 B (half a 128-B line), children's boxes stored in the parent:
//   n0 = (c0.min.x, c0.max.x, c0.min.y, c0.max.y)
//   n1 = (c1.min.x, c1.max.x, c1.min.y, c1.max.y)
//   n2 = (c0.min.z, c0.max.z, c1.min.z, c1.max.z)
//   n3 = bits(child0, child1, -, -);  child >= 0: inner node index
//        child < 0: leaf, ~child = (first_triangle << 3) | (count - 1)
//
// Compressed 8-wide node, 80 B = five uint4 (built by bvh_build.cpp, WideBuilder::emit):
//   w0 = (bits(p.x), bits(p.y), bits(p.z), ex' | ey'<<8 | ez'<<16 | imask<<24)
//        p = origin of the node's quantisation grid, e' = biased float exponent of (grid step / 256),
//        imask bit s = slot s holds an inner node
//   w1 = (first child node, first triangle, meta[0..3], meta[4..7])
//        meta: 0 = empty slot; inner: 0x20 | (24 + slot); leaf: unary(count) << 5 | triangle offset
//   w2 = (qlo.x[0..3], qlo.x[4..7], qlo.y[0..3], qlo.y[4..7])
//   w3 = (qlo.z[0..3], qlo.z[4..7], qhi.x[0..3], qhi.x[4..7])
//   w4 = (qhi.y[0..3], qhi.y[4..7], qhi.z[0..3], qhi.z[4..7])
// Inner children of a node are consecutive nodes (in slot order); the triangles of its leaf
// children are consecutive triangles (in slot order).
#define PT_LEAF_MAX 4
#define PT_STACK 64
#define PT_STACK8 32
#define PT_SENTINEL 0x7fffffff

// Sphere object, replicating ray_object_intersection_test's sphere branch
// (path_tracer.cu:87-98): world->object rows, object->world rows (affine).
struct DevSphere {
  float inv[12]; // row-major 3x4 of world->object
  float m[12];   // row-major 3x4 of object->world
  float cx, cy, cz, radius;
  uint32_t material;
  int32_t object;
  // Conservative world-space bound for the pre-reject in sphere_test (set only when the object
  // transform is a similarity, pre_ok = 1): centre, radius, |centre|_1.
  float wl1;
  uint32_t pre_ok;
  float wx, wy, wz, wr;
};

struct DevMaterial {
  float r, g, b; // albedo
  float param;   // fuzz or refraction index
  int32_t type;
  int32_t pad[3];
};

struct DevScene {
  const float4* nodes; // 4 float4 per node
  const uint4* nodes8; // compressed 8-wide tree, 5 uint4 per node, breadth-first (or null)
  const float4* tris;  // 3 float4 per triangle
  const DevSphere* spheres;
  const DevMaterial* materials;
  uint32_t n_spheres;
  uint32_t n_spheres_before; // spheres[0..n_before) precede the first mesh object
  uint32_t n_nodes;
  uint32_t n_nodes8; // 0 = traverse the binary tree
  uint32_t n_tris;
  float root_lo[3], root_hi[3]; // padded bounds of the whole mesh BVH (classification)
  // Sphere acceleration (scenes with many spheres; the reference scans its objects linearly,
  // path_tracer.cu:118): a binary tree in the mesh BVH's node format over the world bounds of each
  // sphere GROUP (the spheres before / after the first mesh object), leaves of <= 4 spheres
  // (~(first << 3 | count - 1), `first` indexing `spheres`).  Built only for groups of more than
  // PT_SPHERE_BVH_MIN rigidly placed spheres (there the closest hit does not depend on the test
  // order); root < 0 = test the group linearly.
  const float4* sph_nodes;
  int32_t sph_root_before, sph_root_after;
  // 32-byte companion of `nodes` for the traversal kernels (null = walk the 64-byte nodes): the
  // twelve child planes on a 16-bit grid over the root box, always rounded outwards by at least one
  // cell, plus the two child references:  x lo|hi<<16, y, z of child 0; x, y, z of child 1; ref 0;
  // ref 1.  A plane is q_org + q * q_cell.  Half the bytes per inner-node visit through the L1 data
  // pipe, which is what bounds traverse_kernel.
  const uint4* qnodes;
  float q_org[3], q_cell[3];
};
#define PT_SPHERE_BVH_MIN 32

// Path state, all indexed by path id = sample_in_pass * pixels + pixel (64 B/path):
//   ray[2*id]   = (origin.xyz, t_min)  ray[2*id+1] = (direction.xyz, t_max)
//                 interleaved so that one gathered ray is ONE aligned 32-byte sector
//   thr[id]     = (throughput.rgb, bits(rng state))
//   gbuf[id]    = (first-hit normal.xyz, first-hit t)
//   aux[id]     = intersection word: x = bits(t of the closest hit so far, else t_max),
//                 y = code: bit31 triangle (low bits = triangle slot), bit30 "spheres after the
//                 mesh still untested", else sphere index + 1, 0 = nothing hit;
//                 z = BVH node at which traversal starts (classification walks the top levels).
//                 Written by classification, refined by traverse, consumed by shade — it
//                 replaces the reference's 48-byte Intersection record (intersection.hpp:8-14).
struct PathState {
  float4* ray;
  float4* thr;
  float4* gbuf;
  uint4* aux;
};

// Compacted state of the paths parked for one traversal (PT_RNG_PIXEL_STREAM scheduler), indexed
// by slot = position in that iteration's list: ray[2*slot], ray[2*slot+1], thr[slot], aux[slot]
// as in PathState, pid[slot] = the path id the final contribution is written under.
struct ParkBuf {
  float4* ray;
  float4* thr;
  uint4* aux;
  uint32_t* pid;
};

// Ray binning (pt_params.sort_rays; the reference's commented-out sort by material_id,
// path_tracer.cu:439-446): traverse_kernel appends the slot of every ray it finishes to the list
// of its bin — 0 = no triangle hit (sky or a sphere), 1 + material type of the triangle hit
// otherwise — and the following chain launch walks the lists bin after bin, so that the lanes of
// a warp shade the same kind of surface.  list[b * cap + k], count[b]; null list = off.
#define PT_BINS 4
struct BinLists {
  uint32_t* list;
  uint32_t* count; // PT_BINS counters of this iteration
  uint32_t cap;
};

// Exact unsigned division by a run-time constant (Granlund & Montgomery 1994, round-up method):
// the bounce-0 index map divides three times per primary sample, ~20 instructions each as a
// plain `/`.  q = (t + ((n - t) >> sh1)) >> sh2 with t = umulhi(m, n).
struct FastDiv {
  uint32_t m, sh1, sh2;
};
inline FastDiv make_fastdiv(uint32_t d)
{
  uint32_t l = 0;
  while ((1ull << l) < d) ++l; // ceil(log2 d)
  FastDiv f;
  f.m = (uint32_t)(((1ull << 32) * ((1ull << l) - d)) / d + 1);
  f.sh1 = l < 1u ? l : 1u;
  f.sh2 = l > 1u ? l - 1u : 0u;
  return f;
}
#ifdef __CUDACC__
PT_D uint32_t fd_div(uint32_t n, const FastDiv& f)
{
  const uint32_t t = __umulhi(f.m, n);
  return (t + ((n - t) >> f.sh1)) >> f.sh2;
}
#endif

// One wavefront pass = `samples` consecutive iterations of every pixel.
struct PassParams {
  DevCamera cam;
  uint32_t pixels;
  uint32_t tiles_x, tiles_y; // 8x4-pixel warp tiles (tiles_y = tile rows of the rendered band)
  uint32_t tile_y0;          // first tile row of the band (row-band sharding), else 0
  uint32_t pixel_begin, pixel_end; // pixel range of the band (accumulate)
  uint32_t band_pixels;            // pixel_end - pixel_begin: path id = sample * band_pixels + (pixel - pixel_begin)
  FastDiv fd_per_sample, fd_tiles_x, fd_width; // divisors of the bounce-0 index map
  FastDiv fd_samples, fd_sbx;                  // (tile-major orders)
  uint32_t sbx;     // 8x8-tile super-blocks per row (order 2)
  uint32_t order;   // bounce-0 item order: 0 sample-major, 1 tile-major, 2 tile-major in 8x8-tile blocks
  uint32_t samples;
  uint32_t capacity; // path slots per buffer (bounds checks)
  uint32_t first_iteration;
  uint32_t rng_mode;
  uint32_t stream_state; // path-state streams use the evict-first cache policy (PT_STREAM_STATE)
};

} // namespace pt
