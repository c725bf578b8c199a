/*
 * oracle.h — CPU restatement of the reference's hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker.  The product
 * (cuda_path_tracer_b200/libb200pt.so) never links, imports or calls it.
 *
 * Every function restates one piece of LesleyLai/cuda-path-tracer and cites the
 * reference file:line it follows in oracle.c.  It is deliberately written in the
 * reference's own structure (object loop, object-space box tests, per-leaf world
 * transforms, un-culled stack traversal, 1-triangle leaves, running means) and
 * NOT in the structure of the CUDA product, so that agreement means something.
 *
 * Pinning: test_oracle_pinning.py checks this file against (a) the known-answer
 * vectors of the reference's own unit tests (test/aabb_test.cpp,
 * test/transform_test.cpp) and (b) the reference's own host code compiled from
 * /root/reference into oracle/_ref/libref_host.so (BVH builder, intersection
 * routines, transforms), and on the GPU box (c) the reference's CUDA kernels
 * compiled into oracle/_ref/libref_cuda.so.
 */
#ifndef PT_ORACLE_H
#define PT_ORACLE_H

#include "../include/b200pt.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_scene orc_scene;

/* BVHNode of the reference (accelerators/bvh.hpp:17-28), 32 bytes. */
typedef struct orc_bvh_node {
  float min[3];
  float max[3];
  uint32_t first_child_or_primitive;
  uint32_t primitive_count;
} orc_bvh_node;

/* scalar pieces (known-answer testable) */
uint32_t orc_hash(uint32_t a);
uint32_t orc_rng_seed(uint32_t s);
float orc_rng_uniform(uint32_t* state);
void orc_rng_discard(uint32_t* state, uint64_t n);
void orc_aabb_props(const float min3[3], const float max3[3], const float p[3], float extent[3],
                    int* max_extent, float* surface_area, float offset[3]);
/* inverse_transform_ray: ray8 in, ray8 out; m/inv column-major */
void orc_inverse_transform_ray(const float m[16], const float inv[16], const float ray8[8],
                               float out8[8]);
void orc_mat4_inverse(const float m[16], float out[16]);
int orc_ray_triangle(const float ray8[8], const float p0[3], const float p1[3], const float p2[3],
                     pt_hit* rec);
int orc_ray_sphere(const float ray8[8], const float center[3], float radius, pt_hit* rec);
int orc_ray_aabb(const float ray8[8], const float min3[3], const float max3[3]);
void orc_generate_ray(const pt_camera* cam, uint32_t w, uint32_t h, float x, float y, float out8[8]);

/* scene */
orc_scene* orc_scene_create(const pt_scene_desc* desc);
void orc_scene_destroy(orc_scene* s);
uint32_t orc_scene_bvh_size(const orc_scene* s);
const orc_bvh_node* orc_scene_bvh(const orc_scene* s);
void orc_scene_object_aabb(const orc_scene* s, uint32_t object, float min3[3], float max3[3]);

/* closest hit of a ray batch == ray_scene_intersection_test.
 * mode 0: reference traversal (BVH, stack, un-culled); mode 1: brute force over all triangles. */
void orc_trace_batch(const orc_scene* s, const float* rays8, uint64_t n, pt_hit* out, int mode);

/* megakernel-mode render (one RNG stream per pixel): iterations [first, first+n) folded into the
 * running means color3/normal3/depth1 exactly like final_gather. rays_out may be NULL. */
void orc_render_megakernel(const orc_scene* s, const pt_camera* cam, uint32_t w, uint32_t h,
                           int first_iteration, int n_iterations, int max_bounces, float* color3,
                           float* normal3, float* depth1, uint64_t* rays_out);
/* streaming-mode render (compacted-slot re-seeding, stable partition). */
void orc_render_streaming(const orc_scene* s, const pt_camera* cam, uint32_t w, uint32_t h,
                          int first_iteration, int n_iterations, int max_bounces, float* color3,
                          float* normal3, float* depth1, uint64_t* rays_out);

/* A-Trous denoiser. tainted (may be NULL) marks pixels whose value depends, directly or through
 * an earlier iteration, on a read past the end of the buffers (undefined in the reference). */
void orc_denoise(uint32_t w, uint32_t h, const pt_camera* cam, const float* color3,
                 const float* normal3, const float* depth1, int filter_size, float color_weight,
                 float normal_weight, float position_weight, int clamp_fix, float* out3,
                 uint8_t* tainted);

/* preview kernels: kind 0/1 colour, 2 normal, 3 depth */
void orc_tonemap(int kind, uint32_t n_pixels, const float* src, uint8_t* rgba);

int orc_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
