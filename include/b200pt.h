/*
 * b200pt.h — C ABI of the B200-native wavefront path tracer.
 *
 * This is the drop-in boundary for the hot path of LesleyLai/cuda-path-tracer
 * (raygen -> BVH traversal -> triangle/sphere intersection -> shading/scatter ->
 * path compaction -> accumulate, plus the Edge-Avoiding A-Trous denoiser).
 * The reference has no FFI layer: its seam is the C++ class `PathTracer`
 * (src/lib/path_tracer.hpp:60-99) fed by `SceneDescription::build_scene()`
 * (src/lib/scene_description.hpp:48).  Every entry point below names the
 * reference member it replaces.  Plain pointers and sizes only; no C++ or
 * torch types cross this boundary; no entry point calls exit() or throws
 * (the reference's panic()/CUDA_CHECK exit, src/lib/prelude.cpp:5-10).
 *
 * Conventions
 *   - matrices are 4x4 column-major floats (glm::mat4 memory order);
 *   - quaternions are (w, x, y, z) (glm::quat constructor order, camera.hpp:19);
 *   - image buffers are row-major, pixel index = x + y*width, row 0 on top
 *     (src/lib/cuda_utils/indices.cuh:20-26);
 *   - every function returns a pt_status; pt_last_error() gives the message.
 */
#ifndef B200PT_H
#define B200PT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define PT_API
#else
#define PT_API __attribute__((visibility("default")))
#endif

typedef enum pt_status {
  PT_OK = 0,
  PT_ERR_INVALID = 1, /* bad argument / unsupported scene */
  PT_ERR_CUDA = 2,    /* a CUDA runtime call failed */
  PT_ERR_IO = 3,      /* file could not be opened / written */
  PT_ERR_PARSE = 4,   /* scene JSON / OBJ grammar error */
  PT_ERR_NOMEM = 5
} pt_status;

/* Material::Type order of the reference (src/lib/material.hpp:20). */
typedef enum pt_material_type {
  PT_MAT_DIFFUSE = 0,
  PT_MAT_METAL = 1,
  PT_MAT_DIELECTRIC = 2
} pt_material_type;

/* ObjectType order of the reference (src/lib/scene.hpp:12). */
typedef enum pt_object_type { PT_OBJ_SPHERE = 0, PT_OBJ_MESH = 1 } pt_object_type;

/* GPUMethod of the reference (src/lib/path_tracer.hpp:58) names the two RNG
 * disciplines; both are run here by the same wavefront kernels.
 *   PT_RNG_PIXEL_STREAM  one minstd stream per (pixel, iteration), continuous
 *                        over bounces == reference megakernel mode
 *                        (path_tracer.cu:239-264).  Default.
 *   PT_RNG_SLOT_RESEED   re-seeded every bounce from the *compacted slot index*
 *                        and discard(bounce) == reference streaming mode
 *                        (path_tracer.cu:300-301); uses a stable compaction so
 *                        slots equal thrust::stable_partition's. */
typedef enum pt_rng_mode { PT_RNG_PIXEL_STREAM = 0, PT_RNG_SLOT_RESEED = 1 } pt_rng_mode;

/* DisplayBufferType of the reference (src/lib/path_tracer.hpp:19) + denoised. */
typedef enum pt_buffer_kind {
  PT_BUF_FINAL = 0, /* last path_trace / denoise result (path_trace_result_buffer_) */
  PT_BUF_COLOR = 1,
  PT_BUF_NORMAL = 2,
  PT_BUF_DEPTH = 3,
  PT_BUF_DENOISED = 4
} pt_buffer_kind;

typedef struct pt_material {
  int32_t type; /* pt_material_type */
  float albedo[3];
  float fuzz;             /* metal */
  float refraction_index; /* dielectric */
} pt_material;

typedef struct pt_sphere {
  float center[3];
  float radius;
} pt_sphere;

/* One entry per scene object == GPUObject (src/lib/scene.hpp:14-20) plus the
 * material-table index the reference keeps in object_material_indices. */
typedef struct pt_object {
  int32_t type;        /* pt_object_type */
  uint32_t prim_index; /* sphere: index into spheres[]; mesh: index of the mesh when
                          pt_scene_desc.n_meshes > 0, else ignored (one mesh) */
  uint32_t material;   /* index into materials[] */
  float m[16];         /* object -> world */
  float inv[16];       /* world -> object */
} pt_object;

/* Flat scene description == what SceneDescription::build_scene() uploads
 * (src/lib/scene_description.cpp:12-117).  All arrays are host memory and are
 * copied during pt_scene_create; the caller may free them afterwards. */
typedef struct pt_scene_desc {
  const float* positions; /* 3*n_vertices */
  uint64_t n_vertices;
  const uint32_t* indices; /* n_indices = 3*triangles */
  uint64_t n_indices;
  const pt_object* objects;
  uint32_t n_objects;
  const pt_sphere* spheres;
  uint32_t n_spheres;
  const pt_material* materials;
  uint32_t n_materials;
  /* Extension (the reference uploads only its alphabetically-first mesh and every mesh object
   * instances it, scene_description.cpp:95): several meshes in the one index buffer.
   * n_meshes == 0: one mesh covering all indices (reference behaviour, prim_index ignored).
   * Otherwise mesh k is indices[mesh_first_index[k] .. mesh_first_index[k+1]) — n_meshes + 1
   * entries, multiples of 3 — and a mesh object's prim_index selects its mesh. */
  uint32_t n_meshes;
  const uint64_t* mesh_first_index;
} pt_scene_desc;

/* Camera (src/lib/camera.hpp:17-23). vfov in radians. */
typedef struct pt_camera {
  float position[3];
  float rotation[4]; /* w,x,y,z */
  float vfov;
} pt_camera;

typedef struct pt_params {
  int32_t max_depth;      /* reference: static constexpr max_bounces = 50 (path_tracer.cu:27) */
  int32_t rng_mode;       /* pt_rng_mode */
  int32_t max_iterations; /* PathTracer::max_iterations; <=0 means unlimited */
  int32_t samples_per_pass; /* 0 = auto: batch several iterations into one wavefront */
  int32_t profile;        /* 1 = time every kernel with CUDA events (pt_get_stats) */
  int32_t sort_rays;      /* 1 = bin the traversed rays by what they hit (miss / material type)
                             before shading: the reference's commented-out sort by material_id
                             (path_tracer.cu:439-446).  Off by default: measured slower. */
  int32_t lanes;          /* 0 = default (2; PT_LANES overrides), 1 = one pass at a time, 2 = the band's
                             two halves as two concurrent half-passes on two streams (the sparse tail
                             launches of one half overlap the other half's work; same frame) */
  int32_t reserved[1];
} pt_params;

/* EdgeAvoidingATrousDenoiser fields (denoising/edge_avoiding_a_trous_denoiser.hpp:9-12). */
typedef struct pt_denoise_params {
  int32_t filter_size;   /* default 10 -> steps 1,2,4,8 */
  float color_weight;    /* 0.45 */
  float normal_weight;   /* 0.30 */
  float position_weight; /* 0.25 */
  int32_t clamp_fix;     /* 0 = reference tap clamp [0,W]x[0,H] (aliasing kept, reads past
                            the last row are replaced by the last row); 1 = clamp to W-1,H-1 */
  int32_t reserved[3];
} pt_denoise_params;

/* Closest-hit record of pt_trace_batch: the fields of Intersection
 * (src/lib/intersection.hpp:8-14) plus primitive ids for the parity tests. */
typedef struct pt_hit {
  float t; /* < 0: miss */
  float point[3];
  float normal[3];
  uint32_t material;
  uint32_t side;   /* 0 front, 1 back */
  int32_t object;  /* object index, -1 on miss */
  int32_t prim;    /* triangle index in the input index buffer /3, -1 for spheres */
  uint32_t pad;
} pt_hit;

typedef struct pt_stats {
  uint64_t rays;          /* sum over passes and bounces of live paths entering extend */
  uint64_t samples;       /* pixel-samples rendered */
  uint32_t iterations;    /* PathTracer::iteration() */
  uint32_t passes;
  uint64_t kernel_launches;
  double ms_raygen_extend0; /* raygen + classification kernel; profile=1 only */
  double ms_extend;         /* BVH traversal kernel, summed over launches */
  double ms_shade;          /* shade + scatter + classification + compaction kernel */
  double ms_compact;
  double ms_accumulate;
  double ms_denoise;
  double ms_resolve;
  uint64_t n_extend_launches;
  uint64_t n_shade_launches;
  uint32_t max_bounce_reached;
  uint32_t reserved;
  uint64_t rays_traversed; /* subset of `rays` that entered the BVH traversal queue */
} pt_stats;

typedef struct pt_scene_info {
  uint64_t n_triangles;       /* input triangles */
  uint64_t n_world_triangles; /* after baking every mesh instance to world space */
  uint64_t n_bvh_nodes;
  uint32_t bvh_depth;
  uint32_t n_objects, n_spheres, n_materials;
  double build_ms;  /* host BVH build */
  double upload_ms; /* H2D */
  uint64_t device_bytes;
  uint64_t n_bvh8_nodes; /* compressed 8-wide tree (0 = not built) */
  uint32_t bvh8_depth;
  uint32_t device_build; /* 1 = the tree was built on the device (PT_BUILD=lbvh) */
  uint64_t n_bvh_triangles; /* entries of the leaf-ordered triangle array (incl. a null triangle) */
} pt_scene_info;

/* Result of loading a reference scene file (assets/json_parser.cpp:174-224). */
typedef struct pt_scene_file_info {
  pt_camera camera;
  int32_t width, height;
  int32_t spp;
  double load_ms;
} pt_scene_file_info;

typedef struct pt_scene pt_scene;
typedef struct pt_ctx pt_ctx;

PT_API const char* pt_last_error(void);
PT_API int pt_version(void);

/* --- scene: replaces SceneDescription::build_scene + bvh_from_mesh
 *     (scene_description.cpp:12-117, accelerators/bvh.cpp:211-253) ------------- */
PT_API int pt_scene_create(const pt_scene_desc* desc, int device, pt_scene** out);
PT_API int pt_scene_destroy(pt_scene* scene);
PT_API int pt_scene_get_info(const pt_scene* scene, pt_scene_info* info);
/* Copies the device-resident binary tree back: n_bvh_nodes x 16 floats and n_bvh_triangles x 12
 * floats (either pointer may be NULL).  Test hook: a tree built on the device (PT_BUILD=lbvh)
 * is compared node for node with the host restatement of the same algorithm. */
PT_API int pt_scene_copy_bvh(const pt_scene* scene, float* nodes_out, float* tris_out);

/* Host-only half of pt_scene_create (no CUDA call): bake the mesh instances and build both
 * trees; pt_host_bvh_validate walks them and counts structural violations (a triangle that is
 * not referenced exactly once, content outside the stored/decoded child box, a child index out
 * of range).  Used by the CPU tests and the host build benchmark; == bvh_from_mesh
 * (accelerators/bvh.cpp:211-253) + our flattening. */
typedef struct pt_host_bvh pt_host_bvh;
PT_API int pt_host_bvh_build(const pt_scene_desc* desc, int wide, pt_host_bvh** out, pt_scene_info* info);
PT_API int pt_host_bvh_validate(const pt_host_bvh* bvh, uint64_t* violations);
PT_API int pt_host_bvh_arrays(const pt_host_bvh* bvh, const float** nodes, const uint32_t** nodes8,
                              const float** tris);
/* The 32-byte quantised nodes the traversal kernels read (DevScene::qnodes; pt_scene_create builds
 * them the same way for host-built trees up to 512 MB): n_bvh_nodes x 8 words = the twelve child
 * planes as 16-bit grid coordinates (x lo | hi << 16, y, z of child 0; x, y, z of child 1), always
 * at least one whole cell outside the exact plane, then the two child references of the 64-byte
 * node; a plane is org3[axis] + q * cell3[axis].  PT_ERR_INVALID when the tree's bounds are
 * degenerate (the kernels then walk the 64-byte nodes).  Test hook: the CPU suite checks the
 * containment and restates the device's de-quantising slab test. */
PT_API int pt_host_bvh_quantised(pt_host_bvh* bvh, const uint32_t** qnodes, float* org3, float* cell3);
/* Analysis tool: walks the host-built tree for n rays {o, t_min, d, t_max} in the device kernels'
 * visiting order and returns {inner-node visits, leaf visits, triangle tests, rays that hit,
 * deepest stack} in out5.  wide = 0: binary tree, 1: compressed 8-wide tree.  No image is
 * computed; the product path never calls it. */
PT_API int pt_host_bvh_trace_stats(const pt_host_bvh* bvh, const float* rays8, uint64_t n_rays, int wide,
                                   uint64_t* out5);
PT_API int pt_host_bvh_free(pt_host_bvh* bvh);
/* Host-only check of the non-mesh half of scene preparation (no CUDA call): the sphere tables are
 * built exactly as pt_scene_create builds them and the sphere-group trees (more than 32 rigidly
 * placed spheres per group) are walked.  out4 = {spheres, tree nodes, violations, trees built};
 * a violation is a sphere that is not referenced by exactly one leaf or whose world-space bound
 * sticks out of any box on its path from the root.  *hash_out = the description fingerprint that
 * progressive-state files carry. */
PT_API int pt_host_scene_check(const pt_scene_desc* desc, uint64_t* out4, uint64_t* hash_out);

/* replaces read_scene/scene_from_json/load_obj (assets/scene_parser.cpp:6-22,
 * assets/json_parser.cpp:174-224, assets/model_loader.cpp:11-44). */
PT_API int pt_scene_load_file(const char* json_path, int device, pt_scene** out,
                              pt_scene_file_info* info);

/* Host-only half of the above (no CUDA call): parse a scene file into the flat
 * description.  *desc points into memory owned by the returned handle; free it
 * with pt_scene_file_free.  == scene_from_json (assets/json_parser.cpp:174-224). */
typedef struct pt_scene_file pt_scene_file;
PT_API int pt_scene_file_read(const char* json_path, pt_scene_file** out, pt_scene_desc* desc,
                              pt_scene_file_info* info);
PT_API int pt_scene_file_free(pt_scene_file* f);

/* --- integrator context: replaces class PathTracer -------------------------- */
PT_API void pt_params_default(pt_params* p);
PT_API void pt_denoise_params_default(pt_denoise_params* p);

/* PathTracer::create_buffers (path_tracer.cu:559-564). `stream` may be NULL
 * (the context creates its own non-blocking stream) or a cudaStream_t. */
PT_API int pt_ctx_create(const pt_scene* scene, uint32_t width, uint32_t height,
                         const pt_params* params, void* stream, pt_ctx** out);
PT_API int pt_ctx_destroy(pt_ctx* ctx);
/* PathTracer::resize_image (path_tracer.cu:527-545): reallocates and restarts. */
PT_API int pt_ctx_resize(pt_ctx* ctx, uint32_t width, uint32_t height);
/* Row-band sharding of ONE frame over several GPUs (the 1-spp interactive path has no sample
 * range to split): the context renders, accumulates and denoises rows [row_begin, row_end) only.
 * row_begin and row_end are multiples of 4 (row_end may also be the frame height).  The sums
 * buffer stays frame-sized, so neighbours' halo rows are received straight into place;
 * pt_denoise then needs pt_denoise_halo_rows() valid rows beyond either end of the band
 * (2 * (1 + 2 + ... + last step): 62 for filter_size 16), clipped at the frame edge. */
PT_API int pt_ctx_set_rows(pt_ctx* ctx, uint32_t row_begin, uint32_t row_end);
PT_API int pt_denoise_halo_rows(const pt_denoise_params* params, uint32_t* rows);
/* PathTracer::restart (path_tracer.cu:522-525). */
PT_API int pt_ctx_restart(pt_ctx* ctx);
/* PathTracer::iteration(). */
PT_API int pt_ctx_iteration(const pt_ctx* ctx);
PT_API int pt_ctx_set_max_iterations(pt_ctx* ctx, int max_iterations);
PT_API int pt_ctx_set_stream(pt_ctx* ctx, void* stream);

/* PathTracer::path_trace (path_tracer.cu:389-477): one sample per pixel per
 * call, no-op once iteration() >= max_iterations.  Asynchronous. */
PT_API int pt_path_trace(pt_ctx* ctx, const pt_camera* camera);
/* The CLI loop `for i<spp: path_trace` (cli.cpp:96-99) as one call: renders
 * iterations [iteration(), iteration()+n) batching several per wavefront. */
PT_API int pt_render(pt_ctx* ctx, const pt_camera* camera, int n_iterations);
/* Same, for an explicit iteration range (multi-GPU sample-range sharding):
 * accumulates iterations [first, first+n) into the running sums. */
PT_API int pt_render_range(pt_ctx* ctx, const pt_camera* camera, int first_iteration,
                           int n_iterations);
PT_API int pt_sync(pt_ctx* ctx);

/* PathTracer::denoise -> EdgeAvoidingATrousDenoiser::denoise
 * (path_tracer.cu:479-485, denoising/...denoiser.cu:88-116). */
PT_API int pt_denoise(pt_ctx* ctx, const pt_denoise_params* params);

/* PathTracer::send_to_preview (path_tracer.cu:487-520): tonemap `kind` to
 * RGBA8 into dst (width*height*4 bytes; device pointer if dst_is_device). */
PT_API int pt_resolve_rgba8(pt_ctx* ctx, int kind, void* dst, int dst_is_device);
/* Raw float view of a frame buffer (means, like the reference's buffers):
 * COLOR/NORMAL/FINAL/DENOISED -> 3 floats per pixel, DEPTH -> 1 float. */
PT_API int pt_download_f32(pt_ctx* ctx, int kind, float* dst_host);

/* Multi-GPU sample-range sharding: the running SUMS (float4 colour.rgb+count,
 * float4 normal.xyz+depth per pixel; 32 B/pixel) live in one device buffer that
 * a collective can reduce in place.  pt_ctx_sums returns its device pointer and
 * size in floats; pt_ctx_set_sample_count tells the context how many samples
 * the (reduced) sums now hold. */
PT_API int pt_ctx_sums(pt_ctx* ctx, void** device_ptr, uint64_t* n_floats);
PT_API int pt_ctx_set_sample_count(pt_ctx* ctx, int n_samples);
/* Make the context accumulate into a caller-owned device buffer of
 * width*height*8 floats (e.g. a torch tensor that NCCL reduces in place);
 * NULL returns to a context-owned buffer.  The caller zeroes its buffer. */
PT_API int pt_ctx_bind_sums(pt_ctx* ctx, void* device_ptr);
/* Load externally produced colour / normal / depth means as the frame state
 * (denoiser parity tests feed identical inputs to both implementations). */
PT_API int pt_ctx_upload_frame(pt_ctx* ctx, const float* color3, const float* normal3,
                               const float* depth1, const pt_camera* camera);

/* Progressive state on disk (checkpoint + resume of long progressive renders): the
 * running sums, WHICH iterations they hold ([first, first + count), so a shard that rendered
 * pt_render_range(first = K, ...) resumes at K + count and never re-uses seeds) and a fingerprint
 * of what they were rendered from (scene description, max_depth, rng_mode, camera).  load refuses
 * a state of another resolution, scene or parameter set; afterwards pt_render continues where the
 * saved render stopped (same result as an uninterrupted run) and refuses another camera until
 * pt_ctx_restart. */
PT_API int pt_ctx_save_state(pt_ctx* ctx, const char* path);
PT_API int pt_ctx_load_state(pt_ctx* ctx, const char* path);

PT_API int pt_get_stats(pt_ctx* ctx, pt_stats* stats);
PT_API int pt_reset_stats(pt_ctx* ctx);

/* --- multi-GPU group: one process, N devices ---------------------------------------------
 * The reference renders on device 0 only (cuda::init_CUDA / cudaSetDevice(0), src/cli/cli.cpp:71).
 * A group holds one scene replica and one integrator context per device (the BVH is built once
 * on the host and uploaded to each), drives them from one host thread per device and combines
 * their frames over NVLink with NCCL (bound at run time; a one-device group never loads it).
 *   devices == NULL: devices 0 .. n_devices-1;  n_devices <= 0: every visible device.
 * After a render the frame lives in the ROOT context, pt_group_ctx(g, 0): denoise, resolve,
 * download and save it through the ordinary pt_* calls on that context. */
typedef struct pt_group pt_group;
PT_API int pt_group_create(const pt_scene_desc* desc, const int* devices, int n_devices, uint32_t width,
                           uint32_t height, const pt_params* params, pt_group** out);
PT_API int pt_group_destroy(pt_group* group);
PT_API int pt_group_size(const pt_group* group);
PT_API pt_ctx* pt_group_ctx(pt_group* group, int i);
PT_API int pt_group_device(const pt_group* group, int i);
PT_API int pt_group_scene_info(const pt_group* group, pt_scene_info* info);
/* PathTracer::restart for every member; also leaves band mode. */
PT_API int pt_group_restart(pt_group* group);
/* Samples per pixel the root's frame holds == PathTracer::iteration(). */
PT_API int pt_group_iteration(const pt_group* group);
/* Sample-range sharding (progressive / high spp): iterations [first, first+n) are split into
 * one contiguous range per device, rendered with the seeds a single GPU would use, and summed
 * into the root's running sums with ONE ncclReduce of 32 B/pixel.  Equal to pt_render_range on
 * one device up to float re-association of the sums.  Asynchronous like pt_render. */
PT_API int pt_group_render(pt_group* group, const pt_camera* camera, int first_iteration, int n_iterations);
/* Row-band sharding of ONE frame (the 1-spp interactive path): device i renders rows
 * [r_i, r_i+1) (multiples of 4 rows) of every iteration in the range; the bands are gathered in
 * place into the root's frame with ncclSend/ncclRecv.  Identical to the single-device frame. */
PT_API int pt_group_render_bands(pt_group* group, const pt_camera* camera, int first_iteration,
                                 int n_iterations);
PT_API int pt_group_sync(pt_group* group);
/* Sums over the members (rays, launches, kernel times). */
PT_API int pt_group_get_stats(pt_group* group, pt_stats* stats);

/* --- parity hook: closest hit of a ray batch == ray_scene_intersection_test
 *     (path_tracer.cu:110-128).  rays8 = n x {ox,oy,oz,t_min,dx,dy,dz,t_max}
 *     (Ray, src/lib/ray.hpp:8-20), host memory in, host memory out. */
PT_API int pt_trace_batch(const pt_scene* scene, const float* rays8, uint64_t n,
                          pt_hit* hits_out);

/* --- image output: replaces write_image_file (src/lib/image.cpp:9-22) ------- */
PT_API int pt_write_png_rgba8(const char* path, const void* rgba8_host, uint32_t width,
                              uint32_t height);

/* The whole `cuda_pt [-o out.png] [--spp N] <scene>` run (src/main.cpp:9-25,
 * src/cli/cli.cpp:62-115) as a function; the cuda_pt executable is a thin
 * wrapper around it. Returns the process exit code. */
PT_API int pt_cli_main(int argc, char** argv);

#ifdef __cplusplus
}
#endif
#endif /* B200PT_H */
