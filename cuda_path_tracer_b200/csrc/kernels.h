// kernels.h — host-callable launchers of the sm_100a kernels (kernels.cu).
#pragma once
#include "common.cuh"

namespace pt {

struct LaunchEnv {
  cudaStream_t stream;
  int sms; // multiprocessor count; persistent grids are sized sms * resident CTAs
};

// Per-pass device scratch owned by the context.
struct PassBuffers {
  PathState ps;
  ParkBuf park[2];     // compacted parked-path state, alternating between iterations (chain scheduler)
  uint32_t* queue[2];  // ping-pong path-id queues (all live paths of a bounce)
  uint32_t* tq;        // traverse queue: paths whose ray enters the mesh BVH this bounce
  uint32_t* counters;  // [0 .. max_depth]   live paths entering bounce b
  uint32_t* tcounters; // [0 .. max_depth]   length of the traverse queue of bounce b
  uint32_t* work;      // [0 .. max_depth]   persistent-kernel work-fetch cursors
  uint32_t* bin_list;  // PT_BINS x capacity slots (sort_rays only, else null)
  uint32_t* bin_counts; // [0 .. max_depth] x PT_BINS
  uint8_t* flags;      // stable-compaction alive flags (PT_RNG_SLOT_RESEED only)
  uint32_t* block_sums; // stable-compaction block counts / offsets
  unsigned long long* total_rays; // device-side ray counter
  uint32_t capacity;   // paths
};

// experiment switches read once from the environment (wavefront.cu)
int tunable_order();        // PT_ORDER: bounce-0 item order, see PassParams::order
int tunable_stream_state(); // PT_STREAM_STATE
// PT_FINISH: first wavefront bounce finish_kernel may replace (0 = never); PT_FINISH_RAYS: ... once
// at most this many rays are predicted to be parked there
int tunable_finish_after();
int tunable_finish_rays();
void launch_finish(const LaunchEnv& env, const DevScene& sc, const PassBuffers& pb, uint32_t iter, uint32_t max_depth);

// bounce 0: raygen + classification of the primary rays (fills the traverse queue).
void launch_raygen(const LaunchEnv& env, const DevScene& sc, const PassBuffers& pb,
                   const PassParams& pp, uint32_t n_items);
// BVH traversal of the traverse queue of `bounce` (length tcounters[bounce], device side).
void launch_traverse(const LaunchEnv& env, const DevScene& sc, const PassBuffers& pb,
                     const uint32_t* tq, uint32_t bounce);
// traversal of the compacted parked state of chain iteration `iter` (length tcounters[iter]).
void launch_traverse_parked(const LaunchEnv& env, const DevScene& sc, const PassBuffers& pb, uint32_t iter);
// PT_RNG_PIXEL_STREAM scheduler: iter 0 = raygen + in-register chains of simple bounces for
// the primary samples; iter >= 1 = the same for the paths traverse_kernel just served
// (queue[(iter-1)&1], tcounters[iter-1]).  Paths whose next ray needs the BVH are parked in
// queue[iter&1] / tcounters[iter].
void launch_chain(const LaunchEnv& env, const DevScene& sc, const PassBuffers& pb,
                  const PassParams& pp, uint32_t iter, uint32_t n_items_first, uint32_t max_depth);
// shade bounce b over queue q (implicit tile order when bounce == 0): hit rebuild, material,
// scatter, classification of the new ray; survivors -> queue q^1 / counters[b+1] (or flags for
// the stable compaction), BVH candidates -> tq / tcounters[b+1].
void launch_shade(const LaunchEnv& env, const DevScene& sc, const PassBuffers& pb,
                  const PassParams& pp, int q, uint32_t bounce, uint32_t n_items_first,
                  bool last_bounce);
// stable compaction (slot-reseed mode): flags -> queue q^1, counters[bounce+1].
void launch_stable_compact(const LaunchEnv& env, const PassBuffers& pb, const PassParams& pp,
                           int q, uint32_t bounce);
// per-pixel sum of the pass's samples into the running sums; also folds the
// bounce counters into total_rays.
void launch_accumulate(const LaunchEnv& env, const PassBuffers& pb, const PassParams& pp,
                       float4* sum_color, float4* sum_gbuf, uint32_t max_depth);

void launch_resolve_rgba8(const LaunchEnv& env, int kind, const float4* sum_color,
                          const float4* sum_gbuf, const float4* final_rgb, bool final_is_mean,
                          uint32_t pixels, uchar4* out);
void launch_export_f32(const LaunchEnv& env, int kind, const float4* sum_color,
                       const float4* sum_gbuf, const float4* final_rgb, bool final_is_mean,
                       uint32_t pixels, float* out);
void launch_import_frame(const LaunchEnv& env, const float* color3, const float* normal3,
                         const float* depth1, uint32_t pixels, float4* sum_color,
                         float4* sum_gbuf);

struct DenoiseParams {
  float c_phi, n_phi, p_phi;
  int clamp_fix;
};
// mean colour / normal / world position planes from the running sums
void launch_denoise_prepare(const LaunchEnv& env, const DevCamera& cam, const float4* sum_color,
                            const float4* sum_gbuf, float4* color0, float4* normal_depth,
                            float4* position, uint32_t row_lo, uint32_t row_hi);
void launch_atrous(const LaunchEnv& env, const DevCamera& cam, const DenoiseParams& dp,
                   const float4* color_in, const float4* normal_depth, const float4* position,
                   float4* color_out, int step_width, uint32_t row_lo, uint32_t row_hi);

// closest-hit parity hook
struct HitRecord {
  float t, px, py, pz, nx, ny, nz;
  uint32_t material, side;
  int32_t object, prim;
  uint32_t pad;
};
void launch_trace_batch(const LaunchEnv& env, const DevScene& sc, const float4* rays, uint32_t* work,
                        uint32_t n, HitRecord* out);

} // namespace pt
