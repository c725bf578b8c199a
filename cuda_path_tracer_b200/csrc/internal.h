// internal.h — host-side structures behind the opaque C-ABI handles.
#pragma once
#include "../../include/b200pt.h"
#include "bvh_build.h"
#include "kernels.h"

#include <exception>
#include <new>
#include <string>
#include <vector>

namespace pt {

void set_error(const std::string& msg);
int fail(int code, const std::string& msg);
int cuda_fail(cudaError_t e, const char* what);

// No C++ exception crosses the C ABI: host allocations of multi-gigabyte scenes can throw.
template <typename F> int guarded(const char* what, F&& body)
{
  try {
    return body();
  } catch (const std::bad_alloc&) {
    return fail(PT_ERR_NOMEM, std::string(what) + ": out of host memory");
  } catch (const std::exception& e) {
    return fail(PT_ERR_INVALID, std::string(what) + ": " + e.what());
  } catch (...) {
    return fail(PT_ERR_INVALID, std::string(what) + ": unknown failure");
  }
}

#define PT_CUDA(call)                                                                            \
  do {                                                                                           \
    cudaError_t e__ = (call);                                                                    \
    if (e__ != cudaSuccess) return ::pt::cuda_fail(e__, #call);                                  \
  } while (0)

// host math mirroring the glm calls the reference makes (column-major mat4)
struct Mat4 {
  float m[16];
};
Mat4 mat4_identity();
Mat4 mat4_mul(const Mat4& a, const Mat4& b);
Mat4 mat4_translate(float x, float y, float z);
Mat4 mat4_scale(float x, float y, float z);
Mat4 mat4_rotate(float angle_rad, float ax, float ay, float az);
Mat4 mat4_inverse(const Mat4& a);
bool mat4_decompose_trs(const Mat4& a, float pos[3], float quat_wxyz[4]);
void mat4_from_camera(const pt_camera& cam, float out12[12]);
DevCamera make_dev_camera(const pt_camera& cam, uint32_t w, uint32_t h);

// host half of scene creation (context.cpp): built once, uploaded to one or several devices
struct SceneBuild {
  std::vector<DevSphere> spheres; // spheres preceding the first mesh object first
  uint32_t n_spheres_before = 0;
  std::vector<uint32_t> mesh_objects;
  std::vector<DevMaterial> mats;
  FlatBVH bvh;
  uint64_t n_world = 0;
  double build_ms = 0.0;
  bool wide = false, host_built = false;
  uint64_t content_hash = 0; // fingerprint of the description (progressive-state files)
  std::vector<float> sph_nodes; // sphere-group trees (16 floats per node), see DevScene
  int sph_root_before = -1, sph_root_after = -1;
};
int scene_prepare(const pt_scene_desc* desc, bool host_build, SceneBuild& sb);
int scene_upload(const pt_scene_desc* desc, const SceneBuild& sb, DeviceLBVH* dl, int device, pt_scene** out);

// parsed scene file (scene_io.cpp)
struct SceneFile {
  std::vector<float> positions;
  std::vector<uint32_t> indices;
  std::vector<uint64_t> mesh_first_index; // empty = one mesh (reference behaviour)
  std::vector<pt_object> objects;
  std::vector<pt_sphere> spheres;
  std::vector<pt_material> materials;
  pt_camera camera;
  int width = 0, height = 0, spp = 1;
};
int load_scene_file(const char* json_path, SceneFile& out);
int load_obj_file(const char* path, std::vector<float>& positions, std::vector<uint32_t>& indices);

} // namespace pt

struct pt_scene_file {
  pt::SceneFile sf;
};

struct pt_scene {
  int device = 0;
  pt::DevScene dev{};
  void* d_nodes = nullptr;
  void* d_tris = nullptr;
  void* d_nodes8 = nullptr;
  void* d_qnodes = nullptr; // 32-byte quantised companion of d_nodes (DevScene::qnodes), or null
  void* d_spheres = nullptr;
  void* d_sph_nodes = nullptr;
  void* d_materials = nullptr;
  pt_scene_info info{};
  uint64_t content_hash = 0;
};

struct pt_ctx {
  const pt_scene* scene = nullptr;
  uint32_t width = 0, height = 0, pixels = 0;
  pt_params params{};
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int sms = 148;
  int iteration = 0;
  bool usable = true; // false after a failed resize: no frame buffers
  // which iterations the running sums hold: [range_first, range_first + iteration) while
  // range_contiguous (progressive-state files record it; pt_render continues after it)
  int range_first = 0;
  bool range_contiguous = true;
  bool camera_locked = false; // sums came from a state file: rendering on needs the same camera
  uint32_t samples_per_pass = 1;
  uint32_t row_begin = 0, row_end = 0; // rendered / denoised band (row-band sharding); whole frame by default

  // Second lane of the pass (PT_LANES=2, pixel-stream mode): an internal child context with its
  // own stream and path-state buffers that renders the lower half of the band into the SAME sums
  // while this context renders the upper half, so that the sparse tail launches of one half overlap
  // the other half's work.  Each of the two holds buffers for half the frame's rows (rows_cap).
  pt_ctx* lane2 = nullptr;
  bool is_lane = false;
  uint32_t rows_cap = 0; // rows the path-state buffers are sized for (0 = the whole frame)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;

  pt::PassBuffers pb{};
  void* d_state = nullptr;    // one slab for the six PathState planes
  void* d_counters = nullptr; // counters + work cursors (one memset per pass)
  size_t counters_bytes = 0;
  uint32_t* h_counts = nullptr; // pinned look-behind copies of the bounce counters
  std::vector<cudaEvent_t> bounce_events;

  // Parked-list lengths of the last pass whose counts have reached the host (h_counts, copied
  // asynchronously): the next pass uses them to predict how many rays are left after each bounce
  // and hands the tail to finish_kernel at the first bounce where few enough remain.
  std::vector<uint32_t> hist_counts;
  uint32_t hist_known = 0;
  uint64_t hist_paths = 0;
  uint32_t pend_known = 0; // ... of the pass in flight
  uint64_t pend_paths = 0;
  int pend_event = -1;     // bounce event after which h_counts[0 .. pend_known) are complete

  bool own_sums = true;
  float4* d_sums = nullptr; // [0,pixels) colour sums + count, [pixels,2*pixels) normal+depth sums
  // denoiser planes (allocated on first use)
  float4* d_dn[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  float4* final_rgb = nullptr; // last denoise output, or nullptr => colour mean
  uchar4* d_rgba = nullptr;
  float* d_export = nullptr;

  pt_camera last_camera{};
  bool have_camera = false;

  pt_stats stats{};
  std::vector<cudaEvent_t> prof_events; // profile=1: pairs, tagged
  std::vector<cudaEvent_t> prof_pool;   // recycled events
  std::vector<int> prof_tags;
};
