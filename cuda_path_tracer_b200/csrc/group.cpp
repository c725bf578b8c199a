// group.cpp — one process, several B200s: the multi-GPU integrator behind the C ABI.
//
// The reference renders on device 0 only (src/cli/cli.cpp:71).  A pt_group owns one scene replica
// and one integrator context per device (the tree is built ONCE on the host and uploaded to every
// device), drives them from one host thread per device, and combines their frames over NVLink with
// NCCL (loaded at run time; a single-device group needs no NCCL at all):
//
//   pt_group_render        sample-range sharding: device i renders its share of the iteration
//                          range with the seeds a single GPU would use, into its own running
//                          SUMS; one ncclReduce(sum) of 32 B/pixel leaves the frame on device 0.
//                          Sums (not the reference's running means, path_tracer.cu:203-219) are
//                          what makes the shards associative.
//   pt_group_render_bands  row-band sharding of ONE frame (1 spp has no sample range to split):
//                          device i renders rows [r_i, r_{i+1}) of every iteration; the bands are
//                          gathered into device 0's frame-sized sums with ncclSend/ncclRecv.
//
// Denoise / tonemap / download then run on the root context (pt_group_ctx(g, 0)) through the
// ordinary entry points.
#include "internal.h"

#include <dlfcn.h>
#include <nccl.h>

#include <atomic>
#include <chrono>
#include <cstring>
#include <mutex>
#include <thread>

using namespace pt;

namespace {

// NCCL is bound with dlopen so that libb200pt.so has no link-time dependency on it: single-GPU
// users never load it, and a process that already carries NCCL (e.g. under PyTorch) shares it.
struct NcclApi {
  void* handle = nullptr;
  decltype(&ncclCommInitAll) CommInitAll = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclReduce) Reduce = nullptr;
  decltype(&ncclSend) Send = nullptr;
  decltype(&ncclRecv) Recv = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  std::string error;

  bool load()
  {
    if (handle) return true;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (handle) break;
    }
    if (!handle) {
      error = std::string("NCCL not found: ") + (dlerror() ? dlerror() : "dlopen failed");
      return false;
    }
#define PT_NCCL_SYM(field, sym)                                                                    \
  field = reinterpret_cast<decltype(field)>(dlsym(handle, #sym));                                  \
  if (!field) {                                                                                    \
    error = "NCCL symbol missing: " #sym;                                                          \
    return false;                                                                                  \
  }
    PT_NCCL_SYM(CommInitAll, ncclCommInitAll)
    PT_NCCL_SYM(CommDestroy, ncclCommDestroy)
    PT_NCCL_SYM(Reduce, ncclReduce)
    PT_NCCL_SYM(Send, ncclSend)
    PT_NCCL_SYM(Recv, ncclRecv)
    PT_NCCL_SYM(GroupStart, ncclGroupStart)
    PT_NCCL_SYM(GroupEnd, ncclGroupEnd)
    PT_NCCL_SYM(GetErrorString, ncclGetErrorString)
#undef PT_NCCL_SYM
    return true;
  }
};

NcclApi& nccl()
{
  static NcclApi api;
  return api;
}
std::mutex g_nccl_mutex;

double now_ms()
{
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

} // namespace

struct pt_group {
  std::vector<int> devices;
  std::vector<pt_scene*> scenes;
  std::vector<pt_ctx*> ctxs;
  std::vector<ncclComm_t> comms;
  uint32_t width = 0, height = 0;
  int samples = 0; // samples per pixel the root's sums hold
  bool banded = false;
  double last_host_ms = 0.0;
};

// Runs fn(i) on one host thread per device; returns the first failure (its message becomes this
// thread's pt_last_error: the error string is thread-local).
template <typename F> static int for_each_device(pt_group* g, F&& fn)
{
  const int n = (int)g->devices.size();
  std::vector<int> rc(n, PT_OK);
  std::vector<std::string> msg(n);
  auto body = [&](int i) {
    rc[i] = guarded("pt_group worker", [&] { return fn(i); });
    if (rc[i] != PT_OK) msg[i] = pt_last_error();
  };
  if (n == 1) {
    body(0);
  } else {
    std::vector<std::thread> th;
    th.reserve(n);
    for (int i = 0; i < n; ++i) th.emplace_back(body, i);
    for (auto& t : th) t.join();
  }
  for (int i = 0; i < n; ++i)
    if (rc[i] != PT_OK) return fail(rc[i], "device " + std::to_string(g->devices[i]) + ": " + msg[i]);
  return PT_OK;
}

static int nccl_fail(ncclResult_t r, const char* what)
{
  return fail(PT_ERR_CUDA, std::string("NCCL error: ") + nccl().GetErrorString(r) + " in " + what);
}
#define PT_NCCL(call)                                                                              \
  do {                                                                                             \
    ncclResult_t r__ = (call);                                                                     \
    if (r__ != ncclSuccess) return nccl_fail(r__, #call);                                          \
  } while (0)

static int pt_group_create_impl(const pt_scene_desc* desc, const int* devices, int n_devices, uint32_t width,
                                uint32_t height, const pt_params* params, pt_group** out)
{
  if (!desc || !out) return fail(PT_ERR_INVALID, "pt_group_create: null argument");
  *out = nullptr;
  int dev_count = 0;
  PT_CUDA(cudaGetDeviceCount(&dev_count));
  if (n_devices <= 0) n_devices = dev_count; // "all devices"
  if (n_devices < 1 || n_devices > dev_count) return fail(PT_ERR_INVALID, "pt_group_create: no such CUDA device");
  std::vector<int> devs(n_devices);
  for (int i = 0; i < n_devices; ++i) {
    devs[i] = devices ? devices[i] : i;
    if (devs[i] < 0 || devs[i] >= dev_count) return fail(PT_ERR_INVALID, "pt_group_create: no such CUDA device");
    for (int j = 0; j < i; ++j)
      if (devs[j] == devs[i]) return fail(PT_ERR_INVALID, "pt_group_create: a device is listed twice");
  }
  if (params && params->rng_mode == PT_RNG_SLOT_RESEED && n_devices > 1)
    return fail(PT_ERR_INVALID, "pt_group_create: the slot-reseed RNG discipline renders whole frames on one device");

  // the tree is built once on the host; every device gets its own copy (the scene is read-only
  // and a replica is a few hundred MB of 180 GB even at 10 M triangles)
  SceneBuild sb;
  int rc = scene_prepare(desc, true, sb);
  if (rc != PT_OK) return rc;

  auto* g = new pt_group();
  g->devices = devs;
  g->scenes.assign(n_devices, nullptr);
  g->ctxs.assign(n_devices, nullptr);
  g->width = width;
  g->height = height;
  rc = for_each_device(g, [&](int i) {
    int r = scene_upload(desc, sb, nullptr, devs[i], &g->scenes[i]);
    if (r != PT_OK) return r;
    return pt_ctx_create(g->scenes[i], width, height, params, nullptr, &g->ctxs[i]);
  });
  if (rc == PT_OK && n_devices > 1) {
    std::lock_guard<std::mutex> lock(g_nccl_mutex);
    if (!nccl().load()) {
      rc = fail(PT_ERR_INVALID, "pt_group_create: " + nccl().error);
    } else {
      g->comms.assign(n_devices, nullptr);
      const ncclResult_t r = nccl().CommInitAll(g->comms.data(), n_devices, devs.data());
      if (r != ncclSuccess) {
        g->comms.clear();
        rc = nccl_fail(r, "ncclCommInitAll");
      }
    }
  }
  if (rc != PT_OK) {
    const std::string keep = pt_last_error();
    pt_group_destroy(g);
    return fail(rc, keep);
  }
  *out = g;
  return PT_OK;
}

extern "C" {

int pt_group_create(const pt_scene_desc* desc, const int* devices, int n_devices, uint32_t width, uint32_t height,
                    const pt_params* params, pt_group** out)
{
  return guarded("pt_group_create",
                 [&] { return pt_group_create_impl(desc, devices, n_devices, width, height, params, out); });
}

int pt_group_destroy(pt_group* g)
{
  if (!g) return PT_OK;
  for (size_t i = 0; i < g->ctxs.size(); ++i)
    if (g->ctxs[i]) pt_sync(g->ctxs[i]);
  for (size_t i = 0; i < g->comms.size(); ++i) {
    if (!g->comms[i]) continue;
    cudaSetDevice(g->devices[i]);
    nccl().CommDestroy(g->comms[i]);
  }
  for (size_t i = 0; i < g->ctxs.size(); ++i) {
    pt_ctx_destroy(g->ctxs[i]);
    pt_scene_destroy(g->scenes[i]);
  }
  delete g;
  return PT_OK;
}

int pt_group_size(const pt_group* g) { return g ? (int)g->devices.size() : 0; }

pt_ctx* pt_group_ctx(pt_group* g, int i)
{
  if (!g || i < 0 || i >= (int)g->ctxs.size()) return nullptr;
  return g->ctxs[i];
}

int pt_group_device(const pt_group* g, int i)
{
  if (!g || i < 0 || i >= (int)g->devices.size()) return -1;
  return g->devices[i];
}

int pt_group_scene_info(const pt_group* g, pt_scene_info* info)
{
  if (!g || !info || g->scenes.empty()) return fail(PT_ERR_INVALID, "pt_group_scene_info: null argument");
  return pt_scene_get_info(g->scenes[0], info);
}

int pt_group_restart(pt_group* g)
{
  if (!g) return fail(PT_ERR_INVALID, "pt_group_restart: null group");
  for (size_t i = 0; i < g->ctxs.size(); ++i) {
    int rc = pt_ctx_set_rows(g->ctxs[i], 0, g->height);
    if (rc == PT_OK) rc = pt_ctx_restart(g->ctxs[i]);
    if (rc != PT_OK) return rc;
  }
  g->samples = 0;
  g->banded = false;
  return PT_OK;
}

int pt_group_iteration(const pt_group* g) { return g ? g->samples : -1; }

int pt_group_render(pt_group* g, const pt_camera* cam, int first_iteration, int n_iterations)
{
  if (!g || !cam) return fail(PT_ERR_INVALID, "pt_group_render: null argument");
  if (first_iteration < 0 || n_iterations < 0) return fail(PT_ERR_INVALID, "negative iteration range");
  if (g->banded) return fail(PT_ERR_INVALID, "pt_group_render after pt_group_render_bands: call pt_group_restart first");
  if (n_iterations == 0) return PT_OK;
  const int n = (int)g->devices.size();
  const double t0 = now_ms();
  const int base = n_iterations / n, extra = n_iterations % n;
  const int rc = for_each_device(g, [&](int i) {
    pt_ctx* c = g->ctxs[i];
    PT_CUDA(cudaSetDevice(g->devices[i]));
    // the root keeps accumulating (progressive rendering); the others hold this call's share only
    if (i != 0) {
      const int r = pt_ctx_restart(c);
      if (r != PT_OK) return r;
    } else {
      c->range_contiguous = true; // its own share is one more piece of the group's range (set below)
      if (c->iteration == 0) c->range_first = first_iteration;
    }
    const int share = base + (i < extra ? 1 : 0);
    const int first = first_iteration + i * base + std::min(i, extra);
    if (share > 0) {
      const int r = pt_render_range(c, cam, first, share);
      if (r != PT_OK) return r;
    }
    if (n > 1) {
      void* sums = nullptr;
      uint64_t n_floats = 0;
      pt_ctx_sums(c, &sums, &n_floats);
      // in place on the root; NVLink/NVSwitch carries 32 B per pixel per device
      PT_NCCL(nccl().Reduce(sums, sums, (size_t)n_floats, ncclFloat, ncclSum, 0, g->comms[i], c->stream));
    }
    return (int)PT_OK;
  });
  if (rc != PT_OK) return rc;
  // the reduced sums hold the WHOLE range, whatever share the root rendered itself
  pt_ctx* root = g->ctxs[0];
  if (g->samples == 0) {
    root->range_first = first_iteration;
    root->range_contiguous = true;
  } else if (first_iteration != root->range_first + g->samples) {
    root->range_contiguous = false;
  }
  g->samples += n_iterations;
  pt_ctx_set_sample_count(root, g->samples);
  root->final_rgb = nullptr;
  g->last_host_ms = now_ms() - t0;
  return PT_OK;
}

int pt_group_render_bands(pt_group* g, const pt_camera* cam, int first_iteration, int n_iterations)
{
  if (!g || !cam) return fail(PT_ERR_INVALID, "pt_group_render_bands: null argument");
  if (first_iteration < 0 || n_iterations < 0) return fail(PT_ERR_INVALID, "negative iteration range");
  if (!g->banded && g->samples != 0)
    return fail(PT_ERR_INVALID, "pt_group_render_bands after pt_group_render: call pt_group_restart first");
  if (n_iterations == 0) return PT_OK;
  const int n = (int)g->devices.size();
  const double t0 = now_ms();
  // bands of whole 8x4 primary-ray tiles: multiples of 4 rows
  const uint32_t tile_rows = (g->height + 3) / 4;
  std::vector<uint32_t> row(n + 1);
  for (int i = 0; i <= n; ++i) row[i] = std::min<uint32_t>(g->height, (uint32_t)((uint64_t)tile_rows * i / n) * 4u);
  row[n] = g->height;
  const size_t pixels = (size_t)g->width * g->height;
  const int rc = for_each_device(g, [&](int i) {
    pt_ctx* c = g->ctxs[i];
    PT_CUDA(cudaSetDevice(g->devices[i]));
    if (row[i] < row[i + 1]) {
      int r = pt_ctx_set_rows(c, row[i], row[i + 1]);
      if (r == PT_OK) r = pt_render_range(c, cam, first_iteration, n_iterations);
      if (r != PT_OK) return r;
    }
    if (n > 1) {
      // gather: every band's two planes (colour sums + count, normal + depth sums) land in place
      // in the root's frame-sized buffer
      float* sums = (float*)c->d_sums;
      PT_NCCL(nccl().GroupStart());
      for (int j = 1; j < n; ++j) {
        if (row[j] >= row[j + 1]) continue;
        const size_t off = (size_t)row[j] * g->width * 4, cnt = (size_t)(row[j + 1] - row[j]) * g->width * 4;
        if (i == 0) {
          PT_NCCL(nccl().Recv(sums + off, cnt, ncclFloat, j, g->comms[i], c->stream));
          PT_NCCL(nccl().Recv(sums + pixels * 4 + off, cnt, ncclFloat, j, g->comms[i], c->stream));
        } else if (i == j) {
          PT_NCCL(nccl().Send(sums + off, cnt, ncclFloat, 0, g->comms[i], c->stream));
          PT_NCCL(nccl().Send(sums + pixels * 4 + off, cnt, ncclFloat, 0, g->comms[i], c->stream));
        }
      }
      PT_NCCL(nccl().GroupEnd());
    }
    return (int)PT_OK;
  });
  if (rc != PT_OK) return rc;
  g->banded = true;
  g->samples += n_iterations;
  // the root denoises / resolves the WHOLE frame it now holds
  pt_ctx_set_rows(g->ctxs[0], 0, g->height);
  pt_ctx_set_sample_count(g->ctxs[0], g->samples);
  g->ctxs[0]->final_rgb = nullptr;
  g->last_host_ms = now_ms() - t0;
  return PT_OK;
}

int pt_group_sync(pt_group* g)
{
  if (!g) return fail(PT_ERR_INVALID, "pt_group_sync: null group");
  for (size_t i = 0; i < g->ctxs.size(); ++i) {
    const int rc = pt_sync(g->ctxs[i]);
    if (rc != PT_OK) return rc;
  }
  return PT_OK;
}

int pt_group_get_stats(pt_group* g, pt_stats* out)
{
  if (!g || !out) return fail(PT_ERR_INVALID, "pt_group_get_stats: null argument");
  pt_stats total{};
  for (size_t i = 0; i < g->ctxs.size(); ++i) {
    pt_stats s{};
    const int rc = pt_get_stats(g->ctxs[i], &s);
    if (rc != PT_OK) return rc;
    total.rays += s.rays;
    total.rays_traversed += s.rays_traversed;
    total.samples += s.samples;
    total.passes += s.passes;
    total.kernel_launches += s.kernel_launches;
    total.ms_raygen_extend0 += s.ms_raygen_extend0;
    total.ms_extend += s.ms_extend;
    total.ms_shade += s.ms_shade;
    total.ms_compact += s.ms_compact;
    total.ms_accumulate += s.ms_accumulate;
    total.ms_denoise += s.ms_denoise;
    total.ms_resolve += s.ms_resolve;
    total.n_extend_launches += s.n_extend_launches;
    total.n_shade_launches += s.n_shade_launches;
    total.max_bounce_reached = std::max(total.max_bounce_reached, s.max_bounce_reached);
  }
  total.iterations = (uint32_t)g->samples;
  *out = total;
  return PT_OK;
}

} // extern "C"
