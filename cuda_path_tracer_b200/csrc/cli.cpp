// cli.cpp — `cuda_pt [-o out.png] [--spp N] <scene>` as a library function, and
// the PNG writer.  Mirrors src/main.cpp:9-25, src/lib/configurations.cpp:7-45,
// src/cli/cli.cpp:62-115, src/lib/assets/assets.cpp:6-23, src/lib/image.cpp:9-22.
// Additive flags only: --max-depth, --filter-size (enables the denoiser),
// --device, --gpus (the reference hard-codes device 0, cli.cpp:71), --rng-mode, --stats-json.
#include "internal.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <optional>
#include <string>
#include <vector>
#include <zlib.h>

using namespace pt;

namespace {

void put_u32(std::vector<unsigned char>& v, uint32_t x)
{
  v.push_back((unsigned char)(x >> 24));
  v.push_back((unsigned char)(x >> 16));
  v.push_back((unsigned char)(x >> 8));
  v.push_back((unsigned char)x);
}

void put_chunk(std::vector<unsigned char>& out, const char* tag, const unsigned char* data, size_t n)
{
  put_u32(out, (uint32_t)n);
  const size_t start = out.size();
  out.insert(out.end(), tag, tag + 4);
  if (n) out.insert(out.end(), data, data + n);
  const uint32_t crc = (uint32_t)crc32(0L, out.data() + start, (uInt)(n + 4));
  put_u32(out, crc);
}

struct Stopwatch {
  using clock = std::chrono::steady_clock;
  clock::time_point start = clock::now(), last = clock::now();
  std::vector<std::pair<std::string, double>> entries;
  void end_stage(const char* name)
  {
    const auto now = clock::now();
    entries.emplace_back(name, std::chrono::duration<double>(now - last).count());
    last = now;
  }
  void report() const
  {
    std::printf("Elapsed time\n===========\n");
    for (auto& e : entries) std::printf("%s: %gs\n", e.first.c_str(), e.second);
    std::printf("Total: %gs\n\n", std::chrono::duration<double>(last - start).count());
  }
};

// locate_asset_path: walk up from cwd and keep the OUTERMOST "assets" directory
bool locate_asset_path(std::filesystem::path& result)
{
  namespace fs = std::filesystem;
  std::error_code ec;
  const fs::path cur = fs::current_path(ec);
  if (ec) return false;
  bool found = false;
  for (fs::path p = cur; p != cur.root_path(); p = p.parent_path()) {
    const fs::path a = p / "assets";
    if (fs::exists(a, ec) && fs::is_directory(a, ec)) {
      result = a;
      found = true;
    }
  }
  if (found) result = fs::absolute(result);
  return found;
}

void print_help()
{
  std::printf("A Path Tracer written in CUDA\nUsage:\n  cuda_pt [OPTION...] <filename>\n\n"
              "      --filename arg     The name of the scene file\n"
              "  -o, --output arg       Output path tracing result to a file\n"
              "  -h, --help             Print this message\n"
              "      --spp arg          Sample per pixel (if provided, this value will overwrite\n"
              "                         the setting in the scene file\n"
              "      --max-depth arg    Maximum bounces per path (default 50)\n"
              "      --filter-size arg  Run the A-Trous denoiser with this filter size\n"
              "      --rng-mode arg     0 = per-pixel stream (megakernel order), 1 = streaming re-seed\n"
              "      --device arg       CUDA device (default 0)\n"
              "      --gpus arg         Render on this many GPUs of the node (0 = all): sample ranges\n"
              "                         + one NCCL reduce; row bands when spp < gpus\n"
              "      --stats-json arg   Write run statistics to this file\n"
              "      --checkpoint arg   Save the progressive state (sums + iteration) to this file\n"
              "      --resume arg       Continue from a saved progressive state up to --spp\n"
              "      --all-meshes       Load every mesh a scene names (the reference loads the first)\n"
              "      --mesh-cache       Keep a binary copy (<obj>.b200mesh) of every parsed OBJ\n"
              "      --fast-build       Build the BVH on the GPU (LBVH): faster start, slower rays\n\n");
}

} // namespace

static int write_png_rgba8(const char* path, const void* rgba, uint32_t width, uint32_t height)
{
  if (!path || !rgba || !width || !height) return fail(PT_ERR_INVALID, "pt_write_png_rgba8: bad argument");
  const size_t stride = (size_t)width * 4;
  std::vector<unsigned char> raw((stride + 1) * height);
  for (uint32_t y = 0; y < height; ++y) {
    raw[y * (stride + 1)] = 0; // filter: none
    std::memcpy(&raw[y * (stride + 1) + 1], (const unsigned char*)rgba + y * stride, stride);
  }
  uLongf zlen = compressBound((uLong)raw.size());
  std::vector<unsigned char> z(zlen);
  if (compress2(z.data(), &zlen, raw.data(), (uLong)raw.size(), 3) != Z_OK)
    return fail(PT_ERR_IO, "zlib compression failed");
  std::vector<unsigned char> out;
  static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  out.insert(out.end(), sig, sig + 8);
  std::vector<unsigned char> ihdr;
  put_u32(ihdr, width);
  put_u32(ihdr, height);
  ihdr.push_back(8); // bit depth
  ihdr.push_back(6); // RGBA
  ihdr.push_back(0);
  ihdr.push_back(0);
  ihdr.push_back(0);
  put_chunk(out, "IHDR", ihdr.data(), ihdr.size());
  put_chunk(out, "IDAT", z.data(), zlen);
  put_chunk(out, "IEND", nullptr, 0);
  FILE* f = std::fopen(path, "wb");
  if (!f) return fail(PT_ERR_IO, std::string("Failed to write to image file ") + path);
  const size_t w = std::fwrite(out.data(), 1, out.size(), f);
  std::fclose(f);
  if (w != out.size()) return fail(PT_ERR_IO, std::string("Failed to write to image file ") + path);
  return PT_OK;
}

extern "C" int pt_write_png_rgba8(const char* path, const void* rgba, uint32_t width, uint32_t height)
{
  return guarded("pt_write_png_rgba8", [&] { return write_png_rgba8(path, rgba, width, height); });
}

static int cli_main(int argc, char** argv)
{
  std::optional<std::string> filename, output, stats_json, checkpoint, resume;
  std::optional<int> spp;
  int max_depth = 50, filter_size = 0, device = 0, rng_mode = 0, gpus = 1;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    auto value = [&](const char* name) -> const char* {
      if (i + 1 >= argc) {
        std::fprintf(stderr, "Option '%s' is missing an argument\n", name);
        return nullptr;
      }
      return argv[++i];
    };
    auto eq = [&](const char* longname, std::string& out) -> bool {
      const std::string pre = std::string(longname) + "=";
      if (a.rfind(pre, 0) == 0) {
        out = a.substr(pre.size());
        return true;
      }
      return false;
    };
    std::string v;
    if (a == "-h" || a == "--help") {
      print_help();
      return 0;
    } else if (a == "-o" || a == "--output") {
      const char* s = value("output");
      if (!s) return 1;
      output = s;
    } else if (eq("--output", v)) {
      output = v;
    } else if (a == "--spp" || eq("--spp", v)) {
      if (v.empty()) {
        const char* s = value("spp");
        if (!s) return 1;
        v = s;
      }
      spp = std::atoi(v.c_str());
    } else if (a == "--max-depth" || eq("--max-depth", v)) {
      if (v.empty()) { const char* s = value("max-depth"); if (!s) return 1; v = s; }
      max_depth = std::atoi(v.c_str());
    } else if (a == "--filter-size" || eq("--filter-size", v)) {
      if (v.empty()) { const char* s = value("filter-size"); if (!s) return 1; v = s; }
      filter_size = std::atoi(v.c_str());
    } else if (a == "--device" || eq("--device", v)) {
      if (v.empty()) { const char* s = value("device"); if (!s) return 1; v = s; }
      device = std::atoi(v.c_str());
    } else if (a == "--gpus" || eq("--gpus", v)) {
      if (v.empty()) { const char* s = value("gpus"); if (!s) return 1; v = s; }
      gpus = std::atoi(v.c_str());
      if (gpus < 0) {
        std::fprintf(stderr, "Option 'gpus' needs a count >= 0 (0 = every visible GPU)\n");
        return 1;
      }
    } else if (a == "--rng-mode" || eq("--rng-mode", v)) {
      if (v.empty()) { const char* s = value("rng-mode"); if (!s) return 1; v = s; }
      rng_mode = std::atoi(v.c_str());
    } else if (a == "--stats-json" || eq("--stats-json", v)) {
      if (v.empty()) { const char* s = value("stats-json"); if (!s) return 1; v = s; }
      stats_json = v;
    } else if (a == "--checkpoint" || eq("--checkpoint", v)) {
      if (v.empty()) { const char* s = value("checkpoint"); if (!s) return 1; v = s; }
      checkpoint = v;
    } else if (a == "--resume" || eq("--resume", v)) {
      if (v.empty()) { const char* s = value("resume"); if (!s) return 1; v = s; }
      resume = v;
    } else if (a == "--all-meshes") {
      setenv("PT_ALL_MESHES", "1", 1);
    } else if (a == "--fast-build") {
      setenv("PT_BUILD", "lbvh", 1);
    } else if (a == "--mesh-cache") {
      setenv("PT_MESH_CACHE", "1", 1);
    } else if (a == "--filename" || eq("--filename", v)) {
      if (v.empty()) { const char* s = value("filename"); if (!s) return 1; v = s; }
      filename = v;
    } else if (!a.empty() && a[0] == '-') {
      std::fprintf(stderr, "Option '%s' does not exist\n", a.c_str());
      return 1;
    } else if (!filename) {
      filename = a;
    }
  }
  if (!filename) {
    std::fprintf(stderr, "Usage: cuda_pt [options] <filename>\n");
    std::fprintf(stderr, "Run 'cuda_pt --help' for more information");
    return 1;
  }
  std::filesystem::path assets;
  if (!locate_asset_path(assets)) {
    std::fprintf(stderr, "Panic: Cannot find assets directory\n");
    return 1;
  }
  if (!output) {
    std::fprintf(stderr, "cuda_pt: the interactive viewer is not part of this build; pass --output <file.png>\n");
    return 1;
  }

  if (std::filesystem::path(*output).extension() != ".png") {
    // checked before anything is rendered: write_image_file (image.cpp:9-22) fails on these
    std::fprintf(stderr, "%s has an unrecognized extension\n", output->c_str());
    return 1;
  }

  Stopwatch sw;
  std::error_code ec;
  const std::filesystem::path scene_path = std::filesystem::canonical(assets / *filename, ec);
  if (ec) {
    std::fprintf(stderr, "Panic: cannot resolve %s\n", (assets / *filename).string().c_str());
    return 1;
  }
  if (scene_path.extension() != ".json") {
    std::fprintf(stderr, "Panic: Unsupported file extension %s!\n", scene_path.extension().string().c_str());
    return 1;
  }
  // one device: the scene goes straight to it; several: a group (one replica + context per
  // device, one host thread each, NCCL reduce / gather into the root context)
  const bool use_group = gpus != 1;
  pt_scene* scene = nullptr;
  pt_scene_file* sfile = nullptr;
  pt_group* group = nullptr;
  pt_scene_file_info finfo{};
  pt_scene_desc desc{};
  if (use_group) {
    if (pt_scene_file_read(scene_path.string().c_str(), &sfile, &desc, &finfo) != PT_OK) {
      std::fprintf(stderr, "Panic: %s\n", pt_last_error());
      return 1;
    }
  } else if (pt_scene_load_file(scene_path.string().c_str(), device, &scene, &finfo) != PT_OK) {
    std::fprintf(stderr, "Panic: %s\n", pt_last_error());
    return 1;
  }
  const int n_spp = spp ? *spp : finfo.spp;
  sw.end_stage("Scene loading");

  pt_params params;
  pt_params_default(&params);
  params.max_depth = max_depth;
  params.rng_mode = rng_mode;
  pt_ctx* ctx = nullptr;
  auto cleanup = [&] {
    if (group) {
      pt_group_destroy(group);
    } else {
      pt_ctx_destroy(ctx);
      pt_scene_destroy(scene);
    }
    pt_scene_file_free(sfile);
  };
  if (use_group) {
    if (pt_group_create(&desc, nullptr, gpus, (uint32_t)finfo.width, (uint32_t)finfo.height, &params, &group) != PT_OK) {
      std::fprintf(stderr, "Panic: %s\n", pt_last_error());
      cleanup();
      return 1;
    }
    ctx = pt_group_ctx(group, 0);
    gpus = pt_group_size(group);
  } else if (pt_ctx_create(scene, (uint32_t)finfo.width, (uint32_t)finfo.height, &params, nullptr, &ctx) != PT_OK) {
    std::fprintf(stderr, "Panic: %s\n", pt_last_error());
    cleanup();
    return 1;
  }
  pt_sync(ctx);
  std::printf("Start path tracing\n");
  std::printf("spp: %d\n", n_spp);
  std::printf("width: %d, height: %d\n", finfo.width, finfo.height);
  if (group) std::printf("gpus: %d\n", gpus);
  sw.end_stage("Initialization");

  int rc = 0;
  if (resume && pt_ctx_load_state(ctx, resume->c_str()) != PT_OK) rc = 1;
  if (!rc && group) {
    // a resumed root context only renders what is missing
    const int done = pt_ctx_iteration(ctx), todo = std::max(0, n_spp - done);
    if (done > 0) pt_ctx_set_sample_count(ctx, done);
    int r = PT_OK;
    if (todo >= gpus || done > 0) {
      r = pt_group_render(group, &finfo.camera, done, todo);
    } else {
      r = pt_group_render_bands(group, &finfo.camera, 0, todo); // fewer samples than GPUs: split the frame
    }
    if (r != PT_OK || pt_group_sync(group) != PT_OK) rc = 1;
  } else if (!rc) {
    pt_ctx_set_max_iterations(ctx, n_spp);
    // pt_render clips at max_iterations: a resumed context only renders what is missing
    if (pt_render(ctx, &finfo.camera, n_spp) != PT_OK || pt_sync(ctx) != PT_OK) rc = 1;
  }
  sw.end_stage("Path Tracing");
  if (!rc && checkpoint && pt_ctx_save_state(ctx, checkpoint->c_str()) != PT_OK) rc = 1;
  if (!rc && filter_size > 0) {
    pt_denoise_params dp;
    pt_denoise_params_default(&dp);
    dp.filter_size = filter_size;
    if (pt_denoise(ctx, &dp) != PT_OK || pt_sync(ctx) != PT_OK) rc = 1;
    sw.end_stage("Denoising");
  }
  std::vector<unsigned char> rgba((size_t)finfo.width * finfo.height * 4);
  if (!rc && pt_resolve_rgba8(ctx, PT_BUF_FINAL, rgba.data(), 0) != PT_OK) rc = 1;
  // a run that wrote no image is a failed run (write_image_file panics, image.cpp:19-21)
  if (!rc && pt_write_png_rgba8(output->c_str(), rgba.data(), (uint32_t)finfo.width, (uint32_t)finfo.height) != PT_OK)
    rc = 1;
  sw.end_stage("Write image file");
  if (rc) std::fprintf(stderr, "Panic: %s\n", pt_last_error());

  pt_stats st{};
  if (group)
    pt_group_get_stats(group, &st);
  else
    pt_get_stats(ctx, &st);
  if (!rc) {
    std::printf("Done path tracing %s!\n\n", filename->c_str());
    sw.report();
  }
  if (stats_json) {
    if (FILE* f = std::fopen(stats_json->c_str(), "w")) {
      double render_s = 0;
      for (auto& e : sw.entries)
        if (e.first == "Path Tracing") render_s = e.second;
      pt_scene_info si{};
      if (group)
        pt_group_scene_info(group, &si);
      else
        pt_scene_get_info(scene, &si);
      std::fprintf(f,
                   "{\"scene\": \"%s\", \"width\": %d, \"height\": %d, \"spp\": %d, \"max_depth\": %d, "
                   "\"rays\": %llu, \"render_s\": %.6f, \"mrays_per_s\": %.3f, \"triangles\": %llu, "
                   "\"bvh_nodes\": %llu, \"bvh_build_ms\": %.3f, \"launches\": %llu, \"gpus\": %d",
                   filename->c_str(), finfo.width, finfo.height, n_spp, max_depth,
                   (unsigned long long)st.rays, render_s,
                   render_s > 0 ? (double)st.rays / render_s * 1e-6 : 0.0,
                   (unsigned long long)si.n_world_triangles, (unsigned long long)si.n_bvh_nodes,
                   si.build_ms, (unsigned long long)st.kernel_launches, group ? gpus : 1);
      for (auto& e : sw.entries) std::fprintf(f, ", \"%s_s\": %.6f", e.first.c_str(), e.second);
      std::fprintf(f, "}\n");
      std::fclose(f);
    }
  }
  cleanup();
  return rc;
}

extern "C" int pt_cli_main(int argc, char** argv)
{
  // an exception anywhere in the command-line front end is a failed run, like the reference's panic()
  const int rc = guarded("cuda_pt", [&] { return cli_main(argc, argv); });
  if (rc != 0 && std::strstr(pt_last_error(), "cuda_pt: ") == pt_last_error())
    std::fprintf(stderr, "Panic: %s\n", pt_last_error());
  return rc == 0 ? 0 : 1;
}
