// scene_io.cpp — reader for the reference's scene-JSON + OBJ input format.
//
// Grammar and semantics follow src/lib/assets/json_parser.cpp:40-224 (materials,
// surfaces, transform command lists, camera, sampler), scene_description.cpp
// (alphabetical material table :59-66, first-mesh-only :95) and
// assets/model_loader.cpp:11-44 (Assimp: first mesh, triangulated, de-indexed).
// The JSON reader below is a small self-contained recursive-descent parser;
// nlohmann/json and Assimp are not dependencies.
#include "internal.h"

#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <filesystem>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <map>
#include <memory>

namespace pt {
namespace {

// ------------------------------------------------------------- mini JSON
struct JValue;
using JPtr = std::shared_ptr<JValue>;
struct JValue {
  enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
  bool b = false;
  double num = 0.0;
  std::string str;
  std::vector<JPtr> arr;
  std::vector<std::pair<std::string, JPtr>> obj;
  const JValue* find(const std::string& k) const
  {
    for (auto& kv : obj)
      if (kv.first == k) return kv.second.get();
    return nullptr;
  }
};

struct JParser {
  const char* p;
  const char* end;
  std::string err;
  void ws()
  {
    while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) ++p;
  }
  bool fail(const std::string& m)
  {
    if (err.empty()) err = m;
    return false;
  }
  bool parse_string(std::string& out)
  {
    if (p >= end || *p != '"') return fail("expected string");
    ++p;
    while (p < end && *p != '"') {
      if (*p == '\\') {
        ++p;
        if (p >= end) return fail("bad escape");
        switch (*p) {
        case 'n': out += '\n'; break;
        case 't': out += '\t'; break;
        case 'r': out += '\r'; break;
        case 'b': out += '\b'; break;
        case 'f': out += '\f'; break;
        case 'u': {
          if (end - p < 5) return fail("bad \\u escape");
          unsigned cp = (unsigned)std::strtoul(std::string(p + 1, p + 5).c_str(), nullptr, 16);
          if (cp < 0x80) {
            out += (char)cp;
          } else if (cp < 0x800) {
            out += (char)(0xC0 | (cp >> 6));
            out += (char)(0x80 | (cp & 0x3F));
          } else {
            out += (char)(0xE0 | (cp >> 12));
            out += (char)(0x80 | ((cp >> 6) & 0x3F));
            out += (char)(0x80 | (cp & 0x3F));
          }
          p += 4;
          break;
        }
        default: out += *p;
        }
        ++p;
      } else {
        out += *p++;
      }
    }
    if (p >= end) return fail("unterminated string");
    ++p;
    return true;
  }
  int depth = 0; // nesting guard: the parser recurses once per '{' / '['
  struct DepthGuard {
    int& d;
    explicit DepthGuard(int& d_) : d(d_) { ++d; }
    ~DepthGuard() { --d; }
  };
  bool parse(JPtr& out)
  {
    const DepthGuard guard(depth);
    if (depth > 256) return fail("JSON nested deeper than 256 levels");
    ws();
    if (p >= end) return fail("unexpected end of JSON");
    out = std::make_shared<JValue>();
    const char c = *p;
    if (c == '{') {
      out->kind = JValue::Obj;
      ++p;
      ws();
      if (p < end && *p == '}') {
        ++p;
        return true;
      }
      for (;;) {
        ws();
        std::string key;
        if (!parse_string(key)) return false;
        ws();
        if (p >= end || *p != ':') return fail("expected ':'");
        ++p;
        JPtr v;
        if (!parse(v)) return false;
        out->obj.emplace_back(std::move(key), v);
        ws();
        if (p < end && *p == ',') {
          ++p;
          continue;
        }
        if (p < end && *p == '}') {
          ++p;
          return true;
        }
        return fail("expected ',' or '}'");
      }
    }
    if (c == '[') {
      out->kind = JValue::Arr;
      ++p;
      ws();
      if (p < end && *p == ']') {
        ++p;
        return true;
      }
      for (;;) {
        JPtr v;
        if (!parse(v)) return false;
        out->arr.push_back(v);
        ws();
        if (p < end && *p == ',') {
          ++p;
          continue;
        }
        if (p < end && *p == ']') {
          ++p;
          return true;
        }
        return fail("expected ',' or ']'");
      }
    }
    if (c == '"') {
      out->kind = JValue::Str;
      return parse_string(out->str);
    }
    if (!std::strncmp(p, "true", std::min<size_t>(4, end - p)) && end - p >= 4) {
      out->kind = JValue::Bool;
      out->b = true;
      p += 4;
      return true;
    }
    if (!std::strncmp(p, "false", std::min<size_t>(5, end - p)) && end - p >= 5) {
      out->kind = JValue::Bool;
      p += 5;
      return true;
    }
    if (!std::strncmp(p, "null", std::min<size_t>(4, end - p)) && end - p >= 4) {
      p += 4;
      return true;
    }
    char* e = nullptr;
    const double v = std::strtod(p, &e);
    if (e == p) return fail("unexpected character in JSON");
    out->kind = JValue::Num;
    out->num = v;
    p = e;
    return true;
  }
};

bool read_file(const std::string& path, std::string& out)
{
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) return false;
  std::fseek(f, 0, SEEK_END);
  const long n = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  out.resize(n > 0 ? (size_t)n : 0);
  const size_t got = n > 0 ? std::fread(&out[0], 1, (size_t)n, f) : 0;
  std::fclose(f);
  return got == out.size();
}

struct ParseError {
  std::string msg;
};

float num(const JValue* v, const char* what)
{
  if (!v || v->kind != JValue::Num) throw ParseError{std::string("Json Parser: expected number for ") + what};
  return (float)v->num;
}
void vec3(const JValue* v, float out[3])
{
  if (!v || v->kind != JValue::Arr || v->arr.size() != 3) throw ParseError{"Json Parser: vec3 need to be 3d"};
  for (int i = 0; i < 3; ++i) out[i] = num(v->arr[i].get(), "vec3 component");
}
std::string str(const JValue* v, const char* what)
{
  if (!v || v->kind != JValue::Str) throw ParseError{std::string("Json Parser: expected string for ") + what};
  return v->str;
}

// adl_serializer<glm::mat4>::from_json (json_parser.cpp:40-76)
Mat4 transform_command(const JValue& j)
{
  if (j.kind != JValue::Obj) throw ParseError{"Json parser: Unrecognized transform command"};
  if (const JValue* t = j.find("translate")) {
    float v[3];
    vec3(t, v);
    return mat4_translate(v[0], v[1], v[2]);
  }
  if (const JValue* s = j.find("scale")) {
    if (s->kind == JValue::Num) {
      const float f = (float)s->num;
      return mat4_scale(f, f, f);
    }
    float v[3];
    vec3(s, v);
    return mat4_scale(v[0], v[1], v[2]);
  }
  if (const JValue* r = j.find("rotate")) {
    const float angle = num(r, "rotate") * 0.01745329251994329576923690768489f; // glm::radians
    float ax[3];
    vec3(j.find("axis"), ax);
    return mat4_rotate(angle, ax[0], ax[1], ax[2]);
  }
  if (j.find("from") && j.find("at") && j.find("up")) {
    float from[3], at[3], up[3];
    vec3(j.find("from"), from);
    vec3(j.find("at"), at);
    vec3(j.find("up"), up);
    auto norm = [](float* v) {
      const float s = 1.0f / std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
      v[0] *= s, v[1] *= s, v[2] *= s;
    };
    auto cross = [](const float* a, const float* b, float* o) {
      o[0] = a[1] * b[2] - b[1] * a[2];
      o[1] = a[2] * b[0] - b[2] * a[0];
      o[2] = a[0] * b[1] - b[0] * a[1];
    };
    float dir[3] = {from[0] - at[0], from[1] - at[1], from[2] - at[2]};
    norm(dir);
    float left[3], new_up[3];
    cross(up, dir, left);
    norm(left);
    cross(dir, left, new_up);
    norm(new_up);
    Mat4 m = mat4_identity();
    for (int i = 0; i < 3; ++i) {
      m.m[0 + i] = left[i];
      m.m[4 + i] = new_up[i];
      m.m[8 + i] = dir[i];
      m.m[12 + i] = from[i];
    }
    return m;
  }
  throw ParseError{"Json parser: Unrecognized transform command"};
}

// adl_serializer<Transform>::from_json (json_parser.cpp:78-95): M = Mi * M
Mat4 read_transform(const JValue* j)
{
  if (!j) throw ParseError{"Json Parser: surface without transform"};
  Mat4 mat = mat4_identity();
  if (j->kind == JValue::Obj) {
    mat = transform_command(*j);
  } else if (j->kind == JValue::Arr) {
    for (auto& e : j->arr) mat = mat4_mul(transform_command(*e), mat);
  } else {
    throw ParseError{"Json Parser: Transform must be either an object or an array!"};
  }
  return mat;
}

} // namespace

// ------------------------------------------------------------------- OBJ
// Assimp semantics kept: first mesh only (faces up to the first o/g/usemtl
// statement that follows a face), polygons fan-triangulated, one vertex per
// face corner (positions has 3T entries, indices = 0,1,2,...).
// Binary side-car of a parsed OBJ (opt-in, PT_MESH_CACHE=1 / cuda_pt --mesh-cache): the
// 10-M-triangle OBJ of the stress config is ~1 GB of text; the cache is the two arrays verbatim,
// keyed by the source's size and modification time.
namespace {
struct MeshCacheHeader {
  char magic[8]; // "B200MESH"
  uint32_t version, reserved;
  uint64_t source_bytes;
  int64_t source_mtime;
  uint64_t n_floats, n_indices;
};

bool mesh_cache_enabled()
{
  const char* v = std::getenv("PT_MESH_CACHE");
  return v && std::atoi(v) != 0;
}

bool source_key(const char* path, uint64_t& bytes, int64_t& mtime)
{
  std::error_code ec;
  const auto sz = std::filesystem::file_size(path, ec);
  if (ec) return false;
  const auto tm = std::filesystem::last_write_time(path, ec);
  if (ec) return false;
  bytes = (uint64_t)sz;
  mtime = (int64_t)tm.time_since_epoch().count();
  return true;
}

bool read_mesh_cache(const char* path, std::vector<float>& positions, std::vector<uint32_t>& indices)
{
  uint64_t bytes;
  int64_t mtime;
  if (!source_key(path, bytes, mtime)) return false;
  FILE* f = std::fopen((std::string(path) + ".b200mesh").c_str(), "rb");
  if (!f) return false;
  MeshCacheHeader h{};
  bool ok = std::fread(&h, sizeof(h), 1, f) == 1 && !std::memcmp(h.magic, "B200MESH", 8) && h.version == 1 &&
            h.source_bytes == bytes && h.source_mtime == mtime && h.n_indices % 3 == 0 &&
            h.n_floats % 3 == 0 && h.n_floats < (1ull << 34) && h.n_indices < (1ull << 34);
  if (ok) {
    positions.resize(h.n_floats);
    indices.resize(h.n_indices);
    ok = std::fread(positions.data(), 4, h.n_floats, f) == h.n_floats &&
         std::fread(indices.data(), 4, h.n_indices, f) == h.n_indices;
    const uint64_t nv = h.n_floats / 3;
    for (size_t i = 0; ok && i < indices.size(); ++i) ok = indices[i] < nv;
  }
  std::fclose(f);
  if (!ok) {
    positions.clear();
    indices.clear();
  }
  return ok;
}

void write_mesh_cache(const char* path, const std::vector<float>& positions, const std::vector<uint32_t>& indices)
{
  MeshCacheHeader h{};
  std::memcpy(h.magic, "B200MESH", 8);
  h.version = 1;
  if (!source_key(path, h.source_bytes, h.source_mtime)) return;
  h.n_floats = positions.size();
  h.n_indices = indices.size();
  const std::string out = std::string(path) + ".b200mesh";
  FILE* f = std::fopen(out.c_str(), "wb");
  if (!f) return; // read-only asset directory: the cache is best effort
  const bool ok = std::fwrite(&h, sizeof(h), 1, f) == 1 &&
                  std::fwrite(positions.data(), 4, positions.size(), f) == positions.size() &&
                  std::fwrite(indices.data(), 4, indices.size(), f) == indices.size();
  std::fclose(f);
  if (!ok) std::remove(out.c_str());
}
} // namespace

static int parse_obj_file(const char* path, std::vector<float>& positions, std::vector<uint32_t>& indices);
struct TextView {
  const char* p;
  size_t n;
  const char* data() const { return p; }
  size_t size() const { return n; }
};
static int parse_obj_serial(const char* path, const TextView& text, std::vector<float>& positions,
                            std::vector<uint32_t>& indices);

int load_obj_file(const char* path, std::vector<float>& positions, std::vector<uint32_t>& indices)
{
  const bool cache = mesh_cache_enabled();
  if (cache && read_mesh_cache(path, positions, indices)) return PT_OK;
  const int rc = parse_obj_file(path, positions, indices);
  if (rc == PT_OK && cache) write_mesh_cache(path, positions, indices);
  return rc;
}

static int parse_obj_serial(const char* path, const TextView& text, std::vector<float>& positions,
                            std::vector<uint32_t>& indices)
{
  std::vector<float> verts;
  verts.reserve(text.size() / 24);
  positions.clear();
  indices.clear();
  const char* p = text.data();
  const char* end = p + text.size();
  bool have_faces = false, stop = false;
  std::vector<long long> corner;
  auto skip_sp = [&](const char*& q) {
    while (q < end && (*q == ' ' || *q == '\t')) ++q;
  };
  while (p < end && !stop) {
    const char* line_end = (const char*)std::memchr(p, '\n', end - p);
    if (!line_end) line_end = end;
    const char* q = p;
    skip_sp(q);
    if (q < line_end) {
      if (q[0] == 'v' && q + 1 < line_end && (q[1] == ' ' || q[1] == '\t')) {
        ++q;
        float v[3] = {0.f, 0.f, 0.f};
        for (int i = 0; i < 3; ++i) {
          skip_sp(q);
          if (q < line_end && *q == '+') ++q;
          auto r = std::from_chars(q, line_end, v[i]);
          if (r.ec != std::errc()) return fail(PT_ERR_PARSE, std::string("bad vertex in ") + path);
          q = r.ptr;
        }
        verts.push_back(v[0]);
        verts.push_back(v[1]);
        verts.push_back(v[2]);
      } else if (q[0] == 'f' && q + 1 < line_end && (q[1] == ' ' || q[1] == '\t')) {
        ++q;
        corner.clear();
        const long long nv = (long long)(verts.size() / 3);
        for (;;) {
          skip_sp(q);
          if (q >= line_end || *q == '\r' || *q == '#') break;
          long long idx = 0;
          auto r = std::from_chars(q, line_end, idx);
          if (r.ec != std::errc()) return fail(PT_ERR_PARSE, std::string("bad face in ") + path);
          q = r.ptr;
          while (q < line_end && *q != ' ' && *q != '\t' && *q != '\r') ++q; // skip /vt/vn
          if (idx < 0) idx = nv + idx; else idx -= 1;
          if (idx < 0 || idx >= nv) return fail(PT_ERR_PARSE, std::string("face index out of range in ") + path);
          corner.push_back(idx);
        }
        if (corner.size() >= 3) {
          have_faces = true;
          for (size_t k = 1; k + 1 < corner.size(); ++k) {
            const long long tri[3] = {corner[0], corner[k], corner[k + 1]};
            for (int c = 0; c < 3; ++c) {
              indices.push_back((uint32_t)(positions.size() / 3));
              positions.push_back(verts[3 * tri[c] + 0]);
              positions.push_back(verts[3 * tri[c] + 1]);
              positions.push_back(verts[3 * tri[c] + 2]);
            }
          }
        }
      } else if (have_faces && (q[0] == 'o' || q[0] == 'g' || !std::strncmp(q, "usemtl", 6))) {
        stop = true; // aiScene::mMeshes[0] only (model_loader.cpp:22)
      }
    }
    p = line_end + 1;
  }
  if (indices.empty()) return fail(PT_ERR_PARSE, std::string("Unable to load ") + path);
  return PT_OK;
}

// ---- parallel OBJ parser: the 10-M-triangle stress mesh is ~1.5 GB of text, 7 s on one thread.
// The text is cut into chunks at line ends; three parallel sweeps (count, vertices, faces) with
// prefix sums in between reproduce the serial reader exactly: same vertex numbering (negative
// indices count back from the vertices seen so far), same triangle order, same cut-off at the first
// o/g/usemtl statement that follows a face.
namespace {
struct ObjChunk {
  const char* begin;
  const char* end;
  uint64_t n_v = 0, n_tri = 0;
  const char* first_face = nullptr;       // first f line in the chunk
  const char* first_group = nullptr;      // first o/g/usemtl line in the chunk
  const char* group_after_face = nullptr; // first o/g/usemtl line after the chunk's first face
  uint64_t v_base = 0, tri_base = 0;
  bool bad_vertex = false, bad_face = false, range_error = false;
};

inline const char* obj_skip_sp(const char* q, const char* end)
{
  while (q < end && (*q == ' ' || *q == '\t')) ++q;
  return q;
}
enum ObjLine { OBJ_OTHER, OBJ_V, OBJ_F, OBJ_GROUP };
inline ObjLine obj_classify(const char* q, const char* line_end)
{
  if (q >= line_end) return OBJ_OTHER;
  if (q[0] == 'v' && q + 1 < line_end && (q[1] == ' ' || q[1] == '\t')) return OBJ_V;
  if (q[0] == 'f' && q + 1 < line_end && (q[1] == ' ' || q[1] == '\t')) return OBJ_F;
  if (q[0] == 'o' || q[0] == 'g' || ((size_t)(line_end - q) >= 6 && !std::strncmp(q, "usemtl", 6))) return OBJ_GROUP;
  return OBJ_OTHER;
}
// corners of a face line starting after the 'f'; returns false on a malformed index
inline bool obj_face_corners(const char* q, const char* line_end, long long* out, int cap, int& n)
{
  n = 0;
  for (;;) {
    q = obj_skip_sp(q, line_end);
    if (q >= line_end || *q == '\r' || *q == '#') break;
    long long idx = 0;
    auto r = std::from_chars(q, line_end, idx);
    if (r.ec != std::errc()) return false;
    q = r.ptr;
    while (q < line_end && *q != ' ' && *q != '\t' && *q != '\r') ++q; // skip /vt/vn
    if (n < cap) out[n] = idx;
    ++n;
  }
  return true;
}

void obj_count(ObjChunk& c)
{
  c.n_v = c.n_tri = 0;
  c.bad_face = false;
  c.first_face = c.first_group = c.group_after_face = nullptr;
  const char* p = c.begin;
  while (p < c.end) {
    const char* line_end = (const char*)std::memchr(p, '\n', c.end - p);
    if (!line_end) line_end = c.end;
    const char* q = obj_skip_sp(p, c.end);
    switch (obj_classify(q, line_end)) {
    case OBJ_V: ++c.n_v; break;
    case OBJ_F: {
      long long tmp[4];
      int n = 0;
      if (!obj_face_corners(q + 1, line_end, tmp, 0, n)) c.bad_face = true;
      if (n >= 3) {
        c.n_tri += (uint64_t)(n - 2);
        if (!c.first_face) c.first_face = p;
      }
      break;
    }
    case OBJ_GROUP:
      if (!c.first_group) c.first_group = p;
      if (c.first_face && !c.group_after_face) c.group_after_face = p;
      break;
    default: break;
    }
    p = line_end + 1;
  }
}
} // namespace

static int parse_obj_file(const char* path, std::vector<float>& positions, std::vector<uint32_t>& indices)
{
  // the file is mapped, not copied: a 1.5 GB read into a cleared std::string costs a second
  struct Mapping {
    void* p = MAP_FAILED;
    size_t n = 0;
    int fd = -1;
    ~Mapping()
    {
      if (p != MAP_FAILED) munmap(p, n);
      if (fd >= 0) close(fd);
    }
  } map;
  map.fd = open(path, O_RDONLY);
  struct stat st;
  if (map.fd < 0 || fstat(map.fd, &st) != 0 || !S_ISREG(st.st_mode))
    return fail(PT_ERR_IO, std::string("Unable to load ") + path);
  map.n = (size_t)st.st_size;
  if (map.n == 0) return fail(PT_ERR_PARSE, std::string("Unable to load ") + path);
  map.p = mmap(nullptr, map.n, PROT_READ, MAP_PRIVATE, map.fd, 0);
  if (map.p == MAP_FAILED) return fail(PT_ERR_IO, std::string("Unable to load ") + path);
  const TextView text{static_cast<const char*>(map.p), map.n};
  const char* force = std::getenv("PT_OBJ_SERIAL");
  if (text.size() < (1u << 20) || (force && std::atoi(force) != 0)) return parse_obj_serial(path, text, positions, indices);

  const char* base = text.data();
  const char* end = base + text.size();
  int n_chunks = std::max(1, std::min<int>(256, (int)(text.size() >> 19)));
  std::vector<ObjChunk> chunks;
  {
    const char* p = base;
    for (int k = 0; k < n_chunks && p < end; ++k) {
      const char* want = base + (size_t)((double)text.size() * (k + 1) / n_chunks);
      const char* e = k + 1 == n_chunks ? end : (const char*)std::memchr(want, '\n', end - want);
      e = e ? std::min(end, e + 1) : end;
      if (e <= p) continue;
      ObjChunk c;
      c.begin = p;
      c.end = e;
      chunks.push_back(c);
      p = e;
    }
  }
  // sweep 1: counts
#pragma omp parallel for schedule(dynamic, 1)
  for (int k = 0; k < (int)chunks.size(); ++k) obj_count(chunks[k]);
  // cut-off: the first o/g/usemtl statement after the first face ends the (first) mesh
  {
    bool have_faces = false;
    const char* cut = nullptr;
    size_t cut_chunk = 0;
    for (size_t k = 0; k < chunks.size() && !cut; ++k) {
      const ObjChunk& c = chunks[k];
      const char* g = have_faces ? c.first_group : c.group_after_face;
      if (g) {
        cut = g;
        cut_chunk = k;
      }
      if (c.first_face) have_faces = true;
    }
    if (cut) {
      chunks.resize(cut_chunk + 1);
      chunks.back().end = cut;
      obj_count(chunks.back());
    }
  }
  uint64_t nv = 0, nt = 0;
  bool bad_face = false;
  for (auto& c : chunks) {
    c.v_base = nv;
    c.tri_base = nt;
    nv += c.n_v;
    nt += c.n_tri;
    bad_face = bad_face || c.bad_face;
  }
  if (bad_face) return fail(PT_ERR_PARSE, std::string("bad face in ") + path);
  if (nt == 0) return fail(PT_ERR_PARSE, std::string("Unable to load ") + path);
  if (nt * 3 >= (1ull << 32)) return fail(PT_ERR_INVALID, std::string("too many triangles in ") + path);
  std::vector<float> verts(nv * 3);
  positions.resize(nt * 9);
  indices.resize(nt * 3);
  // sweep 2: vertices
#pragma omp parallel for schedule(dynamic, 1)
  for (int k = 0; k < (int)chunks.size(); ++k) {
    ObjChunk& c = chunks[k];
    float* out = verts.data() + c.v_base * 3;
    const char* p = c.begin;
    while (p < c.end) {
      const char* line_end = (const char*)std::memchr(p, '\n', c.end - p);
      if (!line_end) line_end = c.end;
      const char* q = obj_skip_sp(p, c.end);
      if (obj_classify(q, line_end) == OBJ_V) {
        ++q;
        float v[3] = {0.f, 0.f, 0.f};
        for (int i = 0; i < 3; ++i) {
          q = obj_skip_sp(q, line_end);
          if (q < line_end && *q == '+') ++q;
          auto r = std::from_chars(q, line_end, v[i]);
          if (r.ec != std::errc()) {
            c.bad_vertex = true;
            break;
          }
          q = r.ptr;
        }
        out[0] = v[0], out[1] = v[1], out[2] = v[2];
        out += 3;
      }
      p = line_end + 1;
    }
  }
  // sweep 3: faces (fan-triangulated, one vertex per corner)
#pragma omp parallel for schedule(dynamic, 1)
  for (int k = 0; k < (int)chunks.size(); ++k) {
    ObjChunk& c = chunks[k];
    long long seen = (long long)c.v_base; // vertices defined before the current line
    uint64_t tri = c.tri_base;
    std::vector<long long> corner(64);
    const char* p = c.begin;
    while (p < c.end) {
      const char* line_end = (const char*)std::memchr(p, '\n', c.end - p);
      if (!line_end) line_end = c.end;
      const char* q = obj_skip_sp(p, c.end);
      const ObjLine kind = obj_classify(q, line_end);
      if (kind == OBJ_V) {
        ++seen;
      } else if (kind == OBJ_F) {
        int n = 0;
        obj_face_corners(q + 1, line_end, corner.data(), (int)corner.size(), n);
        if (n > (int)corner.size()) {
          corner.resize(n);
          obj_face_corners(q + 1, line_end, corner.data(), (int)corner.size(), n);
        }
        for (int i = 0; i < n; ++i) {
          long long idx = corner[i];
          if (idx < 0) idx = seen + idx; else idx -= 1;
          if (idx < 0 || idx >= seen) c.range_error = true;
          corner[i] = idx;
        }
        if (n >= 3 && !c.range_error) {
          for (int j = 1; j + 1 < n; ++j) {
            const long long t3[3] = {corner[0], corner[j], corner[j + 1]};
            for (int cc = 0; cc < 3; ++cc) {
              const uint64_t o = tri * 3 + (uint64_t)cc;
              indices[o] = (uint32_t)o;
              positions[o * 3 + 0] = verts[3 * t3[cc] + 0];
              positions[o * 3 + 1] = verts[3 * t3[cc] + 1];
              positions[o * 3 + 2] = verts[3 * t3[cc] + 2];
            }
            ++tri;
          }
        }
      }
      p = line_end + 1;
    }
  }
  for (const auto& c : chunks) {
    if (c.bad_vertex) return fail(PT_ERR_PARSE, std::string("bad vertex in ") + path);
    if (c.range_error) return fail(PT_ERR_PARSE, std::string("face index out of range in ") + path);
  }
  return PT_OK;
}

// ------------------------------------------------------------ scene JSON
int load_scene_file(const char* json_path, SceneFile& out)
{
  std::string text;
  if (!read_file(json_path, text))
    return fail(PT_ERR_IO, std::string("Json Parser: Cannot open file ") + json_path);
  JParser jp{text.data(), text.data() + text.size(), {}};
  JPtr root;
  // (`file >> json`, json_parser.cpp:166-167, reads ONE value and leaves what follows unread:
  // text after the document is not an error)
  if (!jp.parse(root) || root->kind != JValue::Obj)
    return fail(PT_ERR_PARSE, "Json Parser: " + (jp.err.empty() ? std::string("root is not an object") : jp.err));

  namespace fs = std::filesystem;
  const fs::path file_dir = fs::path(json_path).parent_path();
  try {
    // read_materials (json_parser.cpp:101-122); GPU table order = std::map order
    // (scene_description.cpp:59-66); duplicates keep the first (try_emplace :151-154)
    const JValue* mats = root->find("materials");
    if (!mats || mats->kind != JValue::Arr) throw ParseError{"Json Parser: materials is not array!"};
    std::map<std::string, pt_material> mat_map;
    for (auto& mj : mats->arr) {
      const std::string name = str(mj->find("name"), "material name");
      const std::string type = str(mj->find("type"), "material type");
      pt_material m{};
      if (type == "lambertian") {
        m.type = PT_MAT_DIFFUSE;
        vec3(mj->find("albedo"), m.albedo);
      } else if (type == "dielectric") {
        m.type = PT_MAT_DIELECTRIC;
        m.refraction_index = num(mj->find("refraction_index"), "refraction_index");
      } else if (type == "metal") {
        m.type = PT_MAT_METAL;
        vec3(mj->find("albedo"), m.albedo);
        m.fuzz = num(mj->find("fuzz"), "fuzz");
      } else {
        throw ParseError{"Json Parser: Unsupported material type " + type};
      }
      mat_map.emplace(name, m);
    }
    std::map<std::string, uint32_t> mat_index;
    for (auto& kv : mat_map) {
      mat_index[kv.first] = (uint32_t)out.materials.size();
      out.materials.push_back(kv.second);
    }

    // read_surfaces (json_parser.cpp:133-159)
    const JValue* surfaces = root->find("surfaces");
    if (!surfaces || surfaces->kind != JValue::Arr) throw ParseError{"Json Parser: surfaces is not array"};
    std::map<std::string, int> mesh_paths; // canonical path -> seen
    std::vector<std::pair<size_t, std::string>> mesh_object_paths; // object index -> its mesh
    for (auto& sj : surfaces->arr) {
      const std::string type = str(sj->find("type"), "surface type");
      const std::string material = str(sj->find("material"), "surface material");
      auto mit = mat_index.find(material);
      if (mit == mat_index.end()) throw ParseError{"Cannot find material " + material};
      pt_object ob{};
      ob.material = mit->second;
      const Mat4 m = read_transform(sj->find("transform"));
      const Mat4 inv = mat4_inverse(m);
      std::memcpy(ob.m, m.m, sizeof(ob.m));
      std::memcpy(ob.inv, inv.m, sizeof(ob.inv));
      if (type == "sphere") {
        ob.type = PT_OBJ_SPHERE;
        ob.prim_index = (uint32_t)out.spheres.size();
        pt_sphere s{};
        s.radius = num(sj->find("radius"), "radius");
        out.spheres.push_back(s);
      } else if (type == "mesh") {
        ob.type = PT_OBJ_MESH;
        const std::string filename = str(sj->find("filename"), "mesh filename");
        std::error_code ec;
        const fs::path canon = fs::canonical(file_dir / filename, ec);
        if (ec) throw ParseError{"Unable to load " + (file_dir / filename).string()};
        mesh_paths[canon.string()] = 1;
        mesh_object_paths.emplace_back(out.objects.size(), canon.string());
      } else {
        throw ParseError{"Json Parser: Not supported surface type " + type};
      }
      out.objects.push_back(ob);
    }
    const char* all_env = std::getenv("PT_ALL_MESHES");
    const bool all_meshes = all_env && std::atoi(all_env) != 0;
    if (!mesh_paths.empty() && !all_meshes) {
      // Only the alphabetically-first mesh is uploaded and every mesh object
      // instances it (scene_description.cpp:95, scene.hpp:33-39).
      if (mesh_paths.size() > 1)
        std::fprintf(stderr,
                     "warning: %zu meshes referenced; like the reference only the first (%s) is used "
                     "(PT_ALL_MESHES=1 / --all-meshes honours every mesh)\n",
                     mesh_paths.size(), mesh_paths.begin()->first.c_str());
      int rc = load_obj_file(mesh_paths.begin()->first.c_str(), out.positions, out.indices);
      if (rc != PT_OK) return rc;
    } else if (!mesh_paths.empty()) {
      // extension: every referenced mesh is loaded (alphabetical order, one shared vertex and
      // index buffer) and each mesh object instances its own
      std::map<std::string, uint32_t> mesh_rank;
      out.mesh_first_index.push_back(0);
      for (auto& kv : mesh_paths) {
        std::vector<float> pos;
        std::vector<uint32_t> idx;
        int rc = load_obj_file(kv.first.c_str(), pos, idx);
        if (rc != PT_OK) return rc;
        const uint32_t base = (uint32_t)(out.positions.size() / 3);
        out.positions.insert(out.positions.end(), pos.begin(), pos.end());
        for (uint32_t i : idx) out.indices.push_back(base + i);
        mesh_rank[kv.first] = (uint32_t)out.mesh_first_index.size() - 1;
        out.mesh_first_index.push_back(out.indices.size());
      }
      for (auto& op : mesh_object_paths) out.objects[op.first].prim_index = mesh_rank[op.second];
    }

    // camera (json_parser.cpp:187-209)
    const JValue* cam = root->find("camera");
    if (!cam || cam->kind != JValue::Obj) throw ParseError{"Json Parser: Camera is not an object!"};
    out.camera = pt_camera{{0.f, 0.f, 0.f}, {1.f, 0.f, 0.f, 0.f}, 1.5707963267948966f};
    if (const JValue* t = cam->find("transform")) {
      const Mat4 m = read_transform(t);
      if (!mat4_decompose_trs(m, out.camera.position, out.camera.rotation))
        throw ParseError{"Json parser: failed to decompose camera transformation!"};
    }
    out.camera.vfov = num(cam->find("vfov"), "vfov") * 0.01745329251994329576923690768489f;
    if (const JValue* r = cam->find("resolution")) {
      if (r->kind != JValue::Arr || r->arr.size() != 2) throw ParseError{"Json Parser: resolution need to be 2d"};
      out.width = (int)num(r->arr[0].get(), "resolution");
      out.height = (int)num(r->arr[1].get(), "resolution");
    }
    out.spp = 1;
    if (const JValue* s = root->find("sampler")) {
      if (s->kind != JValue::Obj) throw ParseError{"Json Parser: Sampler is not an object!"};
      out.spp = (int)num(s->find("samples"), "samples");
    }
  } catch (const ParseError& e) {
    return fail(PT_ERR_PARSE, e.msg);
  }
  return PT_OK;
}

} // namespace pt
