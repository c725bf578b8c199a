// hostmath.cpp — the handful of glm 0.9.9.8 operations the reference's host
// code performs on transforms and cameras (assets/json_parser.cpp:40-95,190-203;
// camera.cpp:5-13; transform.hpp:14-18), restated on plain column-major float[16].
// glm itself is not a dependency of this library.
#include "internal.h"

#include <cmath>
#include <cstring>

namespace pt {

Mat4 mat4_identity()
{
  Mat4 r;
  std::memset(r.m, 0, sizeof(r.m));
  r.m[0] = r.m[5] = r.m[10] = r.m[15] = 1.0f;
  return r;
}

// glm: Result[j] = A[0]*B[j][0] + A[1]*B[j][1] + A[2]*B[j][2] + A[3]*B[j][3]
Mat4 mat4_mul(const Mat4& a, const Mat4& b)
{
  Mat4 r;
  for (int j = 0; j < 4; ++j)
    for (int i = 0; i < 4; ++i) {
      float s = a.m[0 * 4 + i] * b.m[j * 4 + 0];
      s = s + a.m[1 * 4 + i] * b.m[j * 4 + 1];
      s = s + a.m[2 * 4 + i] * b.m[j * 4 + 2];
      s = s + a.m[3 * 4 + i] * b.m[j * 4 + 3];
      r.m[j * 4 + i] = s;
    }
  return r;
}

Mat4 mat4_translate(float x, float y, float z)
{
  Mat4 r = mat4_identity();
  r.m[12] = x, r.m[13] = y, r.m[14] = z;
  return r;
}

Mat4 mat4_scale(float x, float y, float z)
{
  Mat4 r = mat4_identity();
  r.m[0] = x, r.m[5] = y, r.m[10] = z;
  return r;
}

// glm::rotate(angle, axis) applied to the identity
Mat4 mat4_rotate(float angle, float ax, float ay, float az)
{
  const float c = std::cos(angle), s = std::sin(angle);
  const float inv_len = 1.0f / std::sqrt(ax * ax + ay * ay + az * az);
  const float axis[3] = {ax * inv_len, ay * inv_len, az * inv_len};
  const float temp[3] = {(1.0f - c) * axis[0], (1.0f - c) * axis[1], (1.0f - c) * axis[2]};
  Mat4 r = mat4_identity();
  r.m[0] = c + temp[0] * axis[0];
  r.m[1] = temp[0] * axis[1] + s * axis[2];
  r.m[2] = temp[0] * axis[2] - s * axis[1];
  r.m[4] = temp[1] * axis[0] - s * axis[2];
  r.m[5] = c + temp[1] * axis[1];
  r.m[6] = temp[1] * axis[2] + s * axis[0];
  r.m[8] = temp[2] * axis[0] + s * axis[1];
  r.m[9] = temp[2] * axis[1] - s * axis[0];
  r.m[10] = c + temp[2] * axis[2];
  return r;
}

// cofactor inverse (the classic 2x2 sub-determinant formulation)
Mat4 mat4_inverse(const Mat4& a)
{
  const float* m = a.m;
#define M(c, r) m[(c) * 4 + (r)]
  const float c00 = M(2, 2) * M(3, 3) - M(3, 2) * M(2, 3);
  const float c02 = M(1, 2) * M(3, 3) - M(3, 2) * M(1, 3);
  const float c03 = M(1, 2) * M(2, 3) - M(2, 2) * M(1, 3);
  const float c04 = M(2, 1) * M(3, 3) - M(3, 1) * M(2, 3);
  const float c06 = M(1, 1) * M(3, 3) - M(3, 1) * M(1, 3);
  const float c07 = M(1, 1) * M(2, 3) - M(2, 1) * M(1, 3);
  const float c08 = M(2, 1) * M(3, 2) - M(3, 1) * M(2, 2);
  const float c10 = M(1, 1) * M(3, 2) - M(3, 1) * M(1, 2);
  const float c11 = M(1, 1) * M(2, 2) - M(2, 1) * M(1, 2);
  const float c12 = M(2, 0) * M(3, 3) - M(3, 0) * M(2, 3);
  const float c14 = M(1, 0) * M(3, 3) - M(3, 0) * M(1, 3);
  const float c15 = M(1, 0) * M(2, 3) - M(2, 0) * M(1, 3);
  const float c16 = M(2, 0) * M(3, 2) - M(3, 0) * M(2, 2);
  const float c18 = M(1, 0) * M(3, 2) - M(3, 0) * M(1, 2);
  const float c19 = M(1, 0) * M(2, 2) - M(2, 0) * M(1, 2);
  const float c20 = M(2, 0) * M(3, 1) - M(3, 0) * M(2, 1);
  const float c22 = M(1, 0) * M(3, 1) - M(3, 0) * M(1, 1);
  const float c23 = M(1, 0) * M(2, 1) - M(2, 0) * M(1, 1);

  const float f0[4] = {c00, c00, c02, c03};
  const float f1[4] = {c04, c04, c06, c07};
  const float f2[4] = {c08, c08, c10, c11};
  const float f3[4] = {c12, c12, c14, c15};
  const float f4[4] = {c16, c16, c18, c19};
  const float f5[4] = {c20, c20, c22, c23};
  const float v0[4] = {M(1, 0), M(0, 0), M(0, 0), M(0, 0)};
  const float v1[4] = {M(1, 1), M(0, 1), M(0, 1), M(0, 1)};
  const float v2[4] = {M(1, 2), M(0, 2), M(0, 2), M(0, 2)};
  const float v3[4] = {M(1, 3), M(0, 3), M(0, 3), M(0, 3)};
  const float sa[4] = {+1, -1, +1, -1};
  const float sb[4] = {-1, +1, -1, +1};
  Mat4 inv;
  for (int i = 0; i < 4; ++i) {
    inv.m[0 * 4 + i] = (v1[i] * f0[i] - v2[i] * f1[i] + v3[i] * f2[i]) * sa[i];
    inv.m[1 * 4 + i] = (v0[i] * f0[i] - v2[i] * f3[i] + v3[i] * f4[i]) * sb[i];
    inv.m[2 * 4 + i] = (v0[i] * f1[i] - v1[i] * f3[i] + v3[i] * f5[i]) * sa[i];
    inv.m[3 * 4 + i] = (v0[i] * f2[i] - v1[i] * f4[i] + v2[i] * f5[i]) * sb[i];
  }
  const float dot1 = (M(0, 0) * inv.m[0] + M(0, 1) * inv.m[4]) + (M(0, 2) * inv.m[8] + M(0, 3) * inv.m[12]);
#undef M
  const float ood = 1.0f / dot1;
  for (int i = 0; i < 16; ++i) inv.m[i] *= ood;
  return inv;
}

// glm::decompose restricted to what scene_from_json consumes: translation and
// orientation (json_parser.cpp:190-203).  Scale/skew are removed the same way
// (Gram-Schmidt on the upper 3x3 columns) before the quaternion is extracted.
bool mat4_decompose_trs(const Mat4& a, float pos[3], float q[4])
{
  Mat4 L = a;
  if (std::fabs(L.m[15]) < 1e-12f) return false;
  for (int i = 0; i < 16; ++i) L.m[i] /= L.m[15];
  // perspective partition must be trivial for a camera pose
  L.m[3] = L.m[7] = L.m[11] = 0.f;
  L.m[15] = 1.f;
  pos[0] = L.m[12], pos[1] = L.m[13], pos[2] = L.m[14];
  float row[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) row[i][j] = L.m[i * 4 + j];
  auto len = [](const float* v) { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); };
  auto dotf = [](const float* x, const float* y) { return x[0] * y[0] + x[1] * y[1] + x[2] * y[2]; };
  auto scl = [](float* v, float s) { v[0] *= s, v[1] *= s, v[2] *= s; };
  auto comb = [](float* x, const float* y, float a1, float b1) {
    for (int i = 0; i < 3; ++i) x[i] = a1 * x[i] + b1 * y[i];
  };
  float sx = len(row[0]);
  if (sx == 0.f) return false;
  scl(row[0], 1.0f / sx);
  float skz = dotf(row[0], row[1]);
  comb(row[1], row[0], 1.f, -skz);
  float sy = len(row[1]);
  if (sy == 0.f) return false;
  scl(row[1], 1.0f / sy);
  float sky = dotf(row[0], row[2]);
  comb(row[2], row[0], 1.f, -sky);
  float skx = dotf(row[1], row[2]);
  comb(row[2], row[1], 1.f, -skx);
  float sz = len(row[2]);
  if (sz == 0.f) return false;
  scl(row[2], 1.0f / sz);
  const float cr[3] = {row[1][1] * row[2][2] - row[1][2] * row[2][1],
                       row[1][2] * row[2][0] - row[1][0] * row[2][2],
                       row[1][0] * row[2][1] - row[1][1] * row[2][0]};
  if (dotf(row[0], cr) < 0.f)
    for (int i = 0; i < 3; ++i) scl(row[i], -1.f);

  float qx, qy, qz, qw;
  const float trace = row[0][0] + row[1][1] + row[2][2];
  if (trace > 0.f) {
    float root = std::sqrt(trace + 1.0f);
    qw = 0.5f * root;
    root = 0.5f / root;
    qx = root * (row[1][2] - row[2][1]);
    qy = root * (row[2][0] - row[0][2]);
    qz = root * (row[0][1] - row[1][0]);
  } else {
    static const int next[3] = {1, 2, 0};
    int i = 0;
    if (row[1][1] > row[0][0]) i = 1;
    if (row[2][2] > row[i][i]) i = 2;
    const int j = next[i], k = next[j];
    float root = std::sqrt(row[i][i] - row[j][j] - row[k][k] + 1.0f);
    float o[3];
    o[i] = 0.5f * root;
    root = 0.5f / root;
    o[j] = root * (row[i][j] + row[j][i]);
    o[k] = root * (row[i][k] + row[k][i]);
    qw = root * (row[j][k] - row[k][j]);
    qx = o[0], qy = o[1], qz = o[2];
  }
  q[0] = qw, q[1] = qx, q[2] = qy, q[3] = qz;
  return true;
}

// Camera::to_gpu_camera: translate(identity, position) * mat4_cast(rotation)
// (camera.cpp:5-13).  Output: columns 0..2 xyz, then the translation.
void mat4_from_camera(const pt_camera& cam, float out[12])
{
  const float w = cam.rotation[0], x = cam.rotation[1], y = cam.rotation[2], z = cam.rotation[3];
  const float qxx = x * x, qyy = y * y, qzz = z * z;
  const float qxz = x * z, qxy = x * y, qyz = y * z;
  const float qwx = w * x, qwy = w * y, qwz = w * z;
  out[0] = 1.0f - 2.0f * (qyy + qzz);
  out[1] = 2.0f * (qxy + qwz);
  out[2] = 2.0f * (qxz - qwy);
  out[3] = 2.0f * (qxy - qwz);
  out[4] = 1.0f - 2.0f * (qxx + qzz);
  out[5] = 2.0f * (qyz + qwx);
  out[6] = 2.0f * (qxz + qwy);
  out[7] = 2.0f * (qyz - qwx);
  out[8] = 1.0f - 2.0f * (qxx + qyy);
  out[9] = cam.position[0];
  out[10] = cam.position[1];
  out[11] = cam.position[2];
}

DevCamera make_dev_camera(const pt_camera& cam, uint32_t w, uint32_t h)
{
  DevCamera c{};
  mat4_from_camera(cam, c.m);
  const float aspect = (float)w / (float)h;
  c.vp_h = 2.0f * std::tan(cam.vfov / 2.0f);
  c.vp_w = aspect * c.vp_h;
  c.fw_m1 = (float)(w - 1);
  c.fh_m1 = (float)(h - 1);
  c.fh = (float)h;
  c.width = w;
  c.height = h;
  return c;
}

} // namespace pt
