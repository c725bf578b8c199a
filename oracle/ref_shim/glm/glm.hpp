// Minimal glm-compatible header set — TEST INFRASTRUCTURE (oracle/_ref build only).
//
// glm 0.9.9.8 is an un-vendored Conan dependency of the reference (conanfile.txt:6) and is
// absent from this image.  This file restates, from glm's documented semantics, exactly the
// API surface the reference's hot-path sources use (vec3/vec4/mat4/quat and ~20 functions),
// keeping glm's operation order (dot = x+y+z left to right, mat4*vec4 = (m0 v0 + m1 v1) +
// (m2 v2 + m3 v3), normalize = v * (1/sqrt(dot)), min/max = comparison-select) so the
// reference compiled against it computes what it computes against glm up to compiler
// contraction.  It is not used by the product.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <limits>
#include <type_traits>

#if defined(__CUDACC__)
#define GLM_FN __host__ __device__ inline
#define GLM_CE __host__ __device__ constexpr
#else
#define GLM_FN inline
#define GLM_CE constexpr
#endif

namespace glm {

struct vec4;

struct vec3 {
  float x, y, z;
  vec3() = default;
  GLM_CE explicit vec3(float s) : x(s), y(s), z(s) {}
  template <class A, class B, class C,
            class = std::enable_if_t<std::is_arithmetic_v<A> && std::is_arithmetic_v<B> &&
                                     std::is_arithmetic_v<C>>>
  GLM_CE vec3(A a, B b, C c) : x(static_cast<float>(a)), y(static_cast<float>(b)), z(static_cast<float>(c))
  {
  }
  GLM_CE explicit vec3(const vec4& v);
  GLM_CE float& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
  GLM_CE const float& operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
  GLM_CE float& operator[](std::size_t i) { return i == 0 ? x : (i == 1 ? y : z); }
  GLM_CE const float& operator[](std::size_t i) const { return i == 0 ? x : (i == 1 ? y : z); }
  GLM_FN vec3& operator+=(const vec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
  GLM_FN vec3& operator-=(const vec3& o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
  GLM_FN vec3& operator*=(const vec3& o) { x *= o.x; y *= o.y; z *= o.z; return *this; }
  GLM_FN vec3& operator*=(float s) { x *= s; y *= s; z *= s; return *this; }
  GLM_FN vec3& operator/=(float s) { x /= s; y /= s; z /= s; return *this; }
};

struct vec4 {
  float x, y, z, w;
  vec4() = default;
  GLM_CE explicit vec4(float s) : x(s), y(s), z(s), w(s) {}
  template <class A, class B, class C, class D,
            class = std::enable_if_t<std::is_arithmetic_v<A> && std::is_arithmetic_v<B> &&
                                     std::is_arithmetic_v<C> && std::is_arithmetic_v<D>>>
  GLM_CE vec4(A a, B b, C c, D d)
      : x(static_cast<float>(a)), y(static_cast<float>(b)), z(static_cast<float>(c)), w(static_cast<float>(d))
  {
  }
  template <class D, class = std::enable_if_t<std::is_arithmetic_v<D>>>
  GLM_CE vec4(const vec3& v, D d) : x(v.x), y(v.y), z(v.z), w(static_cast<float>(d))
  {
  }
  GLM_CE float& operator[](int i) { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
  GLM_CE const float& operator[](int i) const { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
};

GLM_CE vec3::vec3(const vec4& v) : x(v.x), y(v.y), z(v.z) {}

// ---- vec3 operators
GLM_CE vec3 operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
GLM_CE vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
GLM_CE vec3 operator-(const vec3& a) { return vec3(-a.x, -a.y, -a.z); }
GLM_CE vec3 operator*(const vec3& a, const vec3& b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
GLM_CE vec3 operator*(const vec3& a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
GLM_CE vec3 operator*(float s, const vec3& a) { return vec3(s * a.x, s * a.y, s * a.z); }
GLM_CE vec3 operator/(const vec3& a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
GLM_CE vec3 operator/(const vec3& a, const vec3& b) { return vec3(a.x / b.x, a.y / b.y, a.z / b.z); }
GLM_CE vec3 operator+(const vec3& a, float s) { return vec3(a.x + s, a.y + s, a.z + s); }
GLM_CE vec3 operator-(const vec3& a, float s) { return vec3(a.x - s, a.y - s, a.z - s); }
GLM_CE bool operator==(const vec3& a, const vec3& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
// ---- vec4 operators
GLM_CE vec4 operator+(const vec4& a, const vec4& b) { return vec4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
GLM_CE vec4 operator-(const vec4& a, const vec4& b) { return vec4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
GLM_CE vec4 operator*(const vec4& a, const vec4& b) { return vec4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
GLM_CE vec4 operator*(const vec4& a, float s) { return vec4(a.x * s, a.y * s, a.z * s, a.w * s); }
GLM_CE vec4 operator*(float s, const vec4& a) { return vec4(s * a.x, s * a.y, s * a.z, s * a.w); }
GLM_CE vec4 operator/(const vec4& a, float s) { return vec4(a.x / s, a.y / s, a.z / s, a.w / s); }

// ---- geometric
GLM_CE float dot(const vec3& a, const vec3& b)
{
  const vec3 tmp(a * b);
  return tmp.x + tmp.y + tmp.z;
}
GLM_CE float dot(const vec4& a, const vec4& b)
{
  const vec4 tmp(a * b);
  return (tmp.x + tmp.y) + (tmp.z + tmp.w);
}
GLM_CE vec3 cross(const vec3& x, const vec3& y)
{
  return vec3(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y);
}
GLM_FN float inversesqrt(float x) { return 1.0f / std::sqrt(x); }
GLM_FN float length(const vec3& v) { return std::sqrt(dot(v, v)); }
GLM_FN float distance(const vec3& p0, const vec3& p1) { return length(p1 - p0); }
GLM_FN vec3 normalize(const vec3& v) { return v * inversesqrt(dot(v, v)); }
GLM_CE vec3 reflect(const vec3& I, const vec3& N) { return I - N * dot(N, I) * 2.0f; }
GLM_FN vec3 refract(const vec3& I, const vec3& N, float eta)
{
  const float dotValue(dot(N, I));
  const float k(1.0f - eta * eta * (1.0f - dotValue * dotValue));
  return (k >= 0.0f) ? (eta * I - (eta * dotValue + std::sqrt(k)) * N) : vec3(0.0f);
}

// ---- common
GLM_CE float min(float x, float y) { return (y < x) ? y : x; }
GLM_CE float max(float x, float y) { return (x < y) ? y : x; }
GLM_CE vec3 min(const vec3& a, const vec3& b) { return vec3(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z)); }
GLM_CE vec3 max(const vec3& a, const vec3& b) { return vec3(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z)); }
GLM_CE float clamp(float x, float lo, float hi) { return min(max(x, lo), hi); }
GLM_CE float sign(float x) { return static_cast<float>(0.0f < x) - static_cast<float>(x < 0.0f); }
GLM_FN vec3 pow(const vec3& b, const vec3& e) { return vec3(std::pow(b.x, e.x), std::pow(b.y, e.y), std::pow(b.z, e.z)); }
GLM_CE vec3 mix(const vec3& x, const vec3& y, float a) { return x * (1.0f - a) + y * a; }
GLM_CE vec3 lerp(const vec3& x, const vec3& y, float a) { return mix(x, y, a); }
GLM_CE float radians(float deg) { return deg * 0.01745329251994329576923690768489f; }
template <class T> GLM_CE T pi() { return static_cast<T>(3.14159265358979323846264338327950288); }

// ---- mat4 (column-major)
struct mat4 {
  vec4 value[4];
  mat4() = default;
  GLM_CE explicit mat4(float s)
      : value{vec4(s, 0, 0, 0), vec4(0, s, 0, 0), vec4(0, 0, s, 0), vec4(0, 0, 0, s)}
  {
  }
  GLM_CE explicit mat4(double s) : mat4(static_cast<float>(s)) {}
  GLM_CE mat4(const vec4& a, const vec4& b, const vec4& c, const vec4& d) : value{a, b, c, d} {}
  GLM_CE vec4& operator[](int i) { return value[i]; }
  GLM_CE const vec4& operator[](int i) const { return value[i]; }
};

GLM_CE vec4 operator*(const mat4& m, const vec4& v)
{
  const vec4 Mov0(v[0]);
  const vec4 Mov1(v[1]);
  const vec4 Mul0 = m[0] * Mov0;
  const vec4 Mul1 = m[1] * Mov1;
  const vec4 Add0 = Mul0 + Mul1;
  const vec4 Mov2(v[2]);
  const vec4 Mov3(v[3]);
  const vec4 Mul2 = m[2] * Mov2;
  const vec4 Mul3 = m[3] * Mov3;
  const vec4 Add1 = Mul2 + Mul3;
  const vec4 Add2 = Add0 + Add1;
  return Add2;
}
GLM_CE mat4 operator*(const mat4& m1, const mat4& m2)
{
  const vec4 SrcA0 = m1[0], SrcA1 = m1[1], SrcA2 = m1[2], SrcA3 = m1[3];
  const vec4 SrcB0 = m2[0], SrcB1 = m2[1], SrcB2 = m2[2], SrcB3 = m2[3];
  mat4 Result;
  Result[0] = SrcA0 * SrcB0[0] + SrcA1 * SrcB0[1] + SrcA2 * SrcB0[2] + SrcA3 * SrcB0[3];
  Result[1] = SrcA0 * SrcB1[0] + SrcA1 * SrcB1[1] + SrcA2 * SrcB1[2] + SrcA3 * SrcB1[3];
  Result[2] = SrcA0 * SrcB2[0] + SrcA1 * SrcB2[1] + SrcA2 * SrcB2[2] + SrcA3 * SrcB2[3];
  Result[3] = SrcA0 * SrcB3[0] + SrcA1 * SrcB3[1] + SrcA2 * SrcB3[2] + SrcA3 * SrcB3[3];
  return Result;
}
GLM_CE mat4 operator*(const mat4& m, float s) { return mat4(m[0] * s, m[1] * s, m[2] * s, m[3] * s); }

GLM_CE mat4 transpose(const mat4& m)
{
  return mat4(vec4(m[0][0], m[1][0], m[2][0], m[3][0]), vec4(m[0][1], m[1][1], m[2][1], m[3][1]),
              vec4(m[0][2], m[1][2], m[2][2], m[3][2]), vec4(m[0][3], m[1][3], m[2][3], m[3][3]));
}

GLM_CE mat4 inverse(const mat4& m)
{
  const float Coef00 = m[2][2] * m[3][3] - m[3][2] * m[2][3];
  const float Coef02 = m[1][2] * m[3][3] - m[3][2] * m[1][3];
  const float Coef03 = m[1][2] * m[2][3] - m[2][2] * m[1][3];
  const float Coef04 = m[2][1] * m[3][3] - m[3][1] * m[2][3];
  const float Coef06 = m[1][1] * m[3][3] - m[3][1] * m[1][3];
  const float Coef07 = m[1][1] * m[2][3] - m[2][1] * m[1][3];
  const float Coef08 = m[2][1] * m[3][2] - m[3][1] * m[2][2];
  const float Coef10 = m[1][1] * m[3][2] - m[3][1] * m[1][2];
  const float Coef11 = m[1][1] * m[2][2] - m[2][1] * m[1][2];
  const float Coef12 = m[2][0] * m[3][3] - m[3][0] * m[2][3];
  const float Coef14 = m[1][0] * m[3][3] - m[3][0] * m[1][3];
  const float Coef15 = m[1][0] * m[2][3] - m[2][0] * m[1][3];
  const float Coef16 = m[2][0] * m[3][2] - m[3][0] * m[2][2];
  const float Coef18 = m[1][0] * m[3][2] - m[3][0] * m[1][2];
  const float Coef19 = m[1][0] * m[2][2] - m[2][0] * m[1][2];
  const float Coef20 = m[2][0] * m[3][1] - m[3][0] * m[2][1];
  const float Coef22 = m[1][0] * m[3][1] - m[3][0] * m[1][1];
  const float Coef23 = m[1][0] * m[2][1] - m[2][0] * m[1][1];
  const vec4 Fac0(Coef00, Coef00, Coef02, Coef03);
  const vec4 Fac1(Coef04, Coef04, Coef06, Coef07);
  const vec4 Fac2(Coef08, Coef08, Coef10, Coef11);
  const vec4 Fac3(Coef12, Coef12, Coef14, Coef15);
  const vec4 Fac4(Coef16, Coef16, Coef18, Coef19);
  const vec4 Fac5(Coef20, Coef20, Coef22, Coef23);
  const vec4 Vec0(m[1][0], m[0][0], m[0][0], m[0][0]);
  const vec4 Vec1(m[1][1], m[0][1], m[0][1], m[0][1]);
  const vec4 Vec2(m[1][2], m[0][2], m[0][2], m[0][2]);
  const vec4 Vec3(m[1][3], m[0][3], m[0][3], m[0][3]);
  const vec4 Inv0(Vec1 * Fac0 - Vec2 * Fac1 + Vec3 * Fac2);
  const vec4 Inv1(Vec0 * Fac0 - Vec2 * Fac3 + Vec3 * Fac4);
  const vec4 Inv2(Vec0 * Fac1 - Vec1 * Fac3 + Vec3 * Fac5);
  const vec4 Inv3(Vec0 * Fac2 - Vec1 * Fac4 + Vec2 * Fac5);
  const vec4 SignA(+1, -1, +1, -1);
  const vec4 SignB(-1, +1, -1, +1);
  const mat4 Inverse(Inv0 * SignA, Inv1 * SignB, Inv2 * SignA, Inv3 * SignB);
  const vec4 Row0(Inverse[0][0], Inverse[1][0], Inverse[2][0], Inverse[3][0]);
  const vec4 Dot0(m[0] * Row0);
  const float Dot1 = (Dot0.x + Dot0.y) + (Dot0.z + Dot0.w);
  const float OneOverDeterminant = 1.0f / Dot1;
  return Inverse * OneOverDeterminant;
}

template <class T> GLM_CE T identity() { return T(1.0f); }

GLM_CE mat4 translate(const mat4& m, const vec3& v)
{
  mat4 Result(m);
  Result[3] = m[0] * v[0] + m[1] * v[1] + m[2] * v[2] + m[3];
  return Result;
}
GLM_CE mat4 translate(const vec3& v) { return translate(mat4(1.0f), v); }
GLM_CE mat4 scale(const mat4& m, const vec3& v)
{
  return mat4(m[0] * v[0], m[1] * v[1], m[2] * v[2], m[3]);
}
GLM_CE mat4 scale(const vec3& v) { return scale(mat4(1.0f), v); }
GLM_FN mat4 rotate(const mat4& m, float angle, const vec3& v)
{
  const float a = angle;
  const float c = std::cos(a);
  const float s = std::sin(a);
  const vec3 axis(normalize(v));
  const vec3 temp((1.0f - c) * axis);
  mat4 Rotate;
  Rotate[0][0] = c + temp[0] * axis[0];
  Rotate[0][1] = temp[0] * axis[1] + s * axis[2];
  Rotate[0][2] = temp[0] * axis[2] - s * axis[1];
  Rotate[1][0] = temp[1] * axis[0] - s * axis[2];
  Rotate[1][1] = c + temp[1] * axis[1];
  Rotate[1][2] = temp[1] * axis[2] + s * axis[0];
  Rotate[2][0] = temp[2] * axis[0] + s * axis[1];
  Rotate[2][1] = temp[2] * axis[1] - s * axis[0];
  Rotate[2][2] = c + temp[2] * axis[2];
  mat4 Result;
  Result[0] = m[0] * Rotate[0][0] + m[1] * Rotate[0][1] + m[2] * Rotate[0][2];
  Result[1] = m[0] * Rotate[1][0] + m[1] * Rotate[1][1] + m[2] * Rotate[1][2];
  Result[2] = m[0] * Rotate[2][0] + m[1] * Rotate[2][1] + m[2] * Rotate[2][2];
  Result[3] = m[3];
  return Result;
}
GLM_FN mat4 rotate(float angle, const vec3& v) { return rotate(mat4(1.0f), angle, v); }

// ---- quat (w, x, y, z constructor order; storage irrelevant here)
struct quat {
  float x, y, z, w;
  quat() = default;
  template <class A, class B, class C, class D,
            class = std::enable_if_t<std::is_arithmetic_v<A> && std::is_arithmetic_v<B> &&
                                     std::is_arithmetic_v<C> && std::is_arithmetic_v<D>>>
  GLM_CE quat(A w_, B x_, C y_, D z_)
      : x(static_cast<float>(x_)), y(static_cast<float>(y_)), z(static_cast<float>(z_)), w(static_cast<float>(w_))
  {
  }
};

GLM_CE mat4 mat4_cast(const quat& q)
{
  mat4 Result(1.0f);
  const float qxx(q.x * q.x), qyy(q.y * q.y), qzz(q.z * q.z);
  const float qxz(q.x * q.z), qxy(q.x * q.y), qyz(q.y * q.z);
  const float qwx(q.w * q.x), qwy(q.w * q.y), qwz(q.w * q.z);
  Result[0][0] = 1.0f - 2.0f * (qyy + qzz);
  Result[0][1] = 2.0f * (qxy + qwz);
  Result[0][2] = 2.0f * (qxz - qwy);
  Result[1][0] = 2.0f * (qxy - qwz);
  Result[1][1] = 1.0f - 2.0f * (qxx + qzz);
  Result[1][2] = 2.0f * (qyz + qwx);
  Result[2][0] = 2.0f * (qxz + qwy);
  Result[2][1] = 2.0f * (qyz - qwx);
  Result[2][2] = 1.0f - 2.0f * (qxx + qyy);
  return Result;
}

} // namespace glm
