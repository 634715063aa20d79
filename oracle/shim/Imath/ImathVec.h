// Minimal stand-in for Imath's Vec3 (test infrastructure only; see oracle/README.md).
// Only the members the reference's hot path touches (envutil_payload.cc:145-179,
// twining.h:194-225) are provided. Arithmetic follows Imath's published definitions.
#pragma once
#include <cstddef>
namespace Imath {
template <class T> struct Vec3 {
  T x, y, z;
  Vec3() : x(), y(), z() {}
  Vec3(T a, T b, T c) : x(a), y(b), z(c) {}
  template <class S> Vec3(const Vec3<S>& o) : x(T(o.x)), y(T(o.y)), z(T(o.z)) {}
  T& operator[](std::size_t i) { return (&x)[i]; }
  const T& operator[](std::size_t i) const { return (&x)[i]; }
  Vec3 operator+(const Vec3& o) const { return Vec3(x + o.x, y + o.y, z + o.z); }
  Vec3 operator-(const Vec3& o) const { return Vec3(x - o.x, y - o.y, z - o.z); }
  Vec3 operator-() const { return Vec3(-x, -y, -z); }
  Vec3 operator*(const T& s) const { return Vec3(x * s, y * s, z * s); }
  Vec3 operator/(const T& s) const { return Vec3(x / s, y / s, z / s); }
  T dot(const Vec3& o) const { return x * o.x + y * o.y + z * o.z; }
  T operator^(const Vec3& o) const { return dot(o); }
  Vec3 cross(const Vec3& o) const {
    return Vec3(y * o.z - z * o.y, z * o.x - x * o.z, x * o.y - y * o.x);
  }
};
template <class T> Vec3<T> operator*(const T& s, const Vec3<T>& v) { return Vec3<T>(s * v.x, s * v.y, s * v.z); }
typedef Vec3<float> V3f;
typedef Vec3<double> V3d;
}  // namespace Imath
