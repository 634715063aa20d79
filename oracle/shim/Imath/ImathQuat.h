// Minimal stand-in for Imath::Quat (test infrastructure only).
// v*q follows Imath's operator*(Vec3,Quat): v + 2(q.r*(q.v x v) + q.v x (q.v x v)),
// the same formula the reference spells out in geometry.cc:218-238 ("mulq").
#pragma once
#include "ImathVec.h"
namespace Imath {
template <class T> struct Quat {
  T r;
  Vec3<T> v;
  Quat() : r(1), v(0, 0, 0) {}
  Quat(T s, T i, T j, T k) : r(s), v(i, j, k) {}
  Quat(T s, const Vec3<T>& d) : r(s), v(d) {}
  template <class S> Quat(const Quat<S>& q) : r(T(q.r)), v(T(q.v.x), T(q.v.y), T(q.v.z)) {}
  Quat& invert() {
    T qdot = r * r + (v ^ v);
    r /= qdot;
    v = -v / qdot;
    return *this;
  }
};
template <class T> Vec3<T> operator*(const Vec3<T>& v, const Quat<T>& q) {
  Vec3<T> a = q.v.cross(v);
  Vec3<T> b = q.v.cross(a);
  return v + T(2) * (q.r * a + b);
}
typedef Quat<float> Quatf;
typedef Quat<double> Quatd;
}  // namespace Imath
