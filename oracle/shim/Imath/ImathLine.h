// Minimal stand-in for Imath::Line3 (test infrastructure only). Referenced only by the
// never-instantiated deriv_tangential branch of twining.h:172-231.
#pragma once
#include "ImathVec.h"
namespace Imath {
template <class T> struct Line3 {
  Vec3<T> pos, dir;
  Vec3<T> closestPointTo(const Vec3<T>& point) const { return ((point - pos) ^ dir) * dir + pos; }
};
}  // namespace Imath
