// Minimal stand-in for Imath::Euler<T> (test infrastructure only).
// Only the ZXY static-frame order used at envutil_payload.cc:152 is needed; toQuat()
// restates Imath's published half-angle construction (Shoemake, Graphics Gems IV) for
// an even-parity, non-repeating, static-frame order with axes i,j,k = Z,X,Y.
#pragma once
#include <cmath>
#include "ImathQuat.h"
#include "ImathVec.h"
namespace Imath {
template <class T> struct Euler : public Vec3<T> {
  enum Order { XYZ = 0x0101, XZY = 0x0001, YZX = 0x1101, YXZ = 0x1001, ZXY = 0x2101, ZYX = 0x2001 };
  Order order;
  Euler(T xi, T yi, T zi, Order o = XYZ) : Vec3<T>(xi, yi, zi), order(o) {}
  Quat<T> toQuat() const {
    const bool parityEven = (order & 0x0100) != 0;
    const bool repeated = (order & 0x0010) != 0;
    const bool frameStatic = (order & 0x0001) != 0;
    int i = (order >> 12) & 3;
    int j = parityEven ? (i + 1) % 3 : (i > 0 ? i - 1 : 2);
    int k = parityEven ? (i > 0 ? i - 1 : 2) : (i + 1) % 3;
    Vec3<T> angles = frameStatic ? Vec3<T>(this->x, this->y, this->z) : Vec3<T>(this->z, this->y, this->x);
    if (!parityEven) angles.y = -angles.y;
    T ti = angles.x * T(0.5), tj = angles.y * T(0.5), th = angles.z * T(0.5);
    T ci = std::cos(ti), cj = std::cos(tj), ch = std::cos(th);
    T si = std::sin(ti), sj = std::sin(tj), sh = std::sin(th);
    T cc = ci * ch, cs = ci * sh, sc = si * ch, ss = si * sh;
    T parity = parityEven ? T(1) : T(-1);
    Quat<T> q;
    Vec3<T> a;
    if (repeated) {
      a[i] = cj * (cs + sc);
      a[j] = sj * (cc + ss) * parity;
      a[k] = sj * (cs - sc);
      q.r = cj * (cc - ss);
    } else {
      a[i] = cj * sc - sj * cs;
      a[j] = (cj * ss + sj * cc) * parity;
      a[k] = cj * cs - sj * sc;
      q.r = cj * cc + sj * ss;
    }
    q.v = a;
    return q;
  }
};
typedef Euler<float> Eulerf;
typedef Euler<double> Eulerd;
}  // namespace Imath
