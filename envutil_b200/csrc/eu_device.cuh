// eu_device.cuh - per-pixel device functions of the reprojection pipeline (sm_100a).
//
// Every function restates one functor of the reference in scalar form, one thread per target
// pixel, with the reference's operation order and C-style promotions (float-vector op
// double-scalar is evaluated in double and narrowed, zimt/common.h:278). The translation unit
// is compiled with -fmad=false so that no multiply-add is contracted; the only fused
// operations are the explicit fmaf/fma calls inside include/eu_math.h.
#pragma once
#include <cuda_runtime.h>
#include <float.h>
#include <math.h>
#include <stdint.h>

#include "envutil_b200.h"
#include "eu_math.h"
#include "plan.h"

#define EU_PI_2 1.57079632679489661923
#define EU_PI 3.14159265358979323846

enum { CM_LEFT = 0, CM_RIGHT = 1, CM_TOP = 2, CM_BOTTOM = 3, CM_FRONT = 4, CM_BACK = 5 };

// a job's target / facet with the fields a compiled-in shape fixes (plan.h: eu_render_specs)
// replaced by constants, so that the switches on them below fold away
template <int SP>
__device__ __forceinline__ void dev_spec_target(TargetDev& T) {
  constexpr RenderSpec s = eu_render_specs[SP];
  if constexpr (s.tproj >= 0) T.projection = s.tproj;
  if constexpr (s.tnorm >= 0) T.normalize = s.tnorm;
}
template <int SP>
__device__ __forceinline__ void dev_spec_facet(FacetDev& F) {
  constexpr RenderSpec s = eu_render_specs[SP];
  if constexpr (s.skind >= 0) F.kind = s.skind;
  if constexpr (s.sproj >= 0) F.projection = s.sproj;
  if constexpr (s.bc0 >= 0) F.src.bc0 = s.bc0;
  if constexpr (s.bc1 >= 0) F.src.bc1 = s.bc1;
  if constexpr (s.mask_always >= 0) F.mask_always = s.mask_always;
  if constexpr (SP != 0) F.has_lcp = 0;
  if constexpr (SP != 0) F.fast_div = 3;  // proven for every facet of a job that matches a shape (api.cu)
}
template <int SP>
__device__ __forceinline__ decltype(auto) dev_facet_at(const FacetDev* __restrict__ fa, int i) {
  if constexpr (SP == 0) {
    return (fa[i]);
  } else {
    FacetDev F = fa[i];
    dev_spec_facet<SP>(F);
    return F;
  }
}

__device__ __forceinline__ float dev_norm3(const float v[3]) {  // zimt/xel.h:752-765
  float sqn = v[0] * v[0];
  sqn += v[1] * v[1];
  sqn += v[2] * v[2];
  return sqrtf(sqn);
}

// ------------------------------------------------------------------------------------------
// target side: the seven steppers (stepper.h:517-1578).
//
// Everything a stepper derives from the planar x coordinate alone is the same for a whole
// column, everything derived from planar y alone for a whole row. k_planar_tables evaluates
// those terms once per column / row with the very same functions (so the bits are the ones the
// per-pixel evaluation would produce) and the render kernel only combines them:
//   projection      col.a        col.b        row.a        row.b
//   spherical       sin(px)      cos(px)      sin(py)      cos(py)
//   cylindrical     sin(px)      cos(px)      py           -
//   rectilinear     px           -            py           -
//   fisheye/stereo  px           -            py           -          (not separable)
//   cubemap         px           -            p1           face       p1 = py + (3-face) section_md - refc_md
//   biatan6         tan(px pi/4) -            tan(p1 pi/4) face       (face = y / width, as integer bits)
// `first` is the column term of the same lane in the first vector of the pixel's 512-px segment
// (the cylindrical stepper keeps rcp_length from there, stepper.h:766-769).
// ------------------------------------------------------------------------------------------
struct ColTerm { float a, b; };
struct RowTerm { float a, b; };

__device__ __forceinline__ void dev_col_term(const TargetDev& T, float px, ColTerm& c) {
  c.a = px;
  c.b = 0.0f;
  if (T.projection == EU_SPHERICAL || T.projection == EU_CYLINDRICAL) eu_sincosf(px, &c.a, &c.b);
  else if (T.projection == EU_BIATAN6) c.a = eu_tanf(px * (float)(EU_PI / 4.0));
}
__device__ __forceinline__ void dev_row_term(const TargetDev& T, float py, int y, RowTerm& r) {
  r.a = py;
  r.b = 0.0f;
  if (T.projection == EU_SPHERICAL) {
    eu_sincosf(py, &r.a, &r.b);
  } else if (T.projection == EU_CUBEMAP || T.projection == EU_BIATAN6) {
    int face = y / T.width;  // stepper.h:1289-1294
    float p1 = py + (3 - face) * T.section_md - T.refc_md;
    if (T.projection == EU_BIATAN6) p1 = eu_tanf(p1 * (float)(EU_PI / 4.0));
    r.a = p1;
    r.b = __int_as_float(face);  // the face of the row, for dev_stepper (an integer division per pixel otherwise)
  }
}

__device__ __forceinline__ void dev_stepper(const TargetDev& T, const float* xx, const float* yy,
                                            const float* zz, ColTerm col, RowTerm row, ColTerm first, int y,
                                            float ray[3]) {
  switch (T.projection) {
    case EU_SPHERICAL: {
      float sy = row.a, r = row.b, sx = col.a, z = col.b;
#pragma unroll
      for (int i = 0; i < 3; i++) {
        float xxx = xx[i] * r, yyy = yy[i] * sy, zzz = zz[i] * r;
        ray[i] = xxx * sx + zzz * z + yyy;
      }
      break;
    }
    case EU_CYLINDRICAL: {
      float sx = col.a, z = col.b, py = row.a;
#pragma unroll
      for (int i = 0; i < 3; i++) ray[i] = xx[i] * sx + zz[i] * z + yy[i] * py;
      if (T.normalize) {
        float s0 = first.a, z0 = first.b, f3[3];
#pragma unroll
        for (int i = 0; i < 3; i++) f3[i] = xx[i] * s0 + zz[i] * z0 + yy[i] * py;
        float rcp = 1.0f / dev_norm3(f3);
#pragma unroll
        for (int i = 0; i < 3; i++) ray[i] *= rcp;
      }
      break;
    }
    case EU_RECTILINEAR: {
      float px = col.a, py = row.a;
#pragma unroll
      for (int i = 0; i < 3; i++) {
        float ddd = yy[i] * py + zz[i];
        ray[i] = xx[i] * px + ddd;
      }
      if (T.normalize) {
        float n = dev_norm3(ray);
#pragma unroll
        for (int i = 0; i < 3; i++) ray[i] /= n;
      }
      break;
    }
    case EU_FISHEYE:
    case EU_STEREOGRAPHIC: {
      float px = col.a, py = row.a;
      float sqn = px * px;
      sqn += py * py;
      float nrm = sqrtf(sqn);
      float a;
      if (T.projection == EU_FISHEYE)
        a = (float)(EU_PI_2 - (double)nrm);  // stepper.h:1019-1021
      else
        a = (float)(EU_PI_2 - 2.0 * eu_atan((double)nrm / 2.0));  // stepper.h:1146-1148
      float b = eu_atan2f(px, py);
      float z, r, sx, cy;
      eu_sincosf(a, &z, &r);
      eu_sincosf(b, &sx, &cy);
#pragma unroll
      for (int i = 0; i < 3; i++) ray[i] = xx[i] * r * sx + zz[i] * z + yy[i] * r * cy;
      break;
    }
    default: {  // EU_CUBEMAP, EU_BIATAN6 (stepper.h:1289-1345,1478-1560)
      const int face = __float_as_int(row.b);  // y / width, from the row table
      float p1 = row.a, p0 = col.a;
      float ccc[3], vvv[3];
#pragma unroll
      for (int i = 0; i < 3; i++) {
        // The +-1.0 literals promote these sums to double in the reference: float(double(a) + double(b)) for two
        // floats a, b. That is the float sum a + b, bit for bit - the double sum is exact unless the exponents are
        // more than 29 binades apart, and then both roundings return the larger operand (checked on 4e8 random
        // pairs including subnormals and infinities) - so the conversions and the double adds are left out.
        switch (face) {
          case CM_LEFT: ccc[i] = -xx[i] + p1 * yy[i]; vvv[i] = zz[i]; break;
          case CM_RIGHT: ccc[i] = xx[i] + p1 * yy[i]; vvv[i] = -zz[i]; break;
          case CM_TOP: ccc[i] = -yy[i] - p1 * zz[i]; vvv[i] = -xx[i]; break;
          case CM_BOTTOM: ccc[i] = yy[i] + p1 * zz[i]; vvv[i] = -xx[i]; break;
          case CM_FRONT: ccc[i] = p1 * yy[i] + zz[i]; vvv[i] = xx[i]; break;
          default: ccc[i] = p1 * yy[i] - zz[i]; vvv[i] = -xx[i]; break;
        }
      }
#pragma unroll
      for (int i = 0; i < 3; i++) ray[i] = ccc[i] + p0 * vvv[i];
      if (T.normalize) {
        float n = dev_norm3(ray);
#pragma unroll
        for (int i = 0; i < 3; i++) ray[i] /= n;
      }
    }
  }
}

// The cubemap / biatan6 stepper of a single-facet job with the face's vectors read from RenderParams::cube_tab
// (plan.h): bit for bit what the switch in dev_stepper computes.
__device__ __forceinline__ void dev_stepper_cube_tab(const TargetDev& T, const float (*tab)[12], ColTerm col, RowTerm row,
                                                     float ray[3]) {
  const int face = __float_as_int(row.b);
  const float4 a = *reinterpret_cast<const float4*>(tab[face]);
  const float4 b = *reinterpret_cast<const float4*>(tab[face] + 4);
  const float4 c = *reinterpret_cast<const float4*>(tab[face] + 8);
  const float p1 = row.a, p0 = col.a;
  const float ccc0 = a.x + p1 * b.x, ccc1 = a.y + p1 * b.y, ccc2 = a.z + p1 * b.z;
  ray[0] = ccc0 + p0 * c.x;
  ray[1] = ccc1 + p0 * c.y;
  ray[2] = ccc2 + p0 * c.z;
  if (T.normalize) {
    float n = dev_norm3(ray);
#pragma unroll
    for (int i = 0; i < 3; i++) ray[i] /= n;
  }
}

// rotate(xel_t<float,3>, r3_t<float>), geometry.h:74-82
__device__ __forceinline__ void dev_rot3(const float v[3], const float* __restrict__ m, float out[3]) {
  float o[3];
#pragma unroll
  for (int c = 0; c < 3; c++) o[c] = (v[0] * m[c] + v[1] * m[3 + c]) + v[2] * m[6 + c];
  out[0] = o[0]; out[1] = o[1]; out[2] = o[2];
}

// generic_stepper (stepper.h:353-470) over tf_ex_facet::eval (envutil_payload.cc:1841-1883): the
// bare planar coordinate -> X_to_ray of the target projection (geometry.h:151-567) -> tf3d_t::eval
// (geometry.h:1886-1925). Used for facets with PanoTools translation (TrX/TrY/TrZ).
template <int ORDER>
__device__ __forceinline__ void dev_weights(const float* __restrict__ wmat, float delta, float w[ORDER]);

// inverse_lcp::eval (lens_correction.h:396-405). Its argument is norm(out) / s with a double s: a DOUBLE
// vector, so the reduction to spline coordinates runs in double and is narrowed where the float
// evaluator takes it; clamp gate (NATURAL, zimt/eval.h:2101-2110), cubic 1-D window (eval.h:937-960)
__device__ __forceinline__ float dev_inv_lcp_factor(const InvPlanarDev& P, float radius) {
  double in = (double)radius / P.s;
  in = in / P.rr_max;
  in = sqrt(in);
  in *= (double)(EU_INV_NK - 1);
  float cx = (float)in;
  const float lower = 0.0f, upper = (float)(EU_INV_NK - 1);
  if (cx < lower) cx = lower;
  else if (cx > upper) cx = upper;
  float fl = floorf(cx), t = cx - fl;
  int ix = (int)fl;
  float w[4];
  dev_weights<4>(P.wm, t, w);
  const float* __restrict__ c = P.coef + (ix - 1);
  float sum = __ldg(c);
  sum *= w[0];
#pragma unroll
  for (int i = 1; i < 4; i++) sum += w[i] * __ldg(c + i);
  sum += 1.0f;
  return sum;
}

__device__ __forceinline__ void dev_generic_ray(const TargetDev& T, const InvPlanarDev& IP, const FacetDev& F, float h,
                                                float v, float ray[3]) {
  float in[3];  // RIGHT, DOWN, FORWARD
  if (IP.on) {  // tf22: pto_planar<T, L, true>::eval, environment.h:285-307
    if (IP.has_shear) {  // float vector op double scalar is evaluated in double (gen_simd_type.h:274-316)
      v = (float)(((double)v - IP.shear_t * (double)h) / (1 - IP.shear_t * IP.shear_g));
      h = (float)((double)h - IP.shear_g * (double)v);
    }
    if (IP.has_shift) {  // operator-= narrows its scalar operand first (vector_common.h:302-316)
      h -= IP.h;
      v -= IP.v;
    }
    if (IP.has_lcp) {
      float sqn = h * h;
      sqn += v * v;
      float factor = dev_inv_lcp_factor(IP, sqrtf(sqn));
      h *= factor;
      v *= factor;
    }
  }
  switch (T.projection) {
    case EU_SPHERICAL: {
      float sinlat, coslat, sinlon, coslon;
      eu_sincosf(v, &sinlat, &coslat);
      eu_sincosf(h, &sinlon, &coslon);
      in[0] = sinlon * coslat; in[2] = coslon * coslat; in[1] = sinlat;
      break;
    }
    case EU_CYLINDRICAL: in[2] = eu_cosf(h); in[0] = eu_sinf(h); in[1] = v; break;
    case EU_RECTILINEAR: in[0] = h; in[1] = v; in[2] = 1.0f; break;
    case EU_STEREOGRAPHIC: {
      float r = sqrtf(h * h + v * v);
      float theta = eu_atanf(r / 2.0f) * 2.0f;
      float phi = eu_atan2f(h, -v);
      in[2] = eu_cosf(theta);
      in[1] = -eu_sinf(theta) * eu_cosf(phi);
      in[0] = eu_sinf(theta) * eu_sinf(phi);
      break;
    }
    case EU_FISHEYE: {
      float r = sqrtf(h * h + v * v);
      float phi = eu_atan2f(h, -v);
      in[2] = eu_cosf(r);
      in[1] = -eu_sinf(r) * eu_cosf(phi);
      in[0] = eu_sinf(r) * eu_sinf(phi);
      break;
    }
    default: {  // EU_CUBEMAP, EU_BIATAN6: ir_to_ray_t / ba6_to_ray_t with their default metrics (section 2.0,
                // reference centre 1.0: roll_out_23, geometry.h:1800-1834,660-775,857-990). The double members
                // only ever meet floats in exact operations (a halving, small even integers, 1.0), so float is it
      h += 1.0f;
      v += 6.0f;
      const int section = (int)(v / 2.0f);
      v -= (float)section * 2.0f;
      h -= 1.0f;
      v -= 1.0f;
      if (T.projection == EU_BIATAN6) {
        h = eu_tanf(h * (float)(EU_PI / 4.0));
        v = eu_tanf(v * (float)(EU_PI / 4.0));
      }
      switch (section) {
        case CM_LEFT: in[0] = -1.0f; in[1] = v; in[2] = h; break;
        case CM_RIGHT: in[0] = 1.0f; in[1] = v; in[2] = -h; break;
        case CM_TOP: in[0] = -h; in[1] = -1.0f; in[2] = -v; break;
        case CM_BOTTOM: in[0] = -h; in[1] = 1.0f; in[2] = v; break;
        case CM_FRONT: in[0] = h; in[1] = v; in[2] = 1.0f; break;
        default: in[0] = -h; in[1] = v; in[2] = -1.0f; break;
      }
    }
  }
  float out[3] = {in[0], in[1], in[2]};
  for (int k = 0; k < F.g_nstage; k++) {  // generic_r3: tf3d_t::eval per stage, geometry.h:1886-1925
    const FacetDev::TfStage& S = F.g_st[k];
    if (!S.has_shift) {
      dev_rot3(out, S.ab, out);
      continue;
    }
    dev_rot3(out, S.a, out);
    if (out[2] <= 0.0f) {
      out[0] = 0.0f; out[1] = 0.0f; out[2] = -INFINITY;
    } else {
      out[0] /= out[2];
      out[1] /= out[2];
      out[2] = 1.0f;
#pragma unroll
      for (int c = 0; c < 3; c++) out[c] *= S.dcp;
#pragma unroll
      for (int c = 0; c < 3; c++) out[c] -= S.shift[c];
      dev_rot3(out, S.b, out);
    }
  }
  if (T.normalize) {
    float n = dev_norm3(out);
#pragma unroll
    for (int c = 0; c < 3; c++) out[c] /= n;
  }
  ray[0] = out[0]; ray[1] = out[1]; ray[2] = out[2];
}

// ------------------------------------------------------------------------------------------
// source side
// ------------------------------------------------------------------------------------------

// eu_atan2f(y, x) for an x whose sign bit is clear (a square root of a sum of squares: +0 at least, or NaN - and
// a NaN result does not depend on the skipped step): eu_atan2f without its "x negative" correction.
__device__ __forceinline__ float dev_atan2f_xpos(float y, float x) {
  float ax = eu_fabsf(x), ay = eu_fabsf(y);
  float mx = ax > ay ? ax : ay;
  float mn = ax > ay ? ay : ax;
  float t = (mx == 0.0f) ? 0.0f : mn / mx;
  float p = eu_katanf(t);
  if (ay > ax) p = (EU_PIO2_HI - p) + EU_PIO2_LO;
  return eu_copysignf(p, y);
}

// mount_t::get_coordinate_nomask (environment.h:1077-1110) over the ray_to_X functors
// (geometry.h:277-534) and pto_planar's forward path (environment.h:254-283)
__device__ __forceinline__ void dev_mount_coordinate(const FacetDev& F, const float r[3], float c[2]) {
  switch (F.projection) {
    case EU_RECTILINEAR:
      c[0] = r[0] / r[2];
      c[1] = r[1] / r[2];
      break;
    case EU_SPHERICAL: {
      float s = sqrtf(r[0] * r[0] + r[2] * r[2]);
      c[1] = dev_atan2f_xpos(r[1], s);
      c[0] = eu_atan2f(r[0], r[2]);
      break;
    }
    case EU_CYLINDRICAL: {
      float s = sqrtf(r[0] * r[0] + r[2] * r[2]);
      c[1] = r[1] / s;
      c[0] = eu_atan2f(r[0], r[2]);
      break;
    }
    case EU_STEREOGRAPHIC: {
      float rn = 1.0f / sqrtf(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
      float right = r[0] * rn, down = r[1] * rn, fwd = r[2] * rn;
      float factor = 2.0f / (fwd + 1.0f);
      c[0] = right * factor;
      c[1] = down * factor;
      break;
    }
    default: {  // EU_FISHEYE
      float s = sqrtf(r[0] * r[0] + r[1] * r[1]);
      float rr = (float)EU_PI_2 - eu_atan2f(r[2], s);
      float phi = eu_atan2f(r[1], r[0]);
      float sp, cp;
      eu_sincosf(phi, &sp, &cp);
      c[0] = rr * cp;
      c[1] = rr * sp;
    }
  }
  if (F.has_lcp) {
    float sqn = c[0] * c[0];
    sqn += c[1] * c[1];
    float x = sqrtf(sqn) / F.lcp_s;
    float sum = 0.0f, power = 1.0f;  // eu_polynomial::function, lens_correction.h:94-105
#pragma unroll
    for (int i = 0; i <= 3; i++) {
      sum += F.lcp[3 - i] * power;
      power *= x;
    }
    c[0] *= sum;
    c[1] *= sum;
    if (F.has_shift) {
      c[0] += F.shift_h;
      c[1] += F.shift_v;
    }
    if (F.has_shear) {
      float h0 = (float)((double)c[0] + (double)c[1] * F.shear_g);
      float h1 = (float)((double)c[1] + (double)c[0] * F.shear_t);
      c[0] = h0;
      c[1] = h1;
    }
  }
}

// source_t::test_crd (environment.h:970-978) + the z > 0 test of rectilinear mounts (:1123-1127)
// A ray the generic stepper has invalidated - (0, 0, -inf), normalised to (0, 0, NaN) - misses every mounted image in
// the reference: its NaN is the x86 default NaN, whose SIGN BIT IS SET, so atan2(0, NaN) takes the "x negative"
// branch and yields pi - a latitude no image covers (the other projections produce NaN coordinates, which fail
// every comparison). NVIDIA GPUs produce the positive canonical NaN, atan2(0, NaN) would be 0 and the ray would
// hit the image centre: hence the explicit test. (Found by the random sweep: a `--single` job on a translated
// facet, seed 12 job 136.)
__device__ __forceinline__ bool dev_mount_mask(const FacetDev& F, const float r[3], const float c[2]) {
  bool m = (c[0] >= F.win_x0) && (c[0] <= F.win_x1) && (c[1] >= F.win_y0) && (c[1] <= F.win_y1);
  if (F.projection == EU_RECTILINEAR) m = m && (r[2] > 0.0f);
  else m = m && (r[2] == r[2]);
  return m;
}

// mount_t::get_mask for the synopses that only need to know WHETHER a ray hits (_voronoi_syn tests every facet and
// evaluates one). Rectilinear mounts - photographs, the common case - decide most rays without the two IEEE
// divisions of ray_to_rect_t: a ray with z <= 0 (or NaN) misses, and an approximate quotient (2 ulp) that clears
// the window's edges by a margin of 1e-5 of the edge coordinate (host: FacetDev::win_margin) gives the same
// answer as the correctly rounded one, because rounding is monotonic; only rays inside that margin take the
// exact path. The decision is the reference's for every ray.
__device__ __forceinline__ bool dev_facet_mask(const FacetDev& F, const float r[3]) {
  if (F.mask_always) return true;
  if (F.projection == EU_RECTILINEAR && !F.has_lcp) {
    if (!(r[2] > 0.0f)) return false;
    const float tx = __fdividef(r[0], r[2]), ty = __fdividef(r[1], r[2]);
    const float mx = F.win_margin[0], my = F.win_margin[1];
    if (tx < F.win_x0 - mx || tx > F.win_x1 + mx || ty < F.win_y0 - my || ty > F.win_y1 + my) return false;
    if (tx > F.win_x0 + mx && tx < F.win_x1 - mx && ty > F.win_y0 + my && ty < F.win_y1 - my) return true;
  }
  float c[2];
  dev_mount_coordinate(F, r, c);
  return dev_mount_mask(F, r, c);
}

// coordinate gates, zimt/map.h (vector variants)
__device__ __forceinline__ float dev_vfmod(float lhs, float rhs) {
  float help = lhs;
  help /= rhs;
  help = truncf(help);
  help *= rhs;
  lhs -= help;
  if (fabsf(lhs) >= fabsf(rhs)) lhs = 0.0f;
  return lhs;
}
__device__ __forceinline__ float dev_gate(float c, int bc, float upper) {
  // an axis of extent 1 is gated as CONSTANT with both limits 0: a clamp that always yields 0
  // (build_safe_ev, zimt/eval.h:2060-2068)
  if (bc == EU_BC_CONST0) return 0.0f;
  const float lower = -0.5f;
  float cc = c - lower;
  float w = upper - lower;
  if (bc == EU_BC_PERIODIC) {
    bool below = cc < 0.0f, above = cc >= w;
    if (below || above) {
      float cm = dev_vfmod(cc, w);
      if (below) cm = cm + w;
      if (cm >= w) cm = 0.0f;
      cc = cm;
    }
  } else {
    cc = fabsf(cc);
    if (cc >= w) {
      float cm = dev_vfmod(cc, 2 * w);
      cm -= w;
      cm = fabsf(cm);
      cm = w - cm;
      cc = cm;
    }
  }
  return cc + lower;
}

// one texel at p. TS: floats from one texel to the next (NCH, or 4 = 16-byte texels: one 128-bit
// load). SMEM: 1 = p points into the block's shared-memory tile, 0 = into HBM (read-only path), 2 = into HBM
// that this very kernel is writing (loads that bypass the non-coherent caches: the ordered cubemap fill).
template <int NCH, int TS, int SMEM>
__device__ __forceinline__ void dev_load_texel(const float* __restrict__ p, float v[NCH]) {
  if constexpr (TS == 4) {
    float4 t = SMEM == 1 ? *reinterpret_cast<const float4*>(p)
                         : (SMEM == 2 ? __ldcg(reinterpret_cast<const float4*>(p)) : __ldg(reinterpret_cast<const float4*>(p)));
    v[0] = t.x;
    if constexpr (NCH > 1) v[1] = t.y;
    if constexpr (NCH > 2) v[2] = t.z;
    if constexpr (NCH > 3) v[3] = t.w;
  } else {
#pragma unroll
    for (int c = 0; c < NCH; c++) v[c] = SMEM == 1 ? p[c] : (SMEM == 2 ? __ldcg(p + c) : __ldg(p + c));
  }
}

// b-spline weights for one axis: basis_functor::operator()(result, delta), zimt/basis.h:650-689.
// wmat lives in the kernel parameter block: with ORDER fixed the operands are constant-bank reads.
template <int ORDER>
__device__ __forceinline__ void dev_weights(const float* __restrict__ wmat, float delta, float w[ORDER]) {
  float power = delta;
#pragma unroll
  for (int k = 0; k < ORDER; k++) w[k] = wmat[k];
#pragma unroll
  for (int row = 1; row < ORDER; row++) {
#pragma unroll
    for (int k = 0; k < ORDER; k++) w[k] += power * wmat[row * ORDER + k];
    if (row < ORDER - 1) power *= delta;
  }
}

// ---- opt-in arithmetic variant (libenvutil_b200_fma.so, built with -DEU_CONTRACT_WINDOW) ----------------
// The default library rounds every product and every sum of the window evaluation separately, as the
// reference's parity build does (-ffp-contract=off): that is what makes the output bit-identical, and it
// is 198 of the C2 kernel's 469 instructions per warp. A reference built the usual way (g++ -O3 with FMA
// hardware: -ffp-contract=fast) fuses those pairs itself. With EU_CONTRACT_WINDOW the weights of the
// WINDOW, the window sum and the twining accumulation use fused multiply-adds; rays, source coordinates,
// gates, window positions, face and facet indices stay exactly as they are (same bits), so the variant
// differs from the default by a few ulp of each pixel value and in nothing else.
// WSTD (cubic only): the plan builder has checked that the weight matrix has the cubic b-spline's exact zeros at
// [0][3], [1][1], [1][3] and [2][3] (zimt/basis.h:419-543 yields them as exact zeros). delta is a remainder
// c - floor(c) >= +0 (or NaN), so each skipped term is w + (+0) = w for a w that is never -0, and w[3] = 0 + t = t:
// the same bits with 7 instructions fewer per axis. (A NaN delta still reaches every weight through a later row.)
#ifdef EU_CONTRACT_WINDOW
#define EU_WIN_MULADD(a, b, c) __fmaf_rn((a), (b), (c))
#define EU_WIN_ACC(c, a, b) c = __fmaf_rn((a), (b), c)
template <int ORDER, bool WSTD = false>
__device__ __forceinline__ void dev_window_weights(const float* __restrict__ wmat, float delta, float w[ORDER]) {
  if constexpr (ORDER == 4 && WSTD) {
    const float p1 = delta, p2 = p1 * delta, p3 = p2 * delta;
    w[0] = __fmaf_rn(p3, wmat[12], __fmaf_rn(p2, wmat[8], __fmaf_rn(p1, wmat[4], wmat[0])));
    w[1] = __fmaf_rn(p3, wmat[13], __fmaf_rn(p2, wmat[9], wmat[1]));
    w[2] = __fmaf_rn(p3, wmat[14], __fmaf_rn(p2, wmat[10], __fmaf_rn(p1, wmat[6], wmat[2])));
    w[3] = p3 * wmat[15];
    return;
  }
  float power = delta;
#pragma unroll
  for (int k = 0; k < ORDER; k++) w[k] = wmat[k];
#pragma unroll
  for (int row = 1; row < ORDER; row++) {
#pragma unroll
    for (int k = 0; k < ORDER; k++) w[k] = __fmaf_rn(power, wmat[row * ORDER + k], w[k]);
    if (row < ORDER - 1) power *= delta;
  }
}
#else
#define EU_WIN_MULADD(a, b, c) ((c) + (a) * (b))
#define EU_WIN_ACC(c, a, b) c += (a) * (b)  // the statement as the run-time evaluator has always had it
template <int ORDER, bool WSTD = false>
__device__ __forceinline__ void dev_window_weights(const float* __restrict__ wmat, float delta, float w[ORDER]) {
  if constexpr (ORDER == 4 && WSTD) {
    const float p1 = delta, p2 = p1 * delta, p3 = p2 * delta;
    w[0] = wmat[0] + p1 * wmat[4];
    w[1] = wmat[1];
    w[2] = wmat[2] + p1 * wmat[6];
    w[0] += p2 * wmat[8];
    w[1] += p2 * wmat[9];
    w[2] += p2 * wmat[10];
    w[0] += p3 * wmat[12];
    w[1] += p3 * wmat[13];
    w[2] += p3 * wmat[14];
    w[3] = p3 * wmat[15];
    return;
  }
  dev_weights<ORDER>(wmat, delta, w);
}
#endif

// Where a spline coordinate lands: gates (zimt/eval.h:2039-2164) + split (zimt/basis.h:102-146).
struct Located {
  int ix, iy;    // integral part: the window covers [ix - deg/2, ix - deg/2 + deg] (zimt/eval.h:732)
  float fx, fy;  // remainder fed to the weight functor
};
__device__ __forceinline__ Located dev_locate(const SourceDev& S, int degree, float cx, float cy) {
  cx = dev_gate(cx, S.bc0, S.upper_x);
  cy = dev_gate(cy, S.bc1, S.upper_y);
  Located L;
  if (degree & 1) {
    float f = floorf(cx); L.fx = cx - f; L.ix = (int)f;
    f = floorf(cy); L.fy = cy - f; L.iy = (int)f;
  } else {
    float f = roundf(cx); L.fx = cx - f; L.ix = (int)f;
    f = roundf(cy); L.fy = cy - f; L.iy = (int)f;
  }
  return L;
}

// the window sum proper, weights given (zimt/eval.h:903-996: rows left to right, then down the column of row sums)
template <int NCH, int TS, int DEG, int SMEM>
__device__ __forceinline__ void dev_window_sum_w(const float* __restrict__ p0, int pitch, const float wx[DEG + 1],
                                                 const float wy[DEG + 1], float out[NCH]) {
  constexpr int ORDER = DEG + 1;
  float t[ORDER][ORDER][NCH];
#pragma unroll
  for (int j = 0; j < ORDER; j++) {
    const float* __restrict__ row = p0 + (ptrdiff_t)j * pitch;
#pragma unroll
    for (int i = 0; i < ORDER; i++) dev_load_texel<NCH, TS, SMEM>(row + i * TS, t[j][i]);
  }
#pragma unroll
  for (int j = 0; j < ORDER; j++) {
    float sub[NCH];
#pragma unroll
    for (int c = 0; c < NCH; c++) sub[c] = t[j][0][c] * wx[0];
#pragma unroll
    for (int i = 1; i < ORDER; i++) {
#pragma unroll
      for (int c = 0; c < NCH; c++) sub[c] = EU_WIN_MULADD(wx[i], t[j][i][c], sub[c]);
    }
    if (j == 0) {
#pragma unroll
      for (int c = 0; c < NCH; c++) {
        out[c] = sub[c];
        out[c] *= wy[0];
      }
    } else {
#pragma unroll
      for (int c = 0; c < NCH; c++) out[c] = EU_WIN_MULADD(sub[c], wy[j], out[c]);
    }
  }
}

// evaluator::eval for a fixed degree > 1: window sum in the reference's order
// (zimt/eval.h:903-996). p0 -> texel (ix - deg/2, iy - deg/2); pitch: floats per row.
template <int NCH, int TS, int DEG, int SMEM, bool WSTD = false>
__device__ __forceinline__ void dev_window_sum(const float* __restrict__ p0, int pitch, const float* __restrict__ wmat,
                                               float fx, float fy, float out[NCH]) {
  constexpr int ORDER = DEG + 1;
  float wx[ORDER], wy[ORDER];
  dev_window_weights<ORDER, WSTD>(wmat, fx, wx);
  dev_window_weights<ORDER, WSTD>(wmat, fy, wy);
  dev_window_sum_w<NCH, TS, DEG, SMEM>(p0, pitch, wx, wy, out);
}

template <int NCH, int TS, int SMEM>
__device__ __forceinline__ void dev_eval_linear(const float* __restrict__ p, int pitch, float fx, float fy,
                                                float out[NCH]) {  // _eval_linear, zimt/eval.h:1004-1059
  float p00[NCH], p10[NCH], p01[NCH], p11[NCH];
  dev_load_texel<NCH, TS, SMEM>(p, p00);
  dev_load_texel<NCH, TS, SMEM>(p + TS, p10);
  dev_load_texel<NCH, TS, SMEM>(p + pitch, p01);
  dev_load_texel<NCH, TS, SMEM>(p + pitch + TS, p11);
  float wl0 = 1.0f - fx, wr0 = fx, wl1 = 1.0f - fy, wr1 = fy;
#pragma unroll
  for (int c = 0; c < NCH; c++) {
    float sum = p00[c];
    sum *= wl0;
    sum = EU_WIN_MULADD(p10[c], wr0, sum);
    sum *= wl1;
    float sub = p01[c];
    sub *= wl0;
    sub = EU_WIN_MULADD(p11[c], wr0, sub);
    sum = EU_WIN_MULADD(sub, wr1, sum);
    out[c] = sum;
  }
}

// the window evaluation for a located coordinate; p0 -> texel (ix - degree/2, iy - degree/2).
// DEG >= 0: degree fixed at compile time; DEG < 0: read at run time (all degrees 0..7).
// WSTD: the kernel is compiled for a job shape (the plan builder picks those only for the standard cubic matrix)
template <int NCH, int TS, int DEG, int SMEM, bool WSTD = false>
__device__ __forceinline__ void dev_window_eval(const float* __restrict__ p0, int pitch, int degree,
                                                const float* __restrict__ wmat, float fx, float fy, float out[NCH]) {
  if constexpr (DEG == 1) {
    dev_eval_linear<NCH, TS, SMEM>(p0, pitch, fx, fy, out);
  } else if constexpr (DEG == 3) {
    dev_window_sum<NCH, TS, 3, SMEM, WSTD>(p0, pitch, wmat, fx, fy, out);
  } else {
    switch (degree) {
      case 0: dev_load_texel<NCH, TS, SMEM>(p0, out); break;
      case 1: dev_eval_linear<NCH, TS, SMEM>(p0, pitch, fx, fy, out); break;
      case 2: dev_window_sum<NCH, TS, 2, SMEM>(p0, pitch, wmat, fx, fy, out); break;
      case 3: dev_window_sum<NCH, TS, 3, SMEM>(p0, pitch, wmat, fx, fy, out); break;
      case 4: dev_window_sum<NCH, TS, 4, SMEM>(p0, pitch, wmat, fx, fy, out); break;
      case 5: dev_window_sum<NCH, TS, 5, SMEM>(p0, pitch, wmat, fx, fy, out); break;
      case 6: dev_window_sum<NCH, TS, 6, SMEM>(p0, pitch, wmat, fx, fy, out); break;
      default: dev_window_sum<NCH, TS, 7, SMEM>(p0, pitch, wmat, fx, fy, out); break;
    }
  }
}

// safe evaluator = mapper + evaluator (zimt/eval.h:2039-2164, :1237-1300), gathering from HBM
// I32: the container holds fewer than 2^31 floats, so the window's offset from the core fits 32 bits (the plan
// builder checks that before it picks a kernel compiled this way): one 64-bit multiply-add instead of three
template <int NCH, int TS, int DEG, int SPACE = 0, bool I32 = false>
__device__ __forceinline__ void dev_spline_eval(const SourceDev& S, int degree, const float* __restrict__ wmat,
                                                float cx, float cy, float out[NCH]) {
  if constexpr (DEG >= 0) degree = DEG;
  Located L = dev_locate(S, degree, cx, cy);
  const int h2 = degree / 2;
  const float* p0;
  if constexpr (I32) p0 = S.core + ((L.iy - h2) * S.stride + (L.ix - h2) * TS);
  else p0 = S.core + (ptrdiff_t)(L.iy - h2) * S.stride + (ptrdiff_t)(L.ix - h2) * TS;
  dev_window_eval<NCH, TS, DEG, SPACE, I32>(p0, S.stride, degree, wmat, L.fx, L.fy, out);
}

// eu_atanf for an argument in [-1, 1] or NaN - what dev_cubeface yields: each in-face coordinate is a component
// divided by one of no smaller magnitude, and rounding is monotonic. eu_atanf's |x| > 1 branch (a division and a
// correction by pi/2, include/eu_math.h) never runs for such arguments; these are its remaining operations.
__device__ __forceinline__ float dev_atanf_unit(float x) { return eu_copysignf(eu_katanf(eu_fabsf(x)), x); }

// ray_to_cubeface, geometry.h:1178-1357 (>= ties favour x over y over z)
__device__ __forceinline__ void dev_cubeface(const float c[3], int& face, float in_face[2]) {
  bool m1 = fabsf(c[0]) >= fabsf(c[1]);
  bool m2 = fabsf(c[0]) >= fabsf(c[2]);
  bool m3 = fabsf(c[1]) >= fabsf(c[2]);
  if (m1 && m2) {
    face = c[0] < 0.0f ? CM_LEFT : CM_RIGHT;
    in_face[0] = -c[2] / c[0];
    in_face[1] = c[1] / fabsf(c[0]);
  } else if (!m2 && !m3) {
    face = c[2] < 0.0f ? CM_BACK : CM_FRONT;
    in_face[0] = c[0] / c[2];
    in_face[1] = c[1] / fabsf(c[2]);
  } else {
    face = c[1] < 0.0f ? CM_TOP : CM_BOTTOM;
    in_face[0] = -c[0] / fabsf(c[1]);
    in_face[1] = c[2] / c[1];
  }
}

// x / y for a divisor that is a constant of the facet: q = x * RN(1/y), one exact residual (fma), one correction
// (fma) - Markstein's sequence, three instructions where IEEE division costs ten. It yields the correctly rounded
// quotient RN(x / y), i.e. the very bits of `x / y`, whenever the host has PROVEN that for this y: the set-up code
// runs the sequence over all 2^23 significands of x against the division (api.cu, exact_by_reciprocal; the
// result scales exactly with the exponent of x, and the operands here are far from the subnormal and overflow
// ranges) and sets `ok` only then. Otherwise the division itself is used.
__device__ __forceinline__ float dev_div_const(float x, float y, float rcp, int ok) {
  if (ok) {
    float q = x * rcp;
    float r = __fmaf_rn(-q, y, x);
    return __fmaf_rn(r, rcp, q);
  }
  return x / y;
}

// First half of environment::eval (environment.h:1821-1842): ray -> spline coordinate of the
// facet's source, via mount_t (:1172-1196, md_to_spline :988-1006) or cubemap_view_t
// (:1452-1486). Returns false when the ray misses a mounted image; `face` = cube face or -1.
__device__ __forceinline__ bool dev_facet_coordinate(const FacetDev& F, const float r[3], int& face, float& cx,
                                                     float& cy) {
  face = -1;
  if (F.kind == EU_SRC_MOUNT) {
    float c[2];
    dev_mount_coordinate(F, r, c);
    if (!dev_mount_mask(F, r, c)) return false;
    float ix = (float)((double)c[0] - F.ext_x0);
    ix = dev_div_const(ix, F.ext_w, F.rcp_w, F.fast_div & 1);
    ix *= F.total_w;
    ix -= .5f;
    ix = ix - F.win_xoff;
    float iy = (float)((double)c[1] - F.ext_y0);
    iy = dev_div_const(iy, F.ext_h, F.rcp_h, F.fast_div & 2);
    iy *= F.total_h;
    iy -= .5f;
    iy = iy - F.win_yoff;
    cx = ix;
    cy = iy;
  } else {
    float in_face[2], pk[2];
    dev_cubeface(r, face, in_face);
    if (F.kind == EU_SRC_BIATAN6) {
      in_face[0] = (float)(4.0 / EU_PI) * dev_atanf_unit(in_face[0]);
      in_face[1] = (float)(4.0 / EU_PI) * dev_atanf_unit(in_face[1]);
    }
    pk[0] = in_face[0] + F.refc_md;
    pk[1] = in_face[1] + F.refc_md;
    pk[0] *= F.model_to_px;
    pk[1] *= F.model_to_px;
    pk[1] += (float)(face * F.section_px);
    pk[0] -= .5f;
    pk[1] -= .5f;
    cx = pk[0];
    cy = pk[1];
  }
  return true;
}

template <int NCH>
__device__ __forceinline__ void dev_brighten(const FacetDev& F, float px[NCH]) {
  if (F.brighten != 1.0f) {
    constexpr int NCOL = (NCH == 2 || NCH == 4) ? NCH - 1 : NCH;
#pragma unroll
    for (int i = 0; i < NCOL; i++) px[i] *= F.brighten;
  }
}

// environment::eval, gathering from HBM. Returns the cube face hit, or -1.
template <int NCH, int TS, int DEG, bool I32 = false>
__device__ __forceinline__ int dev_facet_eval(const FacetDev& F, int degree, const float* __restrict__ wmat,
                                              const float r[3], float px[NCH]) {
  int face;
  float cx, cy;
  if (!dev_facet_coordinate(F, r, face, cx, cy)) {
#pragma unroll
    for (int i = 0; i < NCH; i++) px[i] = 0.0f;
    return -1;
  }
  dev_spline_eval<NCH, TS, DEG, 0, I32>(F.src, degree, wmat, cx, cy, px);
  dev_brighten<NCH>(F, px);
  return face;
}

// environment::eval in two halves, for callers that evaluate several facets of IDENTICAL geometry (the exposure
// brackets of one camera position under hdr_merge): where the ray lands is computed once, the window is read
// from each facet's own container. Same operations in the same order as dev_facet_eval, hence the same bits.
template <int DEG>
__device__ __forceinline__ bool dev_facet_locate(const FacetDev& F, int degree, const float r[3], Located& L) {
  int face;
  float cx, cy;
  if (!dev_facet_coordinate(F, r, face, cx, cy)) return false;
  if constexpr (DEG >= 0) degree = DEG;
  L = dev_locate(F.src, degree, cx, cy);
  return true;
}
template <int NCH, int TS, int DEG, bool I32>
__device__ __forceinline__ void dev_facet_window(const FacetDev& F, int degree, const float* __restrict__ wmat,
                                                 const Located& L, float px[NCH]) {
  if constexpr (DEG >= 0) degree = DEG;
  const int h2 = degree / 2;
  const float* p0;
  if constexpr (I32) p0 = F.src.core + ((L.iy - h2) * F.src.stride + (L.ix - h2) * TS);
  else p0 = F.src.core + (ptrdiff_t)(L.iy - h2) * F.src.stride + (ptrdiff_t)(L.ix - h2) * TS;
  dev_window_eval<NCH, TS, DEG, 0, I32>(p0, F.src.stride, degree, wmat, L.fx, L.fy, px);
  dev_brighten<NCH>(F, px);
}

// ---- the general build: any mix of channel counts and texel strides, degree at run time ------
// safe evaluator with the texel stride and the degree read at run time (same operation order as
// the specialised code above; loops instead of unrolled windows)
template <int SNCH>
__device__ __noinline__ void dev_spline_eval_rt(const SourceDev& S, int degree, const float* __restrict__ wmat,
                                                float cx, float cy, float out[SNCH]) {
  Located L = dev_locate(S, degree, cx, cy);
  const int ts = S.tstride, h2 = degree / 2, order = degree + 1;
  const float* __restrict__ p0 = S.core + (ptrdiff_t)(L.iy - h2) * S.stride + (ptrdiff_t)(L.ix - h2) * ts;
  if (degree == 0) {
#pragma unroll
    for (int c = 0; c < SNCH; c++) out[c] = __ldg(p0 + c);
    return;
  }
  if (degree == 1) {  // _eval_linear, zimt/eval.h:1004-1059
    float wl0 = 1.0f - L.fx, wr0 = L.fx, wl1 = 1.0f - L.fy, wr1 = L.fy;
#pragma unroll
    for (int c = 0; c < SNCH; c++) {
      float sum = __ldg(p0 + c);
      sum *= wl0;
      EU_WIN_ACC(sum, __ldg(p0 + ts + c), wr0);
      sum *= wl1;
      float sub = __ldg(p0 + S.stride + c);
      sub *= wl0;
      EU_WIN_ACC(sub, __ldg(p0 + S.stride + ts + c), wr0);
      EU_WIN_ACC(sum, sub, wr1);
      out[c] = sum;
    }
    return;
  }
  float wx[EU_MAX_DEGREE + 1], wy[EU_MAX_DEGREE + 1];
  for (int axis = 0; axis < 2; axis++) {  // basis_functor, zimt/basis.h:650-689
    float* w = axis ? wy : wx;
    float delta = axis ? L.fy : L.fx;
    float power = delta;
    for (int k = 0; k < order; k++) w[k] = wmat[k];
    for (int row = 1; row < order; row++) {
      for (int k = 0; k < order; k++) EU_WIN_ACC(w[k], power, wmat[row * order + k]);
      if (row < order - 1) power *= delta;
    }
  }
#pragma unroll
  for (int c = 0; c < SNCH; c++) {  // _eval, zimt/eval.h:903-996
    float sum = 0.0f;
    for (int j = 0; j < order; j++) {
      const float* __restrict__ row = p0 + (ptrdiff_t)j * S.stride + c;
      float sub = __ldg(row);
      sub *= wx[0];
      for (int i = 1; i < order; i++) EU_WIN_ACC(sub, wx[i], __ldg(row + i * ts));
      if (j == 0) {
        sum = sub;
        sum *= wy[0];
      } else {
        EU_WIN_ACC(sum, sub, wy[j]);
      }
    }
    out[c] = sum;
  }
}

// repix_t, environment.h:1205-1309: a facet pixel of in_n channels as a pixel of OUT channels
template <int OUT>
__device__ __forceinline__ void dev_repix(int in_n, const float in[4], float out[OUT]) {
  if (in_n == OUT) {
#pragma unroll
    for (int i = 0; i < OUT; i++) out[i] = in[i];
    return;
  }
  float o[4] = {0.f, 0.f, 0.f, 0.f};
  switch (in_n) {
    case 1:
      if (OUT == 3) { o[0] = o[1] = o[2] = in[0]; }
      else if (OUT == 2) { o[0] = in[0]; o[1] = 1.0f; }
      else { o[0] = o[1] = o[2] = in[0]; o[3] = 1.0f; }
      break;
    case 2:
      if (OUT == 1) { o[0] = in[0] / in[1]; if (in[1] == 0.0f) o[0] = 0.0f; }
      else if (OUT == 3) { float g = in[0] / in[1]; if (in[1] == 0.0f) g = 0.0f; o[0] = o[1] = o[2] = g; }
      else { o[0] = o[1] = o[2] = in[0]; o[3] = in[1]; }
      break;
    case 3: {
      float sum = in[0];
      sum += in[1];
      sum += in[2];
      if (OUT == 1) o[0] = sum / 3.0f;
      else if (OUT == 2) { o[0] = sum / 3.0f; o[1] = 1.0f; }
      else { o[0] = in[0]; o[1] = in[1]; o[2] = in[2]; o[3] = 1.0f; }
      break;
    }
    default:
      if (OUT == 1) { o[0] = (in[0] + in[1] + in[2]) / 3.0f; o[0] /= in[3]; if (in[3] == 0.0f) o[0] = 0.0f; }
      else if (OUT == 2) { o[0] = (in[0] + in[1] + in[2]) / 3.0f; o[1] = in[3]; }
      else {
        o[0] = in[0] / in[3]; o[1] = in[1] / in[3]; o[2] = in[2] / in[3];
        if (in[3] == 0.0f) o[0] = o[1] = o[2] = 0.0f;
      }
  }
#pragma unroll
  for (int i = 0; i < OUT; i++) out[i] = o[i];
}

// mono_t, environment.h:1325-1384: what replaces repix_t for masked facets (jobs of one or two channels)
template <int OUT>
__device__ __forceinline__ void dev_mono(int in_n, const float in[4], float out[OUT]) {
  float o[4] = {0.f, 0.f, 0.f, 0.f};
  if (in_n == OUT) {
    o[0] = in[0]; o[1] = in[1]; o[2] = in[2]; o[3] = in[3];
  } else if (in_n == 1) {
    o[0] = in[0]; o[1] = 1.0f;
  } else if (in_n == 2) {
    o[0] = in[0] / in[1];
    if (in[1] == 0.0f) o[0] = 0.0f;
  } else if (in_n == 3) {
    o[0] = in[0]; o[1] = 1.0f;
  } else {
    if (OUT == 1) {
      o[0] = in[0];
      o[0] /= in[3];
      if (in[3] == 0.0f) o[0] = 0.0f;
    } else {
      o[0] = in[0]; o[1] = in[3];
    }
  }
#pragma unroll
  for (int i = 0; i < OUT; i++) out[i] = o[i];
}

// environment::eval for any facet: evaluate in the source's channel count, repix, brighten
template <int NCH>
__device__ __forceinline__ int dev_facet_eval_general(const FacetDev& F, int degree, const float* __restrict__ wmat,
                                                      const float r[3], float px[NCH]) {
  int face;
  float cx, cy, sp[4] = {0.f, 0.f, 0.f, 0.f};
  bool hit = dev_facet_coordinate(F, r, face, cx, cy);
  if (hit) {
    switch (F.src.nch) {
      case 1: dev_spline_eval_rt<1>(F.src, degree, wmat, cx, cy, sp); break;
      case 2: dev_spline_eval_rt<2>(F.src, degree, wmat, cx, cy, sp); break;
      case 3: dev_spline_eval_rt<3>(F.src, degree, wmat, cx, cy, sp); break;
      default: dev_spline_eval_rt<4>(F.src, degree, wmat, cx, cy, sp); break;
    }
  }
  if (F.masked) {
    // --mask_for: masking_t paints every channel, alpha_masking_t the colour channels times the interpolated alpha
    // (masking.h:70-139; a miss stays the zero pixel: mount_t clears it after the evaluation)
    if (hit) {
      const int n = F.src.nch;
      if (n == 1 || n == 3) {
        sp[0] = sp[1] = sp[2] = F.paint;
      } else {
        sp[0] = F.paint * sp[n - 1];
        if (n == 4) sp[1] = sp[2] = sp[0];
      }
    }
    dev_mono<NCH>(F.src.nch, sp, px);
  } else {
    dev_repix<NCH>(F.src.nch, sp, px);  // a miss is a zero pixel of the SOURCE type (environment.h:1190-1193)
  }
  dev_brighten<NCH>(F, px);
  return hit ? face : -1;
}

// _hdr_merge_syn::get_quality, envutil_payload.cc:1390-1442
__device__ __forceinline__ float dev_hdr_quality(float grey, float optimum, int kind) {
  bool large = grey > optimum;
  float distance = fabsf(optimum - grey);
  if (kind == EU_HDR_LOW && !large) distance = 0.0f;
  if (kind == EU_HDR_HIGH && large) distance = 0.0f;
  float proximity = optimum - distance;
  return proximity / (optimum * optimum);
}
