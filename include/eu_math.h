/* eu_math.h - the elementary functions of the B200 back-end (numerical contract).
 *
 * envutil's per-pixel results depend on which SIMD back-end supplies sin/cos/tan/atan/atan2:
 * highway's hwy/contrib/math polynomials, Vc's, std::simd's, or libm through zimt's "goading"
 * loops (reference zimt/simd/hwy_simd_type.h:1521-1615 vs zimt/simd/vector_common.h:203-272).
 * The results differ in the last ulp, and because a source coordinate of magnitude ~10^4 texels
 * has an ulp of ~10^-3 texel, a one-ulp difference in an angle is a 1e-4 difference in a pixel
 * of a noisy image. This header therefore SPECIFIES the float32 functions this back-end uses,
 * as fixed sequences of IEEE-754 binary32 operations (+, -, *, /, fma, compare/select): the
 * same inputs give bit-identical outputs on the host and on sm_100a. They are accurate to
 * about 1 ulp (tools/gen_eu_math_coeffs.py derives the coefficients; tests/test_eu_math.py
 * measures the error against libm in double precision).
 *
 * Used by: the CUDA kernels (envutil_b200/csrc), the host-side set-up code, and - as test
 * infrastructure - oracle/eu_math_interpose.c, which substitutes them for libm's
 * sinf/cosf/sincosf/tanf/atanf/atan2f in the "pinned math" build of the reference.
 *
 * Requirements on the including translation unit: floating-point contraction must be OFF
 * (gcc: -ffp-contract=off, nvcc: -fmad=false); fmaf must be a true fused multiply-add.
 */
#ifndef EU_MATH_H
#define EU_MATH_H

#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define EU_HD __host__ __device__ __forceinline__
#else
#define EU_HD static inline
#endif

EU_HD uint32_t eu_f2u(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(f);
#else
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
#endif
}
EU_HD float eu_u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  float f;
  memcpy(&f, &u, 4);
  return f;
#endif
}
EU_HD float eu_fabsf(float x) { return eu_u2f(eu_f2u(x) & 0x7fffffffu); }
EU_HD float eu_copysignf(float mag, float sgn) {
  return eu_u2f((eu_f2u(mag) & 0x7fffffffu) | (eu_f2u(sgn) & 0x80000000u));
}

/* ---- argument reduction x = k*(pi/2) + r, |r| <= pi/4 (Cody-Waite, three binary32 parts of
 * pi/2, exact products through fma). Adequate for |x| up to ~1e4; the pipeline only ever
 * passes angles of a few pi. k is obtained with the 1.5*2^23 rounding constant, i.e.
 * round-to-nearest-even of x*2/pi, as plain arithmetic. */
EU_HD float eu_reduce_pio2f(float x, int32_t* quadrant) {
  const float TWO_OVER_PI = 0x1.45f306p-1f;
  const float PIO2_1 = 0x1.921fb6p+0f;
  const float PIO2_2 = -0x1.777a5cp-25f;
  const float PIO2_3 = -0x1.ee59dap-50f;
  const float MAGIC = 12582912.0f; /* 1.5 * 2^23 */
  float t = fmaf(x, TWO_OVER_PI, MAGIC);
  float k = t - MAGIC;
  *quadrant = (int32_t)eu_f2u(t); /* low bits of the integer part live in the mantissa */
  float r = fmaf(-k, PIO2_1, x);
  r = fmaf(-k, PIO2_2, r);
  r = fmaf(-k, PIO2_3, r);
  return r;
}

/* sin and cos of a reduced argument |r| <= pi/4 */
EU_HD float eu_ksinf(float r) {
  const float S0 = -0x1.555556p-3f, S1 = 0x1.111108p-7f, S2 = -0x1.a00f1ep-13f, S3 = 0x1.6cbaf8p-19f;
  float u = r * r;
  float p = fmaf(u, S3, S2);
  p = fmaf(u, p, S1);
  p = fmaf(u, p, S0);
  return fmaf(r * u, p, r);
}
EU_HD float eu_kcosf(float r) {
  const float C0 = 0x1.555556p-5f, C1 = -0x1.6c16b8p-10f, C2 = 0x1.a010a4p-16f, C3 = -0x1.241246p-22f;
  float u = r * r;
  float p = fmaf(u, C3, C2);
  p = fmaf(u, p, C1);
  p = fmaf(u, p, C0);
  return fmaf(u * u, p, fmaf(u, -0.5f, 1.0f));
}

EU_HD void eu_sincosf(float x, float* s, float* c) {
  int32_t q;
  float r = eu_reduce_pio2f(x, &q);
  float sr = eu_ksinf(r);
  float cr = eu_kcosf(r);
  float ss = (q & 1) ? cr : sr;
  float cc = (q & 1) ? sr : cr;
  if (q & 2) ss = -ss;
  if ((q + 1) & 2) cc = -cc;
  *s = ss;
  *c = cc;
}
EU_HD float eu_sinf(float x) {
  float s, c;
  eu_sincosf(x, &s, &c);
  return s;
}
EU_HD float eu_cosf(float x) {
  float s, c;
  eu_sincosf(x, &s, &c);
  return c;
}
/* tan of a reduced argument |r| <= pi/4: r + r^3 * T(r^2); odd quadrants give -1/tan(r) */
EU_HD float eu_ktanf(float r) {
  const float T0 = 0x1.555564p-2f, T1 = 0x1.110ccap-3f, T2 = 0x1.bafdbcp-5f, T3 = 0x1.5b561ap-6f,
              T4 = 0x1.6a07f8p-7f, T5 = -0x1.77bbcp-13f, T6 = 0x1.277c48p-8f;
  float u = r * r;
  float p = fmaf(u, T6, T5);
  p = fmaf(u, p, T4);
  p = fmaf(u, p, T3);
  p = fmaf(u, p, T2);
  p = fmaf(u, p, T1);
  p = fmaf(u, p, T0);
  return fmaf(r * u, p, r);
}
EU_HD float eu_tanf(float x) {
  int32_t q;
  float r = eu_reduce_pio2f(x, &q);
  float t = eu_ktanf(r);
  return (q & 1) ? (-1.0f / t) : t;
}

/* atan of t in [0,1]: t + t^3 * A(t^2) */
EU_HD float eu_katanf(float t) {
  const float A0 = -0x1.55553cp-2f, A1 = 0x1.9991aep-3f, A2 = -0x1.241e96p-3f, A3 = 0x1.c07a8p-4f,
              A4 = -0x1.57db3cp-4f, A5 = 0x1.d99f9ap-5f, A6 = -0x1.fc8baap-6f, A7 = 0x1.634176p-7f,
              A8 = -0x1.d1f9f4p-10f;
  float u = t * t;
  float p = fmaf(u, A8, A7);
  p = fmaf(u, p, A6);
  p = fmaf(u, p, A5);
  p = fmaf(u, p, A4);
  p = fmaf(u, p, A3);
  p = fmaf(u, p, A2);
  p = fmaf(u, p, A1);
  p = fmaf(u, p, A0);
  return fmaf(t * u, p, t);
}

#define EU_PIO2_HI 0x1.921fb6p+0f
#define EU_PIO2_LO -0x1.777a5cp-25f
#define EU_PI_HI 0x1.921fb6p+1f
#define EU_PI_LO -0x1.777a5cp-24f

EU_HD float eu_atanf(float x) {
  float a = eu_fabsf(x);
  int big = a > 1.0f;
  float t = big ? 1.0f / a : a;
  float p = eu_katanf(t);
  if (big) p = (EU_PIO2_HI - p) + EU_PIO2_LO;
  return eu_copysignf(p, x);
}

EU_HD float eu_atan2f(float y, float x) {
  float ax = eu_fabsf(x), ay = eu_fabsf(y);
  float mx = ax > ay ? ax : ay;
  float mn = ax > ay ? ay : ax;
  float t = (mx == 0.0f) ? 0.0f : mn / mx;
  float p = eu_katanf(t);
  if (ay > ax) p = (EU_PIO2_HI - p) + EU_PIO2_LO;
  if (eu_f2u(x) & 0x80000000u) p = (EU_PI_HI - p) + EU_PI_LO;
  return eu_copysignf(p, y);
}


/* ---- double-precision atan. One per-pixel call site of the reference works in double:
 * stereographic_stepper computes  a = M_PI_2 - 2.0 * atan ( norm(planar) / 2.0 )  with a
 * double-promoted operand (reference stepper.h:1146-1151, promotion rules zimt/common.h:278).
 * The set-up arithmetic (get_vfov / get_step, envutil_basic.cc:50-156) uses the same function for
 * its double atan, so that host, device and the pinned reference build see one definition.
 * Specified here as a fixed sequence of binary64 operations so host and device agree bit for
 * bit: |x|>1 -> pi/2 - atan(1/|x|); t>tan(pi/8) -> pi/4 + atan((t-1)/(t+1)); then the Taylor
 * series of atan on |u| <= tan(pi/8) with 24 terms (truncation error < 1e-20). */
EU_HD double eu_atan(double x) {
  double a = fabs(x);
  int big = a > 1.0;
  double t = big ? 1.0 / a : a;
  int mid = t > 0.41421356237309503;
  double u = mid ? (t - 1.0) / (t + 1.0) : t;
  double s = u * u;
  double q = -1.0 / 47.0;
  q = fma(q, s, 1.0 / 45.0);
  q = fma(q, s, -1.0 / 43.0);
  q = fma(q, s, 1.0 / 41.0);
  q = fma(q, s, -1.0 / 39.0);
  q = fma(q, s, 1.0 / 37.0);
  q = fma(q, s, -1.0 / 35.0);
  q = fma(q, s, 1.0 / 33.0);
  q = fma(q, s, -1.0 / 31.0);
  q = fma(q, s, 1.0 / 29.0);
  q = fma(q, s, -1.0 / 27.0);
  q = fma(q, s, 1.0 / 25.0);
  q = fma(q, s, -1.0 / 23.0);
  q = fma(q, s, 1.0 / 21.0);
  q = fma(q, s, -1.0 / 19.0);
  q = fma(q, s, 1.0 / 17.0);
  q = fma(q, s, -1.0 / 15.0);
  q = fma(q, s, 1.0 / 13.0);
  q = fma(q, s, -1.0 / 11.0);
  q = fma(q, s, 1.0 / 9.0);
  q = fma(q, s, -1.0 / 7.0);
  q = fma(q, s, 1.0 / 5.0);
  q = fma(q, s, -1.0 / 3.0);
  double r = fma(u * s, q, u);
  if (mid) r = (0x1.921fb54442d18p-1 + r) + 0x1.1a62633145c07p-55;
  if (big) r = (0x1.921fb54442d18p+0 - r) + 0x1.1a62633145c07p-54;
  return copysign(r, x);
}

#endif /* EU_MATH_H */
