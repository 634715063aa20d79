/* eu_oracle.c - TEST INFRASTRUCTURE (see eu_oracle.h). Plain-C, scalar restatement of the
 * reference's per-pixel reprojection path. Each function cites the reference code it follows.
 *
 * Arithmetic rules followed throughout (they decide the last bit of every result):
 *  - the reference's SIMD types follow C promotion: float-vector (op) double-scalar is computed
 *    in double and narrowed on assignment (zimt/common.h:278, zimt/simd/gen_simd_type.h:274-321);
 *    compound assignments and comparisons against a scalar narrow the scalar to float first
 *    (zimt/simd/vector_common.h:302-316, zimt/simd/vector_mask.h:50-74);
 *  - the strict reference build uses no fused multiply-add (-ffp-contract=off, no -march);
 *    this file must be compiled with -ffp-contract=off as well;
 *  - sin/cos/tan/atan/atan2 on floats are the functions of include/eu_math.h, the same ones
 *    oracle/_ref/envutil_ref_pm is linked against.
 */
#include "eu_oracle.h"

#include <float.h>
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "eu_math.h"

#define ORC_MAX_DEGREE 7
#define ORC_SEGMENT 512 /* WIELDING_SEGMENT_SIZE, zimt/bill.h:69 */
#define ORC_LANES 16    /* zimt/simd.h:106-123 for the goading back-end */

enum { BC_PERIODIC = 0, BC_REFLECT = 1, BC_NATURAL = 2, BC_MIRROR = 3 };
enum { KIND_MOUNT = 0, KIND_CUBEMAP = 1, KIND_BIATAN6 = 2 };
/* face_index_t, envutil_basic.h:56-64 */
enum { CM_LEFT = 0, CM_RIGHT = 1, CM_TOP = 2, CM_BOTTOM = 3, CM_FRONT = 4, CM_BACK = 5 };

/* ------------------------------------------------------------------------------------------
 * set-up arithmetic
 * ---------------------------------------------------------------------------------------- */

/* get_vfov, envutil_basic.cc:50-110 (the CUBEMAP case falls through to default) */
static double orc_get_vfov(int prj, int w, int h, double hfov) {
  switch (prj) {
    case EU_RECTILINEAR: return 2.0 * eu_atan(h * tan(hfov / 2.0) / w);
    case EU_CYLINDRICAL: {
      double ppr = w / hfov;
      double hr = h / ppr;
      return 2.0 * eu_atan(hr / 2.0);
    }
    case EU_STEREOGRAPHIC: {
      double wr = 2.0 * tan(hfov / 4.0);
      double ppr = w / wr;
      double hr = h / ppr;
      return 4.0 * eu_atan(hr / 2.0);
    }
    case EU_SPHERICAL:
    case EU_FISHEYE: return hfov * h / w;
    default: return hfov;
  }
}

/* get_step, envutil_basic.cc:112-156 */
double orc_get_step(int prj, int w, int h, double hfov) {
  (void)h;
  switch (prj) {
    case EU_RECTILINEAR:
    case EU_CUBEMAP: return eu_atan(2.0 * tan(hfov / 2.0) / w);
    case EU_BIATAN6:
    case EU_SPHERICAL:
    case EU_CYLINDRICAL:
    case EU_FISHEYE: return hfov / w;
    case EU_STEREOGRAPHIC: return eu_atan(4.0 * tan(hfov / 4.0) / w);
    default: return 0.0;
  }
}

/* get_extent, envutil_basic.cc:158-226 */
void orc_get_extent(int prj, int w, int h, double hfov, double e[4]) {
  double ax = -hfov / 2.0, bx = hfov / 2.0;
  double by = orc_get_vfov(prj, w, h, hfov) / 2.0, ay = -by;
  switch (prj) {
    case EU_SPHERICAL:
    case EU_FISHEYE: e[0] = ax; e[1] = bx; e[2] = ay; e[3] = by; break;
    case EU_CYLINDRICAL: e[0] = ax; e[1] = bx; e[2] = tan(ay); e[3] = tan(by); break;
    case EU_RECTILINEAR: e[0] = tan(ax); e[1] = tan(bx); e[2] = tan(ay); e[3] = tan(by); break;
    case EU_STEREOGRAPHIC:
      e[0] = 2.0 * tan(ax / 2.0); e[1] = 2.0 * tan(bx / 2.0);
      e[2] = 2.0 * tan(ay / 2.0); e[3] = 2.0 * tan(by / 2.0);
      break;
    case EU_CUBEMAP:
    case EU_BIATAN6: e[0] = tan(ax); e[1] = tan(bx); e[2] = 6 * e[0]; e[3] = 6 * e[1]; break;
    default: e[0] = e[1] = e[2] = e[3] = 0.0;
  }
}

/* rotate_3d / make_r3_t, envutil_payload.cc:136-218. Imath::Eulerf(roll,pitch,yaw,ZXY)
 * .toQuat() in FLOAT (static frame, even parity, axes i,j,k = Z,X,Y), optional
 * Quat::invert(), then rows = e_k * Quat<double>(q). */
void orc_rotation(double roll_d, double pitch_d, double yaw_d, int inverse, double m[9]) {
  float ti = (float)roll_d * 0.5f, tj = (float)pitch_d * 0.5f, th = (float)yaw_d * 0.5f;
  float ci, cj, ch, si, sj, sh;
  eu_sincosf(ti, &si, &ci);
  eu_sincosf(tj, &sj, &cj);
  eu_sincosf(th, &sh, &ch);
  float cc = ci * ch, cs = ci * sh, sc = si * ch, ss = si * sh;
  float q[3], qr;
  q[2] = cj * sc - sj * cs;
  q[0] = cj * ss + sj * cc;
  q[1] = cj * cs - sj * sc;
  qr = cj * cc + sj * ss;
  /* Imath::Quat<T> q with T = double (make_r3_t is called with the double members of args /
   * facet_spec): the float quaternion is widened first, invert() then runs in double */
  double r = qr, v0 = q[0], v1 = q[1], v2 = q[2];
  if (inverse) {
    double d = r * r + ((v0 * v0 + v1 * v1) + v2 * v2);
    r /= d;
    v0 = -v0 / d;
    v1 = -v1 / d;
    v2 = -v2 / d;
  }
  for (int k = 0; k < 3; k++) {
    double e[3] = {0, 0, 0};
    e[k] = 1.0;
    double a[3] = {v1 * e[2] - v2 * e[1], v2 * e[0] - v0 * e[2], v0 * e[1] - v1 * e[0]};
    double b[3] = {v1 * a[2] - v2 * a[1], v2 * a[0] - v0 * a[2], v0 * a[1] - v1 * a[0]};
    for (int c = 0; c < 3; c++) m[3 * k + c] = e[c] + 2.0 * (r * a[c] + b[c]);
  }
}

/* rotate(r3, r3), geometry.h:79-97: row_i = lhs[i][0]*rhs[0] + lhs[i][1]*rhs[1] + lhs[i][2]*rhs[2] */
static void mat_mul(const double a[9], const double b[9], double m[9]) {
  for (int i = 0; i < 3; i++)
    for (int c = 0; c < 3; c++)
      m[3 * i + c] = (a[3 * i] * b[c] + a[3 * i + 1] * b[3 + c]) + a[3 * i + 2] * b[6 + c];
}

/* B-spline basis of degree n at x2/2 (what zimt/basis.h calls bspline_basis_2), by the
 * Cox-de Boor recursion on the cardinal spline, in long double. */
static long double basis2(int x2, int n) {
  /* N_n(t) on knots 0..n+1, centred: b(x) = N_n(x + (n+1)/2) */
  long double t = (long double)x2 / 2.0L + (long double)(n + 1) / 2.0L;
  long double v[ORC_MAX_DEGREE + 3];
  for (int i = 0; i <= n; i++) v[i] = (t >= i && t < i + 1) ? 1.0L : 0.0L;
  for (int k = 1; k <= n; k++)
    for (int i = 0; i <= n - k; i++)
      v[i] = ((t - i) / k) * v[i] + ((i + k + 1 - t) / k) * v[i + 1];
  return v[0];
}

/* poles of the b-spline prefilter: roots z, |z|<1, of sum_k b_n(k) z^(k+m) (what
 * zimt/poles.h tabulates), largest magnitude first (poles.h:1311-1334). Newton iteration in
 * long double on the deflated polynomial, polished on the full one. */
int orc_poles(int degree, long double* poles) {
  int m = degree / 2;
  if (degree < 2 || degree > ORC_MAX_DEGREE) return 0;
  long double c[2 * ORC_MAX_DEGREE + 1];
  int n = 2 * m;
  for (int k = -m; k <= m; k++) c[k + m] = basis2(2 * k, degree);
  /* all roots are real and negative, in reciprocal pairs: scan for sign changes in (-1,0) */
  int found = 0;
  long double prev_x = -1.0L, prev_v = 0;
  {
    long double v = 0;
    for (int i = n; i >= 0; i--) v = v * prev_x + c[i];
    prev_v = v;
  }
  const int steps = 200000;
  for (int s = 1; s <= steps && found < m; s++) {
    long double x = -1.0L + (long double)s / steps;
    long double v = 0;
    for (int i = n; i >= 0; i--) v = v * x + c[i];
    if ((v < 0) != (prev_v < 0)) {
      long double lo = prev_x, hi = x;
      for (int it = 0; it < 200; it++) {
        long double mid = 0.5L * (lo + hi), vm = 0;
        for (int i = n; i >= 0; i--) vm = vm * mid + c[i];
        if ((vm < 0) == (prev_v < 0)) lo = mid; else hi = mid;
      }
      long double r = 0.5L * (lo + hi);
      for (int it = 0; it < 4; it++) { /* Newton polish */
        long double f = 0, d = 0;
        for (int i = n; i >= 0; i--) { d = d * r + f; f = f * r + c[i]; }
        if (d != 0) r -= f / d;
      }
      poles[found++] = r;
    }
    prev_x = x;
    prev_v = v;
  }
  /* scan went from -1 upward, so poles[] is already sorted by descending magnitude */
  return found;
}

/* basis_functor::calculate_weight_matrix, zimt/basis.h:419-543 (derivative 0), narrowed to
 * float as the evaluator's basis_functor<float> holds it. m[row*(degree+1)+k]. */
void orc_weight_matrix(int degree, float* mtx) {
  int order = degree + 1;
  long double der_line[ORC_MAX_DEGREE + 1];
  long double faculty = 1;
  for (int row = 0; row < order; row++) {
    if (row > 1) faculty *= row;
    int first = 0;
    int mm = degree - row;
    if (mm == 0) {
      der_line[0] = 1;
      first = 1;
    } else if (degree & 1) {
      for (int x2 = -mm + 1; x2 <= mm - 1; x2 += 2) der_line[first++] = basis2(x2, mm);
    } else {
      for (int x2 = -mm; x2 <= mm; x2 += 2) der_line[first++] = basis2(x2, mm);
    }
    for (int p = first; p < order; p++) der_line[p] = 0;
    for (int d = mm; d < degree; d++) {
      int put = first, pick = first - 1;
      while (pick >= 0) {
        der_line[put] = der_line[pick] - der_line[put];
        --put;
        --pick;
      }
      der_line[put] = -der_line[put];
      first++;
    }
    for (int k = 0; k < order; k++) mtx[row * order + k] = (float)(der_line[k] / faculty);
  }
}

/* metrics_t, cubemap.h:233-400 */
typedef struct {
  int face_px, section_px, left_frame_px, right_frame_px;
  double model_to_px, px_to_model, section_md, refc_md;
} cm_metrics_t;

static void cm_metrics(int face_px, double face_fov, int support_min, int tile_px, cm_metrics_t* m) {
  double overscan_md = 0.0, diameter_md = 2.0;
  if (face_fov > M_PI_2) {
    double radius_md = tan(face_fov / 2.0);
    diameter_md = 2.0 * radius_md;
    overscan_md = radius_md - 1.0;
  }
  m->face_px = face_px;
  m->model_to_px = (double)face_px / diameter_md;
  m->px_to_model = diameter_md / (double)face_px;
  long inherent = (long)trunc(m->model_to_px * overscan_md);
  long additional = inherent < support_min ? support_min - inherent : 0;
  long px_min = face_px + 2 * additional;
  long n_tiles = px_min / tile_px;
  if (n_tiles * tile_px < px_min) n_tiles++;
  m->section_px = (int)(n_tiles * tile_px);
  long frame_total = m->section_px - face_px;
  m->left_frame_px = (int)(frame_total / 2);
  m->right_frame_px = (int)(frame_total - m->left_frame_px);
  m->section_md = m->px_to_model * m->section_px;
  m->refc_md = m->px_to_model * ((double)m->left_frame_px + (double)face_px / 2.0);
}

/* ------------------------------------------------------------------------------------------
 * recursive prefilter (zimt/recursive.h) on one line of floats with stride
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int bc, npoles;
  float pole[ORC_MAX_DEGREE / 2 + 1];
  long double pole_x[ORC_MAX_DEGREE / 2 + 1];
  int horizon[ORC_MAX_DEGREE / 2 + 1];
  float gain;
} iir_t;

/* iir_filter ctor, recursive.h:774-862; overall_gain :93-103 */
static void iir_setup(iir_t* f, int bc, int degree, long double tolerance) {
  f->bc = bc;
  f->npoles = degree / 2;
  orc_poles(degree, f->pole_x);
  long double lambda = 1;
  for (int k = 0; k < f->npoles; k++) {
    f->pole[k] = (float)f->pole_x[k];
    if (tolerance > 0)
      f->horizon[k] = (int)ceill(logl(tolerance) / logl(fabsl(f->pole_x[k])));
    else
      f->horizon[k] = INT_MAX;
    lambda *= (1 - f->pole_x[k]) * (1 - 1 / f->pole_x[k]);
  }
  f->gain = (float)lambda;
}

#define C(n) c[(size_t)(n) * st]

/* initial causal coefficient, recursive.h:321-583 */
static float iir_icc(const iir_t* f, const float* c, size_t st, int M, int k) {
  float z = f->pole[k], zn, z2n, iz, Sum;
  int n, hz = f->horizon[k];
  switch (f->bc) {
    case BC_MIRROR:
      if (hz < M) {
        zn = z; Sum = C(0);
        for (n = 1; n < hz; n++) { Sum += zn * C(n); zn *= z; }
      } else {
        zn = z; iz = 1.0f / z;
        z2n = (float)powl(f->pole_x[k], (long double)(M - 1));
        Sum = C(0) + z2n * C(M - 1);
        z2n *= z2n * iz;
        for (n = 1; n <= M - 2; n++) { Sum += (zn + z2n) * C(n); zn *= z; z2n *= iz; }
        Sum /= (1.0f - zn * zn);
      }
      return Sum;
    case BC_NATURAL:
      if (hz < M) {
        float c02 = C(0) + C(0);
        zn = z; Sum = C(0);
        for (n = 1; n < hz; n++) { Sum += zn * (c02 - C(n)); zn *= z; }
        return Sum;
      } else {
        zn = z; iz = 1.0f / z;
        z2n = (float)powl(f->pole_x[k], (long double)(M - 1));
        Sum = ((1.0f + z) / (1.0f - z)) * (C(0) - z2n * C(M - 1));
        z2n *= z2n * iz;
        for (n = 1; n <= M - 2; n++) { Sum -= (zn - z2n) * C(n); zn *= z; z2n *= iz; }
        return Sum / (1.0f - zn * zn);
      }
    case BC_REFLECT:
      if (hz < M) {
        zn = z; Sum = C(0);
        for (n = 0; n < hz; n++) { Sum += zn * C(n); zn *= z; }
        return Sum;
      } else {
        zn = z; iz = 1.0f / z;
        z2n = (float)powl(f->pole_x[k], (long double)(2 * M));
        Sum = 0;
        for (n = 0; n < M - 1; n++) { Sum += (zn + z2n) * C(n); zn *= z; z2n *= iz; }
        Sum += (zn + z2n) * C(n);
        return C(0) + Sum / (1.0f - zn * zn);
      }
    case BC_PERIODIC:
    default:
      if (hz < M) {
        zn = z; Sum = C(0);
        for (n = M - 1; n > (M - hz); n--) { Sum += zn * C(n); zn *= z; }
      } else {
        zn = z; Sum = C(0);
        for (n = M - 1; n > 0; n--) { Sum += zn * C(n); zn *= z; }
        Sum /= (1.0f - zn);
      }
      return Sum;
  }
}

/* initial anticausal coefficient, recursive.h:360-583 */
static float iir_iacc(const iir_t* f, const float* c, size_t st, int M, int k) {
  float z = f->pole[k], zn, Sum;
  switch (f->bc) {
    case BC_MIRROR: return (z / (z * z - 1.0f)) * (C(M - 1) + z * C(M - 2));
    case BC_NATURAL: return -(z / ((1.0f - z) * (1.0f - z))) * (C(M - 1) - z * C(M - 2));
    case BC_REFLECT: return C(M - 1) / (1.0f - 1.0f / z);
    case BC_PERIODIC:
    default:
      if (f->horizon[k] < M) {
        zn = z; Sum = C(M - 1) * z;
        for (int n = 0; n < f->horizon[k]; n++) { zn *= z; Sum += zn * C(n); }
        Sum = -Sum;
      } else {
        zn = z; Sum = C(M - 1);
        for (int n = 0; n < M - 1; n++) { Sum += zn * C(n); zn *= z; }
        Sum = z * Sum / (zn - 1.0f);
      }
      return Sum;
  }
}

/* solve_gain_inlined, recursive.h:631-729 (in place) */
static void iir_line(const iir_t* f, float* c, size_t st, int M) {
  if (M == 1 || f->npoles < 1) return;
  float p = f->pole[0], g = f->gain;
  float X = g * iir_icc(f, c, st, M, 0);
  C(0) = X;
  for (int n = 1; n < M; n++) { X = g * C(n) + p * X; C(n) = X; }
  X = iir_iacc(f, c, st, M, 0);
  C(M - 1) = X;
  for (int n = M - 2; n >= 0; n--) { X = p * (X - C(n)); C(n) = X; }
  for (int k = 1; k < f->npoles; k++) {
    p = f->pole[k];
    X = iir_icc(f, c, st, M, k);
    C(0) = X;
    for (int n = 1; n < M; n++) { X = C(n) + p * X; C(n) = X; }
    X = iir_iacc(f, c, st, M, k);
    C(M - 1) = X;
    for (int n = M - 2; n >= 0; n--) { X = p * (X - C(n)); C(n) = X; }
  }
}
#undef C

/* ------------------------------------------------------------------------------------------
 * staged source
 * ---------------------------------------------------------------------------------------- */
struct orc_source {
  int kind, projection, nch, degree;
  int w, h;           /* core shape */
  int cw, chh;        /* container shape */
  int lx, ly;         /* left frames */
  float* container;
  float* core;        /* texel (0,0) of the core */
  size_t stride;      /* floats per container row */
  int bc0, bc1;
  cm_metrics_t cm;
  float wmat[(ORC_MAX_DEGREE + 1) * (ORC_MAX_DEGREE + 1)];
};

/* get_left_brace_size / get_right_brace_size, zimt/bspline.h:305-372 */
static int left_brace(int degree, int bc) {
  int b = degree / 2;
  if (bc == BC_REFLECT) b++;
  else if (degree & 1) b++;
  if (bc == BC_PERIODIC && !(degree & 1)) b++;
  return b;
}
static int right_brace(int degree, int bc) {
  int b = degree / 2;
  if (bc == BC_REFLECT && !(degree & 1)) b++;
  if (degree & 1) b++;
  if (bc == BC_PERIODIC) b++;
  return b;
}

static float* texel(const orc_source_t* s, int x, int y) { /* core coordinates */
  return s->core + (ptrdiff_t)y * (ptrdiff_t)s->stride + (ptrdiff_t)x * s->nch;
}

/* bracer::apply for one axis, zimt/brace.h:151-338, copying boundary conditions. The reference fills the two
 * braces in lock-step from the core outwards, each slice from the slice its mirror image (or period)
 * points at - which, when a brace is wider than the core, is a brace slice filled a few steps earlier.
 * The closed form of that is folding the index with period 2m (REFLECT) or m (PERIODIC). */
static int brace_src(int i, int m, int bc) { /* core index that container slice i (core coordinates) copies */
  if (bc == BC_PERIODIC) {
    i %= m;
    return i < 0 ? i + m : i;
  }
  i %= 2 * m;
  if (i < 0) i += 2 * m;
  return i < m ? i : 2 * m - 1 - i;
}
static void brace_axis(orc_source_t* s, int axis, int bc, int lsz, int rsz) {
  int nch = s->nch;
  if (axis == 0) {
    int m = s->w;
    for (int Y = -s->ly; Y < s->chh - s->ly; Y++) {
      for (int k = 0; k < lsz; k++) memcpy(texel(s, -1 - k, Y), texel(s, brace_src(-1 - k, m, bc), Y), sizeof(float) * nch);
      for (int k = 0; k < rsz; k++) memcpy(texel(s, m + k, Y), texel(s, brace_src(m + k, m, bc), Y), sizeof(float) * nch);
    }
  } else {
    int m = s->h;
    size_t rowb = sizeof(float) * s->stride;
    for (int k = 0; k < lsz; k++) memcpy(texel(s, -s->lx, -1 - k), texel(s, -s->lx, brace_src(-1 - k, m, bc)), rowb);
    for (int k = 0; k < rsz; k++) memcpy(texel(s, -s->lx, m + k), texel(s, -s->lx, brace_src(m + k, m, bc)), rowb);
  }
}

/* zimt::prefilter, prefilter.h:125-198: axis 0 then axis 1, channels independent */
static void prefilter_2d(float* base, size_t stride, int nch, int w, int h, int bc0, int bc1, int degree,
                         long double tolerance) {
  if (degree <= 1) return;
  iir_t f0, f1;
  iir_setup(&f0, bc0, degree, tolerance);
  iir_setup(&f1, bc1, degree, tolerance);
#pragma omp parallel for schedule(static)
  for (int y = 0; y < h; y++)
    for (int c = 0; c < nch; c++) iir_line(&f0, base + (size_t)y * stride + c, (size_t)nch, w);
#pragma omp parallel for schedule(static)
  for (int x = 0; x < w; x++)
    for (int c = 0; c < nch; c++) iir_line(&f1, base + (size_t)x * nch + c, stride, h);
}

/* spherical_prefilter, environment.h:356-522 */
static void spherical_prefilter(orc_source_t* s, int degree, int ry) {
  int w = s->w, h = s->h, nch = s->nch;
  iir_t f;
  if (degree > 1) {
    iir_setup(&f, BC_PERIODIC, degree, (long double)0.0001);
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; y++)
      for (int c = 0; c < nch; c++) iir_line(&f, texel(s, 0, y) + c, (size_t)nch, w);
    /* vertical pass over [left half column top->bottom ; right half column bottom->top] */
    int half = w / 2;
#pragma omp parallel
    {
      float* line = (float*)malloc(sizeof(float) * 2 * (size_t)h);
#pragma omp for schedule(static)
      for (int x = 0; x < half; x++)
        for (int c = 0; c < nch; c++) {
          for (int y = 0; y < h; y++) {
            line[y] = texel(s, x, y)[c];
            line[h + y] = texel(s, x + half, h - 1 - y)[c];
          }
          iir_line(&f, line, 1, 2 * h);
          for (int y = 0; y < h; y++) {
            texel(s, x, y)[c] = line[y];
            texel(s, x + half, h - 1 - y)[c] = line[h + y];
          }
        }
      free(line);
    }
  }
  /* brace rows across the poles: row -1-k of one half = row k of the other half */
  int half = w / 2;
  for (int k = 0; k < s->ly; k++) {
    if (k >= h) break;
    memcpy(texel(s, 0, -1 - k), texel(s, half, k), sizeof(float) * nch * half);
    memcpy(texel(s, half, -1 - k), texel(s, 0, k), sizeof(float) * nch * half);
  }
  for (int k = 0; k < ry; k++) {
    if (k >= h) break;
    memcpy(texel(s, 0, h + k), texel(s, half, h - 1 - k), sizeof(float) * nch * half);
    memcpy(texel(s, half, h + k), texel(s, 0, h - 1 - k), sizeof(float) * nch * half);
  }
  brace_axis(s, 0, BC_PERIODIC, s->lx, s->cw - s->lx - s->w);
}

/* ---- coordinate gates, zimt/map.h (vector variants) ---- */
static float v_fmod(float lhs, float rhs) {
  float help = lhs;
  help /= rhs;
  help = truncf(help);
  help *= rhs;
  lhs -= help;
  if (fabsf(lhs) >= fabsf(rhs)) lhs = 0;
  return lhs;
}
static float gate_mirror(float c, float lower, float upper) {
  float cc = c - lower;
  float w = upper - lower;
  cc = fabsf(cc);
  if (cc >= w) {
    float cm = v_fmod(cc, 2 * w);
    cm -= w;
    cm = fabsf(cm);
    cm = w - cm;
    cc = cm;
  }
  return cc + lower;
}
static float gate_periodic(float c, float lower, float upper) {
  float cc = c - lower;
  float w = upper - lower;
  int below = cc < 0, above = cc >= w;
  if (below || above) {
    float cm = v_fmod(cc, w);
    if (below) cm = cm + w;
    if (cm >= w) cm = 0;
    cc = cm;
  }
  return cc + lower;
}

/* Optional instrumentation ("recompute the distinct source texels touched exactly from the
 * oracle's tap addresses"): one byte per CONTAINER texel of a registered source, set for every texel a
 * spline window reads. Racing threads all store the same value. */
#define ORC_TOUCH_MAX 64
static struct { const orc_source_t* src; unsigned char* map; } g_touch[ORC_TOUCH_MAX];
static int g_touch_n = 0;
void orc_touch_map(const orc_source_t* src, unsigned char* map) { /* src NULL: forget every map */
  if (!src) { g_touch_n = 0; return; }
  for (int i = 0; i < g_touch_n; i++)
    if (g_touch[i].src == src) { g_touch[i].map = map; return; }
  if (g_touch_n < ORC_TOUCH_MAX) { g_touch[g_touch_n].src = src; g_touch[g_touch_n].map = map; g_touch_n++; }
}
size_t orc_touch_map_size(const orc_source_t* src) { return (size_t)src->cw * src->chh; }
static void touch_window(const orc_source_t* s, int ix, int iy, int degree) {
  if (g_touch_n == 0) return;
  unsigned char* map = NULL;
  for (int i = 0; i < g_touch_n; i++)
    if (g_touch[i].src == s) map = g_touch[i].map;
  if (!map) return;
  int h2 = degree / 2;
  for (int j = 0; j <= degree; j++)
    for (int i = 0; i <= degree; i++)
      map[(size_t)(iy - h2 + j + s->ly) * s->cw + (size_t)(ix - h2 + i + s->lx)] = 1;
}

/* Opt-in arithmetic of the library's libenvutil_b200_fma.so build (include/envutil_b200.h, eu_render_arithmetic):
 * fused multiply-adds in the window evaluation and the twining accumulation of a RENDER (never while a
 * source is staged). Not the reference's arithmetic - the reference parity build rounds every product -
 * but what the contracted kernels are checked against bit for bit; the CPU tests state how far it is from
 * the pinned arithmetic. orc_set_arithmetic(1) asks for it, orc_render applies it for its own duration. */
static int g_contract_req = 0, g_contract = 0;
void orc_set_arithmetic(int contracted) { g_contract_req = contracted != 0; }
static inline float win_muladd(float a, float b, float c) { return g_contract ? fmaf(a, b, c) : c + a * b; }

/* safe evaluator = mapper + evaluator (zimt/eval.h:2039-2164, :1237-1300, :903-1059).
 * degree: the evaluator's degree (spline_degree + shift). crd in spline coordinates. */
static void spline_eval(const orc_source_t* s, int degree, const float* wmat, float cx, float cy, float* out) {
  int nch = s->nch;
  /* gates: PERIODIC / REFLECT limits are -0.5 .. N-0.5 (zimt/bspline.h:233-286) */
  float ux = (float)((long double)(s->w - 1) + 0.5L), uy = (float)((long double)(s->h - 1) + 0.5L);
  cx = (s->bc0 == BC_PERIODIC) ? gate_periodic(cx, -0.5f, ux) : gate_mirror(cx, -0.5f, ux);
  cy = (s->bc1 == BC_PERIODIC) ? gate_periodic(cy, -0.5f, uy) : gate_mirror(cy, -0.5f, uy);
  /* an axis of extent 1 is gated as CONSTANT with both limits 0: a clamp that always yields 0
   * (build_safe_ev, zimt/eval.h:2060-2064) */
  if (s->w == 1) cx = 0.0f;
  if (s->h == 1) cy = 0.0f;
  /* split, zimt/basis.h:102-146 */
  float fx, fy;
  int ix, iy;
  if (degree & 1) {
    float f = floorf(cx); fx = cx - f; ix = (int)f;
    f = floorf(cy); fy = cy - f; iy = (int)f;
  } else {
    float f = roundf(cx); fx = cx - f; ix = (int)f;
    f = roundf(cy); fy = cy - f; iy = (int)f;
  }
  touch_window(s, ix, iy, degree);
  if (degree == 0) {
    const float* p = texel(s, ix, iy);
    for (int c = 0; c < nch; c++) out[c] = p[c];
    return;
  }
  if (degree == 1) { /* _eval_linear, eval.h:1004-1059 */
    float wl0 = 1.0f - fx, wr0 = fx, wl1 = 1.0f - fy, wr1 = fy;
    const float* p00 = texel(s, ix, iy);
    const float* p10 = texel(s, ix + 1, iy);
    const float* p01 = texel(s, ix, iy + 1);
    const float* p11 = texel(s, ix + 1, iy + 1);
    for (int c = 0; c < nch; c++) {
      float sum = p00[c];
      sum *= wl0;
      sum = win_muladd(p10[c], wr0, sum);
      sum *= wl1;
      float sub = p01[c];
      sub *= wl0;
      sub = win_muladd(p11[c], wr0, sub);
      sum = win_muladd(sub, wr1, sum);
      out[c] = sum;
    }
    return;
  }
  /* weights, basis.h:650-689 */
  int order = degree + 1;
  float wx[ORC_MAX_DEGREE + 1], wy[ORC_MAX_DEGREE + 1];
  for (int axis = 0; axis < 2; axis++) {
    float* w = axis ? wy : wx;
    float delta = axis ? fy : fx;
    float power = delta;
    for (int k = 0; k < order; k++) w[k] = wmat[k];
    for (int row = 1; row < order; row++) {
      for (int k = 0; k < order; k++) w[k] = win_muladd(power, wmat[row * order + k], w[k]);
      if (row < order - 1) power *= delta;
    }
  }
  /* _eval, eval.h:903-996: window offsets k - degree/2 per axis (:732) */
  int h2 = degree / 2;
  for (int c = 0; c < nch; c++) {
    float sum = 0;
    for (int j = 0; j < order; j++) {
      const float* row = texel(s, ix - h2, iy - h2 + j) + c;
      float sub = row[0];
      sub *= wx[0];
      for (int i = 1; i < order; i++) sub = win_muladd(wx[i], row[(size_t)i * nch], sub);
      if (j == 0) {
        sum = sub;
        sum *= wy[0];
      } else {
        sum = win_muladd(sub, wy[j], sum);
      }
    }
    out[c] = sum;
  }
}

/* ray_to_cubeface, geometry.h:1178-1357 */
static void ray_to_cubeface(const float c[3], int* face, float in_face[2]) {
  int m1 = fabsf(c[0]) >= fabsf(c[1]);
  int m2 = fabsf(c[0]) >= fabsf(c[2]);
  int m3 = fabsf(c[1]) >= fabsf(c[2]);
  if (m1 && m2) {
    *face = c[0] < 0 ? CM_LEFT : CM_RIGHT;
    in_face[0] = -c[2] / c[0];
    in_face[1] = c[1] / fabsf(c[0]);
  } else if (!m2 && !m3) {
    *face = c[2] < 0 ? CM_BACK : CM_FRONT;
    in_face[0] = c[0] / c[2];
    in_face[1] = c[1] / fabsf(c[2]);
  } else {
    *face = c[1] < 0 ? CM_TOP : CM_BOTTOM;
    in_face[0] = -c[0] / fabsf(c[1]);
    in_face[1] = c[2] / c[1];
  }
}

/* cubemap_t::fill_support, cubemap.h:607-911: 1-px mirrored ring, then the four frame
 * stripes of every section by bilinear reprojection from the IR itself, in the reference's
 * order (later stripes see pixels written by earlier ones). */
static void cubemap_fill_support(orc_source_t* s) {
  const cm_metrics_t* m = &s->cm;
  int S = m->section_px, F = m->face_px, L = m->left_frame_px, R = m->right_frame_px, nch = s->nch;
  if (L == 0 && R == 0) return;
  size_t tb = sizeof(float) * nch;
  for (int face = 0; face < 6; face++) { /* mirror_around, :607-660 */
    int oy = face * S + L, ox = L;
    int cmin = L > 0 ? -1 : 0, cmax = R > 0 ? F : F - 1;
    for (int x = cmin; x <= cmax; x++) {
      if (L) memcpy(texel(s, ox + x, oy - 1), texel(s, ox + x, oy), tb);
      if (R) memcpy(texel(s, ox + x, oy + F), texel(s, ox + x, oy + F - 1), tb);
    }
    for (int y = cmin; y <= cmax; y++) {
      if (L) memcpy(texel(s, ox - 1, oy + y), texel(s, ox, oy + y), tb);
      if (R) memcpy(texel(s, ox + F, oy + y), texel(s, ox + F - 1, oy + y), tb);
    }
  }
  int ithird = (int)(m->model_to_px * 2);
  int ishift = S - 1;
  float* rowbuf = (float*)malloc(tb * S);
  for (int face = 0; face < 6; face++) {
    int win[4][4] = {{0, 0, S, L}, {0, S - R, S, S}, {0, L, L, S - R}, {L + F, L, S, S - R}};
    int on[4] = {L > 0, R > 0, L > 0, R > 0};
    for (int st = 0; st < 4; st++) {
      if (!on[st]) continue;
      int x0 = win[st][0], y0 = win[st][1], x1 = win[st][2], y1 = win[st][3];
      /* zimt::process works through a line in vectors of 16 pixels: a vector is evaluated from the IR as
       * it is, then stored (zimt/wielding.h:317-455). That matters for odd face widths only: the face is
       * then not centred in its section (left frame = right frame - 1) while the pixel-to-ray step
       * assumes it is (ishift = S - 1), so the first frame row below the TOP/BOTTOM faces (first column
       * right of the LEFT/RIGHT faces) maps onto its OWN section's edge and averages in the frame texel
       * before it - already rewritten if that lies in an earlier vector of the line, not yet if in the same.
       * (The LEFT/RIGHT column reads the line ABOVE, which another thread of the reference is working on at
       * the same time: a race in the reference; this restatement takes the lines in order. Odd face widths
       * are therefore not pinned for the texels that depend on that column - see DESIGN.md.) */
      for (int y = y0; y < y1; y++) {
        for (int v0 = x0; v0 < x1; v0 += ORC_LANES) {
          int v1 = v0 + ORC_LANES < x1 ? v0 + ORC_LANES : x1;
          for (int x = v0; x < v1; x++) {
            int c0 = 2 * x - ishift, c1 = 2 * y - ishift;
            float ray[3];
            switch (face) { /* fill_frame_t::eval, :733-780 */
              case CM_FRONT: ray[0] = (float)c0; ray[1] = (float)c1; ray[2] = (float)ithird; break;
              case CM_BACK: ray[0] = (float)(-c0); ray[1] = (float)c1; ray[2] = (float)(-ithird); break;
              case CM_RIGHT: ray[0] = (float)ithird; ray[1] = (float)c1; ray[2] = (float)(-c0); break;
              case CM_LEFT: ray[0] = (float)(-ithird); ray[1] = (float)c1; ray[2] = (float)c0; break;
              case CM_BOTTOM: ray[0] = (float)(-c0); ray[1] = (float)ithird; ray[2] = (float)c1; break;
              default: ray[0] = (float)(-c0); ray[1] = (float)(-ithird); ray[2] = (float)(-c1); break;
            }
            int fv;
            float in_face[2], pk[2];
            ray_to_cubeface(ray, &fv, in_face);
            /* metrics_t::get_pickup_coordinate_px, cubemap.h:401-411: refc_md is a double here */
            pk[0] = (float)((double)in_face[0] + m->refc_md);
            pk[1] = (float)((double)in_face[1] + m->refc_md);
            pk[0] *= (float)m->model_to_px;
            pk[1] *= (float)m->model_to_px;
            pk[1] += (float)(fv * S);
            pk[0] -= .5f;
            pk[1] -= .5f;
            spline_eval(s, 1, NULL, pk[0], pk[1], rowbuf + (size_t)(x - v0) * nch);
          }
          memcpy(texel(s, v0, face * S + y), rowbuf, tb * (v1 - v0));
        }
      }
    }
  }
  free(rowbuf);
}

orc_source_t* orc_source_create(const eu_facet_t* f, const eu_opts_t* o, const float* pixels) {
  orc_source_t* s = (orc_source_t*)calloc(1, sizeof(*s));
  int degree = o->spline_degree;
  int pdeg = o->prefilter_degree < 0 ? degree : o->prefilter_degree;
  s->projection = f->projection;
  s->nch = f->nchannels;
  s->degree = degree;
  int nch = s->nch;
  if (degree > 1) orc_weight_matrix(degree, s->wmat);
  if (f->projection == EU_CUBEMAP || f->projection == EU_BIATAN6) {
    /* cubemap_t ctor + load, cubemap.h:548-580,1147-1233 */
    s->kind = f->projection == EU_CUBEMAP ? KIND_CUBEMAP : KIND_BIATAN6;
    cm_metrics(f->width, f->hfov, o->support_min, o->tile_size, &s->cm);
    int S = s->cm.section_px, F = f->width, L = s->cm.left_frame_px;
    s->w = s->cw = S;
    s->h = s->chh = 6 * S;
    s->lx = s->ly = 0;
    s->bc0 = s->bc1 = BC_REFLECT;
    s->stride = (size_t)S * nch;
    s->container = (float*)calloc((size_t)S * 6 * S * nch, sizeof(float));
    s->core = s->container;
    for (int face = 0; face < 6; face++)
      for (int y = 0; y < F; y++)
        memcpy(texel(s, L, face * S + L + y), pixels + ((size_t)(face * F + y) * F) * nch, sizeof(float) * nch * F);
    cubemap_fill_support(s);
    if (pdeg > 1)
      for (int face = 0; face < 6; face++) /* cubemap_t::prefilter, :921-946 */
        prefilter_2d(texel(s, 0, face * S), s->stride, nch, S, S, BC_NATURAL, BC_NATURAL, pdeg,
                     (long double)FLT_EPSILON);
    return s;
  }
  /* source_t ctor, environment.h:594-950 */
  s->kind = KIND_MOUNT;
  /* a 'W' window: the raster on disk is the window, the geometry refers to the total size
   * (envutil_main.cc:754-786, environment.h:596-601) */
  int ww = f->window_width > 0 ? f->window_width : f->width;
  int wh = f->window_height > 0 ? f->window_height : f->height;
  s->w = ww;
  s->h = wh;
  s->bc0 = BC_REFLECT;
  s->bc1 = BC_REFLECT;
  if ((f->projection == EU_SPHERICAL || f->projection == EU_CYLINDRICAL) && fabs(f->hfov - 2.0 * M_PI) < .000001)
    s->bc0 = BC_PERIODIC;
  s->lx = left_brace(degree, s->bc0);
  s->ly = left_brace(degree, s->bc1);
  int rx = right_brace(degree, s->bc0), ry = right_brace(degree, s->bc1);
  s->cw = s->w + s->lx + rx;
  s->chh = s->h + s->ly + ry;
  s->stride = (size_t)s->cw * nch;
  s->container = (float*)calloc((size_t)s->cw * s->chh * nch, sizeof(float));
  s->core = s->container + (size_t)s->ly * s->stride + (size_t)s->lx * nch;
  for (int y = 0; y < s->h; y++)
    memcpy(texel(s, 0, y), pixels + (size_t)y * s->w * nch, sizeof(float) * nch * s->w);
  if (f->projection == EU_SPHERICAL && fabs(f->hfov - 2.0 * M_PI) < .000001 && f->width == 2 * f->height) {
    spherical_prefilter(s, pdeg, ry);
  } else {
    prefilter_2d(s->core, s->stride, nch, s->w, s->h, s->bc0, s->bc1, pdeg, (long double)FLT_EPSILON);
    brace_axis(s, 0, s->bc0, s->lx, rx);
    brace_axis(s, 1, s->bc1, s->ly, ry);
  }
  return s;
}

/* fill_polygon, envutil_basic.cc:236-321 (non-zero winding scan-line fill; clears alpha) */
static void fill_polygon_clear(const float* px, const float* py, int N, int w, int h, float* alpha) {
  int* nodeX = (int*)malloc(sizeof(int) * (N + 1));
  int* dir = (int*)malloc(sizeof(int) * (N + 1));
  for (int pixelY = 0; pixelY < h; pixelY++) {
    int nodes = 0, j = N - 1;
    for (int i = 0; i < N; i++) {
      int cross = 0;
      if (py[i] < (float)pixelY && py[j] >= (float)pixelY) cross = 1;
      else if (py[j] < (float)pixelY && py[i] >= (float)pixelY) cross = -1;
      if (cross) {
        nodeX[nodes] = (int)(px[i] + (pixelY - py[i]) / (py[j] - py[i]) * (px[j] - px[i]));
        dir[nodes++] = cross;
      }
      j = i;
    }
    int i = 0;
    while (i < nodes - 1) {
      if (nodeX[i] > nodeX[i + 1]) {
        int sw = nodeX[i]; nodeX[i] = nodeX[i + 1]; nodeX[i + 1] = sw;
        sw = dir[i]; dir[i] = dir[i + 1]; dir[i + 1] = sw;
        if (i) i--;
      } else {
        i++;
      }
    }
    int w_ord = 0;
    for (i = 0; i < nodes; i++) {
      w_ord += dir[i];
      if (!w_ord) continue;
      if (i + 1 >= nodes) break; /* the reference reads nodeX[i+1] unchecked; a closed polygon never gets here */
      if (nodeX[i] >= w) break;
      if (nodeX[i + 1] > 0) {
        if (nodeX[i] < 0) nodeX[i] = 0;
        if (nodeX[i + 1] > w) nodeX[i + 1] = w;
        for (int x = nodeX[i]; x < nodeX[i + 1]; x++) alpha[(size_t)pixelY * w + x] = 0.0f;
      }
    }
  }
  free(nodeX);
  free(dir);
}

static int reflect_index(int i, int w) { /* zimt/extrapolate.h:141-153 */
  if (i < 0) i = -1 - i;
  if (i >= w) {
    i %= 2 * w;
    if (i >= w) i = 2 * w - i - 1;
  }
  return i;
}

/* the 0/1 plane of a masked / cropped facet, feathered (environment.h:711-843). The binomial
 * kernel applied to 0/1 data yields multiples of 1/256: every partial sum is exact in float, the
 * order of summation of the reference's circular-buffer FIR (zimt/convolve.h) does not matter. */
static float* build_alpha_plane(const eu_facet_t* f, const eu_alpha_spec_t* a) {
  int w = f->window_width > 0 ? f->window_width : f->width, h = f->window_height > 0 ? f->window_height : f->height;
  float* alpha = (float*)malloc(sizeof(float) * (size_t)w * h);
  for (size_t i = 0; i < (size_t)w * h; i++) alpha[i] = 1.0f;
  const float* xy = a->mask_xy;
  for (int m = 0; m < a->n_masks; m++) {
    int n = a->mask_sizes[m];
    float* vx = (float*)malloc(sizeof(float) * n);
    float* vy = (float*)malloc(sizeof(float) * n);
    for (int k = 0; k < n; k++) { vx[k] = xy[2 * k]; vy[k] = xy[2 * k + 1]; }
    fill_polygon_clear(vx, vy, n, w, h, alpha);
    free(vx); free(vy);
    xy += 2 * n;
  }
  if (a->has_crop) {
    float ca = (float)(fabs((double)(a->crop_x1 - a->crop_x0)) / 2.0);
    float cb = (float)(fabs((double)(a->crop_y1 - a->crop_y0)) / 2.0);
    if (f->projection == EU_FISHEYE) { /* elliptic crop, :746-772 */
      float mx = (float)((a->crop_x0 + a->crop_x1) / 2.0);
      float my = (float)((a->crop_y0 + a->crop_y1) / 2.0);
      for (int y = 0; y < h; y++) {
        float dy = fabsf((float)y - my);
        if (dy > cb) {
          for (int x = 0; x < w; x++) alpha[(size_t)y * w + x] = 0.0f;
          continue;
        }
        float xmargin = (float)sqrt((double)(ca * ca) * (1.0 - (double)((dy * dy) / (cb * cb))));
        for (int x = 0; x < w; x++) {
          float dx = fabsf((float)x - mx);
          if (dx > xmargin) alpha[(size_t)y * w + x] = 0.0f;
        }
      }
    } else { /* rectangular crop, :773-790 */
      for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++)
          if (x < a->crop_x0 || x >= a->crop_x1 || y < a->crop_y0 || y >= a->crop_y1) alpha[(size_t)y * w + x] = 0.0f;
    }
  }
  static const float k5[5] = {1.0f / 16.0f, 4.0f / 16.0f, 6.0f / 16.0f, 4.0f / 16.0f, 1.0f / 16.0f};
  float* tmp = (float*)malloc(sizeof(float) * (size_t)w * h);
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      float s = 0.0f;
      for (int j = 0; j < 5; j++) s += k5[j] * alpha[(size_t)y * w + reflect_index(x - 2 + j, w)];
      tmp[(size_t)y * w + x] = s;
    }
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      float s = 0.0f;
      for (int j = 0; j < 5; j++) s += k5[j] * tmp[(size_t)reflect_index(y - 2 + j, h) * w + x];
      alpha[(size_t)y * w + x] = s;
    }
  free(tmp);
  return alpha;
}

orc_source_t* orc_source_create_alpha(const eu_facet_t* f, const eu_opts_t* o, const float* pixels,
                                      const eu_alpha_spec_t* a) {
  int w = f->window_width > 0 ? f->window_width : f->width, h = f->window_height > 0 ? f->window_height : f->height;
  int C = f->nchannels, nat = a->native_nchannels;
  float* alpha = build_alpha_plane(f, a);
  float* px = (float*)malloc(sizeof(float) * (size_t)w * h * C);
  for (size_t i = 0; i < (size_t)w * h; i++) {
    for (int c = 0; c < C; c++) {
      float v = c < nat ? pixels[i * nat + c] : 1.0f; /* added alpha channel is 1 (:698-709) */
      px[i * C + c] = v * alpha[i];                   /* v3 = v1 * v2, every channel (:867-872) */
    }
  }
  free(alpha);
  orc_source_t* s = orc_source_create(f, o, px);
  free(px);
  return s;
}

void orc_source_free(orc_source_t* s) {
  if (!s) return;
  free(s->container);
  free(s);
}

const float* orc_source_container(const orc_source_t* s, int32_t shape[4]) {
  shape[0] = s->cw; shape[1] = s->chh; shape[2] = s->lx; shape[3] = s->ly;
  return s->container;
}

/* ------------------------------------------------------------------------------------------
 * per-job derived state
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  const orc_source_t* src;
  int projection, kind;
  float xx[3], yy[3], zz[3]; /* stepper basis rows, narrowed (stepper ctor args) */
  /* source_t, environment.h:594-645,970-1006 */
  double ext_x0, ext_y0;
  float ext_w, ext_h;
  float win_x0, win_x1, win_y0, win_y1;
  int total_w, total_h;
  int win_xoff, win_yoff;
  int has_lcp, has_shift, has_shear;
  float lcp[4], lcp_s, sh_h, sh_v;
  double shear_g, shear_t;
  float refc_md, model_to_px;
  int section_px;
  float recip_step, brighten, optimum;
  int hdr_kind; /* 0 LOW 1 MIDDLE 2 HIGH */
  int mask_always;
  /* generic_stepper + tf_ex_facet + generic_r3 + tf3d_t for facets with PanoTools translation
   * (envutil_payload.cc:1628-1883, geometry.h:1850-1942); float matrices, rows as r3_t holds them */
  int generic;
  float g_t2m[9], g_m2s[9], g_shift[3], g_dcp;
  /* 'single' jobs whose target facet has lens correction / shift / shear / translation: generic_r3(ft, fs)
   * is up to two tf3d_t in sequence (envutil_payload.cc:1716-1760); a tf3d_t without shift is one rotation */
  int g_nstage;
  struct { int has_shift; float a[9], b[9], ab[9], shift[3], dcp; } g_st[2];
  int masked;   /* --mask_for: 0 normal, else the colour channels are `paint` (facet_spec::masked, masking.h) */
  float paint;
} facet_ctx;

#define INV_SZ 100            /* knot count passed by pto_planar's ctor, environment.h:250 */
#define INV_NK (INV_SZ + 4)
/* pto_planar<float, L, true> of the target facet (environment.h:240-309): the inverse of shear, shift
 * and lens polynomial; inverse_lcp's spline (lens_correction.h:273-406) */
typedef struct {
  int on, has_shear, has_shift, has_lcp;
  double shear_g, shear_t, s, rr_max;
  float h, v;
  float coef[2 + INV_NK + 2]; /* braced NATURAL cubic spline, core at [2] */
} inv_planar_t;

typedef struct {
  inv_planar_t inv;
  int projection, width, height, normalize; /* width x height: the raster rendered (the crop, if any) */
  int full_w, full_h, off_x, off_y;         /* the target the steppers are built for; the crop's origin */
  float fx0, fx1, fy0, fy1, delta;
  float bias_x, bias_y; /* of the biased steppers r10 / r01 */
  float section_md, refc_md;
} target_ctx;

/* stepper_base ctor, stepper.h:294-306 */
static void target_setup(const eu_target_t* t, target_ctx* T) {
  double e[4];
  int w = t->width, h = t->height;
  orc_get_extent(t->projection, w, h, t->hfov, e);
  float a0 = (float)e[0], a1 = (float)e[1], b0 = (float)e[2], b1 = (float)e[3];
  memset(&T->inv, 0, sizeof(T->inv));
  T->projection = t->projection;
  /* cropped output, envutil_payload.cc:440-443,470-474: the raster has the crop's size and the
   * discrete coordinates handed to the steppers are offset by the crop's origin */
  T->width = t->crop_width > 0 ? t->crop_width : w;
  T->height = t->crop_width > 0 ? t->crop_height : h;
  T->full_w = w;
  T->full_h = h;
  T->off_x = t->crop_width > 0 ? t->crop_x0 : 0;
  T->off_y = t->crop_width > 0 ? t->crop_y0 : 0;
  T->fx1 = (float)(a1 / (2.0 * w));
  T->fx0 = (float)(a0 / (2.0 * w));
  T->fy1 = (float)(b1 / (2.0 * h));
  T->fy0 = (float)(b0 / (2.0 * h));
  T->bias_x = .25f * (a1 - a0) / (float)w;
  T->bias_y = .25f * (b1 - b0) / (float)h;
  T->delta = (float)ORC_LANES * (a1 - a0) / (float)w;
  T->section_md = a1 - a0;                    /* stepper.h:1265 */
  T->refc_md = (float)((a1 - a0) / 2.0);      /* stepper.h:1266 */
}

/* stepper_base::init + increase, stepper.h:324-350, as zimt::process drives them
 * (zimt/wielding.h:317-455): init at the start of each 512-px segment, then += delta per
 * 16-px vector. */
static float planar_x(const target_ctx* T, int x, float bias) {
  int seg0 = (x / ORC_SEGMENT) * ORC_SEGMENT;
  int r = x - seg0, lane = r % ORC_LANES, v = r / ORC_LANES;
  float ll0 = (float)(2 * lane) + (float)((seg0 + T->off_x) * 2 + 1);
  float p = bias + ll0 * T->fx1 + ((float)(2 * T->full_w) - ll0) * T->fx0;
  for (int i = 0; i < v; i++) p += T->delta;
  return p;
}
static float planar_y(const target_ctx* T, int y, float bias) {
  int ll1 = (y + T->off_y) * 2 + 1;
  return bias + ll1 * T->fy1 + (float)(2 * T->full_h - ll1) * T->fy0;
}

static float norm3(const float v[3]) {
  float sqn = v[0] * v[0];
  sqn += v[1] * v[1];
  sqn += v[2] * v[2];
  return sqrtf(sqn);
}

/* rotate(xel_t<float,3>, r3_t<float>), geometry.h:74-82: (v0*m0 + v1*m1) + v2*m2, per component */
static void rot3f(const float v[3], const float m[9], float out[3]) {
  float o[3];
  for (int c = 0; c < 3; c++) o[c] = (v[0] * m[c] + v[1] * m[3 + c]) + v[2] * m[6 + c];
  out[0] = o[0]; out[1] = o[1]; out[2] = o[2];
}
/* rotate(r3_t<float>, r3_t<float>), geometry.h:84-91 */
static void matmulf(const float a[9], const float b[9], float m[9]) {
  for (int i = 0; i < 3; i++) rot3f(a + 3 * i, b, m + 3 * i);
}

/* generic_stepper::init/increase (stepper.h:353-470) over tf_ex_facet::eval
 * (envutil_payload.cc:1841-1883): planar -> X_to_ray of the target projection (geometry.h:151-567)
 * -> tf3d_t::eval (geometry.h:1886-1925). */
/* eu_polynomial<double, 4>, lens_correction.h:86-175 */
static double poly4(const double cf[5], double x) {
  double sum = 0.0, power = 1.0;
  for (int i = 0; i <= 4; i++) { sum += cf[4 - i] * power; power *= x; }
  return sum;
}
static double poly4_deriv(const double dcf[5], double x) {
  double sum = 0.0, power = 1.0;
  for (int i = 0; i < 4; i++) { sum += dcf[4 - i - 1] * power; power *= x; }
  return sum;
}
static int poly4_inverse(const double cf[5], const double dcf[5], double desired, double* x) { /* Newton, :133-170 */
  const double tolerance = 100 * DBL_EPSILON;
  double current = *x, result, difference = 0.0, last_difference = DBL_MAX;
  for (int count = 0; count < 16; count++) {
    result = poly4(cf, current);
    difference = desired - result;
    if (last_difference == difference) break;
    if (fabs(difference) <= tolerance) break;
    last_difference = difference;
    current = current + difference / poly4_deriv(dcf, current);
  }
  if (fabs(difference) < tolerance) { *x = current; return 1; }
  return 0;
}
/* inverse_lcp ctor, lens_correction.h:341-386: knots of the factor-minus-one spline, NATURAL cubic,
 * prefiltered and braced like any zimt::bspline<float, 1> */
static int inv_lcp_setup(inv_planar_t* P, double a, double b, double c, double r_max_in) {
  const int sz = INV_SZ, nk = INV_NK;
  double cf[5] = {a, b, c, 1.0 - (a + b + c), 0.0}, dcf[5];
  {
    size_t power = 4;
    for (int i = 0; i <= 4; i++) { dcf[i] = cf[i] * power; --power; }
  }
  double r_max = r_max_in * ((sz + 3.0) / sz);
  P->rr_max = poly4(cf, r_max);
  float* core = P->coef + 2;
  for (int i = 0; i < nk; i++) {
    double notch = (double)i / (nk - 1);
    notch *= notch;
    notch *= P->rr_max;
    double out = i * r_max / sz;
    if (!poly4_inverse(cf, dcf, notch, &out)) return -1;
    core[i] = (float)(notch == 0.0 ? 1.0 / poly4_deriv(dcf, 0.0) : (out / notch) - 1);
  }
  iir_t f;
  iir_setup(&f, BC_NATURAL, 3, (long double)FLT_EPSILON);
  iir_line(&f, core, 1, nk);
  for (int k = 0; k < 2; k++) { /* bracer, NATURAL: twice the pivot minus the mirrored source, brace.h:254-266,299-311 */
    core[-1 - k] = core[0] + core[0] - core[1 + k];
    core[nk + k] = core[nk - 1] + core[nk - 1] - core[nk - 2 - k];
  }
  return 0;
}
/* inverse_lcp::eval, lens_correction.h:396-405, called with norm(out) / s - a DOUBLE vector (float vector
 * over double scalar) - so the reduction to spline coordinates runs in double and is narrowed when the
 * float evaluator takes it; clamp gate (NATURAL, eval.h:2101-2110), cubic 1-D evaluation (eval.h:937-960) */
static float inv_lcp_factor(const inv_planar_t* P, float radius) {
  double in = (double)radius / P->s;
  in = in / P->rr_max;
  in = sqrt(in);
  in *= (INV_NK - 1);
  float cx = (float)in;
  const float lower = 0.0f, upper = (float)(INV_NK - 1);
  if (cx < lower) cx = lower;
  else if (cx > upper) cx = upper;
  float fl = floorf(cx), t = cx - fl;
  int ix = (int)fl;
  float wm[16], w[4];
  orc_weight_matrix(3, wm);
  float power = t;
  for (int k = 0; k < 4; k++) w[k] = wm[k];
  for (int row = 1; row < 4; row++) {
    for (int k = 0; k < 4; k++) w[k] += power * wm[row * 4 + k];
    if (row < 3) power *= t;
  }
  const float* c = P->coef + 2 + ix - 1;
  float sum = c[0];
  sum *= w[0];
  for (int i = 1; i < 4; i++) sum += w[i] * c[i];
  sum += 1;
  return sum;
}

static void generic_ray(const target_ctx* T, const facet_ctx* F, float h, float v, float ray[3]) {
  float in[3]; /* RIGHT, DOWN, FORWARD */
  if (T->inv.on) { /* tf22: pto_planar<T, L, true>::eval, environment.h:285-307 */
    const inv_planar_t* P = &T->inv;
    if (P->has_shear) { /* float vector op double scalar is evaluated in double (gen_simd_type.h:274-316) */
      v = (float)(((double)v - P->shear_t * (double)h) / (1 - P->shear_t * P->shear_g));
      h = (float)((double)h - P->shear_g * (double)v);
    }
    if (P->has_shift) { /* operator-= narrows its scalar operand first (vector_common.h:302-316) */
      h -= P->h;
      v -= P->v;
    }
    if (P->has_lcp) {
      float sqn = h * h;
      sqn += v * v;
      float factor = inv_lcp_factor(P, sqrtf(sqn));
      h *= factor;
      v *= factor;
    }
  }
  switch (T->projection) {
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
    default: { /* EU_CUBEMAP, EU_BIATAN6: ir_to_ray_t / ba6_to_ray_t as roll_out_23 default-constructs them
                * (geometry.h:1800-1834,660-775,857-990): section_md 2.0, refc_md 1.0, ul2c = {1, 6} */
      float c0 = h + 1.0f, c1 = v + 6.0f;
      int section = (int)((double)c1 / 2.0);
      c1 = (float)((double)c1 - (double)section * 2.0);
      c0 -= 1.0f;
      c1 -= 1.0f;
      if (T->projection == EU_BIATAN6) {
        c0 = eu_tanf(c0 * (float)(M_PI / 4));
        c1 = eu_tanf(c1 * (float)(M_PI / 4));
      }
      switch (section) {
        case CM_LEFT: in[0] = -1.0f; in[1] = c1; in[2] = c0; break;
        case CM_RIGHT: in[0] = 1.0f; in[1] = c1; in[2] = -c0; break;
        case CM_TOP: in[0] = -c0; in[1] = -1.0f; in[2] = -c1; break;
        case CM_BOTTOM: in[0] = -c0; in[1] = 1.0f; in[2] = c1; break;
        case CM_FRONT: in[0] = c0; in[1] = c1; in[2] = 1.0f; break;
        default: in[0] = -c0; in[1] = c1; in[2] = -1.0f; break;
      }
    }
  }
  float out[3];
  if (F->g_nstage > 0) { /* generic_r3(ft, fs): tf3d_t::eval per stage, geometry.h:1886-1925 */
    out[0] = in[0]; out[1] = in[1]; out[2] = in[2];
    for (int k = 0; k < F->g_nstage; k++) {
      if (!F->g_st[k].has_shift) {
        rot3f(out, F->g_st[k].ab, out);
        continue;
      }
      rot3f(out, F->g_st[k].a, out);
      if (out[2] <= 0.0f) {
        out[0] = 0.0f; out[1] = 0.0f; out[2] = -INFINITY;
      } else {
        out[0] /= out[2];
        out[1] /= out[2];
        out[2] = 1.0f;
        for (int c = 0; c < 3; c++) out[c] *= F->g_st[k].dcp;
        for (int c = 0; c < 3; c++) out[c] -= F->g_st[k].shift[c];
        rot3f(out, F->g_st[k].b, out);
      }
    }
  } else {
  rot3f(in, F->g_t2m, out);
  if (out[2] <= 0.0f) {
    out[0] = 0.0f; out[1] = 0.0f; out[2] = -INFINITY;
  } else {
    out[0] /= out[2];
    out[1] /= out[2];
    out[2] = 1.0f;
    for (int c = 0; c < 3; c++) out[c] *= F->g_dcp;
    for (int c = 0; c < 3; c++) out[c] -= F->g_shift[c];
    rot3f(out, F->g_m2s, out);
  }
  }
  if (T->normalize) {
    float n = norm3(out);
    for (int c = 0; c < 3; c++) out[c] /= n;
  }
  ray[0] = out[0]; ray[1] = out[1]; ray[2] = out[2];
}

/* the seven steppers, stepper.h:517-1578. (px,py) planar coordinate, (x,y) discrete target
 * coordinate (needed by the cube steppers and by the cylindrical stepper's per-segment
 * rcp_length). */
static void stepper_ray(const target_ctx* T, const facet_ctx* F, float px, float py, int x, int y, float bias_x,
                        float ray[3]) {
  const float *xx = F->xx, *yy = F->yy, *zz = F->zz;
  if (F->generic) {
    generic_ray(T, F, px, py, ray);
    return;
  }
  switch (T->projection) {
    case EU_SPHERICAL: {
      float sy, r, sx, z;
      eu_sincosf(py, &sy, &r);
      eu_sincosf(px, &sx, &z);
      for (int i = 0; i < 3; i++) {
        float xxx = xx[i] * r, yyy = yy[i] * sy, zzz = zz[i] * r;
        ray[i] = xxx * sx + zzz * z + yyy;
      }
      break;
    }
    case EU_CYLINDRICAL: {
      float sx, z;
      eu_sincosf(px, &sx, &z);
      for (int i = 0; i < 3; i++) ray[i] = xx[i] * sx + zz[i] * z + yy[i] * py;
      if (T->normalize) {
        /* rcp_length is taken from the first vector of the segment, same lane (:766-769) */
        int seg0 = (x / ORC_SEGMENT) * ORC_SEGMENT, lane = (x - seg0) % ORC_LANES;
        float p0 = planar_x(T, seg0 + lane, bias_x), s0, z0, first[3];
        eu_sincosf(p0, &s0, &z0);
        for (int i = 0; i < 3; i++) first[i] = xx[i] * s0 + zz[i] * z0 + yy[i] * py;
        float rcp = 1.0f / norm3(first);
        for (int i = 0; i < 3; i++) ray[i] *= rcp;
      }
      break;
    }
    case EU_RECTILINEAR: {
      for (int i = 0; i < 3; i++) {
        float ddd = yy[i] * py + zz[i];
        ray[i] = xx[i] * px + ddd;
      }
      if (T->normalize) {
        float n = norm3(ray);
        for (int i = 0; i < 3; i++) ray[i] /= n;
      }
      break;
    }
    case EU_FISHEYE:
    case EU_STEREOGRAPHIC: {
      float sqn = px * px;
      sqn += py * py;
      float nrm = sqrtf(sqn);
      float a;
      if (T->projection == EU_FISHEYE)
        a = (float)(M_PI_2 - (double)nrm); /* :1019-1021, double promotion */
      else
        a = (float)(M_PI_2 - 2.0 * eu_atan((double)nrm / 2.0)); /* :1146-1148 */
      float b = eu_atan2f(px, py);
      float z, r, sx, cy;
      eu_sincosf(a, &z, &r);
      eu_sincosf(b, &sx, &cy);
      for (int i = 0; i < 3; i++) ray[i] = xx[i] * r * sx + zz[i] * z + yy[i] * r * cy;
      break;
    }
    case EU_CUBEMAP:
    case EU_BIATAN6: {
      int face = y / T->width;
      float p1 = py + (3 - face) * T->section_md - T->refc_md;
      float p0 = px;
      if (T->projection == EU_BIATAN6) {
        p1 = eu_tanf(p1 * (float)(M_PI / 4.0));
        p0 = eu_tanf(p0 * (float)(M_PI / 4.0));
      }
      float ccc[3], vvv[3];
      for (int i = 0; i < 3; i++) {
        switch (face) { /* :1304-1331; the +-1.0 factors promote the sum to double */
          case CM_LEFT: ccc[i] = (float)(-1.0 * xx[i] + (double)(p1 * yy[i])); vvv[i] = zz[i]; break;
          case CM_RIGHT: ccc[i] = (float)(1.0 * xx[i] + (double)(p1 * yy[i])); vvv[i] = -zz[i]; break;
          case CM_TOP: ccc[i] = (float)(-1.0 * yy[i] - (double)(p1 * zz[i])); vvv[i] = -xx[i]; break;
          case CM_BOTTOM: ccc[i] = (float)(1.0 * yy[i] + (double)(p1 * zz[i])); vvv[i] = -xx[i]; break;
          case CM_FRONT: ccc[i] = (float)((double)(p1 * yy[i]) + 1.0 * zz[i]); vvv[i] = xx[i]; break;
          default: ccc[i] = (float)((double)(p1 * yy[i]) - 1.0 * zz[i]); vvv[i] = -xx[i]; break;
        }
      }
      for (int i = 0; i < 3; i++) ray[i] = ccc[i] + p0 * vvv[i];
      if (T->normalize) {
        float n = norm3(ray);
        for (int i = 0; i < 3; i++) ray[i] /= n;
      }
      break;
    }
    default: ray[0] = ray[1] = 0; ray[2] = 1;
  }
}

/* mount_t::get_coordinate_nomask, environment.h:1077-1110 + geometry.h:277-534 + pto_planar */
static void mount_coordinate(const facet_ctx* F, const float r[3], float c[2]) {
  switch (F->projection) {
    case EU_RECTILINEAR: c[0] = r[0] / r[2]; c[1] = r[1] / r[2]; break;
    case EU_SPHERICAL: {
      float s = sqrtf(r[0] * r[0] + r[2] * r[2]);
      c[1] = eu_atan2f(r[1], s);
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
    default: { /* FISHEYE */
      float s = sqrtf(r[0] * r[0] + r[1] * r[1]);
      float rr = (float)M_PI_2 - eu_atan2f(r[2], s);
      float phi = eu_atan2f(r[1], r[0]);
      c[0] = rr * eu_cosf(phi);
      c[1] = rr * eu_sinf(phi);
    }
  }
  if (F->has_lcp) { /* pto_planar::eval forward, environment.h:254-283; lcp lens_correction.h:94-105 */
    float sqn = c[0] * c[0];
    sqn += c[1] * c[1];
    float x = sqrtf(sqn) / F->lcp_s;
    float sum = 0.0f, power = 1.0f;
    for (int i = 0; i <= 3; i++) {
      sum += F->lcp[3 - i] * power;
      power *= x;
    }
    c[0] *= sum;
    c[1] *= sum;
    if (F->has_shift) {
      c[0] += F->sh_h;
      c[1] += F->sh_v;
    }
    if (F->has_shear) {
      float h0 = (float)((double)c[0] + (double)c[1] * F->shear_g);
      float h1 = (float)((double)c[1] + (double)c[0] * F->shear_t);
      c[0] = h0;
      c[1] = h1;
    }
  }
}

static int mount_mask(const facet_ctx* F, const float r[3], const float c[2]) {
  int m = (c[0] >= F->win_x0) && (c[0] <= F->win_x1) && (c[1] >= F->win_y0) && (c[1] <= F->win_y1);
  if (F->projection == EU_RECTILINEAR) m = m && (r[2] > 0.0f);
  return m;
}

/* environment::get_mask, environment.h:1567,1700-1760 */
static int facet_mask(const facet_ctx* F, const float r[3]) {
  if (F->mask_always) return 1;
  float c[2];
  mount_coordinate(F, r, c);
  return mount_mask(F, r, c);
}

/* repix_t, environment.h:1205-1309: channel-count adaptation between a facet and the target */
static void repix(int in_n, int out_n, const float* in, float* out) {
  if (in_n == out_n) {
    for (int i = 0; i < in_n; i++) out[i] = in[i];
    return;
  }
  switch (in_n) {
    case 1:
      if (out_n == 3) { out[0] = out[1] = out[2] = in[0]; }
      else if (out_n == 2) { out[0] = in[0]; out[1] = 1.0f; }
      else { out[0] = out[1] = out[2] = in[0]; out[3] = 1.0f; }
      break;
    case 2:
      if (out_n == 1) { out[0] = in[0] / in[1]; if (in[1] == 0.0f) out[0] = 0.0f; }
      else if (out_n == 3) { float g = in[0] / in[1]; if (in[1] == 0.0f) g = 0.0f; out[0] = out[1] = out[2] = g; }
      else { out[0] = out[1] = out[2] = in[0]; out[3] = in[1]; }
      break;
    case 3: {
      float sum = in[0];  /* xel_t::sum(): ((in0 + in1) + in2), zimt/xel.h */
      sum += in[1];
      sum += in[2];
      if (out_n == 1) out[0] = sum / 3.0f;
      else if (out_n == 2) { out[0] = sum / 3.0f; out[1] = 1.0f; }
      else { out[0] = in[0]; out[1] = in[1]; out[2] = in[2]; out[3] = 1.0f; }
      break;
    }
    default:
      if (out_n == 1) { out[0] = (in[0] + in[1] + in[2]) / 3.0f; out[0] /= in[3]; if (in[3] == 0.0f) out[0] = 0.0f; }
      else if (out_n == 2) { out[0] = (in[0] + in[1] + in[2]) / 3.0f; out[1] = in[3]; }
      else {
        out[0] = in[0] / in[3]; out[1] = in[1] / in[3]; out[2] = in[2] / in[3];
        if (in[3] == 0.0f) out[0] = out[1] = out[2] = 0.0f;
      }
  }
}

/* mono_t, environment.h:1325-1384: replaces repix_t for masked facets; one- and two-channel jobs only */
static void mono(int in_n, int out_n, const float* in, float* out) {
  if (in_n == out_n) {
    for (int i = 0; i < in_n; i++) out[i] = in[i];
  } else if (in_n == 1) {
    out[0] = in[0]; out[1] = 1.0f;
  } else if (in_n == 2) {
    out[0] = in[0] / in[1];
    if (in[1] == 0.0f) out[0] = 0.0f;
  } else if (in_n == 3) {
    out[0] = in[0];
    if (out_n == 2) out[1] = 1.0f;
  } else if (out_n == 1) {
    out[0] = in[0];
    out[0] /= in[3];
    if (in[3] == 0.0f) out[0] = 0.0f;
  } else {
    out[0] = in[0]; out[1] = in[3];
  }
}

/* environment::eval, environment.h:1821-1842 over mount_t::eval :1172-1196 or
 * cubemap_view_t::eval :1473-1486, then repix_t to the job's channel count (:1862-1960), then
 * brighten on the colour channels. Returns the cube face (or -1). */
static int facet_eval(const facet_ctx* F, int nch, const float r[3], float* px) {
  const orc_source_t* s = F->src;
  int snch = s->nch, face = -1;
  float sp[4] = {0, 0, 0, 0};
  int hit = 1;
  if (F->kind == KIND_MOUNT) {
    float c[2];
    mount_coordinate(F, r, c);
    if (!mount_mask(F, r, c)) {
      hit = 0; /* mount_t::eval zeroes the source-typed pixel (:1190-1193); repix then sees zeros */
    } else {
      /* source_t::md_to_spline, :988-1006 */
      float ix = (float)((double)c[0] - F->ext_x0);
      ix /= F->ext_w;
      ix *= (float)F->total_w;
      ix -= .5f;
      ix = ix - (float)F->win_xoff; /* crd_spl = image_crd - window offset, environment.h:1003-1005 */
      float iy = (float)((double)c[1] - F->ext_y0);
      iy /= F->ext_h;
      iy *= (float)F->total_h;
      iy -= .5f;
      iy = iy - (float)F->win_yoff;
      spline_eval(s, s->degree, s->wmat, ix, iy, sp);
    }
  } else {
    float in_face[2], pk[2];
    ray_to_cubeface(r, &face, in_face);
    if (F->kind == KIND_BIATAN6) {
      in_face[0] = (float)(4.0 / M_PI) * eu_atanf(in_face[0]);
      in_face[1] = (float)(4.0 / M_PI) * eu_atanf(in_face[1]);
    }
    /* cubemap_view_t::get_pickup_coordinate_px, :1452-1461 (float members) */
    pk[0] = in_face[0] + F->refc_md;
    pk[1] = in_face[1] + F->refc_md;
    pk[0] *= F->model_to_px;
    pk[1] *= F->model_to_px;
    pk[1] += (float)(face * F->section_px);
    pk[0] -= .5f;
    pk[1] -= .5f;
    spline_eval(s, s->degree, s->wmat, pk[0], pk[1], sp);
  }
  if (F->masked) {
    /* masking_t / alpha_masking_t instead of the evaluator (environment.h:947-961,1585-1587, masking.h:70-139): every
     * channel is `paint`, or - with an alpha channel - the colour channels are paint * alpha; a miss stays zero */
    if (hit) {
      if (snch == 1 || snch == 3) {
        for (int i = 0; i < snch; i++) sp[i] = F->paint;
      } else {
        sp[0] = F->paint * sp[snch - 1];
        if (snch == 4) sp[1] = sp[2] = sp[0];
      }
    }
    mono(snch, nch, sp, px);
  } else {
    repix(snch, nch, sp, px);
  }
  if (F->brighten != 1.0f) {
    int ncol = (nch == 2 || nch == 4) ? nch - 1 : nch;
    for (int i = 0; i < ncol; i++) px[i] *= F->brighten;
  }
  return hit ? face : -1;
}

/* _hdr_merge_syn::get_quality, envutil_payload.cc:1390-1442 */
static float hdr_quality(float grey, float optimum, int kind) {
  int large = grey > optimum;
  float distance = fabsf(optimum - grey);
  if (kind == 0 && !large) distance = 0.0f;
  if (kind == 2 && large) distance = 0.0f;
  float proximity = optimum - distance;
  return proximity / (optimum * optimum);
}

/* one synopsis evaluation for a set of per-facet rays: single facet, _voronoi_syn
 * (envutil_payload.cc:818-956) or _hdr_merge_syn (:1500-1622) */
static int synopsis(int mode, int nf, const facet_ctx* F, float (*rays)[3], int nch, float* px) {
  if (mode == 0) return facet_eval(&F[0], nch, rays[0], px);
  if (mode == 1) {
    int champion = -1;
    float max_z = -FLT_MAX;
    if (facet_mask(&F[0], rays[0])) {
      champion = 0;
      max_z = rays[0][2] * F[0].recip_step;
    }
    for (int i = 1; i < nf; i++) {
      if (!facet_mask(&F[i], rays[i])) continue;
      float cz = rays[i][2] * F[i].recip_step;
      if (cz > max_z) {
        max_z = cz;
        champion = i;
      }
    }
    if (champion < 0) {
      for (int c = 0; c < nch; c++) px[c] = 0.0f;
    } else {
      facet_eval(&F[champion], nch, rays[champion], px);
    }
    return champion;
  }
  /* hdr_merge (envutil_payload.cc:1500-1622); with alpha the colour is de-associated for the
   * weighted sum, the alpha is the maximum seen, and the result is re-associated */
  float acc[4] = {0, 0, 0, 0}, qsum = 0.0f, p[4];
  int na = (nch == 2 || nch == 4) ? nch - 1 : -1; /* alpha channel index, or -1 */
  for (int i = 0; i < nf; i++) {
    facet_eval(&F[i], nch, rays[i], p);
    float grey = (nch <= 2) ? p[0] : fmaxf(p[0], fmaxf(p[1], p[2]));
    /* std::max(r, std::max(g,b)) returns its first argument on ties - same value */
    float q = hdr_quality(grey, F[i].optimum, F[i].hdr_kind);
    if (na >= 0) q = p[na] * q; /* get_quality(grey, alpha, ...), :1400-1408 */
    qsum += q;
    if (na < 0) {
      for (int c = 0; c < nch; c++) acc[c] += p[c] * q;
    } else {
      for (int c = 0; c < na; c++) {
        float v = 0.0f;
        if (p[na] > 0.000001f) v = p[c] / p[na];
        acc[c] += v * q;
      }
      acc[na] = fmaxf(acc[na], p[na]);
    }
  }
  int ncol = na >= 0 ? na : nch;
  for (int c = 0; c < ncol; c++) {
    acc[c] /= qsum;
    if (!(qsum > 0.0f)) acc[c] = 0.0f;
    if (na >= 0) acc[c] *= acc[na];
  }
  for (int c = 0; c < nch; c++) px[c] = acc[c];
  return -1;
}

/* _voronoi_syn_plus::operator() (envutil_payload.cc:964-1233) for one zimt vector of `nl` lanes
 * (a 16-pixel run of a row): facets that the ray hits, sorted by z * recip_step (stable, strict
 * >), composited front to back as associated alpha. The reference takes a shortcut per VECTOR:
 * if every lane's front facet is the last facet that was hit at all (`next_best`) and every
 * lane's alpha is >= 1, the front facet's pixels are the result - which differs from the
 * composite where a b-spline overshoots alpha beyond 1. Hence the group-wise restatement. */
static void voronoi_plus_group(int nf, const facet_ctx* F, int nl, float (*rays)[64][3], int nch, float (*out)[4],
                               int* idx) {
  static __thread int ids[ORC_LANES][64];
  static __thread float zs[ORC_LANES][64];
  int cnt[ORC_LANES];
  int next_best = -1;
  for (int l = 0; l < nl; l++) cnt[l] = 0;
  for (int i = 0; i < nf; i++) {
    int any = 0;
    for (int l = 0; l < nl; l++) {
      if (!facet_mask(&F[i], rays[l][i])) continue;
      any = 1;
      float z = rays[l][i][2] * F[i].recip_step;
      int k = cnt[l]++;
      zs[l][k] = z;
      ids[l][k] = i;
      while (k > 0 && zs[l][k] > zs[l][k - 1]) { /* masked_swap while strictly greater */
        float tz = zs[l][k]; zs[l][k] = zs[l][k - 1]; zs[l][k - 1] = tz;
        int ti = ids[l][k]; ids[l][k] = ids[l][k - 1]; ids[l][k - 1] = ti;
        k--;
      }
    }
    if (any) next_best = i;
  }
  for (int l = 0; l < nl; l++) {
    for (int c = 0; c < nch; c++) out[l][c] = 0.0f;
    if (idx) idx[l] = cnt[l] ? ids[l][0] : -1;
  }
  if (next_best < 0) return;
  int all_top = 1;
  for (int l = 0; l < nl; l++)
    if (!(cnt[l] > 0 && ids[l][0] == next_best)) all_top = 0;
  if (all_top) {
    float help[ORC_LANES][4];
    int opaque = 1;
    for (int l = 0; l < nl; l++) {
      facet_eval(&F[next_best], nch, rays[l][next_best], help[l]);
      if (!(help[l][nch - 1] >= 1.0f)) opaque = 0;
    }
    if (opaque) {
      for (int l = 0; l < nl; l++)
        for (int c = 0; c < nch; c++) out[l][c] = help[l][c];
      return;
    }
  }
  for (int l = 0; l < nl; l++) {
    float help[4];
    for (int k = 0; k < cnt[l]; k++) {
      facet_eval(&F[ids[l][k]], nch, rays[l][ids[l][k]], help);
      if (k == 0) {
        for (int c = 0; c < nch; c++) out[l][c] = help[c];
      } else {
        for (int c = 0; c < nch; c++) out[l][c] += (1.0f - out[l][nch - 1]) * help[c];
      }
    }
  }
}

static void r3f(double roll, double pitch, double yaw, int inverse, float m[9]) { /* make_r3_t -> r3_t<float> */
  double d[9];
  orc_rotation(roll, pitch, yaw, inverse, d);
  for (int i = 0; i < 9; i++) m[i] = (float)d[i];
}
/* tf3d_t ctor, geometry.h:1865-1880 */
static void tf3d_stage(facet_ctx* F, int k, const float a[9], const float b[9], const float shift[3], float dcp) {
  memcpy(F->g_st[k].a, a, sizeof(float) * 9);
  memcpy(F->g_st[k].b, b, sizeof(float) * 9);
  matmulf(a, b, F->g_st[k].ab);
  for (int c = 0; c < 3; c++) F->g_st[k].shift[c] = shift[c];
  F->g_st[k].has_shift = (shift[0] != 0 || shift[1] != 0 || shift[2] != 0);
  F->g_st[k].dcp = dcp;
}
/* a translation given in model space, taken to the CS of its translation plane (envutil_payload.cc:1683-1695) */
static void plane_shift(const eu_facet_t* f, const float tp[9], float sh[3]) {
  sh[0] = (float)f->tr_x; sh[1] = (float)f->tr_y; sh[2] = (float)f->tr_z;
  if (f->tp_y != 0 || f->tp_p != 0 || f->tp_r != 0) { /* rotate(xel_t<double,3>(shift), r): in double */
    double sd[3] = {sh[0], sh[1], sh[2]}, o[3];
    for (int c = 0; c < 3; c++) o[c] = (sd[0] * tp[c] + sd[1] * tp[3 + c]) + sd[2] * tp[6 + c];
    for (int c = 0; c < 3; c++) sh[c] = (float)o[c];
  }
}
/* generic_r3(ft, fs), envutil_payload.cc:1636-1760: target facet ft in the camera position */
static void generic_r3_setup(const eu_facet_t* ft, const eu_facet_t* fs, facet_ctx* F) {
  float r_camera[9], rt_tp[9], rt_tpi[9], rs_tp[9], rs_tpi[9], r_facet[9], m1[9], m2[9];
  r3f(ft->roll, ft->pitch, ft->yaw, 0, r_camera);
  r3f(ft->tp_r, ft->tp_p, ft->tp_y, 1, rt_tp);
  r3f(ft->tp_r, ft->tp_p, ft->tp_y, 0, rt_tpi);
  r3f(fs->tp_r, fs->tp_p, fs->tp_y, 1, rs_tp);
  r3f(fs->tp_r, fs->tp_p, fs->tp_y, 0, rs_tpi);
  r3f(fs->roll, fs->pitch, fs->yaw, 1, r_facet);
  int have_ttp = (ft->tr_x != 0 || ft->tr_y != 0 || ft->tr_z != 0);
  int have_stp = (fs->tr_x != 0 || fs->tr_y != 0 || fs->tr_z != 0);
  float shift_t[3], shift_s[3];
  plane_shift(ft, rt_tp, shift_t);
  float dcp = (float)(1.0 - shift_t[2]);
  for (int c = 0; c < 3; c++) shift_t[c] = -shift_t[c];
  plane_shift(fs, rs_tp, shift_s);
  if (have_ttp) {
    matmulf(r_camera, rt_tp, m1);
    if (have_stp) {
      tf3d_stage(F, 0, m1, rt_tpi, shift_t, dcp);
      matmulf(rs_tpi, r_facet, m2);
      tf3d_stage(F, 1, rs_tp, m2, shift_s, 1.0f);
      F->g_nstage = 2;
    } else {
      matmulf(rt_tpi, r_facet, m2);
      tf3d_stage(F, 0, m1, m2, shift_t, dcp);
      F->g_nstage = 1;
    }
  } else if (have_stp) {
    matmulf(r_camera, rs_tp, m1);
    matmulf(rs_tpi, r_facet, m2);
    tf3d_stage(F, 0, m1, m2, shift_s, 1.0f);
    F->g_nstage = 1;
  } else { /* rotate_t(rotate(r_camera, r_facet)) */
    const float zero[3] = {0.f, 0.f, 0.f};
    float id[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    tf3d_stage(F, 0, r_camera, r_facet, zero, 1.0f);
    (void)id;
    F->g_nstage = 1;
  }
}

static int facet_setup(const eu_target_t* t, const eu_opts_t* o, const eu_facet_t* f, const orc_source_t* src,
                       facet_ctx* F, const eu_facet_t* ft) {
  memset(F, 0, sizeof(*F));
  F->src = src;
  F->projection = f->projection;
  F->kind = src->kind;
  /* basis = R_camera * R_facet^-1, envutil_payload.cc:1923-1948 */
  double cam[9], fct[9], m[9];
  orc_rotation(t->roll, t->pitch, t->yaw, 0, cam);
  orc_rotation(f->roll, f->pitch, f->yaw, 1, fct);
  mat_mul(cam, fct, m);
  for (int i = 0; i < 3; i++) {
    F->xx[i] = (float)m[i];
    F->yy[i] = (float)m[3 + i];
    F->zz[i] = (float)m[6 + i];
  }
  double e[4];
  orc_get_extent(f->projection, f->width, f->height, f->hfov, e);
  F->ext_x0 = e[0];
  F->ext_y0 = e[2];
  F->ext_w = (float)(e[1] - e[0]);
  F->ext_h = (float)(e[3] - e[2]);
  F->total_w = f->width;
  F->total_h = f->height;
  { /* window extent, environment.h:607-618: BOTH axes are scaled by widths (sic) */
    double wx = e[1] - e[0], wy = e[3] - e[2];
    int ww = f->window_width > 0 ? f->window_width : f->width;
    int xo = f->window_width > 0 ? f->window_x_offset : 0, yo = f->window_width > 0 ? f->window_y_offset : 0;
    double px0 = (double)xo / f->width, py0 = (double)yo / f->width;
    double px1 = (double)(xo + ww) / f->width, py1 = (double)(yo + ww) / f->width;
    F->win_x0 = (float)(e[0] + px0 * wx);
    F->win_y0 = (float)(e[2] + py0 * wy);
    F->win_x1 = (float)(e[0] + px1 * wx);
    F->win_y1 = (float)(e[2] + py1 * wy);
    F->win_xoff = xo;
    F->win_yoff = yo;
  }
  F->mask_always = (src->kind != KIND_MOUNT) || (f->projection == EU_FISHEYE && f->hfov >= M_PI * 2.0);
  /* process_geometry, envutil_basic.h:499-543 */
  F->has_lcp = (f->a != 0.0 || f->b != 0.0 || f->c != 0.0);
  F->has_shift = (f->h != 0.0 || f->v != 0.0);
  F->has_shear = (f->shear_g != 0.0 || f->shear_t != 0.0);
  {
    double dv = fabs(e[3] - e[2]) / 2.0, dh = fabs(e[1] - e[0]) / 2.0;
    double sref = dh < dv ? dh : dv;
    double factor = fabs(e[1] - e[0]) / f->width;
    float a = (float)f->a, b = (float)f->b, c = (float)f->c;
    F->lcp[0] = a; F->lcp[1] = b; F->lcp[2] = c;
    F->lcp[3] = 1.0f - (a + b + c);
    F->lcp_s = (float)sref;
    F->sh_h = (float)(f->h * factor);
    F->sh_v = (float)(f->v * factor);
    F->shear_g = f->shear_g;
    F->shear_t = f->shear_t;
  }
  if (src->kind != KIND_MOUNT) {
    F->refc_md = (float)src->cm.refc_md;
    F->model_to_px = (float)src->cm.model_to_px;
    F->section_px = src->cm.section_px;
  }
  F->generic = (f->tr_x != 0 || f->tr_y != 0 || f->tr_z != 0); /* has_translation, envutil_basic.h:505 */
  if (F->generic) { /* generic_r3(ft, fs) with an untranslated target, envutil_payload.cc:1640-1716 */
    double d[9];
    float r_camera[9], rs_tp[9], rs_tpi[9], r_facet[9];
    orc_rotation(t->roll, t->pitch, t->yaw, 0, d);
    for (int i = 0; i < 9; i++) r_camera[i] = (float)d[i];
    orc_rotation(f->tp_r, f->tp_p, f->tp_y, 1, d);
    for (int i = 0; i < 9; i++) rs_tp[i] = (float)d[i];
    orc_rotation(f->tp_r, f->tp_p, f->tp_y, 0, d);
    for (int i = 0; i < 9; i++) rs_tpi[i] = (float)d[i];
    orc_rotation(f->roll, f->pitch, f->yaw, 1, d);
    for (int i = 0; i < 9; i++) r_facet[i] = (float)d[i];
    float sh[3] = {(float)f->tr_x, (float)f->tr_y, (float)f->tr_z};
    if (f->tp_y != 0 || f->tp_p != 0 || f->tp_r != 0) { /* rotate(xel_t<double,3>(shift), rs_tp): in double */
      double sd[3] = {sh[0], sh[1], sh[2]}, o[3];
      for (int c = 0; c < 3; c++) o[c] = (sd[0] * rs_tp[c] + sd[1] * rs_tp[3 + c]) + sd[2] * rs_tp[6 + c];
      for (int c = 0; c < 3; c++) sh[c] = (float)o[c];
    }
    matmulf(r_camera, rs_tp, F->g_t2m);
    matmulf(rs_tpi, r_facet, F->g_m2s);
    for (int c = 0; c < 3; c++) F->g_shift[c] = sh[c];
    F->g_dcp = 1.0f;
  }
  if (ft) { /* 'single' target with lens correction / translation: every facet steps generically (:2063-2068) */
    F->generic = 1;
    generic_r3_setup(ft, f, F);
  }
  double step = orc_get_step(f->projection, f->width, f->height, f->hfov);
  F->recip_step = (float)(1.0 / step);
  F->brighten = (float)(f->brighten == 0.0 ? 1.0 : f->brighten);
  F->masked = f->masked != 0;
  F->paint = f->masked == 2 ? 1.0f : 0.0f;
  (void)o;
  return 0;
}

static int orc_render_impl(const eu_target_t* t, const eu_opts_t* o, int nf, const eu_facet_t* facets,
                           orc_source_t* const* sources, const eu_tap_t* taps, int n_taps, int row0, int row1,
                           float* out, int32_t* index_out, int n_threads);
int orc_render(const eu_target_t* t, const eu_opts_t* o, int nf, const eu_facet_t* facets,
               orc_source_t* const* sources, const eu_tap_t* taps, int n_taps, int row0, int row1, float* out,
               int32_t* index_out, int n_threads) {
  g_contract = g_contract_req; /* the opt-in arithmetic holds for the render only, see orc_set_arithmetic */
  int rc = orc_render_impl(t, o, nf, facets, sources, taps, n_taps, row0, row1, out, index_out, n_threads);
  g_contract = 0;
  return rc;
}
static int orc_render_impl(const eu_target_t* t, const eu_opts_t* o, int nf, const eu_facet_t* facets,
                           orc_source_t* const* sources, const eu_tap_t* taps, int n_taps, int row0, int row1,
                           float* out, int32_t* index_out, int n_threads) {
  if (nf < 1 || nf > 64) return EU_ERR_ARGUMENT;
  target_ctx T;
  target_setup(t, &T);
  const eu_facet_t* ft = NULL;
  if (t->single > 0 && t->single <= nf) {
    const eu_facet_t* cand = &facets[t->single - 1];
    if (cand->has_2d_tf || cand->has_translation) ft = cand;
  }
  if (ft && ft->has_2d_tf) { /* tf22 = pto_planar<float, L, true>(ft) */
    T.inv.on = 1;
    T.inv.has_shear = ft->has_shear;
    T.inv.has_shift = ft->has_shift;
    T.inv.has_lcp = ft->has_lcp;
    T.inv.shear_g = ft->shear_g;
    T.inv.shear_t = ft->shear_t;
    T.inv.s = ft->s;
    T.inv.h = (float)ft->shift_h;
    T.inv.v = (float)ft->shift_v;
    if (ft->has_lcp && inv_lcp_setup(&T.inv, ft->a, ft->b, ft->c, ft->r_max) != 0) return -1;
  }
  facet_ctx* F = (facet_ctx*)calloc((size_t)nf, sizeof(facet_ctx));
  for (int i = 0; i < nf; i++) facet_setup(t, o, &facets[i], sources[i], &F[i], ft);
  int nch = t->nchannels;
  int mode = 0;
  int first = 0;
  if (nf > 1 && o->solo < 0) mode = (o->synopsis == EU_SYN_HDR_MERGE) ? 2 : 1;
  if (nf > 1 && o->solo >= 0) first = o->solo;
  /* normalize: envutil_payload.cc:2105,2118 (false: one facet, no twining), else true */
  T.normalize = !(mode == 0 && n_taps == 0);
  if (mode == 2) { /* _hdr_merge_syn ctor, :1354-1375 */
    float lowest = 100000.0f, highest = -1.0f;
    int lo = -1, hi = -1;
    for (int i = 0; i < nf; i++) {
      double br = (float)(facets[i].brighten == 0.0 ? 1.0 : facets[i].brighten);
      F[i].optimum = (float)(0.5f * br);
      if (br < lowest) { lowest = (float)br; lo = i; }
      if (br > highest) { highest = (float)br; hi = i; }
    }
    for (int i = 0; i < nf; i++) F[i].hdr_kind = (i == lo) ? 0 : (i == hi) ? 2 : 1;
  }
  int nfe = mode == 0 ? 1 : nf;
  const facet_ctx* FE = mode == 0 ? &F[first] : F;
#ifdef _OPENMP
  if (n_threads <= 0) n_threads = omp_get_max_threads();
#else
  n_threads = 1;
#endif
  if (mode == 1 && (nch == 2 || nch == 4)) mode = 3; /* roll_out: voronoi_syn_plus for alpha, :2306-2311 */
  const float unbrighten = (t->gain == 0.0) ? 1.0f : (float)t->gain;
#pragma omp parallel for schedule(dynamic, 4) num_threads(n_threads)
  for (int y = row0; y < row1; y++) {
    /* one zimt vector at a time: 16 consecutive pixels of the row (segments are multiples of 16) */
    float (*rays)[64][3] = (float (*)[64][3])malloc(sizeof(float) * ORC_LANES * 64 * 3);
    float (*r10)[64][3] = (float (*)[64][3])malloc(sizeof(float) * ORC_LANES * 64 * 3);
    float (*r01)[64][3] = (float (*)[64][3])malloc(sizeof(float) * ORC_LANES * 64 * 3);
    float (*sub)[64][3] = (float (*)[64][3])malloc(sizeof(float) * ORC_LANES * 64 * 3);
    for (int g0 = 0; g0 < T.width; g0 += ORC_LANES) {
      int nl = T.width - g0 < ORC_LANES ? T.width - g0 : ORC_LANES;
      float res[ORC_LANES][4], help[ORC_LANES][4];
      int ids[ORC_LANES], hid[ORC_LANES];
      float p0y = planar_y(&T, y, 0.0f), p1y = planar_y(&T, y, T.bias_y);
      for (int l = 0; l < nl; l++) {
        int x = g0 + l;
        float p0x = planar_x(&T, x, 0.0f);
        for (int i = 0; i < nfe; i++) stepper_ray(&T, &FE[i], p0x, p0y, x, y, 0.0f, rays[l][i]);
        if (n_taps) {
          float p1x = planar_x(&T, x, T.bias_x);
          for (int i = 0; i < nfe; i++) {
            stepper_ray(&T, &FE[i], p1x, p0y, x, y, T.bias_x, r10[l][i]);
            stepper_ray(&T, &FE[i], p0x, p1y, x, y, 0.0f, r01[l][i]);
          }
        }
      }
      if (n_taps == 0) {
        if (mode == 3) voronoi_plus_group(nfe, FE, nl, rays, nch, res, ids);
        else
          for (int l = 0; l < nl; l++) ids[l] = synopsis(mode, nfe, FE, rays[l], nch, res[l]);
      } else {
        /* deriv_stepper stepper.h:1606-1694; twine_t twining.h:106-263; synopsis_t
         * envutil_payload.cc:647-690 */
        for (int l = 0; l < nl; l++) {
          ids[l] = -1;
          for (int c = 0; c < 4; c++) res[l][c] = 0.0f;
        }
        for (int k = 0; k < n_taps; k++) {
          float cx = taps[k].x * 4.0f, cy = taps[k].y * 4.0f, cw = taps[k].w;
          for (int l = 0; l < nl; l++)
            for (int i = 0; i < nfe; i++)
              for (int c = 0; c < 3; c++) {
                float du = r10[l][i][c] - rays[l][i][c], dv = r01[l][i][c] - rays[l][i][c];
                sub[l][i][c] = rays[l][i][c] + cx * du + cy * dv;
              }
          if (mode == 3) voronoi_plus_group(nfe, FE, nl, sub, nch, help, hid);
          else
            for (int l = 0; l < nl; l++) hid[l] = synopsis(mode, nfe, FE, sub[l], nch, help[l]);
          for (int l = 0; l < nl; l++) {
            if (k == 0) ids[l] = hid[l];
            for (int c = 0; c < nch; c++) res[l][c] = win_muladd(cw, help[l][c], res[l][c]);
          }
        }
      }
      for (int l = 0; l < nl; l++) {
        float* px = out + ((size_t)(y - row0) * T.width + g0 + l) * nch;
        if (unbrighten != 1.0f) { /* amplify_type in work(), envutil_payload.cc:481-511 */
          int ncol = (nch == 2 || nch == 4) ? nch - 1 : nch;
          for (int c = 0; c < ncol; c++) res[l][c] *= unbrighten;
        }
        for (int c = 0; c < nch; c++) px[c] = res[l][c];
        if (index_out) index_out[(size_t)(y - row0) * T.width + g0 + l] = ids[l];
      }
    }
    free(rays); free(r10); free(r01); free(sub);
  }
  free(F);
  return 0;
}

/* ---- tethered output: lut_based_tf + to_screen_t (envutil_payload.cc:221-413) --------------------------------
 * The knots: `std::function<float(float)> fn = RGB2sRGB<double, double>` is called with the double x = i / 255.0,
 * so x is narrowed to float on the way in and the result on the way out; `double y = fn(x) * 255.0` is stored
 * into the float core (:262-267). Degree 1 needs no prefilter. The evaluator is the safe 1-D one: clamp to
 * [0, 255] (NATURAL, zimt/eval.h:2101-2110), floor/remainder split, _eval_linear level 0 (zimt/eval.h:1037-1059):
 * sum = c[i] * (1 - t); sum += c[i + 1] * t. c[256] is the NATURAL brace 2 c[255] - c[254]; it only ever meets t = 0. */
void orc_screen_lut(float lut[257]) {
  for (int i = 0; i < 256; i++) {
    double x = i / 255.0;
    double v = (double)(float)x;
    double r = 1.055 * pow(v, 0.41666666666666667) - 0.055;
    if (v <= 0.0031308) r = 12.92 * v;
    float fr = (float)r;
    double y = (double)fr * 255.0;
    lut[i] = (float)y;
  }
  lut[256] = 2.0f * lut[255] - lut[254];
}

static uint32_t screen_channel(const float* lut, float in) {
  float c = in * 255.0f;
  if (c < 0.0f) c = 0.0f;
  else if (c > 255.0f) c = 255.0f;
  float fl = floorf(c), t = c - fl;
  int i = (int)fl;
  float wl = 1.0f - t, wr = t;
  float sum = lut[i];
  sum *= wl;
  float help = lut[i + 1];
  help = help * wr;
  sum += help;
  return (uint32_t)sum;
}

void orc_to_screen(const float* px, int nch, size_t n, uint32_t* out) {
  float lut[257];
  orc_screen_lut(lut);
  for (size_t k = 0; k < n; k++) {
    const float* p = px + k * (size_t)nch;
    uint32_t c[4];
    for (int j = 0; j < nch; j++) c[j] = screen_channel(lut, p[j]);
    uint32_t o;
    switch (nch) {
      case 1: o = 0xFF000000u | (c[0] << 16) | (c[0] << 8) | c[0]; break;
      case 2: o = (c[1] << 24) | (c[0] << 16) | (c[0] << 8) | c[0]; break;
      case 3: o = 0xFF000000u | (c[2] << 16) | (c[1] << 8) | c[0]; break;
      default: o = (c[3] << 24) | (c[2] << 16) | (c[1] << 8) | c[0]; break;
    }
    out[k] = o;
  }
}
