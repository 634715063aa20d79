// host_setup.cc - host-side set-up arithmetic of the B200 back-end (no CUDA in this file).
//
// Everything here runs once per job on the host and produces the scalars the kernels consume.
// It restates the reference's arithmetic in the reference's own precision (double set-up,
// float quaternion) because those scalars define the float inputs of every per-pixel result:
//   get_vfov/get_step/get_extent   reference envutil_basic.cc:50-226
//   facet / target set-up           reference envutil_main.cc:483-510,935-976,1199-1232,
//                                   envutil_basic.h:499-543 (process_geometry)
//   Euler -> quaternion -> rows     reference envutil_payload.cc:136-218 (Imath ZXY order)
//   basis = R_camera * R_facet^-1   reference envutil_payload.cc:1923-1948, geometry.h:79-97
//   make_spread / twine_setup       reference envutil_main.cc:1253-1355,1405-1616
//   metrics_t                       reference cubemap.h:233-400
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

#include "envutil_b200.h"
#include "eu_math.h"
#include "host_setup.h"

extern "C" {

double eu_get_vfov(int projection, int width, int height, double hfov) {
  double vfov = 0.0;
  switch (projection) {
    case EU_RECTILINEAR:
      vfov = 2.0 * eu_atan(height * tan(hfov / 2.0) / width);
      break;
    case EU_CYLINDRICAL: {
      double pixels_per_rad = width / hfov;
      double h_rad = height / pixels_per_rad;
      vfov = 2.0 * eu_atan(h_rad / 2.0);
      break;
    }
    case EU_STEREOGRAPHIC: {
      double w_rad = 2.0 * tan(hfov / 4.0);
      double pixels_per_rad = width / w_rad;
      double h_rad = height / pixels_per_rad;
      vfov = 4.0 * eu_atan(h_rad / 2.0);
      break;
    }
    case EU_SPHERICAL:
    case EU_FISHEYE:
      vfov = hfov * height / width;
      break;
    default:
      // the reference's CUBEMAP/BIATAN6 case has no break and falls into its default
      // (envutil_basic.cc:85-94); the cubemap extent does not use the value
      vfov = hfov;
      break;
  }
  return vfov;
}

double eu_get_step(int projection, int width, int height, double hfov) {
  (void)height;
  double step = 0.0;
  switch (projection) {
    case EU_RECTILINEAR:
    case EU_CUBEMAP:
      step = eu_atan(2.0 * tan(hfov / 2.0) / width);
      break;
    case EU_BIATAN6:
    case EU_SPHERICAL:
    case EU_CYLINDRICAL:
    case EU_FISHEYE:
      step = hfov / width;
      break;
    case EU_STEREOGRAPHIC:
      step = eu_atan(4.0 * tan(hfov / 4.0) / width);
      break;
    default:
      break;
  }
  return step;
}

void eu_get_extent(int projection, int width, int height, double hfov, double ext[4]) {
  double x0, x1, y0, y1;
  double alpha_x = -hfov / 2.0;
  double beta_x = hfov / 2.0;
  double beta_y = eu_get_vfov(projection, width, height, hfov) / 2.0;
  double alpha_y = -beta_y;
  switch (projection) {
    case EU_SPHERICAL:
    case EU_FISHEYE:
      x0 = alpha_x; x1 = beta_x; y0 = alpha_y; y1 = beta_y;
      break;
    case EU_CYLINDRICAL:
      x0 = alpha_x; x1 = beta_x; y0 = tan(alpha_y); y1 = tan(beta_y);
      break;
    case EU_RECTILINEAR:
      x0 = tan(alpha_x); x1 = tan(beta_x); y0 = tan(alpha_y); y1 = tan(beta_y);
      break;
    case EU_STEREOGRAPHIC:
      x0 = 2.0 * tan(alpha_x / 2.0); x1 = 2.0 * tan(beta_x / 2.0);
      y0 = 2.0 * tan(alpha_y / 2.0); y1 = 2.0 * tan(beta_y / 2.0);
      break;
    case EU_CUBEMAP:
    case EU_BIATAN6:
      x0 = tan(alpha_x); x1 = tan(beta_x); y0 = 6 * x0; y1 = 6 * x1;
      break;
    default:
      x0 = x1 = y0 = y1 = 0.0;
      break;
  }
  ext[0] = x0; ext[1] = x1; ext[2] = y0; ext[3] = y1;
}

int eu_facet_prepare(eu_facet_t* f) {
  if (!f || f->width <= 0 || f->height <= 0 || f->projection < 0 || f->projection >= EU_PRJ_NONE ||
      !(f->hfov > 0.0) || f->nchannels < 1 || f->nchannels > 4)
    return EU_ERR_ARGUMENT;
  if (f->window_width <= 0 || f->window_height <= 0) {  // no 'W' window given: the whole image
    f->window_width = f->width;
    f->window_height = f->height;
    f->window_x_offset = f->window_y_offset = 0;
  }
  f->step = eu_get_step(f->projection, f->width, f->height, f->hfov);
  double e[4];
  eu_get_extent(f->projection, f->width, f->height, f->hfov, e);
  f->x0 = e[0]; f->x1 = e[1]; f->y0 = e[2]; f->y1 = e[3];
  // process_geometry (envutil_basic.h:499-543)
  f->has_shift = (f->h != 0.0 || f->v != 0.0);
  f->has_lcp = (f->a != 0.0 || f->b != 0.0 || f->c != 0.0);
  f->has_shear = (f->shear_g != 0.0 || f->shear_t != 0.0);
  f->has_2d_tf = (f->has_shift || f->has_lcp || f->has_shear);
  f->has_translation = (f->tr_x != 0 || f->tr_y != 0 || f->tr_z != 0);
  double dv = fabs(f->y1 - f->y0) / 2.0;
  double dh = fabs(f->x1 - f->x0) / 2.0;
  f->s = (dh < dv) ? dh : dv;
  double aspect = (dh >= dv) ? dh / dv : dv / dh;
  f->r_max = sqrt(1 + aspect * aspect);
  f->d = 1.0 - (f->a + f->b + f->c);
  double factor = fabs(f->x1 - f->x0) / f->width;
  f->shift_h = f->h * factor;
  f->shift_v = f->v * factor;
  // the reference really adds y0 twice instead of squaring it (envutil_basic.h:531-534)
  double d1 = f->x0 * f->x0 + f->y0 + f->y0;
  double d2 = f->x1 * f->x1 + f->y0 + f->y0;
  double d3 = f->x0 * f->x0 + f->y1 + f->y1;
  double d4 = f->x1 * f->x1 + f->y1 + f->y1;
  d1 = std::max(d1, d2);
  d1 = std::max(d1, d3);
  d1 = std::max(d1, d4);
  f->cap_radius = sqrt(d1);
  if (f->brighten == 0.0) f->brighten = 1.0;
  return EU_OK;
}

int eu_target_prepare(eu_target_t* t) {
  if (!t || t->projection < 0 || t->projection >= EU_PRJ_NONE) return EU_ERR_ARGUMENT;
  if (t->width == 0) t->width = 1024;  // envutil_main.cc:483-486
  if (t->projection == EU_CUBEMAP || t->projection == EU_BIATAN6) {
    t->height = 6 * t->width;
    if (t->hfov < M_PI_2 - 1e-12) return EU_ERR_ARGUMENT;  // assert(hfov >= 90), :499-503
  }
  if (t->projection == EU_SPHERICAL && t->height == 0) {
    if (t->width & 1) ++t->width;
    t->height = t->width / 2;
  }
  if (t->height == 0) t->height = t->width;
  if (t->width <= 0 || t->height <= 0 || !(t->hfov > 0.0)) return EU_ERR_ARGUMENT;
  double e[4];
  eu_get_extent(t->projection, t->width, t->height, t->hfov, e);
  t->x0 = e[0]; t->x1 = e[1]; t->y0 = e[2]; t->y1 = e[3];
  t->step = (t->x1 - t->x0) / t->width;
  return EU_OK;
}

// Imath::Eulerf(roll, pitch, yaw, ZXY).toQuat(): static frame, even parity, axes i,j,k = Z,X,Y.
// The quaternion is FLOAT even when the caller works in double (envutil_payload.cc:152).
void eu_rotation_matrix(double roll_d, double pitch_d, double yaw_d, int inverse, double m[9]) {
  float roll = float(roll_d), pitch = float(pitch_d), yaw = float(yaw_d);
  float ti = roll * 0.5f, tj = pitch * 0.5f, th = yaw * 0.5f;
  // sin/cos: the back-end's own binary32 functions (include/eu_math.h)
  float ci, cj, ch, si, sj, sh;
  eu_sincosf(ti, &si, &ci);
  eu_sincosf(tj, &sj, &cj);
  eu_sincosf(th, &sh, &ch);
  float cc = ci * ch, cs = ci * sh, sc = si * ch, ss = si * sh;
  float qv[3], qr;
  qv[2] = cj * sc - sj * cs;          // a[i], i = Z
  qv[0] = (cj * ss + sj * cc) * 1.0f; // a[j], j = X
  qv[1] = cj * cs - sj * sc;          // a[k], k = Y
  qr = cj * cc + sj * ss;
  // the reference holds the quaternion as Imath::Quat<double> (rotate_3d<double,1>): the float
  // result of toQuat() is widened, and Quat::invert() then runs in double
  double r = qr, v0 = qv[0], v1 = qv[1], v2 = qv[2];
  if (inverse) {
    double qdot = r * r + ((v0 * v0 + v1 * v1) + v2 * v2);
    r /= qdot;
    v0 = -v0 / qdot;
    v1 = -v1 / qdot;
    v2 = -v2 / qdot;
  }
  // rows = e_k * Quat<double>(q):  v + 2 (q.r (q.v x v) + q.v x (q.v x v))
  for (int k = 0; k < 3; k++) {
    double e[3] = {0.0, 0.0, 0.0};
    e[k] = 1.0;
    double a[3] = {v1 * e[2] - v2 * e[1], v2 * e[0] - v0 * e[2], v0 * e[1] - v1 * e[0]};
    double b[3] = {v1 * a[2] - v2 * a[1], v2 * a[0] - v0 * a[2], v0 * a[1] - v1 * a[0]};
    for (int c = 0; c < 3; c++) m[3 * k + c] = e[c] + 2.0 * (r * a[c] + b[c]);
  }
}

// rotate(r_camera, r_facet): row_i = sum_j cam[i][j] * fct[j]   (geometry.h:79-97)
void eu_facet_basis(const eu_target_t* t, const eu_facet_t* f, double m[9]) {
  double cam[9], fct[9];
  eu_rotation_matrix(t->roll, t->pitch, t->yaw, 0, cam);
  eu_rotation_matrix(f->roll, f->pitch, f->yaw, 1, fct);
  for (int i = 0; i < 3; i++)
    for (int c = 0; c < 3; c++)
      m[3 * i + c] = (cam[3 * i + 0] * fct[0 + c] + cam[3 * i + 1] * fct[3 + c]) + cam[3 * i + 2] * fct[6 + c];
}

static void make_spread(std::vector<eu_tap_t>& trg, int w, int h, float d, float sigma, float threshold) {
  if (w <= 2) w = 2;
  if (h <= 0) h = w;
  float wgt = 1.0 / (w * h);
  double x0 = -(w - 1.0) / (2.0 * w);
  double dx = 1.0 / w;
  double y0 = -(h - 1.0) / (2.0 * h);
  double dy = 1.0 / h;
  trg.clear();
  sigma *= -x0;
  double sum = 0.0;
  for (int y = 0; y < h; y++) {
    for (int x = 0; x < w; x++) {
      float wf = 1.0;
      if (sigma > 0.0) {
        double wx = (x0 + x * dx) / sigma;
        double wy = (y0 + y * dy) / sigma;
        wf = exp(-sqrt(wx * wx + wy * wy));
      }
      eu_tap_t v = {float(d * (x0 + x * dx)), float(d * (y0 + y * dy)), wf * wgt};
      trg.push_back(v);
      sum += wf * wgt;
    }
  }
  double th_sum = 0.0;
  bool renormalize = false;
  if (sigma != 0.0) {
    for (auto& v : trg) {
      v.w /= sum;
      if (v.w >= threshold) {
        th_sum += v.w;
      } else {
        renormalize = true;
        v.w = 0.0f;
      }
    }
    if (renormalize)
      for (auto& v : trg) v.w /= th_sum;
  }
  if (renormalize) {
    std::vector<eu_tap_t> help = trg;
    trg.clear();
    for (auto v : help)
      if (v.w > 0.0f) trg.push_back(v);
  }
}

int eu_make_spread(const eu_target_t* t, const eu_opts_t* o, int n_facets, const eu_facet_t* facets, int twine,
                   double twine_width, double twine_density, double twine_sigma, double twine_threshold,
                   int twine_max, eu_tap_t* taps, int max_taps, int* twine_out) {
  if (!o) return EU_ERR_ARGUMENT;
  return eu_make_spread_ex(t, n_facets, facets, o->spline_degree, o->solo, twine, twine_width, twine_density,
                           twine_sigma, twine_threshold, twine_max, taps, max_taps, twine_out);
}

int eu_cubemap_metrics(int face_px, double hfov, int support_min, int tile_size, int32_t out_i[4],
                       double out_d[4]) {
  eu_cubemap_metrics_t m;
  int rc = eu_compute_cubemap_metrics(face_px, hfov, support_min, tile_size, &m);
  if (rc != EU_OK) return rc;
  out_i[0] = m.section_px; out_i[1] = m.left_frame_px; out_i[2] = m.right_frame_px; out_i[3] = m.n_tiles;
  out_d[0] = m.refc_md; out_d[1] = m.model_to_px; out_d[2] = m.section_md; out_d[3] = m.px_to_model;
  return EU_OK;
}

// The knots of to_screen_t's transfer function (lut_based_tf's constructor, envutil_payload.cc:243-267, with
// fn = RGB2sRGB<double, double> :221-231 and amplify = 255). fn is held as std::function<float(float)>, so the knot
// position i / 255.0 is narrowed to float on the way in and the curve's value on the way out, before the
// multiplication by 255 (in double) and the store into the float core. lut[256] is the NATURAL brace value
// (2 c[255] - c[254], zimt/brace.h:254-266): the evaluator only ever multiplies it by 0.
void eu_screen_lut(float lut[257]) {
  for (int i = 0; i < 256; i++) {
    const double x = i / 255.0;
    const double v = (double)(float)x;
    double r = 1.055 * std::pow(v, 0.41666666666666667) - 0.055;
    if (v <= 0.0031308) r = 12.92 * v;
    const float fr = (float)r;
    lut[i] = (float)((double)fr * 255.0);
  }
  lut[256] = 2.0f * lut[255] - lut[254];
}

}  // extern "C"

// arguments::twine_setup (envutil_main.cc:1405-1616) incl. its quirks: `solo > 0` (not >= 0),
// twine 1 still yields the 2x2 kernel of make_spread, twine_width is replaced by the
// magnification when a bilinear source is magnified.
int eu_make_spread_ex(const eu_target_t* t, int n_facets, const eu_facet_t* facets, int spline_degree,
                      int solo, int twine, double twine_width, double twine_density, double twine_sigma,
                      double twine_threshold, int twine_max, eu_tap_t* taps, int max_taps, int* twine_out) {
  if (!t || n_facets < 1 || !facets) return EU_ERR_ARGUMENT;
  if (twine != -1) {
    if (twine < 0) twine = 0;
    if (twine > 0 && !(twine_width > 0.0)) return EU_ERR_ARGUMENT;
  } else {
    double smallest_step = std::numeric_limits<double>::max();
    if (n_facets == 1 || solo > 0) {
      smallest_step = facets[n_facets == 1 ? 0 : solo].step;
    } else {
      for (int i = 0; i < n_facets; i++) smallest_step = std::min(facets[i].step, smallest_step);
    }
    double mag = smallest_step / t->step;
    if (mag > 1.0) {
      if (spline_degree > 1) {
        if (n_facets > 1) twine = 3;
        else if (mag < 2.0) twine = 2;
        else twine = 1;
      } else {
        twine = std::min(5, int(1.0 + mag));
        twine_width = mag;
      }
    } else {
      twine = int(1.0 + 1.0 / mag);
      twine = std::min(twine_max, twine);
      twine_width = 1.0;
    }
  }
  if (float(twine_density) != 1.0f) twine = int(std::round(twine * twine_density));
  if (twine_out) *twine_out = twine;
  if (twine == 0) return 0;
  std::vector<eu_tap_t> spread;
  make_spread(spread, twine, twine, float(twine_width), float(twine_sigma), float(twine_threshold));
  if (int(spread.size()) > max_taps) return EU_ERR_ARGUMENT;
  for (size_t i = 0; i < spread.size(); i++) taps[i] = spread[i];
  return int(spread.size());
}

// metrics_t (cubemap.h:233-400)
int eu_compute_cubemap_metrics(int face_px, double face_fov, int support_min, int tile_px,
                               eu_cubemap_metrics_t* m) {
  if (face_px <= 0 || tile_px <= 0 || (tile_px & (tile_px - 1)) != 0 || support_min < 0) return EU_ERR_ARGUMENT;
  if (face_fov < M_PI_2) return EU_ERR_ARGUMENT;
  double overscan_md = 0.0, radius_md = 1.0, diameter_md = 2.0;
  if (face_fov > M_PI_2) {
    radius_md = tan(face_fov / 2.0);
    diameter_md = 2.0 * radius_md;
    overscan_md = radius_md - 1.0;
  }
  m->face_px = face_px;
  m->model_to_px = double(face_px) / diameter_md;
  m->px_to_model = diameter_md / double(face_px);
  double px_overscan = m->model_to_px * overscan_md;
  long inherent_support_px = (long)std::trunc(px_overscan);
  long additional = 0;
  if (inherent_support_px < support_min) additional = support_min - inherent_support_px;
  long px_min = face_px + 2 * additional;
  long n_tiles = px_min / tile_px;
  if (n_tiles * tile_px < px_min) n_tiles++;
  m->n_tiles = int(n_tiles);
  m->section_px = int(n_tiles * tile_px);
  long frame_total = m->section_px - face_px;
  m->left_frame_px = int(frame_total / 2);
  m->right_frame_px = int(frame_total - m->left_frame_px);
  m->section_md = m->px_to_model * m->section_px;
  double refc_px = double(m->left_frame_px) + double(face_px) / 2.0;
  m->refc_md = m->px_to_model * refc_px;
  return EU_OK;
}

// fill_polygon, envutil_basic.cc:236-321: scan-line fill with the non-zero winding rule; node x
// positions are truncated to int exactly as there
static void fill_polygon_clear(const float* px, const float* py, int N, int w, int h, unsigned char* plane) {
  std::vector<int> nodeX(N + 1), dir(N + 1);
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
        std::swap(nodeX[i], nodeX[i + 1]);
        std::swap(dir[i], dir[i + 1]);
        if (i) i--;
      } else {
        i++;
      }
    }
    int w_ord = 0;
    for (i = 0; i < nodes; i++) {
      w_ord += dir[i];
      if (!w_ord) continue;
      if (i + 1 >= nodes) break;
      if (nodeX[i] >= w) break;
      if (nodeX[i + 1] > 0) {
        if (nodeX[i] < 0) nodeX[i] = 0;
        if (nodeX[i + 1] > w) nodeX[i + 1] = w;
        for (int x = nodeX[i]; x < nodeX[i + 1]; x++) plane[(size_t)pixelY * w + x] = 0;
      }
    }
  }
}

void eu_build_alpha_mask(const eu_facet_t* f, const eu_alpha_spec_t* a, unsigned char* plane) {
  const int w = f->window_width, h = f->window_height;
  memset(plane, 1, (size_t)w * h);
  const float* xy = a->mask_xy;
  for (int m = 0; m < a->n_masks; m++) {
    int n = a->mask_sizes[m];
    std::vector<float> vx(n), vy(n);
    for (int k = 0; k < n; k++) {
      vx[k] = xy[2 * k];
      vy[k] = xy[2 * k + 1];
    }
    if (n >= 3) fill_polygon_clear(vx.data(), vy.data(), n, w, h, plane);
    xy += 2 * n;
  }
  if (a->has_crop) {
    float ca = (float)(fabs((double)(a->crop_x1 - a->crop_x0)) / 2.0);
    float cb = (float)(fabs((double)(a->crop_y1 - a->crop_y0)) / 2.0);
    if (f->projection == EU_FISHEYE) {  // elliptic crop, environment.h:746-772
      float mx = (float)((a->crop_x0 + a->crop_x1) / 2.0);
      float my = (float)((a->crop_y0 + a->crop_y1) / 2.0);
      for (int y = 0; y < h; y++) {
        float dy = fabsf((float)y - my);
        if (dy > cb) {
          memset(plane + (size_t)y * w, 0, w);
          continue;
        }
        float xmargin = (float)sqrt((double)(ca * ca) * (1.0 - (double)((dy * dy) / (cb * cb))));
        for (int x = 0; x < w; x++) {
          float dx = fabsf((float)x - mx);
          if (dx > xmargin) plane[(size_t)y * w + x] = 0;
        }
      }
    } else {  // rectangular crop, environment.h:773-790
      for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++)
          if (x < a->crop_x0 || x >= a->crop_x1 || y < a->crop_y0 || y >= a->crop_y1) plane[(size_t)y * w + x] = 0;
    }
  }
}
