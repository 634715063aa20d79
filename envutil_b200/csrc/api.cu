// api.cu - device side of the C ABI (include/envutil_b200.h): source staging, the asset cache,
// plan building and kernel launches. This is what a `cuda_dispatch : dispatch_base` adapter
// calls in place of the reference's payload() body (envutil_payload.cc:2408-2436):
//   eu_source_upload  ~ environment<...>(fct) construction     (envutil_payload.cc:1934,
//                       environment.h:594-950 / cubemap.h:548-946,1147-1233)
//   eu_render         ~ fuse<>() + work()                       (envutil_payload.cc:1885-2284,425-579)
//   eu_cycle          ~ conclude_cycle()                        (environment.h:224, envutil_payload.cc:2433)
// There is no CPU fallback: every entry point fails with EU_ERR_NO_DEVICE / EU_ERR_CUDA when
// the GPU is not usable.
#include <cuda_runtime.h>

#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "bspline_consts.h"
#include "envutil_b200.h"
#include "host_setup.h"
#include "kernels.h"
#include "plan.h"

struct eu_source {
  std::string key;
  int kind, projection, nch, degree;
  int w, h;              // core
  int cw, chh;           // container
  int lx, ly, rx, ry;    // brace
  int bc0, bc1;
  float* container;      // device: chh rows of `pitch` floats, cw texels of `tstride` floats each
  int tstride;           // nch, or 4 for the padded RGB layout
  int pitch;             // floats per container row, a multiple of 4 (16-byte row granules for bulk copies)
  eu_cubemap_metrics_t cm;
  long last_used_cycle;
  int refs;
  bool foreign_use;  // rendered from on a caller's stream: release has to synchronise the device
  bool reserved = false;  // container handed out by eu_source_reserve (eu_source_write_rect / eu_source_commit)
};

namespace {

#define EU_TABLE_SLOTS 8
#define EU_PLANAR_ENTRIES 6
#define EU_SLOT_FACETS_BYTES (sizeof(FacetDev) * EU_MAX_FACETS)
#define EU_SLOT_TAPS_BYTES (sizeof(float) * 3 * EU_MAX_TAPS)
#define EU_SLOT_BYTES (EU_SLOT_FACETS_BYTES + EU_SLOT_TAPS_BYTES + sizeof(float) * (EU_INV_NK + 8))

struct Context {
  bool up = false;
  int device = -1;
  cudaStream_t stream = nullptr;       // staging + render
  cudaStream_t up_stream = nullptr;    // H2D of asynchronous uploads
  cudaStream_t down_stream = nullptr;  // D2H of asynchronous jobs
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  struct JobSlot {
    float* d_out = nullptr;
    size_t cap = 0;
    cudaEvent_t start = nullptr, rendered = nullptr, done = nullptr;
    bool pending = false;
    int launches = 0;
  } jobs[EU_MAX_JOBS_IN_FLIGHT];
  int next_job = 0;
  std::vector<eu_source*> sources;
  std::map<std::string, eu_source*> by_key;
  long cycle = 0;
  // Per-job tables. A plan's facet array, tap list and inverse-lens spline live in ONE slot of a small ring;
  // the slot is rewritten on the stream the next plan renders on only after that stream has waited for the
  // event recorded behind the slot's last consumer, so jobs on different streams (eu_render_rows on a
  // caller's stream, eu_render_async on the library's) never see each other's tables.
  struct TableSlot {
    unsigned char* mem = nullptr;  // FacetDev[EU_MAX_FACETS] | taps[3 * EU_MAX_TAPS] | invcoef[EU_INV_NK + 8]
    cudaEvent_t used = nullptr;
    bool in_use = false;
  } slots[EU_TABLE_SLOTS];
  int next_slot = 0;
  // stepper tables (k_planar_tables) of the most recent targets: pipelines that alternate between targets
  // (the two stages of BASELINE configs[4]) find theirs again instead of draining the device
  struct PlanarEntry {
    float2* buf = nullptr;  // column terms [2][W], row terms [2][H], then the bare planar coordinates
    size_t cap = 0;
    TargetDev key;
    bool valid = false;
    long stamp = 0;
    cudaEvent_t used = nullptr;
    bool in_use = false;
  } planar[EU_PLANAR_ENTRIES];
  long planar_clock = 0;
  float* d_out = nullptr;
  size_t out_cap = 0;
  int32_t* d_index = nullptr;
  size_t index_cap = 0;
  float* d_screen_lut = nullptr;  // eu_screen_lut's table, uploaded on first use
  // eu_source_write_rect into 16-byte-texel containers: the rectangle lands in one of these scratch buffers (H2D on a
  // copy stream of its own) and a kernel on the caller's stream widens it into the container. With a ring of them
  // the copies of consecutive rectangles follow each other without waiting for the kernels in between.
  struct RectSlot {
    float* buf = nullptr;
    size_t cap = 0;
    cudaEvent_t copied = nullptr, freed = nullptr;
    bool used = false;
  } rect_ring[3];
  int next_rect = 0;
  cudaStream_t rect_stream = nullptr;
};
Context g;
thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CK(call)                                                                                    \
  do {                                                                                              \
    cudaError_t e_ = (call);                                                                        \
    if (e_ != cudaSuccess)                                                                          \
      return fail(EU_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

int need_up() {
  if (!g.up) return fail(EU_ERR_STATE, "eu_init has not been called");
  return EU_OK;
}

// Large device buffers (containers, upload staging) come from CUDA's stream-ordered pool with
// the release threshold lifted: a freed container is handed to the next upload instead of going
// back to the driver (cudaMalloc/cudaFree of a few hundred MB cost milliseconds and synchronise).
cudaError_t pool_alloc(float** p, size_t n_floats, cudaStream_t st = nullptr) {
  return cudaMallocAsync((void**)p, n_floats * sizeof(float), st ? st : g.stream);
}
void pool_free(void* p) {
  if (p) cudaFreeAsync(p, g.stream);
}

template <typename T>
int grow(T*& p, size_t& cap, size_t need) {
  if (need <= cap) return EU_OK;
  if (p) CK(cudaFree(p));
  p = nullptr;
  cap = 0;
  CK(cudaMalloc(&p, need * sizeof(T)));
  cap = need;
  return EU_OK;
}

// get_left_brace_size / get_right_brace_size, zimt/bspline.h:305-372
int left_brace(int degree, int bc) {
  int b = degree / 2;
  if (bc == EU_BC_REFLECT) b++;
  else if (degree & 1) b++;
  if (bc == EU_BC_PERIODIC && !(degree & 1)) b++;
  return b;
}
int right_brace(int degree, int bc) {
  int b = degree / 2;
  if (bc == EU_BC_REFLECT && !(degree & 1)) b++;
  if (degree & 1) b++;
  if (bc == EU_BC_PERIODIC) b++;
  return b;
}

// iir_filter ctor (zimt/recursive.h:774-862), overall gain (:93-103); everything is held in
// long double and narrowed at use (:650-662). M: line length (decides which init formula runs).
void iir_setup(IirDev& f, int bc, int degree, long double tolerance, int M) {
  memset(&f, 0, sizeof(f));
  f.bc = bc;
  f.npoles = degree / 2;
  long double lambda = 1.0L;
  for (int k = 0; k < f.npoles; k++) {
    long double p = eu_bspline_poles[degree][k];
    f.pole[k] = (float)p;
    f.horizon[k] = tolerance > 0 ? (int)ceill(logl(tolerance) / logl(fabsl(p))) : INT_MAX;
    lambda *= (1.0L - p) * (1.0L - 1.0L / p);
    long double e = (bc == EU_BC_REFLECT) ? (long double)(2 * M) : (long double)(M - 1);
    f.pole_pow[k] = (float)powl(p, e);
  }
  f.gain = (float)lambda;
}

bool is_full_sphere(const eu_facet_t* f) {  // environment.h:905-907
  return f->projection == EU_SPHERICAL && fabs(f->hfov - 2.0 * M_PI) < .000001 && f->width == 2 * f->height;
}

void source_dev(const eu_source* s, SourceDev& d) {
  d.tstride = s->tstride;
  d.stride = s->pitch;
  d.core = s->container + (size_t)s->ly * d.stride + (size_t)s->lx * s->tstride;
  d.nch = s->nch;
  d.w = s->w;
  d.h = s->h;
  d.bc0 = s->w == 1 ? EU_BC_CONST0 : s->bc0;  // zimt gates an axis of extent 1 as CONSTANT (zimt/eval.h:2060-2068)
  d.bc1 = s->h == 1 ? EU_BC_CONST0 : s->bc1;
  // limits of the safe evaluator's gates: -0.5 .. N-0.5 (zimt/bspline.h:233-286)
  d.upper_x = (float)((long double)(s->w - 1) + 0.5L);
  d.upper_y = (float)((long double)(s->h - 1) + 0.5L);
}

void free_source(eu_source* s) {
  if (!s) return;
  pool_free(s->container);
  if (!s->key.empty()) {
    auto it = g.by_key.find(s->key);
    if (it != g.by_key.end() && it->second == s) g.by_key.erase(it);
  }
  for (size_t i = 0; i < g.sources.size(); i++)
    if (g.sources[i] == s) {
      g.sources.erase(g.sources.begin() + i);
      break;
    }
  delete s;
}

bool known_source(eu_source_h s) {
  for (auto p : g.sources)
    if (p == s) return true;
  return false;
}

// dev_div_const (eu_device.cuh): is  q = x * rcp;  r = fma(-q, y, x);  fma(r, rcp, q)  with rcp = RN(1 / y) the correctly
// rounded x / y for EVERY x? Checked by exhaustion over the 2^23 significands of one binade of x (the three
// operations and the division all scale exactly with the exponent of x, away from the subnormal and overflow
// ranges; the sign is symmetric). About 10 ms per distinct divisor, remembered.
bool exact_by_reciprocal(float y) {
  static std::map<uint32_t, bool> known;
  uint32_t key;
  memcpy(&key, &y, 4);
  auto it = known.find(key);
  if (it != known.end()) return it->second;
  bool ok = std::isfinite(y) && y != 0.0f && fabsf(y) > 1e-18f && fabsf(y) < 1e18f;
  if (ok) {
    const float rcp = 1.0f / y;
    for (uint32_t m = 0; m < (1u << 23) && ok; m++) {
      const uint32_t bits = 0x3f800000u | m;  // 1.0 <= x < 2.0
      float x;
      memcpy(&x, &bits, 4);
      const float q = x * rcp;
      const float r = fmaf(-q, y, x);
      ok = fmaf(r, rcp, q) == x / y;
    }
  }
  known[key] = ok;
  return ok;
}

// facet -> FacetDev: what source_t / mount_t / cubemap_view_t / environment hold
// (environment.h:594-645,970-1006,1428-1461,1786-1860), narrowed as the functors narrow it
int facet_dev(const eu_target_t* t, const eu_facet_t* f, const eu_source* s, FacetDev& F, const eu_facet_t* ft = nullptr) {
  memset(&F, 0, sizeof(F));
  if (s->projection != f->projection || s->nch != f->nchannels ||
      (s->kind == EU_SRC_MOUNT && (s->w != f->window_width || s->h != f->window_height)))
    return fail(EU_ERR_ARGUMENT, "facet description does not match its staged source");
  source_dev(s, F.src);
  F.kind = s->kind;
  F.projection = f->projection;
  double m[9];
  eu_facet_basis(t, f, m);
  for (int i = 0; i < 3; i++) {
    F.xx[i] = (float)m[i];
    F.yy[i] = (float)m[3 + i];
    F.zz[i] = (float)m[6 + i];
  }
  F.ext_x0 = f->x0;
  F.ext_y0 = f->y0;
  F.ext_w = (float)(f->x1 - f->x0);
  F.ext_h = (float)(f->y1 - f->y0);
  F.total_w = (float)f->width;
  F.total_h = (float)f->height;
  F.rcp_w = 1.0f / F.ext_w;
  F.rcp_h = 1.0f / F.ext_h;
  F.fast_div = 0;
  if (s->kind == EU_SRC_MOUNT) F.fast_div = (exact_by_reciprocal(F.ext_w) ? 1 : 0) | (exact_by_reciprocal(F.ext_h) ? 2 : 0);
  {  // window extent, environment.h:607-618: x0 + (offset / total_width) * (x1 - x0) etc., in double,
     // narrowed for the float compares of test_crd (:970-978). BOTH axes use widths (sic).
    double wx = f->x1 - f->x0, wy = f->y1 - f->y0;
    double px0 = (double)f->window_x_offset / f->width, py0 = (double)f->window_y_offset / f->width;
    double px1 = (double)(f->window_x_offset + f->window_width) / f->width;
    double py1 = (double)(f->window_y_offset + f->window_width) / f->width;
    F.win_x0 = (float)(f->x0 + px0 * wx);
    F.win_y0 = (float)(f->y0 + py0 * wy);
    F.win_x1 = (float)(f->x0 + px1 * wx);
    F.win_y1 = (float)(f->y0 + py1 * wy);
    F.win_xoff = (float)f->window_x_offset;
    F.win_yoff = (float)f->window_y_offset;
    F.win_margin[0] = 1e-5f * fmaxf(fabsf(F.win_x0), fabsf(F.win_x1)) + 1e-30f;
    F.win_margin[1] = 1e-5f * fmaxf(fabsf(F.win_y0), fabsf(F.win_y1)) + 1e-30f;
  }
  F.mask_always = (s->kind != EU_SRC_MOUNT) || (f->projection == EU_FISHEYE && f->hfov >= M_PI * 2.0);
  F.has_lcp = f->has_lcp;
  F.has_shift = f->has_shift;
  F.has_shear = f->has_shear;
  float a = (float)f->a, b = (float)f->b, c = (float)f->c;
  F.lcp[0] = a;
  F.lcp[1] = b;
  F.lcp[2] = c;
  F.lcp[3] = 1.0f - (a + b + c);
  F.lcp_s = (float)f->s;
  F.shift_h = (float)f->shift_h;
  F.shift_v = (float)f->shift_v;
  F.shear_g = f->shear_g;
  F.shear_t = f->shear_t;
  if (s->kind != EU_SRC_MOUNT) {
    F.refc_md = (float)s->cm.refc_md;
    F.model_to_px = (float)s->cm.model_to_px;
    F.section_px = s->cm.section_px;
  }
  F.recip_step = (float)(1.0 / f->step);
  F.brighten = (float)(f->brighten == 0.0 ? 1.0 : f->brighten);
  F.hdr_optimum = 0.0f;
  F.hdr_kind = EU_HDR_MIDDLE;
  F.masked = f->masked != 0 ? 1 : 0;
  F.paint = f->masked == 2 ? 1.0f : 0.0f;
  // generic_r3, envutil_payload.cc:1636-1809: a facet with translation - or every facet, when the job is a
  // 'single' on a facet with lens correction / translation (`ft`) - gets its rays from the generic stepper
  F.generic = (f->has_translation || ft) ? 1 : 0;
  F.g_nstage = 0;
  if (F.generic) {
    // the matrices are r3_t<float>, built from make_r3_t's double rows narrowed element by element
    auto rot = [](double r, double p, double y, int inv, float out[9]) {
      double d[9];
      eu_rotation_matrix(r, p, y, inv, d);
      for (int i = 0; i < 9; i++) out[i] = (float)d[i];
    };
    auto mul = [](const float a[9], const float b[9], float o[9]) {  // rotate(r3_t, r3_t), geometry.h:84-91
      for (int i = 0; i < 3; i++)
        for (int c = 0; c < 3; c++) o[3 * i + c] = (a[3 * i] * b[c] + a[3 * i + 1] * b[3 + c]) + a[3 * i + 2] * b[6 + c];
    };
    auto plane_shift = [](const eu_facet_t* q, const float tp[9], float sh[3]) {  // :1683-1695,1709-1716
      sh[0] = (float)q->tr_x; sh[1] = (float)q->tr_y; sh[2] = (float)q->tr_z;
      if (q->tp_y != 0 || q->tp_p != 0 || q->tp_r != 0) {  // rotate(xel_t<double,3>(shift), r): in double
        double sd[3] = {sh[0], sh[1], sh[2]}, od[3];
        for (int c = 0; c < 3; c++) od[c] = (sd[0] * tp[c] + sd[1] * tp[3 + c]) + sd[2] * tp[6 + c];
        for (int c = 0; c < 3; c++) sh[c] = (float)od[c];
      }
    };
    auto stage = [&](int k, const float a[9], const float b[9], const float shift[3], float dcp) {  // tf3d_t ctor
      FacetDev::TfStage& S = F.g_st[k];
      memcpy(S.a, a, sizeof(S.a));
      memcpy(S.b, b, sizeof(S.b));
      mul(a, b, S.ab);
      for (int c = 0; c < 3; c++) S.shift[c] = shift[c];
      S.has_shift = (shift[0] != 0 || shift[1] != 0 || shift[2] != 0) ? 1 : 0;
      S.dcp = dcp;
    };
    float r_camera[9], rt_tp[9], rt_tpi[9], rs_tp[9], rs_tpi[9], r_facet[9], m1[9], m2[9];
    if (ft) rot(ft->roll, ft->pitch, ft->yaw, 0, r_camera);
    else rot(t->roll, t->pitch, t->yaw, 0, r_camera);
    rot(f->tp_r, f->tp_p, f->tp_y, 1, rs_tp);
    rot(f->tp_r, f->tp_p, f->tp_y, 0, rs_tpi);
    rot(f->roll, f->pitch, f->yaw, 1, r_facet);
    const bool have_ttp = ft && (ft->tr_x != 0 || ft->tr_y != 0 || ft->tr_z != 0);
    const bool have_stp = (f->tr_x != 0 || f->tr_y != 0 || f->tr_z != 0);
    float shift_t[3] = {0.f, 0.f, 0.f}, shift_s[3], dcp = 1.0f;
    if (ft) {
      rot(ft->tp_r, ft->tp_p, ft->tp_y, 1, rt_tp);
      rot(ft->tp_r, ft->tp_p, ft->tp_y, 0, rt_tpi);
      plane_shift(ft, rt_tp, shift_t);
      dcp = (float)(1.0 - shift_t[2]);
      for (int c = 0; c < 3; c++) shift_t[c] = -shift_t[c];
    }
    plane_shift(f, rs_tp, shift_s);
    if (have_ttp) {
      mul(r_camera, rt_tp, m1);
      if (have_stp) {
        stage(0, m1, rt_tpi, shift_t, dcp);
        mul(rs_tpi, r_facet, m2);
        stage(1, rs_tp, m2, shift_s, 1.0f);
        F.g_nstage = 2;
      } else {
        mul(rt_tpi, r_facet, m2);
        stage(0, m1, m2, shift_t, dcp);
        F.g_nstage = 1;
      }
    } else if (have_stp) {
      mul(r_camera, rs_tp, m1);
      mul(rs_tpi, r_facet, m2);
      stage(0, m1, m2, shift_s, 1.0f);
      F.g_nstage = 1;
    } else {  // rotate_t(rotate(r_camera, r_facet))
      const float zero[3] = {0.f, 0.f, 0.f};
      stage(0, r_camera, r_facet, zero, 1.0f);
      F.g_nstage = 1;
    }
  }
  return EU_OK;
}

// stepper_base ctor, stepper.h:294-306 (extents narrowed to float first, factors in double)
// the raster a job produces: the target, or its crop (eu_target_t::crop_*)
int out_width(const eu_target_t* t) { return t->crop_width > 0 ? t->crop_width : t->width; }
int out_height(const eu_target_t* t) { return t->crop_width > 0 ? t->crop_height : t->height; }

void target_dev(const eu_target_t* t, bool normalize, TargetDev& T) {
  memset(&T, 0, sizeof(T));
  int w = t->width, h = t->height;
  float a0 = (float)t->x0, a1 = (float)t->x1, b0 = (float)t->y0, b1 = (float)t->y1;
  T.projection = t->projection;
  T.width = out_width(t);
  T.height = out_height(t);
  T.full_w = w;
  T.full_h = h;
  T.off_x = t->crop_width > 0 ? t->crop_x0 : 0;
  T.off_y = t->crop_width > 0 ? t->crop_y0 : 0;
  T.normalize = normalize;
  T.fx1 = (float)(a1 / (2.0 * w));
  T.fx0 = (float)(a0 / (2.0 * w));
  T.fy1 = (float)(b1 / (2.0 * h));
  T.fy0 = (float)(b0 / (2.0 * h));
  T.bias_x = .25f * (a1 - a0) / (float)w;
  T.bias_y = .25f * (b1 - b0) / (float)h;
  T.delta = (float)EU_LANES * (a1 - a0) / (float)w;
  T.section_md = a1 - a0;
  T.refc_md = (float)((a1 - a0) / 2.0);
  T.unbrighten = (t->gain == 0.0) ? 1.0f : (float)t->gain;
}

// the facet a single-facet job renders (solo), else facet 0
int first_of(int nf, const eu_opts_t* o) { return (nf > 1 && o->solo >= 0 && o->solo < nf) ? o->solo : 0; }

struct Plan {
  RenderParams P;
  int launches;
  int slot = -1;    // Context::slots entry holding this plan's tables (-1: none needed)
  int planar = -1;  // Context::planar entry holding its stepper tables
};

// after the plan's last kernel has been enqueued on `st`: later plans may reuse its tables once this point
// of `st` has been reached
cudaError_t plan_done(const Plan& plan, cudaStream_t st) {
  if (plan.slot >= 0) {
    cudaError_t e = cudaEventRecord(g.slots[plan.slot].used, st);
    if (e != cudaSuccess) return e;
    g.slots[plan.slot].in_use = true;
  }
  if (plan.planar >= 0) {
    cudaError_t e = cudaEventRecord(g.planar[plan.planar].used, st);
    if (e != cudaSuccess) return e;
    g.planar[plan.planar].in_use = true;
  }
  return cudaSuccess;
}

// the plan's table slot, ready to be written on `cs`
int take_slot(Plan& plan, cudaStream_t cs, unsigned char** mem) {
  if (plan.slot < 0) {
    plan.slot = g.next_slot;
    g.next_slot = (g.next_slot + 1) % EU_TABLE_SLOTS;
    Context::TableSlot& S = g.slots[plan.slot];
    if (!S.mem) {
      CK(cudaMalloc(&S.mem, EU_SLOT_BYTES));
      CK(cudaEventCreateWithFlags(&S.used, cudaEventDisableTiming));
    }
    if (S.in_use) CK(cudaStreamWaitEvent(cs, S.used, 0));
  }
  *mem = g.slots[plan.slot].mem;
  return EU_OK;
}

// pto_planar<float, L, true>: the inverse planar transformation of a 'single' job's target facet. The knots
// of inverse_lcp's spline (lens_correction.h:341-386) are found on the host - Newton's method in double,
// eu_polynomial<double, 4>::inverse (:133-170) - and become b-spline coefficients on the device with the
// very kernels that prefilter rasters (one line of EU_INV_NK floats, NATURAL), then the NATURAL brace.
int inverse_planar(const eu_facet_t* ft, cudaStream_t cs, float* d_invcoef, InvPlanarDev& IP) {
  memset(&IP, 0, sizeof(IP));
  IP.on = 1;
  IP.has_shear = ft->has_shear;
  IP.has_shift = ft->has_shift;
  IP.has_lcp = ft->has_lcp;
  IP.shear_g = ft->shear_g;
  IP.shear_t = ft->shear_t;
  IP.s = ft->s;
  IP.h = (float)ft->shift_h;
  IP.v = (float)ft->shift_v;
  for (int i = 0; i < 16; i++) IP.wm[i] = (float)eu_bspline_weights[3][i / 4][i % 4];
  const int sz = EU_INV_SZ, nk = EU_INV_NK;
  const double cf[5] = {ft->a, ft->b, ft->c, 1.0 - (ft->a + ft->b + ft->c), 0.0};
  double dcf[5];
  {
    size_t power = 4;
    for (int i = 0; i <= 4; i++) { dcf[i] = cf[i] * power; --power; }
  }
  auto fn = [&](double x) {
    double sum = 0.0, power = 1.0;
    for (int i = 0; i <= 4; i++) { sum += cf[4 - i] * power; power *= x; }
    return sum;
  };
  auto der = [&](double x) {
    double sum = 0.0, power = 1.0;
    for (int i = 0; i < 4; i++) { sum += dcf[4 - i - 1] * power; power *= x; }
    return sum;
  };
  const double r_max = ft->r_max * ((sz + 3.0) / sz);
  IP.rr_max = fn(r_max);
  float knots[EU_INV_NK];
  for (int i = 0; i < nk; i++) {
    double notch = (double)i / (nk - 1);
    notch *= notch;
    notch *= IP.rr_max;
    double current = i * r_max / sz, result, difference = 0.0, last_difference = DBL_MAX;
    const double tolerance = 100 * DBL_EPSILON;
    for (int count = 0; count < 16; count++) {
      result = fn(current);
      difference = notch - result;
      if (last_difference == difference) break;
      if (fabs(difference) <= tolerance) break;
      last_difference = difference;
      current = current + difference / der(current);
    }
    if (!(fabs(difference) < tolerance))  // the reference asserts here (lens_correction.h:186-190)
      return fail(EU_ERR_ARGUMENT, "the lens polynomial a=%g b=%g c=%g cannot be inverted", ft->a, ft->b, ft->c);
    knots[i] = (float)(notch == 0.0 ? 1.0 / der(0.0) : (current / notch) - 1);
  }
  float* core = d_invcoef + 4;  // 16-byte aligned start of the line; the brace uses core[-2..-1] and core[nk..nk+1]
  CK(cudaMemcpyAsync(core, knots, sizeof(knots), cudaMemcpyHostToDevice, cs));  // pageable: staged before return
  IirDev f;
  iir_setup(f, EU_BC_NATURAL, 3, (long double)FLT_EPSILON, nk);
  CK(eu_launch_iir_x(core, EU_INV_NK + 4, 1, nk, 1, f, cs));
  CK(eu_launch_brace_natural_1d(core, nk, 2, cs));
  IP.coef = core;
  return EU_OK;
}

// builds the plan and uploads the per-job tables; everything is enqueued on g.stream
// `cs`: the stream the render will be launched on. The per-job facet array and tap list live in
// one device buffer each; they are rewritten ON THAT STREAM, i.e. after every render enqueued
// before (another job's kernel may still be reading them) and before this job's.
int build_plan(const eu_target_t* t, const eu_opts_t* o, int nf, const eu_facet_t* facets,
               const eu_source_h* sources, const eu_tap_t* taps, int n_taps, cudaStream_t cs, Plan& plan) {
  if (!t || !o || !facets || !sources) return fail(EU_ERR_ARGUMENT, "null argument");
  if (nf < 1 || nf > EU_MAX_FACETS) return fail(EU_ERR_ARGUMENT, "facet count %d out of range 1..%d", nf, EU_MAX_FACETS);
  if (n_taps < 0 || n_taps > EU_MAX_TAPS) return fail(EU_ERR_ARGUMENT, "tap count %d out of range", n_taps);
  if (n_taps > 0 && !taps) return fail(EU_ERR_ARGUMENT, "taps missing");
  if (t->width <= 0 || t->height <= 0) return fail(EU_ERR_ARGUMENT, "target not prepared");
  if (t->crop_width > 0) {
    // unreachable through the reference's surface: a crop comes from a PTO p-line, whose projection codes
    // (envutil_main.cc:590-611) do not include cubemaps
    if (t->projection == EU_CUBEMAP || t->projection == EU_BIATAN6)
      return fail(EU_ERR_UNSUPPORTED, "cropped output of a cubemap target");
    if (t->crop_height <= 0 || t->crop_x0 < 0 || t->crop_y0 < 0 || t->crop_x0 + t->crop_width > t->width ||
        t->crop_y0 + t->crop_height > t->height)
      return fail(EU_ERR_ARGUMENT, "crop %dx%d+%d+%d does not lie inside the %dx%d target", t->crop_width, t->crop_height,
                  t->crop_x0, t->crop_y0, t->width, t->height);
  }
  if (o->spline_degree < 0 || o->spline_degree > EU_MAX_DEGREE)
    return fail(EU_ERR_ARGUMENT, "spline degree %d out of range 0..%d", o->spline_degree, EU_MAX_DEGREE);
  int nch = t->nchannels;
  if (nch < 1 || nch > 4) return fail(EU_ERR_ARGUMENT, "%d-channel target", nch);
  for (int i = 0; i < nf; i++) {
    if (!known_source(sources[i])) return fail(EU_ERR_ARGUMENT, "source %d is not a live handle", i);
    if (sources[i]->degree != o->spline_degree)
      return fail(EU_ERR_ARGUMENT, "source %d was staged for degree %d, job asks for %d", i, sources[i]->degree,
                  o->spline_degree);
    sources[i]->last_used_cycle = g.cycle;
  }
  RenderParams& P = plan.P;
  memset(&P, 0, sizeof(P));
  plan.launches = 0;
  int mode = EU_MODE_SINGLE, first = 0;
  if (nf > 1 && o->solo < 0) mode = (o->synopsis == EU_SYN_HDR_MERGE) ? EU_MODE_HDR : EU_MODE_VORONOI;
  if (nf > 1 && o->solo >= 0) {
    if (o->solo >= nf) return fail(EU_ERR_ARGUMENT, "solo %d >= facet count %d", o->solo, nf);
    first = o->solo;
  }
  // roll_out: 2/4-channel panoramas composite with alpha (voronoi_syn_plus, envutil_payload.cc:2306-2311)
  if (mode == EU_MODE_VORONOI && (nch == 2 || nch == 4)) mode = EU_MODE_VORONOI_PLUS;
  // normalize: envutil_payload.cc:2105,2118 (false for one facet without twining), else true
  target_dev(t, !(mode == EU_MODE_SINGLE && n_taps == 0), P.trg);
  P.mode = mode;
  P.degree = o->spline_degree;
  P.n_taps = n_taps;
  P.nch = nch;
  P.tstride = sources[first_of(nf, o)]->tstride;
  // a 'single' job on a facet with lens correction / shift / shear / translation (fuse(), :2058-2068)
  const eu_facet_t* ft = nullptr;
  if (t->single > 0) {
    if (t->single > nf) return fail(EU_ERR_ARGUMENT, "target.single %d is beyond the facet count %d", t->single, nf);
    const eu_facet_t* cand = &facets[t->single - 1];
    if (cand->has_2d_tf || cand->has_translation) ft = cand;
  }
  std::vector<FacetDev> F(nf);
  for (int i = 0; i < nf; i++) {
    int rc = facet_dev(t, &facets[i], sources[i], F[i], ft);
    if (rc) return rc;
  }
  if (ft && ft->has_2d_tf) {  // tf22 = pto_planar<float, L, true>(ft), envutil_payload.cc:1864
    unsigned char* mem = nullptr;
    int rc = take_slot(plan, cs, &mem);
    if (rc) return rc;
    rc = inverse_planar(ft, cs, reinterpret_cast<float*>(mem + EU_SLOT_FACETS_BYTES + EU_SLOT_TAPS_BYTES), P.inv);
    if (rc) return rc;
    plan.launches += 2;
  }
  for (int i = 1; i < nf; i++) {  // same_geometry: FacetDev::same_geom
    FacetDev a = F[i - 1], b = F[i];
    a.src.core = b.src.core = nullptr;
    a.brighten = b.brighten = 0.0f;
    a.hdr_optimum = b.hdr_optimum = 0.0f;
    a.hdr_kind = b.hdr_kind = 0;
    a.same_geom = b.same_geom = 0;
    F[i].same_geom = memcmp(&a, &b, sizeof(FacetDev)) == 0 ? 1 : 0;
  }
  if (mode == EU_MODE_HDR) {  // _hdr_merge_syn ctor, envutil_payload.cc:1354-1375
    float lowest = 100000.0f, highest = -1.0f;
    int lo = -1, hi = -1;
    for (int i = 0; i < nf; i++) {
      double br = (float)(facets[i].brighten == 0.0 ? 1.0 : facets[i].brighten);
      F[i].hdr_optimum = (float)(0.5f * br);
      if (br < lowest) { lowest = (float)br; lo = i; }
      if (br > highest) { highest = (float)br; hi = i; }
    }
    for (int i = 0; i < nf; i++) F[i].hdr_kind = (i == lo) ? EU_HDR_LOW : (i == hi) ? EU_HDR_HIGH : EU_HDR_MIDDLE;
  }
  P.f0 = F[first];
  {  // RenderParams::cube_tab: the face vectors of the cubemap / biatan6 steppers for f0, signs folded in
    const float *xx = P.f0.xx, *yy = P.f0.yy, *zz = P.f0.zz;
    const float* rows[6][3] = {{xx, yy, zz}, {xx, yy, zz}, {yy, zz, xx}, {yy, zz, xx}, {zz, yy, xx}, {zz, yy, xx}};
    // CM_LEFT -xx + p1 yy, zz | RIGHT xx + p1 yy, -zz | TOP -yy - p1 zz, -xx | BOTTOM yy + p1 zz, -xx |
    // FRONT p1 yy + zz, xx | BACK p1 yy - zz, -xx      (stepper.h:1304-1331)
    const float sign[6][3] = {{-1, 1, 1}, {1, 1, -1}, {-1, -1, -1}, {1, 1, -1}, {1, 1, 1}, {-1, 1, -1}};
    for (int f = 0; f < 6; f++)
      for (int k = 0; k < 3; k++) {
        for (int i = 0; i < 3; i++) P.cube_tab[f][4 * k + i] = sign[f][k] < 0 ? -rows[f][k][i] : rows[f][k][i];
        P.cube_tab[f][4 * k + 3] = 0.0f;
      }
  }
  P.use_tiles = (o->reserved[1] & EU_OPT_NO_TILES) ? 0 : 1;
  P.src_cw = sources[first]->cw;
  P.src_ch = sources[first]->chh;
  P.src_lx = sources[first]->lx;
  P.src_ly = sources[first]->ly;
  P.src_base = sources[first]->container;
  P.n_facets = mode == EU_MODE_SINGLE ? 1 : nf;
  if (mode != EU_MODE_SINGLE) {
    unsigned char* mem = nullptr;
    int rc = take_slot(plan, cs, &mem);
    if (rc) return rc;
    // pageable source: the call returns once the data sit in the driver's staging buffer
    CK(cudaMemcpyAsync(mem, F.data(), sizeof(FacetDev) * nf, cudaMemcpyHostToDevice, cs));
    P.facets = reinterpret_cast<const FacetDev*>(mem);
  }
  if (n_taps > 0) {
    unsigned char* mem = nullptr;
    int rc = take_slot(plan, cs, &mem);
    if (rc) return rc;
    std::vector<float> tp(3 * (size_t)n_taps);
    for (int k = 0; k < n_taps; k++) {  // bias 4 = 1/0.25, twining.h:106-121
      tp[3 * k] = taps[k].x * 4.0f;
      tp[3 * k + 1] = taps[k].y * 4.0f;
      tp[3 * k + 2] = taps[k].w;
    }
    CK(cudaMemcpyAsync(mem + EU_SLOT_FACETS_BYTES, tp.data(), sizeof(float) * tp.size(), cudaMemcpyHostToDevice, cs));
    P.taps = reinterpret_cast<const float*>(mem + EU_SLOT_FACETS_BYTES);
    if (n_taps <= EU_INLINE_TAPS) {
      memcpy(P.ptaps, tp.data(), sizeof(float) * tp.size());
      P.taps_inline = 1;
    }
  }
  // stepper tables: recomputed only for a target that is not among the recent ones
  const int ow = out_width(t), oh = out_height(t);
  {
    const size_t need = 3 * (size_t)(ow + oh);  // float2 terms + the bare planar coordinate per entry, in float2 units
    TargetDev key = P.trg;
    key.normalize = 0;  // does not enter the tables
    int hit = -1, victim = 0;
    for (int i = 0; i < EU_PLANAR_ENTRIES; i++) {
      Context::PlanarEntry& E = g.planar[i];
      if (E.valid && memcmp(&E.key, &key, sizeof(TargetDev)) == 0) hit = i;
      if (!E.valid) { if (g.planar[victim].valid) victim = i; }
      else if (g.planar[victim].valid && E.stamp < g.planar[victim].stamp) victim = i;
    }
    if (hit < 0) {
      Context::PlanarEntry& E = g.planar[victim];
      if (!E.used) CK(cudaEventCreateWithFlags(&E.used, cudaEventDisableTiming));
      if (E.in_use) CK(cudaStreamWaitEvent(g.stream, E.used, 0));  // its last reader, on whatever stream that was
      E.valid = false;
      if (need > E.cap) {
        if (E.buf) CK(cudaFreeAsync(E.buf, g.stream));
        E.buf = nullptr;
        E.cap = 0;
        CK(cudaMallocAsync((void**)&E.buf, need * sizeof(float2), g.stream));
        E.cap = need;
      }
      CK(eu_launch_planar_tables(P.trg, E.buf, E.buf + 2 * (size_t)ow, reinterpret_cast<float*>(E.buf + 2 * (size_t)(ow + oh)),
                                 g.stream));
      plan.launches++;
      E.key = key;
      E.valid = true;
      hit = victim;
    }
    Context::PlanarEntry& E = g.planar[hit];
    E.stamp = ++g.planar_clock;
    plan.planar = hit;
    P.col_tab = E.buf;
    P.row_tab = E.buf + 2 * (size_t)ow;
    P.planar_raw = reinterpret_cast<const float*>(E.buf + 2 * (size_t)(ow + oh));
  }
  P.col0 = 0;
  P.col1 = ow;
  P.arith = (o->reserved[1] & EU_OPT_CONTRACTED) ? 1 : 0;
  // the specialised kernels assume every facet they touch has the job's channel count and one
  // texel stride; anything else (and translation) runs the general build
  P.any_generic = 0;
  for (int i = 0; i < nf; i++) {
    if (mode == EU_MODE_SINGLE && i != first) continue;
    if (F[i].generic || sources[i]->nch != nch || sources[i]->tstride != P.tstride) P.any_generic = 1;
    if (F[i].masked) {  // --mask_for: the general build paints (dev_facet_eval_general)
      P.any_generic = 1;
      if (sources[i]->nch != nch && nch > 2)
        return fail(EU_ERR_ARGUMENT, "a masked facet of %d channels in a job of %d: mono_t converts to one or two channels only "
                    "(environment.h:1338-1339)", sources[i]->nch, nch);
    }
  }
  // a compiled-in job shape (plan.h: eu_render_specs): the target and every facet that is
  // evaluated must agree with all values the entry fixes
  P.spec = 0;
  for (int sp = 1; sp < EU_N_SPECS && !P.spec; sp++) {
    const RenderSpec& rs = eu_render_specs[sp];
    bool fits = (rs.tproj < 0 || rs.tproj == P.trg.projection) && (rs.tnorm < 0 || rs.tnorm == P.trg.normalize) &&
                n_taps <= EU_INLINE_TAPS;  // their twining filter travels in the parameter block
    for (int i = 0; i < nf && fits; i++) {
      if (mode == EU_MODE_SINGLE && i != first) continue;
      const FacetDev& f = F[i];
      // the kernels compiled for a shape also assume 32-bit window offsets and proven reciprocal divisions
      const bool small = (long long)sources[i]->pitch * sources[i]->chh < (1ll << 31);
      fits = small && (f.kind != EU_SRC_MOUNT || f.fast_div == 3) && !f.has_lcp && !f.generic &&
             (rs.skind < 0 || rs.skind == f.kind) &&
             (rs.sproj < 0 || rs.sproj == f.projection) && (rs.bc0 < 0 || rs.bc0 == f.src.bc0) &&
             (rs.bc1 < 0 || rs.bc1 == f.src.bc1) && (rs.mask_always < 0 || rs.mask_always == f.mask_always);
    }
    if (fits) P.spec = sp;
  }
  if (o->reserved[1] & EU_OPT_NO_SHAPES) P.spec = 0;  // back-end option: general kernels only
  {
    int d = o->spline_degree;
    for (int row = 0; row <= d; row++)
      for (int k = 0; k <= d; k++) P.wmat[row * (d + 1) + k] = (float)eu_bspline_weights[d][row][k];
    // the shape kernels skip the cubic matrix's zero terms (dev_window_weights<4, true>): only if they ARE +0
    if (d == 3) {
      const int z[4] = {3, 5, 7, 11};
      for (int i : z) {
        uint32_t bits;
        memcpy(&bits, &P.wmat[i], 4);
        if (bits != 0u) P.spec = 0;
      }
    }
  }
  return EU_OK;
}

// source_t ctor, environment.h:594-950: shape, boundary conditions and brace of a mounted image's container
void mount_layout(const eu_facet_t* f, int degree, eu_source* s) {
  const int nch = f->nchannels;
  s->kind = EU_SRC_MOUNT;
  s->w = f->window_width;   // the raster handed in is the window ('W' clause, envutil_main.cc:754-786);
  s->h = f->window_height;  // the geometry refers to the total size
  s->bc0 = s->bc1 = EU_BC_REFLECT;
  if ((f->projection == EU_SPHERICAL || f->projection == EU_CYLINDRICAL) && fabs(f->hfov - 2.0 * M_PI) < .000001)
    s->bc0 = EU_BC_PERIODIC;
  s->lx = left_brace(degree, s->bc0);
  s->ly = left_brace(degree, s->bc1);
  s->rx = right_brace(degree, s->bc0);
  s->ry = right_brace(degree, s->bc1);
  s->cw = s->w + s->lx + s->rx;
  s->chh = s->h + s->ly + s->ry;
  s->pitch = (s->cw * nch + 3) & ~3;
}

// the core of a mounted image is in place: prefilter (degree > 1) and brace
int mount_finish(const eu_facet_t* f, int pdeg, eu_source* s, cudaStream_t st, int* launches) {
  // a reserved RGB source in the 16-byte layout is braced as 4-float texels (it is never prefiltered: that layout is
  // chosen for degree <= 1 only)
  const int nch = (s->tstride == 4 && f->nchannels == 3) ? 4 : f->nchannels, stride = s->pitch;
  if (nch != f->nchannels && pdeg > 1) return fail(EU_ERR_UNSUPPORTED, "16-byte RGB texels cannot be prefiltered in place");
  float* core = s->container + (size_t)s->ly * stride + (size_t)s->lx * nch;
  bool sphere = is_full_sphere(f);
  if (sphere && (s->ly > s->h || s->ry > s->h)) return fail(EU_ERR_ARGUMENT, "image too small for its brace");
  if (pdeg > 1) {
    IirDev fx, fy;
    if (sphere) {  // spherical_prefilter, environment.h:356-522 (tolerance 1e-4)
      iir_setup(fx, EU_BC_PERIODIC, pdeg, (long double)0.0001, s->w);
      iir_setup(fy, EU_BC_PERIODIC, pdeg, (long double)0.0001, 2 * s->h);
      CK(eu_launch_iir_x(core, stride, nch, s->w, s->h, fx, st));
      CK(eu_launch_iir_y_spherical(core, stride, nch, s->w, s->h, fy, st));
    } else {  // zimt::prefilter, prefilter.h:125-198
      iir_setup(fx, s->bc0, pdeg, (long double)FLT_EPSILON, s->w);
      iir_setup(fy, s->bc1, pdeg, (long double)FLT_EPSILON, s->h);
      CK(eu_launch_iir_x(core, stride, nch, s->w, s->h, fx, st));
      CK(eu_launch_iir_y(core, stride, nch, s->w, s->h, 1, fy, st));
    }
    *launches += 2;
  }
  CK(eu_launch_brace(core, stride, nch, s->w, s->h, s->lx, s->rx, s->ly, s->ry, s->bc0, s->bc1, sphere ? 1 : 0, st));
  *launches += 1;
  return EU_OK;
}

// `pixels` is a device pointer (kind = DeviceToDevice) or a host pointer (HostToDevice): the
// raster is copied straight into its place inside the container, there is no staging copy.
// `cpst`: the stream the placement copies run on (the staging stream itself, or the upload stream
// of an asynchronous upload; the two are ordered by events around the copies).
int stage_on_device(const eu_facet_t* f, const eu_opts_t* o, const float* d_pixels, cudaMemcpyKind kind,
                    cudaStream_t st, cudaStream_t cpst, eu_source* s, int* launches, float* copy_ms) {
  auto copies_end = [&]() -> cudaError_t {  // the staging kernels start after the copy
    cudaError_t e = cudaEventRecord(g.ev[3], cpst);
    if (e == cudaSuccess && cpst != st) e = cudaStreamWaitEvent(st, g.ev[3], 0);
    return e;
  };
  int degree = o->spline_degree;
  int pdeg = o->prefilter_degree < 0 ? degree : o->prefilter_degree;
  if (pdeg > EU_MAX_DEGREE) return fail(EU_ERR_ARGUMENT, "prefilter degree %d out of range", pdeg);
  int nch = f->nchannels;
  s->projection = f->projection;
  s->nch = nch;
  s->degree = degree;
  s->tstride = nch;
  *launches = 0;
  size_t tb = sizeof(float) * nch;
  if (f->projection == EU_CUBEMAP || f->projection == EU_BIATAN6) {
    // cubemap_t ctor + load, cubemap.h:548-580,1147-1233
    if (f->height != 6 * f->width) return fail(EU_ERR_ARGUMENT, "cubemap input must be 1:6 (got %dx%d)", f->width, f->height);
    s->kind = f->projection == EU_CUBEMAP ? EU_SRC_CUBEMAP : EU_SRC_BIATAN6;
    int rc = eu_compute_cubemap_metrics(f->width, f->hfov, o->support_min, o->tile_size, &s->cm);
    if (rc) return fail(rc, "bad cubemap metrics (face %d px, hfov %g, support %d, tile %d)", f->width, f->hfov,
                        o->support_min, o->tile_size);
    int S = s->cm.section_px, Fpx = f->width, L = s->cm.left_frame_px, R = s->cm.right_frame_px;
    // The spline windows of rays that hit a face edge reach degree/2 + 1 texels into the support frame. With
    // --support_min / --tile_size chosen so that there is less (e.g. 0 and 1: no frame at all) the reference
    // reads the unset brace of its b-spline object - undefined there (found by the randomised sweep: values
    // of 1e25) - and the kernels would read outside the container: refused.
    {
      const int need = degree / 2 + 1;
      // faces wider than 90 degrees carry support inside the image (inherent_support_px, cubemap.h:300-301)
      const int inherent = f->hfov > M_PI_2 ? (int)std::trunc(s->cm.model_to_px * (tan(f->hfov / 2.0) - 1.0)) : 0;
      if (L + inherent < need || R + inherent < need)
        return fail(EU_ERR_UNSUPPORTED, "cubemap support frame of %d/%d px is narrower than the %d px a degree-%d "
                    "spline needs (raise --support_min)", L, R, need, degree);
    }
    s->w = s->cw = S;
    s->h = s->chh = 6 * S;
    s->lx = s->ly = s->rx = s->ry = 0;
    s->bc0 = s->bc1 = EU_BC_REFLECT;
    s->pitch = (S * nch + 3) & ~3;
    const int pitch = s->pitch;
    size_t n = (size_t)pitch * 6 * S;
    // allocate, clear and fill on the copy stream: an asynchronous upload does not wait for the
    // staging stream (which may still be busy with the previous job), only the other way round
    CK(pool_alloc(&s->container, n, cpst));
    // even face widths: the placement copies write every face texel and the support fill every frame texel
    // before anything reads it, so the container need not be cleared (321 MB at C2's size). Odd widths: the
    // fill reads frame texels it has not written yet (stage.cu, k_cm_fill_ordered) - zero, as in the reference.
    if ((Fpx & 1) || L == 0 || R == 0) CK(cudaMemsetAsync(s->container, 0, n * sizeof(float), cpst));
    CK(cudaEventRecord(g.ev[2], cpst));  // start of the placement copies (timing of the blocking upload)
    if (kind == cudaMemcpyDeviceToDevice) {  // one kernel for the six faces
      CK(eu_launch_cubemap_place(d_pixels, s->container, pitch, nch, Fpx, S, L, cpst));
      ++*launches;
    } else {
      for (int face = 0; face < 6; face++)
        CK(cudaMemcpy2DAsync(s->container + (size_t)(face * S + L) * pitch + (size_t)L * nch, (size_t)pitch * sizeof(float),
                             d_pixels + (size_t)face * Fpx * Fpx * nch, (size_t)Fpx * tb, (size_t)Fpx * tb, Fpx, kind, cpst));
    }
    CK(copies_end());
    int nl = 0;
    CK(eu_launch_cubemap_support(s->container, pitch, nch, Fpx, S, L, R, s->cm.refc_md, s->cm.model_to_px, &nl, st));
    *launches += nl;
    if (pdeg > 1) {  // cubemap_t::prefilter, cubemap.h:921-946: per section, NATURAL, both axes
      IirDev fx;
      iir_setup(fx, EU_BC_NATURAL, pdeg, (long double)FLT_EPSILON, S);
      CK(eu_launch_iir_x(s->container, pitch, nch, S, 6 * S, fx, st));
      CK(eu_launch_iir_y(s->container, pitch, nch, S, S, 6, fx, st));
      *launches += 2;
    }
    return EU_OK;
  }
  mount_layout(f, degree, s);
  size_t n = (size_t)s->pitch * s->chh;
  CK(pool_alloc(&s->container, n, cpst));
  int stride = s->pitch;
  float* core = s->container + (size_t)s->ly * stride + (size_t)s->lx * nch;
  CK(cudaEventRecord(g.ev[2], cpst));
  CK(cudaMemcpy2DAsync(core, (size_t)stride * sizeof(float), d_pixels, (size_t)s->w * tb, (size_t)s->w * tb, s->h, kind,
                       cpst));
  CK(copies_end());
  return mount_finish(f, pdeg, s, st, launches);
}

// 16-byte texel layout for RGB sources: one LDG.128 per tap instead of three LDG.32. Measured on B200 with the
// kernels compiled for a job shape (profiles/r02c_configs_variants.jsonl): bilinear jobs gain 12-17 % (C3a 1.18 ->
// 1.02 ms, C3b 1.25 -> 1.03, C4 2.09 -> 1.84), the footprint-staged cubic kernel nothing - so it is the default for
// degree <= 1 and off for higher degrees. o->reserved[0]: 0 = that rule, 1 = always, 2 = never.
int maybe_pad(const eu_opts_t* o, eu_source* s, cudaStream_t st, int* launches) {
  if (s->nch != 3) return EU_OK;
  const bool want = o->reserved[0] == 1 || (o->reserved[0] == 0 && s->degree <= 1);
  if (!want) return EU_OK;
  size_t ntex = (size_t)s->cw * s->chh;
  float* padded = nullptr;
  CK(pool_alloc(&padded, ntex * 4));
  cudaError_t e = eu_launch_pad_texels(s->container, s->pitch, padded, s->cw, s->cw, s->chh, s->nch, st);
  if (e != cudaSuccess) {
    pool_free(padded);
    return fail(EU_ERR_CUDA, "pad kernel: %s", cudaGetErrorString(e));
  }
  pool_free(s->container);
  s->container = padded;
  s->tstride = 4;
  s->pitch = s->cw * 4;
  ++*launches;
  return EU_OK;
}

}  // namespace

extern "C" {

const char* eu_last_error(void) { return g_err; }

// both arithmetics are built in; eu_opts_t.reserved[1] bit 4 (EU_OPT_CONTRACTED) selects per job
int eu_render_arithmetic(void) { return 2; }

int eu_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int eu_init(int device_id) {
  if (g.up) {
    if (device_id == g.device) return EU_OK;
    return fail(EU_ERR_STATE, "already initialised on device %d", g.device);
  }
  int n = eu_device_count();
  if (n <= 0) return fail(EU_ERR_NO_DEVICE, "no CUDA device is visible");
  if (device_id < 0 || device_id >= n) return fail(EU_ERR_ARGUMENT, "device %d out of range 0..%d", device_id, n - 1);
  CK(cudaSetDevice(device_id));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device_id));
  // arch-specific ('a') code is not forward compatible: sm_100a SASS runs on compute capability 10.0 only
  if (prop.major != 10 || prop.minor != 0)
    return fail(EU_ERR_NO_DEVICE, "device %d is sm_%d%d; this library carries sm_100a code only", device_id, prop.major,
                prop.minor);
  CK(cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&g.up_stream, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&g.down_stream, cudaStreamNonBlocking));
  for (auto& j : g.jobs) {
    CK(cudaEventCreate(&j.start));
    CK(cudaEventCreate(&j.rendered));
    CK(cudaEventCreate(&j.done));
  }
  {
    cudaMemPool_t mp;
    CK(cudaDeviceGetDefaultMemPool(&mp, device_id));
    uint64_t keep = UINT64_MAX;
    CK(cudaMemPoolSetAttribute(mp, cudaMemPoolAttrReleaseThreshold, &keep));
  }
  for (auto& e : g.ev) CK(cudaEventCreate(&e));
  g.device = device_id;
  g.up = true;
  return EU_OK;
}

void eu_shutdown(void) {
  if (!g.up) return;
  cudaStreamSynchronize(g.stream);
  while (!g.sources.empty()) free_source(g.sources.back());
  for (auto& S : g.slots) {
    cudaFree(S.mem);
    if (S.used) cudaEventDestroy(S.used);
  }
  for (auto& E : g.planar) {
    if (E.buf) cudaFreeAsync(E.buf, g.stream);
    if (E.used) cudaEventDestroy(E.used);
  }
  cudaFree(g.d_out);
  cudaFree(g.d_index);
  cudaFree(g.d_screen_lut);
  g.d_screen_lut = nullptr;
  for (auto& R : g.rect_ring) {
    cudaFree(R.buf);
    if (R.copied) cudaEventDestroy(R.copied);
    if (R.freed) cudaEventDestroy(R.freed);
    R = Context::RectSlot();
  }
  if (g.rect_stream) cudaStreamDestroy(g.rect_stream);
  g.rect_stream = nullptr;
  g.next_rect = 0;
  cudaDeviceSynchronize();
  for (auto& e : g.ev) cudaEventDestroy(e);
  for (auto& j : g.jobs) {
    cudaFree(j.d_out);
    cudaEventDestroy(j.start);
    cudaEventDestroy(j.rendered);
    cudaEventDestroy(j.done);
  }
  cudaStreamDestroy(g.up_stream);
  cudaStreamDestroy(g.down_stream);
  {
    cudaMemPool_t mp;
    if (cudaDeviceGetDefaultMemPool(&mp, g.device) == cudaSuccess) cudaMemPoolTrimTo(mp, 0);
  }
  cudaStreamDestroy(g.stream);
  g = Context();
}

// what every entry point that stages a raster demands of its description (eu_source_upload*, eu_source_reserve)
static int check_raster(const eu_facet_t* f, const eu_opts_t* o) {
  if (f->width <= 0 || f->height <= 0 || f->nchannels < 1 || f->nchannels > 4)
    return fail(EU_ERR_ARGUMENT, "bad raster description %dx%dx%d", f->width, f->height, f->nchannels);
  if (o->spline_degree < 0 || o->spline_degree > EU_MAX_DEGREE)
    return fail(EU_ERR_ARGUMENT, "spline degree %d out of range", o->spline_degree);
  if (f->window_width <= 0 || f->window_height <= 0 || f->window_x_offset < 0 || f->window_y_offset < 0 ||
      f->window_x_offset + f->window_width > f->width || f->window_y_offset + f->window_height > f->height)
    return fail(EU_ERR_ARGUMENT, "facet window %dx%d+%d+%d does not lie inside %dx%d (run eu_facet_prepare)",
                f->window_width, f->window_height, f->window_x_offset, f->window_y_offset, f->width, f->height);
  if ((f->projection == EU_CUBEMAP || f->projection == EU_BIATAN6) &&
      (f->window_width != f->width || f->window_height != f->height))
    return fail(EU_ERR_ARGUMENT, "cubemaps cannot be windowed");
  return EU_OK;
}

static int upload_common(const char* asset_key, const eu_facet_t* f, const eu_opts_t* o, const float* pixels,
                         cudaMemcpyKind kind, cudaStream_t caller, eu_source_h* out, eu_timing_t* t,
                         bool async = false) {
  int rc = need_up();
  if (rc) return rc;
  if (!f || !o || !pixels || !out) return fail(EU_ERR_ARGUMENT, "null argument");
  rc = check_raster(f, o);
  if (rc) return rc;
  cudaStream_t st = g.stream;
  if (kind == cudaMemcpyDeviceToDevice) {  // order our stream after the caller's work on the raster
    CK(cudaEventRecord(g.ev[2], caller));
    CK(cudaStreamWaitEvent(st, g.ev[2], 0));
  }
  // until it is registered the source is owned here: any early return frees the container
  struct Guard {
    eu_source* s;
    ~Guard() {
      if (s) {
        cudaStreamSynchronize(g.up_stream);  // placement copies of a failed upload may still be in flight
        pool_free(s->container);
        delete s;
      }
    }
  } guard{new eu_source()};
  eu_source* s = guard.s;
  s->container = nullptr;
  s->last_used_cycle = g.cycle;
  s->refs = 1;
  s->foreign_use = false;
  int launches = 0;
  float copy_ms = 0;
  CK(cudaEventRecord(g.ev[0], st));
  // Host rasters ALWAYS travel on the upload stream, also for the blocking entry point: the staging /
  // render stream then never issues a PCIe copy itself. (Measured: once it had - blocking calls before
  // pipelined ones - the pipelined uploads and downloads stopped overlapping, 9 -> 20-50 ms per frame.)
  rc = stage_on_device(f, o, pixels, kind, st, (async || kind == cudaMemcpyHostToDevice) ? g.up_stream : st, s, &launches,
                       &copy_ms);
  if (rc == EU_OK) rc = maybe_pad(o, s, st, &launches);
  if (rc != EU_OK) return rc;
  CK(cudaEventRecord(g.ev[1], st));
  if (!async) CK(cudaStreamSynchronize(st));  // renders are enqueued on the same stream, i.e. behind the staging
  if (t && !async) {
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, g.ev[0], g.ev[1]));
    CK(cudaEventElapsedTime(&copy_ms, g.ev[2], g.ev[3]));
    bool host = kind == cudaMemcpyHostToDevice;
    // staging kernels (a device-side placement copy included). Host rasters arrive on the upload stream
    // and the staging stream waits for ev[3], the end of the copies: the kernels are what lies between
    // that event and ev[1] (the allocation and the clear happen on the upload stream, before ev[2]).
    if (host) CK(cudaEventElapsedTime(&ms, g.ev[3], g.ev[1]));
    t->render_ms = ms;
    t->h2d_ms = host ? copy_ms : 0.0f;
    t->d2h_ms = 0;
    t->launches = launches;
    t->shape = 0;
  }
  guard.s = nullptr;  // from here on the registry owns it
  g.sources.push_back(s);
  if (asset_key && *asset_key) {
    s->key = asset_key;
    auto it = g.by_key.find(s->key);
    if (it != g.by_key.end()) {  // replaced: the old object stays alive until released / aged out
      it->second->key.clear();
      it->second = s;
    } else {
      g.by_key[s->key] = s;
    }
  }
  *out = s;
  return EU_OK;
}

int eu_source_upload_device(const char* asset_key, const eu_facet_t* f, const eu_opts_t* o, const float* d_pixels,
                            void* cuda_stream, eu_source_h* out, eu_timing_t* t) {
  return upload_common(asset_key, f, o, d_pixels, cudaMemcpyDeviceToDevice, (cudaStream_t)cuda_stream, out, t);
}

int eu_source_upload(const char* asset_key, const eu_facet_t* f, const eu_opts_t* o, const float* pixels,
                     eu_source_h* out, eu_timing_t* t) {
  return upload_common(asset_key, f, o, pixels, cudaMemcpyHostToDevice, nullptr, out, t);
}

int eu_source_upload_alpha(const char* asset_key, const eu_facet_t* f, const eu_opts_t* o, const float* pixels,
                           const eu_alpha_spec_t* a, eu_source_h* out, eu_timing_t* t) {
  int rc = need_up();
  if (rc) return rc;
  if (!f || !o || !pixels || !a || !out) return fail(EU_ERR_ARGUMENT, "null argument");
  const int nat = a->native_nchannels, C = f->nchannels;
  if (nat < 1 || nat > 4 || !(C == nat || (C == nat + 1 && (nat == 1 || nat == 3))) || (C != 2 && C != 4))
    return fail(EU_ERR_ARGUMENT, "masked facets carry alpha: %d native channels cannot become %d", nat, C);
  if (f->projection == EU_CUBEMAP || f->projection == EU_BIATAN6)
    return fail(EU_ERR_ARGUMENT, "masks and lens crop apply to single images, not to cubemaps");
  if (a->n_masks < 0 || (a->n_masks > 0 && (!a->mask_sizes || !a->mask_xy))) return fail(EU_ERR_ARGUMENT, "bad mask list");
  if (f->width <= 0 || f->height <= 0 || f->window_width <= 0 || f->window_height <= 0)
    return fail(EU_ERR_ARGUMENT, "bad raster description (run eu_facet_prepare)");
  const int w = f->window_width, h = f->window_height;  // the raster on hand is the window
  const size_t n = (size_t)w * h;
  std::vector<unsigned char> plane(n);
  eu_build_alpha_mask(f, a, plane.data());  // host: polygons and crop are a few scan lines each
  cudaStream_t st = g.stream;
  struct Scratch {  // freed (stream-ordered) on every path out of this function
    unsigned char* mask = nullptr;
    float *raw = nullptr, *a = nullptr, *b = nullptr, *px = nullptr;
    ~Scratch() {
      if (mask) cudaFreeAsync(mask, g.stream);
      pool_free(raw);
      pool_free(a);
      pool_free(b);
      pool_free(px);
    }
  } d;
  CK(cudaMallocAsync((void**)&d.mask, n, st));
  CK(pool_alloc(&d.raw, n * nat));
  CK(pool_alloc(&d.a, n));
  CK(pool_alloc(&d.b, n));
  CK(pool_alloc(&d.px, n * C));
  CK(cudaEventRecord(g.ev[2], st));
  CK(cudaMemcpyAsync(d.mask, plane.data(), n, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d.raw, pixels, n * nat * sizeof(float), cudaMemcpyHostToDevice, st));
  CK(cudaEventRecord(g.ev[3], st));
  CK(eu_launch_alpha_apply(d.mask, d.a, d.b, d.raw, nat, d.px, C, w, h, st));
  CK(cudaStreamSynchronize(st));  // plane goes out of scope; ev[2]/ev[3] are reused by the staging below
  float h2d = 0;
  CK(cudaEventElapsedTime(&h2d, g.ev[2], g.ev[3]));
  rc = upload_common(asset_key, f, o, d.px, cudaMemcpyDeviceToDevice, st, out, t);
  if (rc == EU_OK && t) {
    t->h2d_ms = h2d;
    t->launches += 3;
  }
  return rc;
}

eu_source_h eu_source_find(const char* asset_key) {
  if (!g.up || !asset_key) return nullptr;
  auto it = g.by_key.find(asset_key);
  if (it == g.by_key.end()) return nullptr;
  it->second->last_used_cycle = g.cycle;
  return it->second;
}

int eu_source_release(eu_source_h s) {
  int rc = need_up();
  if (rc) return rc;
  if (!known_source(s)) return fail(EU_ERR_ARGUMENT, "not a live source handle");
  // the container is freed in stream order on the library stream, behind every render enqueued
  // there; only renders on a caller's stream (eu_render_rows) need a device-wide wait
  if (s->foreign_use) CK(cudaDeviceSynchronize());
  free_source(s);
  return EU_OK;
}

// asset_handler_t::conclude_cycle, environment.h:200-227: keyed assets that were not used in
// the cycle that ends now are dropped; unkeyed sources belong to the caller.
int eu_cycle(void) {
  int rc = need_up();
  if (rc) return rc;
  CK(cudaDeviceSynchronize());
  std::vector<eu_source*> drop;
  for (auto s : g.sources)
    if (!s->key.empty() && s->last_used_cycle < g.cycle) drop.push_back(s);
  for (auto s : drop) free_source(s);
  g.cycle++;
  return EU_OK;
}

size_t eu_source_container_floats(eu_source_h s, int32_t shape[4]) {
  if (!g.up || !known_source(s)) return 0;
  if (shape) {
    shape[0] = s->cw;
    shape[1] = s->chh;
    shape[2] = s->lx;
    shape[3] = s->ly;
  }
  return (size_t)s->cw * s->chh * s->nch;
}

int eu_source_download(eu_source_h s, float* out) {
  int rc = need_up();
  if (rc) return rc;
  if (!known_source(s) || !out) return fail(EU_ERR_ARGUMENT, "bad argument");
  CK(cudaStreamSynchronize(g.stream));
  // strip the row padding (and the texel padding of the 16-byte layout): the caller gets the
  // reference's container layout, cw*chh texels of nch floats
  if (s->tstride == s->nch) {
    CK(cudaMemcpy2D(out, (size_t)s->cw * s->nch * sizeof(float), s->container, (size_t)s->pitch * sizeof(float),
                    (size_t)s->cw * s->nch * sizeof(float), s->chh, cudaMemcpyDeviceToHost));
  } else {
    size_t ntex = (size_t)s->cw * s->chh;
    CK(cudaMemcpy2D(out, s->nch * sizeof(float), s->container, s->tstride * sizeof(float), s->nch * sizeof(float), ntex,
                    cudaMemcpyDeviceToHost));
  }
  return EU_OK;
}

static const float* g_peer_asked = nullptr;  // last d_out looked up with cudaPointerGetAttributes ...
static int g_peer_answer = 0;                // ... and whether it is another GPU's memory

int eu_render_rows(const eu_target_t* t, const eu_opts_t* o, int n_facets, const eu_facet_t* facets,
                   const eu_source_h* sources, const eu_tap_t* taps, int n_taps, int row0, int row1, float* d_out,
                   void* cuda_stream, eu_timing_t* timing) {
  if (!t) return fail(EU_ERR_ARGUMENT, "null argument");
  return eu_render_rows_pitched(t, o, n_facets, facets, sources, taps, n_taps, row0, row1, d_out,
                                out_width(t) * t->nchannels, cuda_stream, timing);
}

int eu_render_rows_pitched(const eu_target_t* t, const eu_opts_t* o, int n_facets, const eu_facet_t* facets,
                           const eu_source_h* sources, const eu_tap_t* taps, int n_taps, int row0, int row1,
                           float* d_out, int out_pitch_floats, void* cuda_stream, eu_timing_t* timing) {
  if (!t) return fail(EU_ERR_ARGUMENT, "null argument");
  return eu_render_rect_pitched(t, o, n_facets, facets, sources, taps, n_taps, row0, row1, 0, out_width(t), d_out,
                                out_pitch_floats, 0, cuda_stream, timing);
}

int eu_render_rect_pitched(const eu_target_t* t, const eu_opts_t* o, int n_facets, const eu_facet_t* facets,
                           const eu_source_h* sources, const eu_tap_t* taps, int n_taps, int row0, int row1, int col0,
                           int col1, float* d_out, int out_pitch_floats, int out_texel_floats, void* cuda_stream,
                           eu_timing_t* timing) {
  int rc = need_up();
  if (rc) return rc;
  Plan plan;
  rc = build_plan(t, o, n_facets, facets, sources, taps, n_taps, (cudaStream_t)cuda_stream, plan);
  if (rc) return rc;
  if (row0 < 0 || row1 > out_height(t) || row0 >= row1) return fail(EU_ERR_ARGUMENT, "bad row band [%d,%d)", row0, row1);
  if (col0 < 0 || col1 > out_width(t) || col0 >= col1 || (col0 & 31))
    return fail(EU_ERR_ARGUMENT, "bad column range [%d,%d): it must lie inside the raster and start at a multiple of 32", col0, col1);
  plan.P.col0 = col0;
  plan.P.col1 = col1;
  if (out_texel_floats == 0) out_texel_floats = t->nchannels;
  if (out_texel_floats != t->nchannels && !(t->nchannels == 3 && out_texel_floats == 4))
    return fail(EU_ERR_ARGUMENT, "output texels of %d floats for a %d-channel job", out_texel_floats, t->nchannels);
  if (out_texel_floats == 4 && t->nchannels == 3 && ((out_pitch_floats & 3) || (reinterpret_cast<uintptr_t>(d_out) & 15)))
    return fail(EU_ERR_ARGUMENT, "16-byte output texels need a 16-byte aligned raster");
  plan.P.out_tstride = out_texel_floats;
  if (!d_out) return fail(EU_ERR_ARGUMENT, "null output");
  if (out_pitch_floats < out_width(t) * (out_texel_floats ? out_texel_floats : t->nchannels))
    return fail(EU_ERR_ARGUMENT, "output pitch %d is shorter than a row", out_pitch_floats);
  cudaStream_t caller = (cudaStream_t)cuda_stream;
  if (caller != g.stream)
    for (int i = 0; i < n_facets; i++) sources[i]->foreign_use = true;
  // tables were enqueued on the library stream; the render runs on the caller's stream
  CK(cudaEventRecord(g.ev[2], g.stream));
  CK(cudaStreamWaitEvent(caller, g.ev[2], 0));
  plan.P.row0 = row0;
  plan.P.row1 = row1;
  plan.P.out = d_out;
  plan.P.out_pitch = out_pitch_floats;
  plan.P.index_out = nullptr;
  {  // a frame opened with eu_frame_open lives on another GPU: wide stores for the link
    // a band is rendered to the same address frame after frame: the driver is asked once per address
    // (the answer is dropped whenever a frame is opened, closed or freed)
    if (d_out != g_peer_asked) {
      cudaPointerAttributes pa;
      g_peer_answer = 0;
      if (cudaPointerGetAttributes(&pa, d_out) == cudaSuccess) {
        if (pa.type == cudaMemoryTypeDevice && pa.device != g.device) g_peer_answer = 1;
      } else {
        cudaGetLastError();
      }
      g_peer_asked = d_out;
    }
    plan.P.wide_stores = g_peer_answer;
    if (o->reserved[1] & EU_OPT_NARROW_STORES) plan.P.wide_stores = 0;  // back-end option: 4-byte stores everywhere
  }
  if (timing) CK(cudaEventRecord(g.ev[0], caller));
  int shape = 0;
  CK(eu_launch_render(plan.P, caller, &shape));
  CK(plan_done(plan, caller));
  plan.launches++;
  if (timing) {
    CK(cudaEventRecord(g.ev[1], caller));
    CK(cudaEventSynchronize(g.ev[1]));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, g.ev[0], g.ev[1]));
    timing->render_ms = ms;
    timing->h2d_ms = timing->d2h_ms = 0;
    timing->launches = plan.launches;
    timing->shape = shape;
  }
  return EU_OK;
}

int eu_render(const eu_target_t* t, const eu_opts_t* o, int n_facets, const eu_facet_t* facets,
              const eu_source_h* sources, const eu_tap_t* taps, int n_taps, float* out, eu_timing_t* timing) {
  int rc = need_up();
  if (rc) return rc;
  if (!t || !out) return fail(EU_ERR_ARGUMENT, "null argument");
  if (t->width <= 0 || t->height <= 0 || t->nchannels < 1) return fail(EU_ERR_ARGUMENT, "target not prepared");
  size_t n = (size_t)out_width(t) * out_height(t) * t->nchannels;
  rc = grow(g.d_out, g.out_cap, n);
  if (rc) return rc;
  eu_timing_t tm;
  rc = eu_render_rows(t, o, n_facets, facets, sources, taps, n_taps, 0, out_height(t), g.d_out, g.stream, &tm);
  if (rc) return rc;
  // the download runs on the download stream, for the same reason as the upload above
  CK(cudaEventRecord(g.ev[1], g.stream));
  CK(cudaStreamWaitEvent(g.down_stream, g.ev[1], 0));
  CK(cudaEventRecord(g.ev[2], g.down_stream));
  CK(cudaMemcpyAsync(out, g.d_out, n * sizeof(float), cudaMemcpyDeviceToHost, g.down_stream));
  CK(cudaEventRecord(g.ev[3], g.down_stream));
  CK(cudaStreamSynchronize(g.down_stream));
  if (timing) {
    *timing = tm;
    CK(cudaEventElapsedTime(&timing->d2h_ms, g.ev[2], g.ev[3]));
  }
  return EU_OK;
}

int eu_source_upload_async(const char* asset_key, const eu_facet_t* f, const eu_opts_t* o, const float* pixels,
                           eu_source_h* out) {
  return upload_common(asset_key, f, o, pixels, cudaMemcpyHostToDevice, nullptr, out, nullptr, true);
}

struct eu_job {
  int slot;
};
static eu_job g_job_handles[EU_MAX_JOBS_IN_FLIGHT];

int eu_render_async(const eu_target_t* t, const eu_opts_t* o, int n_facets, const eu_facet_t* facets,
                    const eu_source_h* sources, const eu_tap_t* taps, int n_taps, float* out, eu_job_h* job) {
  int rc = need_up();
  if (rc) return rc;
  if (!t || !out || !job) return fail(EU_ERR_ARGUMENT, "null argument");
  if (t->width <= 0 || t->height <= 0 || t->nchannels < 1) return fail(EU_ERR_ARGUMENT, "target not prepared");
  const int slot = g.next_job;
  Context::JobSlot& J = g.jobs[slot];
  if (J.pending) return fail(EU_ERR_STATE, "%d jobs are in flight already: eu_job_wait one first", EU_MAX_JOBS_IN_FLIGHT);
  const size_t n = (size_t)out_width(t) * out_height(t) * t->nchannels;
  if (n > J.cap) {
    CK(cudaDeviceSynchronize());
    if (J.d_out) CK(cudaFree(J.d_out));
    J.d_out = nullptr;
    J.cap = 0;
    CK(cudaMalloc(&J.d_out, n * sizeof(float)));
    J.cap = n;
  }
  Plan plan;
  rc = build_plan(t, o, n_facets, facets, sources, taps, n_taps, g.stream, plan);
  if (rc) return rc;
  plan.P.row0 = 0;
  plan.P.row1 = out_height(t);
  plan.P.out = J.d_out;
  plan.P.out_pitch = out_width(t) * t->nchannels;
  plan.P.out_tstride = t->nchannels;
  plan.P.index_out = nullptr;
  CK(cudaEventRecord(J.start, g.stream));
  CK(eu_launch_render(plan.P, g.stream));
  CK(plan_done(plan, g.stream));
  CK(cudaEventRecord(J.rendered, g.stream));
  CK(cudaStreamWaitEvent(g.down_stream, J.rendered, 0));
  CK(cudaMemcpyAsync(out, J.d_out, n * sizeof(float), cudaMemcpyDeviceToHost, g.down_stream));
  CK(cudaEventRecord(J.done, g.down_stream));
  // the slot's buffer is rewritten only after its download: the render stream waits when it wraps
  J.pending = true;
  J.launches = plan.launches + 1;
  g.next_job = (slot + 1) % EU_MAX_JOBS_IN_FLIGHT;
  CK(cudaStreamWaitEvent(g.stream, g.jobs[g.next_job].done, 0));  // no-op for a slot that was never used
  g_job_handles[slot].slot = slot;
  *job = &g_job_handles[slot];
  return EU_OK;
}

int eu_job_wait(eu_job_h job, eu_timing_t* timing) {
  int rc = need_up();
  if (rc) return rc;
  if (!job || job->slot < 0 || job->slot >= EU_MAX_JOBS_IN_FLIGHT) return fail(EU_ERR_ARGUMENT, "not a job handle");
  Context::JobSlot& J = g.jobs[job->slot];
  if (!J.pending) return fail(EU_ERR_STATE, "job is not pending");
  CK(cudaEventSynchronize(J.done));
  J.pending = false;
  if (timing) {
    CK(cudaEventElapsedTime(&timing->render_ms, J.start, J.rendered));
    CK(cudaEventElapsedTime(&timing->d2h_ms, J.rendered, J.done));  // includes waiting for the download stream
    timing->h2d_ms = 0;
    timing->launches = J.launches;
    timing->shape = 0;
  }
  return EU_OK;
}

// ---- tethered output (to_screen_t, envutil_payload.cc:298-413) --------------------------------------
static int screen_lut_ready() {
  if (g.d_screen_lut) return EU_OK;
  float lut[257];
  eu_screen_lut(lut);
  CK(cudaMalloc(&g.d_screen_lut, sizeof(lut)));
  // first use only: a blocking copy from the stack, then a device-wide synchronise, so that the table is in place
  // for whichever stream reads it first (a pageable copy this small may return before its DMA has finished)
  CK(cudaMemcpy(g.d_screen_lut, lut, sizeof(lut), cudaMemcpyHostToDevice));
  CK(cudaDeviceSynchronize());
  return EU_OK;
}

int eu_to_screen_device(const float* d_pixels, int nchannels, size_t n_pixels, uint32_t* d_out, void* cuda_stream) {
  int rc = need_up();
  if (rc) return rc;
  if (!d_pixels || !d_out) return fail(EU_ERR_ARGUMENT, "null argument");
  if (nchannels < 1 || nchannels > 4) return fail(EU_ERR_ARGUMENT, "%d channels", nchannels);
  if ((nchannels == 4 && (reinterpret_cast<uintptr_t>(d_pixels) & 15)) || (nchannels == 2 && (reinterpret_cast<uintptr_t>(d_pixels) & 7)))
    return fail(EU_ERR_ARGUMENT, "pixels of %d channels must be %d-byte aligned", nchannels, nchannels * 4);
  rc = screen_lut_ready();
  if (rc) return rc;
  CK(eu_launch_to_screen(d_pixels, nchannels, n_pixels, g.d_screen_lut, d_out, (cudaStream_t)cuda_stream));
  return EU_OK;
}

int eu_render_screen(const eu_target_t* t, const eu_opts_t* o, int n_facets, const eu_facet_t* facets,
                     const eu_source_h* sources, const eu_tap_t* taps, int n_taps, uint32_t* out, eu_timing_t* timing) {
  int rc = need_up();
  if (rc) return rc;
  if (!t || !out) return fail(EU_ERR_ARGUMENT, "null argument");
  if (t->width <= 0 || t->height <= 0 || t->nchannels < 1) return fail(EU_ERR_ARGUMENT, "target not prepared");
  const size_t npx = (size_t)out_width(t) * out_height(t);
  rc = grow(g.d_out, g.out_cap, npx * t->nchannels);
  if (rc) return rc;
  rc = grow(g.d_index, g.index_cap, npx);
  if (rc) return rc;
  rc = screen_lut_ready();
  if (rc) return rc;
  eu_target_t tt = *t;
  tt.gain = 0.0;  // work() skips the un-brighten stage when it runs tethered (envutil_payload.cc:491)
  eu_timing_t tm;
  rc = eu_render_rows(&tt, o, n_facets, facets, sources, taps, n_taps, 0, out_height(t), g.d_out, g.stream, &tm);
  if (rc) return rc;
  uint32_t* d_scr = reinterpret_cast<uint32_t*>(g.d_index);
  CK(eu_launch_to_screen(g.d_out, t->nchannels, npx, g.d_screen_lut, d_scr, g.stream));
  tm.launches += 1;
  CK(cudaEventRecord(g.ev[1], g.stream));
  CK(cudaStreamWaitEvent(g.down_stream, g.ev[1], 0));
  CK(cudaEventRecord(g.ev[2], g.down_stream));
  CK(cudaMemcpyAsync(out, d_scr, npx * sizeof(uint32_t), cudaMemcpyDeviceToHost, g.down_stream));
  CK(cudaEventRecord(g.ev[3], g.down_stream));
  CK(cudaStreamSynchronize(g.down_stream));
  if (timing) {
    *timing = tm;
    CK(cudaEventElapsedTime(&timing->d2h_ms, g.ev[2], g.ev[3]));
  }
  return EU_OK;
}

int eu_debug_planes(const eu_target_t* t, const eu_opts_t* o, int n_facets, const eu_facet_t* facets,
                    const eu_source_h* sources, int32_t* index_out) {
  int rc = need_up();
  if (rc) return rc;
  if (!t || !index_out) return fail(EU_ERR_ARGUMENT, "null argument");
  Plan plan;
  rc = build_plan(t, o, n_facets, facets, sources, nullptr, 0, g.stream, plan);
  if (rc) return rc;
  size_t n = (size_t)out_width(t) * out_height(t);
  rc = grow(g.d_index, g.index_cap, n);
  if (rc) return rc;
  plan.P.row0 = 0;
  plan.P.row1 = out_height(t);
  plan.P.out = nullptr;
  plan.P.out_pitch = out_width(t) * t->nchannels;
  plan.P.out_tstride = t->nchannels;
  plan.P.index_out = g.d_index;
  CK(eu_launch_render(plan.P, g.stream));
  CK(plan_done(plan, g.stream));
  CK(cudaMemcpyAsync(index_out, g.d_index, n * sizeof(int32_t), cudaMemcpyDeviceToHost, g.stream));
  CK(cudaStreamSynchronize(g.stream));
  return EU_OK;
}

int eu_debug_tie_plane(const eu_target_t* t, const eu_opts_t* o, int n_facets, const eu_facet_t* facets,
                       const eu_source_h* sources, int ulps, unsigned char* tie_out) {
  int rc = need_up();
  if (rc) return rc;
  if (!t || !tie_out || ulps < 0) return fail(EU_ERR_ARGUMENT, "bad argument");
  Plan plan;
  rc = build_plan(t, o, n_facets, facets, sources, nullptr, 0, g.stream, plan);
  if (rc) return rc;
  size_t n = (size_t)out_width(t) * out_height(t);
  unsigned char* d_tie = nullptr;
  CK(cudaMallocAsync((void**)&d_tie, n, g.stream));
  plan.P.row0 = 0;
  plan.P.row1 = out_height(t);
  cudaError_t e = eu_launch_tie_plane(plan.P, d_tie, ulps, g.stream);
  if (e == cudaSuccess) e = plan_done(plan, g.stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(tie_out, d_tie, n, cudaMemcpyDeviceToHost, g.stream);
  cudaFreeAsync(d_tie, g.stream);
  if (e != cudaSuccess) return fail(EU_ERR_CUDA, "tie plane: %s", cudaGetErrorString(e));
  CK(cudaStreamSynchronize(g.stream));
  return EU_OK;
}

// ---- a source whose raster is produced on the device, in place -----------------------------
int eu_source_reserve(const char* asset_key, const eu_facet_t* f, const eu_opts_t* o, eu_source_h* out, float** d_core,
                      int* pitch_floats, int* texel_floats) {
  int rc = need_up();
  if (rc) return rc;
  if (!f || !o || !out || !d_core || !pitch_floats || !texel_floats) return fail(EU_ERR_ARGUMENT, "null argument");
  if (f->projection == EU_CUBEMAP || f->projection == EU_BIATAN6)
    return fail(EU_ERR_UNSUPPORTED, "eu_source_reserve is for single images (a cubemap's faces are re-arranged on upload)");
  rc = check_raster(f, o);
  if (rc) return rc;
  eu_source* s = new eu_source();
  s->container = nullptr;
  s->last_used_cycle = g.cycle;
  s->refs = 1;
  s->foreign_use = false;
  s->projection = f->projection;
  s->nch = f->nchannels;
  s->degree = o->spline_degree;
  s->tstride = f->nchannels;
  mount_layout(f, o->spline_degree, s);
  // the texel layout follows the same rule as an uploaded source's (maybe_pad): RGB for degree <= 1 in 16-byte texels
  if (s->nch == 3 && (o->reserved[0] == 1 || (o->reserved[0] == 0 && o->spline_degree <= 1))) {
    s->tstride = 4;
    s->pitch = s->cw * 4;
  }
  cudaError_t e = pool_alloc(&s->container, (size_t)s->pitch * s->chh);
  // rows the caller never writes are zero, not whatever the pool held (a prefilter would spread NaNs)
  if (e == cudaSuccess) e = cudaMemsetAsync(s->container, 0, (size_t)s->pitch * s->chh * sizeof(float), g.stream);
  if (e != cudaSuccess) {
    pool_free(s->container);
    delete s;
    return fail(EU_ERR_CUDA, "container: %s", cudaGetErrorString(e));
  }
  s->reserved = true;
  g.sources.push_back(s);
  if (asset_key && *asset_key) {
    s->key = asset_key;
    auto it = g.by_key.find(s->key);
    if (it != g.by_key.end()) {  // replaced: the old object stays alive until released / aged out
      it->second->key.clear();
      it->second = s;
    } else {
      g.by_key[s->key] = s;
    }
  }
  *out = s;
  *d_core = s->container + (size_t)s->ly * s->pitch + (size_t)s->lx * s->tstride;
  *pitch_floats = s->pitch;
  *texel_floats = s->tstride;
  return EU_OK;
}

int eu_source_commit(eu_source_h s, const eu_facet_t* f, const eu_opts_t* o, void* cuda_stream, eu_timing_t* t) {
  int rc = need_up();
  if (rc) return rc;
  if (!s || !f || !o) return fail(EU_ERR_ARGUMENT, "null argument");
  if (!known_source(s)) return fail(EU_ERR_ARGUMENT, "not a live source handle");
  if (s->kind != EU_SRC_MOUNT || !s->reserved) return fail(EU_ERR_ARGUMENT, "not a reserved single-image source");
  // the description must be the one the container was laid out for (texel stride, shape, boundary conditions)
  eu_source probe;
  mount_layout(f, o->spline_degree, &probe);
  if (f->nchannels != s->nch || f->projection != s->projection || o->spline_degree != s->degree || probe.w != s->w ||
      probe.h != s->h || probe.bc0 != s->bc0 || probe.bc1 != s->bc1 || probe.cw != s->cw || probe.chh != s->chh)
    return fail(EU_ERR_ARGUMENT, "eu_source_commit: the facet / options differ from those given to eu_source_reserve");
  int pdeg = o->prefilter_degree < 0 ? o->spline_degree : o->prefilter_degree;
  if (pdeg > EU_MAX_DEGREE) return fail(EU_ERR_ARGUMENT, "prefilter degree %d out of range", pdeg);
  // prefilter and brace run on the caller's stream, behind the work that wrote the rows and before whatever renders
  // from the source: commits of different sources on different streams do not meet (they used to hop through the
  // library's own stream, which chained them)
  cudaStream_t st = (cudaStream_t)cuda_stream;
  int launches = 0;
  if (t) CK(cudaEventRecord(g.ev[0], st));
  rc = mount_finish(f, pdeg, s, st, &launches);
  if (rc) return rc;
  if (t) CK(cudaEventRecord(g.ev[1], st));
  if (t) {  // blocking and timed; with t == NULL the call only enqueues
    CK(cudaStreamSynchronize(st));
    CK(cudaEventElapsedTime(&t->render_ms, g.ev[0], g.ev[1]));
    t->h2d_ms = t->d2h_ms = 0;
    t->launches = launches;
    t->shape = 0;
  }
  return EU_OK;
}

int eu_source_write_rect(eu_source_h s, const float* pixels, size_t src_pitch_floats, int row0, int row1, int col0,
                         int col1, void* cuda_stream) {
  int rc = need_up();
  if (rc) return rc;
  if (!s || !pixels) return fail(EU_ERR_ARGUMENT, "null argument");
  if (!known_source(s) || !s->reserved) return fail(EU_ERR_ARGUMENT, "not a source from eu_source_reserve");
  if (row0 < 0 || row1 > s->h || row0 >= row1 || col0 < 0 || col1 > s->w || col0 >= col1)
    return fail(EU_ERR_ARGUMENT, "rectangle [%d,%d) x [%d,%d) does not lie inside the %dx%d raster", row0, row1, col0, col1,
                s->w, s->h);
  const size_t wb = (size_t)(col1 - col0) * s->nch * sizeof(float);
  if (src_pitch_floats * sizeof(float) < wb) return fail(EU_ERR_ARGUMENT, "source pitch shorter than the rectangle's rows");
  float* dst = s->container + (size_t)(s->ly + row0) * s->pitch + (size_t)(s->lx + col0) * s->tstride;
  cudaStream_t st = (cudaStream_t)cuda_stream;  // NULL is the legacy default stream, as for eu_render_rows
  cudaPointerAttributes pa;
  cudaMemcpyKind kind = cudaMemcpyHostToDevice;
  if (cudaPointerGetAttributes(&pa, pixels) == cudaSuccess) {
    if (pa.type == cudaMemoryTypeDevice) kind = cudaMemcpyDeviceToDevice;
  } else {
    cudaGetLastError();
  }
  if (s->tstride != s->nch) {  // 16-byte texels: the rectangle lands in a scratch buffer and is widened on the device
    const size_t rowf = (size_t)(col1 - col0) * s->nch;
    const size_t need = rowf * (size_t)(row1 - row0);
    if (!g.rect_stream) CK(cudaStreamCreateWithFlags(&g.rect_stream, cudaStreamNonBlocking));
    Context::RectSlot& R = g.rect_ring[g.next_rect];
    g.next_rect = (g.next_rect + 1) % 3;
    if (!R.copied) {
      CK(cudaEventCreateWithFlags(&R.copied, cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&R.freed, cudaEventDisableTiming));
    }
    if (need > R.cap) {  // grows to the largest rectangle seen; the kernel that last read the old buffer has to finish
      if (R.used) CK(cudaEventSynchronize(R.freed));
      if (R.buf) CK(cudaFree(R.buf));
      R.buf = nullptr;
      R.cap = 0;
      CK(cudaMalloc((void**)&R.buf, need * sizeof(float)));
      R.cap = need;
      R.used = false;
    }
    // device rasters may have been produced by earlier work on the caller's stream: their copy stays on it
    cudaStream_t cs = kind == cudaMemcpyDeviceToDevice ? st : g.rect_stream;
    if (R.used) CK(cudaStreamWaitEvent(cs, R.freed, 0));  // the widening kernel of three rectangles ago
    // a rectangle whose rows follow each other in the caller's buffer travels as ONE contiguous copy: measured on a B200
    // box, the pitched form of the same bytes is 10 % slower alone and overlaps a concurrent download far worse
    // (configs[4], one GPU: upload || download 58.7 ms against 48.9 ms, profiles/r02h_probe_overlap.txt)
    cudaError_t e = src_pitch_floats == rowf
                        ? cudaMemcpyAsync(R.buf, pixels, wb * (size_t)(row1 - row0), kind, cs)
                        : cudaMemcpy2DAsync(R.buf, rowf * sizeof(float), pixels, src_pitch_floats * sizeof(float), wb, row1 - row0, kind, cs);
    if (e == cudaSuccess) e = cudaEventRecord(R.copied, cs);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(st, R.copied, 0);  // the caller's stream: everything after this call sees the texels
    if (e == cudaSuccess) e = eu_launch_pad_texels(R.buf, (int)rowf, dst, s->pitch / 4, col1 - col0, row1 - row0, s->nch, st);
    if (e == cudaSuccess) e = cudaEventRecord(R.freed, st);
    R.used = true;
    if (e != cudaSuccess) return fail(EU_ERR_CUDA, "write_rect: %s", cudaGetErrorString(e));
    return EU_OK;
  }
  CK(cudaMemcpy2DAsync(dst, (size_t)s->pitch * sizeof(float), pixels, src_pitch_floats * sizeof(float), wb, row1 - row0, kind,
                       st));
  return EU_OK;
}

// ---- frames shared between the processes of one box (CUDA IPC) -----------------------------
int eu_frame_alloc(size_t n_floats, float** d_frame) {
  int rc = need_up();
  if (rc) return rc;
  if (!d_frame || n_floats == 0) return fail(EU_ERR_ARGUMENT, "null argument");
  // cudaMalloc, not the stream-ordered pool: pool memory cannot be exported with cudaIpcGetMemHandle
  CK(cudaMalloc((void**)d_frame, n_floats * sizeof(float)));
  return EU_OK;
}

int eu_frame_free(float* d_frame) {
  int rc = need_up();
  if (rc) return rc;
  g_peer_asked = nullptr;
  CK(cudaDeviceSynchronize());
  CK(cudaFree(d_frame));
  return EU_OK;
}

int eu_frame_export(const float* d_frame, unsigned char handle[EU_FRAME_HANDLE_BYTES]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == EU_FRAME_HANDLE_BYTES, "handle size");
  int rc = need_up();
  if (rc) return rc;
  if (!d_frame || !handle) return fail(EU_ERR_ARGUMENT, "null argument");
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, const_cast<float*>(d_frame)));
  memcpy(handle, &h, sizeof(h));
  return EU_OK;
}

int eu_frame_open(const unsigned char handle[EU_FRAME_HANDLE_BYTES], float** d_frame) {
  int rc = need_up();
  if (rc) return rc;
  if (!d_frame || !handle) return fail(EU_ERR_ARGUMENT, "null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  void* p = nullptr;
  CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  g_peer_asked = nullptr;
  *d_frame = static_cast<float*>(p);
  return EU_OK;
}

int eu_frame_close(float* d_frame) {
  int rc = need_up();
  if (rc) return rc;
  g_peer_asked = nullptr;
  CK(cudaDeviceSynchronize());  // no store of ours may still be in flight towards the owner
  CK(cudaIpcCloseMemHandle(d_frame));
  return EU_OK;
}

}  // extern "C"
