// plan.h - POD job description handed from the host plan builder to the sm_100a kernels.
//
// The host computes every per-job scalar in the reference's precision (see plan.cc) and
// narrows it to what the reference's functors hold (mostly float); the kernels only do the
// per-pixel work. Field comments cite the reference member each value replaces.
#pragma once
#include <stdint.h>

#define EU_MAX_FACETS 64
#define EU_MAX_TAPS 1024
#define EU_MAX_DEGREE 7

enum { EU_SRC_MOUNT = 0, EU_SRC_CUBEMAP = 1, EU_SRC_BIATAN6 = 2 };
enum { EU_GATE_PERIODIC = 0, EU_GATE_MIRROR = 1, EU_GATE_CLAMP = 2 };
enum { EU_MODE_SINGLE = 0, EU_MODE_VORONOI = 1, EU_MODE_HDR = 2 };
enum { EU_HDR_MIDDLE = 0, EU_HDR_LOW = 1, EU_HDR_HIGH = 2 };

// staged source in HBM: interleaved float texels, row-major container = core + brace frame
// (zimt::bspline container, zimt/bspline.h:305-428; cubemap IR: cubemap.h:548-580)
struct SourceDev {
  const float* core;  // texel (0,0) of the core
  int32_t stride_y;   // floats from one container row to the next
  int32_t nch;        // floats per texel
  int32_t w, h;       // core shape
};

struct FacetDev {
  SourceDev src;
  int32_t kind;        // EU_SRC_*
  int32_t projection;  // eu_projection_t of the facet
  float bx[3], by[3], bz[3];  // basis rows narrowed to float (stepper ctor args, stepper.h:579)
  // mount_t / source_t (environment.h:594-1006)
  double ext_x0, ext_y0;   // total_extent.x0 / .y0 (subtracted in double, :992,:997)
  float ext_w, ext_h;      // float(x1-x0), float(y1-y0)              (:993,:998)
  float total_w, total_h;  // float(total_width/height)               (:994,:999)
  float win_xoff, win_yoff;
  float win_x0, win_x1, win_y0, win_y1;  // window_extent as float thresholds equivalent to the
                                         // reference's float-vs-double compares (:970-978)
  int32_t mask_always;     // get_mask yields all-true (cubemaps, fisheye >= 360: :1567,:1741)
  int32_t has_lcp, has_shift, has_shear;  // pto_planar (environment.h:240-284)
  float lcp_a, lcp_b, lcp_c, lcp_d, lcp_s, shift_h, shift_v, shear_g, shear_t;
  // cubemap_view_t (environment.h:1396-1486)
  float refc_md, model_to_px;
  int32_t section_px;
  // safe-evaluator gates (zimt/eval.h:2039-2164, zimt/map.h)
  int32_t gate_x, gate_y;
  float lower_x, upper_x, lower_y, upper_y;
  // environment (environment.h:1786-1860)
  float recip_step, brighten;
  // _hdr_merge_syn (envutil_payload.cc:1354-1375)
  float hdr_optimum;
  int32_t hdr_kind;
};

struct TargetDev {
  int32_t projection, width, height, normalize;
  float fx0, fx1, fy0, fy1;  // stepper_base scaling factors (stepper.h:299-302)
  float delta;               // 16 * (a1-a0)/W                (stepper.h:305)
  float bias_x[3], bias_y[3];  // r00, r10, r01 of deriv_stepper (stepper.h:1606-1625)
  float section_md, refc_md;   // cubemap/biatan6 steppers    (stepper.h:1265-1266)
};

struct RenderParams {
  TargetDev trg;
  int32_t n_facets;
  int32_t mode;       // EU_MODE_*
  int32_t degree;     // spline degree of the evaluator
  int32_t n_taps;     // 0: plain rays (ninputs 3), else twining (ninputs 9)
  int32_t nch;
  int32_t row0, row1;  // rows rendered by this launch
  int32_t want_index;  // debug plane instead of pixels
  float* out;          // first float of row `row0`
  int32_t* index_out;
};
