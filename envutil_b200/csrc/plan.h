// plan.h - POD job description handed from the host plan builder to the sm_100a kernels.
//
// The host computes every per-job scalar in the reference's precision (api.cu) and narrows
// it to what the reference's functors hold (mostly float); the kernels only do the per-pixel
// work. Field comments cite the reference member each value replaces.
#pragma once
#include <stdint.h>

#define EU_MAX_FACETS 64
#define EU_SMEM_FACETS 24  // synopsis jobs with up to this many facets keep them in shared memory
#define EU_MAX_TAPS 1024
#define EU_INLINE_TAPS 64  // twining filters up to 8 x 8 travel in the kernel parameter block
#define EU_MAX_DEGREE 7
#define EU_SEGMENT 512  // WIELDING_SEGMENT_SIZE (zimt/bill.h:69)
#define EU_LANES 16     // zimt vector width of the reference build we track (zimt/simd.h:106-123)

enum { EU_SRC_MOUNT = 0, EU_SRC_CUBEMAP = 1, EU_SRC_BIATAN6 = 2 };
// CONST0: the gate of an axis of extent 1 (zimt/eval.h:2060-2068), never a boundary condition of the prefilter
enum { EU_BC_PERIODIC = 0, EU_BC_REFLECT = 1, EU_BC_NATURAL = 2, EU_BC_MIRROR = 3, EU_BC_CONST0 = 4 };
// VORONOI_PLUS: alpha compositing of the z-sorted facets (_voronoi_syn_plus), 2/4-channel jobs
enum { EU_MODE_SINGLE = 0, EU_MODE_VORONOI = 1, EU_MODE_HDR = 2, EU_MODE_VORONOI_PLUS = 3 };
enum { EU_HDR_LOW = 0, EU_HDR_MIDDLE = 1, EU_HDR_HIGH = 2 };

// staged source in HBM: interleaved float texels, row-major container = core + brace frame
// (zimt::bspline container, zimt/bspline.h:305-428; cubemap IR: cubemap.h:548-580)
struct SourceDev {
  const float* core;  // texel (0,0) of the core
  int32_t stride;     // floats from one container row to the next
  int32_t tstride;    // floats from one texel to the next (= nch, or 4 for the padded layout)
  int32_t nch;        // channels per texel
  int32_t w, h;       // core shape
  int32_t bc0, bc1;   // EU_BC_* of the two axes (gates: zimt/eval.h:2039-2164)
  float upper_x, upper_y;  // gate limits N-0.5 (zimt/bspline.h:262-286); lower is -0.5
};

struct FacetDev {
  SourceDev src;
  int32_t kind;        // EU_SRC_*
  int32_t projection;  // eu_projection_t of the facet
  float xx[3], yy[3], zz[3];  // basis rows narrowed to float (stepper ctor args, stepper.h:579)
  // mount_t / source_t (environment.h:594-1006)
  double ext_x0, ext_y0;   // total_extent.x0 / .y0 (subtracted in double, :992,:997)
  float ext_w, ext_h;      // float(x1-x0), float(y1-y0)              (:993,:998)
  float total_w, total_h;  // float(total_width/height)               (:994,:999)
  float win_x0, win_x1, win_y0, win_y1;  // window_extent narrowed for the float compares (:970-978)
  float win_xoff, win_yoff;  // window offset in pixels, subtracted after md_to_spline (:1003-1005)
  float win_margin[2];       // dev_facet_mask: how far an approximate coordinate must clear the window's edges (x, y)
  float rcp_w, rcp_h;        // RN(1 / ext_w), RN(1 / ext_h) for dev_div_const
  int32_t fast_div;          // bit 0 / 1: the reciprocal sequence is proven exact for ext_w / ext_h (api.cu)
  int32_t mask_always;     // get_mask yields all-true (cubemaps, fisheye >= 360: :1567,:1741)
  int32_t has_lcp, has_shift, has_shear;  // pto_planar (environment.h:240-284)
  float lcp[4], lcp_s, shift_h, shift_v;
  double shear_g, shear_t;
  // cubemap_view_t (environment.h:1396-1486)
  float refc_md, model_to_px;
  int32_t section_px;
  // environment (environment.h:1786-1860)
  float recip_step, brighten;
  // _hdr_merge_syn (envutil_payload.cc:1354-1375)
  float hdr_optimum;
  int32_t hdr_kind;
  // 1: everything that decides where a ray lands in this facet equals the PREVIOUS facet of the job (exposure
  // brackets of one camera position): the kernels reuse that facet's window position (api.cu: same_geometry)
  int32_t same_geom;
  // --mask_for (masking.h:70-139): masked != 0 -> the colour channels are `paint` (times alpha, if the source has one)
  int32_t masked;
  float paint;
  // generic_stepper + tf_ex_facet + generic_r3 + tf3d_t: facets with PanoTools translation
  // (envutil_payload.cc:1628-1883, geometry.h:1850-1942). Float matrices, rows as r3_t holds them.
  // 'single' jobs on a facet with lens correction / translation put every facet on the generic stepper
  // and generic_r3(ft, fs) becomes up to two tf3d_t in sequence (envutil_payload.cc:1716-1760); a stage
  // without shift is the single rotation ab = rotate(a, b).
  int32_t generic, g_nstage;
  struct TfStage {
    int32_t has_shift;
    float a[9], b[9], ab[9], shift[3], dcp;
  } g_st[2];
};

// pto_planar<float, L, true> of a 'single' job's target facet (environment.h:240-309): the inverse of
// shear, shift and lens polynomial applied to the planar target coordinate before it becomes a ray;
// inverse_lcp's spline (lens_correction.h:273-406): EU_INV_NK knots, NATURAL cubic, braced by 2
#define EU_INV_SZ 100
#define EU_INV_NK (EU_INV_SZ + 4)
struct InvPlanarDev {
  int32_t on, has_shear, has_shift, has_lcp;
  double shear_g, shear_t, s, rr_max;
  float h, v;
  const float* coef;  // device memory: coefficient of knot 0 (two brace values before, two after the last)
  float wm[16];       // cubic weight matrix (the job's own degree may differ)
};

struct TargetDev {
  int32_t projection, width, height, normalize;  // width x height: the raster that is rendered (the crop, if any)
  int32_t full_w, full_h, off_x, off_y;  // the target the steppers are built for, and the crop's origin in it
  float fx0, fx1, fy0, fy1;  // stepper_base scaling factors (stepper.h:299-302)
  float delta;               // 16 * (a1-a0)/W                (stepper.h:305)
  float bias_x, bias_y;      // bias of deriv_stepper's r10 / r01 (stepper.h:303-304,1606-1625)
  float section_md, refc_md; // cubemap/biatan6 steppers    (stepper.h:1265-1266)
  float unbrighten;          // --single: colour channels of the result are multiplied by this (work(), :481-511)
};

// Job shapes with kernels compiled for them: where the table gives a value the kernel folds the
// reference's run-time switch on it (stepper of the target projection, ray -> source coordinate
// functor, boundary gates) into straight-line code; -1 = read from the job as usual. A job
// matches an entry if its target and EVERY facet it evaluates agree with all fixed values and no
// facet has lens correction (pto_planar). Entry 0 = nothing fixed. Same arithmetic either way.
struct RenderSpec {
  int tproj, tnorm;        // TargetDev.projection / .normalize
  int skind, sproj;        // FacetDev.kind / .projection
  int bc0, bc1;            // SourceDev.bc0 / .bc1
  int mask_always;         // FacetDev.mask_always
};
#define EU_N_SPECS 8
constexpr RenderSpec eu_render_specs[EU_N_SPECS] = {
    {-1, -1, -1, -1, -1, -1, -1},
    {0 /*spherical*/, -1, EU_SRC_CUBEMAP, -1, EU_BC_REFLECT, EU_BC_REFLECT, 1},      // 1: cubemap -> spherical
    {0 /*spherical*/, -1, EU_SRC_BIATAN6, -1, EU_BC_REFLECT, EU_BC_REFLECT, 1},      // 2: biatan6 -> spherical
    {2 /*rectilinear*/, 0, EU_SRC_MOUNT, 0, EU_BC_PERIODIC, EU_BC_REFLECT, -1},     // 3: full lat/lon -> rectilinear view
    {6 /*biatan6*/, 0, EU_SRC_MOUNT, 0, EU_BC_PERIODIC, EU_BC_REFLECT, -1},         // 4: full lat/lon -> biatan6 cubemap
    {4 /*fisheye*/, -1, EU_SRC_MOUNT, 0, EU_BC_PERIODIC, EU_BC_REFLECT, -1},        // 5: full lat/lon -> fisheye
    {2 /*rectilinear*/, 1, EU_SRC_MOUNT, 2, EU_BC_REFLECT, EU_BC_REFLECT, 0},       // 6: rectilinear facets -> rectilinear (hdr_merge of brackets)
    {0 /*spherical*/, -1, EU_SRC_MOUNT, 2, EU_BC_REFLECT, EU_BC_REFLECT, 0},        // 7: rectilinear facets -> spherical panorama
};

struct RenderParams {
  TargetDev trg;
  FacetDev f0;              // the facet of single-facet jobs (constant bank)
  // cubemap / biatan6 TARGETS, single-facet jobs: per cube face the constant part and the direction of the face's
  // stepper for f0 (stepper.h:1304-1331: ccc = +-row_a + p1 * +-row_b, vvv = +-row_c of the facet's basis) as
  // A[3], pad, B[3], pad, C[3], pad with the signs folded in: ray = (A + p1 * B) + p0 * C, the same products and
  // sums with the same operands as the switch (a + (-b) is a - b, -(p1 * b) is p1 * (-b)): one indexed constant load
  // per vector instead of three switches on the face per pixel
  alignas(16) float cube_tab[6][12];
  InvPlanarDev inv;         // 'single' jobs: inverse planar transformation of the target facet
  float wmat[64];           // (degree+1)^2 weight matrix (zimt/basis.h:419-543), float, packed
  const FacetDev* facets;   // all facets (global memory), used by the synopsis modes
  const float* taps;        // n_taps x (x*4, y*4, w)  (twining.h:106-121)
  float ptaps[3 * EU_INLINE_TAPS];  // the same, inside the parameter block, for filters of up to EU_INLINE_TAPS taps
  const float2* col_tab;    // [2][width]: per-column stepper terms, plain and x-biased (eu_device.cuh)
  const float2* row_tab;    // [2][height]: per-row stepper terms, plain and y-biased
  const float* planar_raw;  // [2][width] then [2][height]: the bare planar coordinates (generic steppers)
  int32_t n_facets;
  int32_t mode;       // EU_MODE_*
  int32_t degree;     // spline degree of the evaluator
  int32_t n_taps;     // 0: plain rays (ninputs 3), else twining (ninputs 9)
  int32_t taps_inline;  // the first n_taps entries of `taps` are also in `ptaps` (constant bank: uniform loads)
  int32_t nch;
  int32_t tstride;    // floats per texel in HBM: nch, or 4 (padded RGB); same for all facets
  int32_t any_generic;  // the general build is needed: some facet uses the generic stepper (translation) or
                        // differs from the job in channel count / texel stride
  int32_t spec;       // index into eu_render_specs the job matches (0: none)
  int32_t use_tiles;  // 1: stage the gather footprint in shared memory where the kernel supports it
  int32_t src_cw, src_ch;  // container shape of f0's source in texels (tile path)
  int32_t src_lx, src_ly;  // its left / top brace: core texel (0,0) is container texel (lx, ly)
  const float* src_base;   // first float of f0's container (256-byte aligned, rows 16-byte aligned)
  int32_t row0, row1;  // rows rendered by this launch
  int32_t col0, col1;  // columns rendered by this launch (col0 a multiple of 32; the whole width unless a caller narrows it)
  int32_t arith;       // 0: every product and sum rounded separately; 1: fused multiply-adds in the window evaluation
  float* out;          // first float of row `row0`
  int32_t out_pitch;   // floats from one output row to the next (width * nch for a dense band)
  int32_t out_tstride; // floats from one output pixel to the next: nch, or 4 for RGB rendered into a 16-byte-texel container
  int32_t wide_stores; // 1: `out` is another GPU's memory - RGB pixels leave as 128-bit stores (dev_store_pixel)
  int32_t* index_out;  // optional index plane (face / winning facet)
};
