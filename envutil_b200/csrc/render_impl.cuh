// render_impl.cuh - the fused render kernel: target ray -> (rotation folded into the stepper
// basis) -> source coordinate -> gates -> b-spline window -> twining / synopsis -> brighten ->
// store. One thread per target pixel, 32x8-pixel tiles so that the threads of a block gather
// from a compact source footprint; no intermediate ray or coordinate buffer ever touches HBM.
//
// Replaces zimt::process + get_t/act_t/put_t of the reference (envutil_payload.cc:425-579,
// zimt/wielding.h:155-463) for all stepper x source x synopsis combinations of `fuse`
// (envutil_payload.cc:1885-2284).
//
// Template parameters: NCH channels, TS floats per texel in HBM (NCH, or 4 for padded RGB),
// MODE (single facet / voronoi / hdr_merge), TWINE, DEG (1, 3, or -1 = degree read at run time).
// Included by render_c*.cu, one translation unit per (NCH, TS) so that they compile in parallel.
#pragma once
#include <limits.h>

#include "eu_device.cuh"
#include "kernels.h"

#define TILE_X 32
#define TILE_Y 8

// Both arithmetics live in one library (eu_opts_t.reserved[1] bit 4 selects per job): the render translation
// units are compiled twice, the second time with -DEU_CONTRACT_WINDOW (eu_device.cuh: fused multiply-adds in
// the window evaluation and the twining accumulation). Kernels and launchers of the two builds must not share
// symbols, so everything below sits in a namespace named after the arithmetic and the exported launchers
// carry a suffix (render.cu picks by RenderParams::arith).
#ifdef EU_CONTRACT_WINDOW
#define EU_ARITH_NS eu_contracted
#define EU_ARITH_FN(name) name##_fma
#else
#define EU_ARITH_NS eu_exact
#define EU_ARITH_FN(name) name
#endif

namespace EU_ARITH_NS {

__device__ __forceinline__ int first_lane_column(int x) {
  int seg0 = (x / EU_SEGMENT) * EU_SEGMENT;
  return seg0 + (x - seg0) % EU_LANES;
}

// The pixel store (zimt/put.h:122-135: interleaved NCH-tuples). RGB pixels are 12 bytes: written
// per lane that is three 4-byte stores scattered over the warp's 384 contiguous bytes. Into local
// HBM that is the fastest form (L2 merges the sectors; measured on C2: 0.554 ms against 0.580 ms
// for the variant below). When `out` is a peer GPU's frame (P.wide_stores, NVLink) and the warp is
// complete and its span 16-byte aligned, the pixels go through a 384-byte shared-memory slot of the
// warp and leave as 24 128-bit stores: full 32-byte sectors per request on the link.
template <int NCH>
__device__ __forceinline__ void dev_store_pixel(const RenderParams& P, const TargetDev& T, int x, int y,
                                                const float px[NCH], float* wslot) {
  if constexpr (NCH == 3) {
    if (P.out_tstride == 4) {  // RGB into a container of 16-byte texels (stage one of a two-stage job): one 128-bit store
      float* d4 = P.out + (size_t)(y - P.row0) * P.out_pitch + (size_t)x * 4;
      *reinterpret_cast<float4*>(d4) = make_float4(px[0], px[1], px[2], 0.0f);
      return;
    }
  }
  float* dst = P.out + (size_t)(y - P.row0) * P.out_pitch + (size_t)x * NCH;
  if constexpr (NCH == 4) {
    if (((P.out_pitch & 3) | (int)(reinterpret_cast<uintptr_t>(P.out) & 15)) == 0) {
      *reinterpret_cast<float4*>(dst) = make_float4(px[0], px[1], px[2], px[3]);
    } else {
      dst[0] = px[0]; dst[1] = px[1]; dst[2] = px[2]; dst[3] = px[3];
    }
  } else if constexpr (NCH == 3) {
    const int lane = threadIdx.x;  // TILE_X == 32: a warp is one row of the tile
    const bool vec = P.wide_stores && ((P.out_pitch & 3) == 0) && ((reinterpret_cast<uintptr_t>(P.out) & 15) == 0) &&
                     (x - lane + TILE_X <= P.col1);  // warp-uniform: the whole warp renders
    if (vec) {
      wslot[lane * 3] = px[0];
      wslot[lane * 3 + 1] = px[1];
      wslot[lane * 3 + 2] = px[2];
      __syncwarp();
      if (lane < 24)
        reinterpret_cast<float4*>(dst - lane * 3)[lane] = reinterpret_cast<const float4*>(wslot)[lane];
    } else {
      dst[0] = px[0]; dst[1] = px[1]; dst[2] = px[2];
    }
  } else {
#pragma unroll
    for (int c = 0; c < NCH; c++) dst[c] = px[c];
  }
}

// the job's target and first facet as the kernel sees them: the parameter block itself (SP 0), or
// copies with the fields of the compiled-in shape replaced by constants
template <int SP>
struct SpecView {
  TargetDev t;
  FacetDev f;
  __device__ __forceinline__ explicit SpecView(const RenderParams& P) : t(P.trg), f(P.f0) {
    dev_spec_target<SP>(t);
    dev_spec_facet<SP>(f);
  }
  __device__ __forceinline__ const TargetDev& trg() const { return t; }
  __device__ __forceinline__ const FacetDev& f0() const { return f; }
};
template <>
struct SpecView<0> {
  const RenderParams& p;
  __device__ __forceinline__ explicit SpecView(const RenderParams& P) : p(P) {}
  __device__ __forceinline__ const TargetDev& trg() const { return p.trg; }
  __device__ __forceinline__ const FacetDev& f0() const { return p.f0; }
};

// the ray of one facet for a pixel: the projection's own stepper, or the generic stepper for
// facets with translation. which: 0 = r00, 1 = r10 (x-biased), 2 = r01 (y-biased)
struct PixelTerms {
  ColTerm col, colb, first, firstb;
  RowTerm row, rowb;
  float px, pxb, py, pyb;  // bare planar coordinates (generic steppers only)
};
// ctab: RenderParams::cube_tab when F is the job's only facet (single-facet jobs), else null
template <bool GEN, int WHICH>
__device__ __forceinline__ void dev_facet_ray(const TargetDev& T, const InvPlanarDev& IP, const FacetDev& F,
                                              const PixelTerms& t, int y, float r[3], const float (*ctab)[12] = nullptr) {
  if constexpr (GEN) {
    if (F.generic) {
      dev_generic_ray(T, IP, F, WHICH == 1 ? t.pxb : t.px, WHICH == 2 ? t.pyb : t.py, r);
      return;
    }
  }
  if (ctab != nullptr && T.projection >= EU_CUBEMAP) {
    dev_stepper_cube_tab(T, ctab, WHICH == 1 ? t.colb : t.col, WHICH == 2 ? t.rowb : t.row, r);
    return;
  }
  dev_stepper(T, F.xx, F.yy, F.zz, WHICH == 1 ? t.colb : t.col, WHICH == 2 ? t.rowb : t.row,
              WHICH == 1 ? t.firstb : t.first, y, r);
}

// environment::eval of one facet: the specialised evaluator, or - in the general build - the one
// that copes with any channel count / texel stride / degree
template <int NCH, int TS, int DEG, bool GEN, bool I32 = false>
__device__ __forceinline__ int dev_eval_facet(const RenderParams& P, const FacetDev& F, const float r[3],
                                              float px[NCH]) {
  if constexpr (GEN) return dev_facet_eval_general<NCH>(F, P.degree, P.wmat, r, px);
  else return dev_facet_eval<NCH, TS, DEG, I32>(F, P.degree, P.wmat, r, px);
}

// tap k of the twining filter: from the parameter block (constant bank; kernels compiled for a job shape - the plan
// builder only picks them for filters of up to EU_INLINE_TAPS taps), else from global memory
template <bool INLINE>
__device__ __forceinline__ void dev_tap(const RenderParams& P, int k, float& cx, float& cy, float& cw) {
  if constexpr (INLINE) {
    cx = P.ptaps[3 * k]; cy = P.ptaps[3 * k + 1]; cw = P.ptaps[3 * k + 2];
  } else {
    cx = __ldg(P.taps + 3 * k); cy = __ldg(P.taps + 3 * k + 1); cw = __ldg(P.taps + 3 * k + 2);
  }
}

// One synopsis evaluation for rays produced by `ray_of(i, ray)`; returns the index-plane value.
//   single facet / _voronoi_syn (envutil_payload.cc:818-956) / _hdr_merge_syn (:1500-1622) /
//   _voronoi_syn_plus (:964-1233)
// `fa`: the facet array - the block's shared-memory copy (PF, up to EU_SMEM_FACETS facets: every
// lane reads the same field, one broadcast wavefront instead of a global load) or global memory.
// `active`: this lane renders a pixel. Only VORONOI_PLUS needs it: the reference takes a shortcut
// per 16-lane zimt vector there, which is voted on by the half-warp, so every lane of the warp
// must walk through the code (a warp is 32 consecutive pixels of one row = two zimt vectors).
template <int NCH, int TS, int MODE, int DEG, bool GEN, int SP, typename RayFn>
__device__ __forceinline__ int dev_synopsis(const RenderParams& P, const FacetDev& f0, const FacetDev* __restrict__ fa,
                                            RayFn ray_of, bool active, float px[NCH]) {
  if constexpr (MODE == EU_MODE_SINGLE) {
    float r[3];
    ray_of(0, r);
    return dev_eval_facet<NCH, TS, DEG, GEN, SP != 0>(P, f0, r, px);
  } else if constexpr (MODE == EU_MODE_VORONOI) {
    int champion = -1;
    float max_z = -FLT_MAX, best[3] = {0.f, 0.f, 0.f};
    for (int i = 0; i < P.n_facets; i++) {
      const FacetDev& F = dev_facet_at<SP>(fa, i);
      float r[3];
      if constexpr (!GEN) {
        // A rectilinear mount only sees rays with z > 0 (the first thing dev_facet_mask tests, environment.h:1123-1127):
        // z alone is computed before the branch - of this first evaluation only r[2] is live - and the facets behind the
        // camera (half of a full panorama's) cost a third of a ray. The second evaluation reuses z.
        if (F.projection == EU_RECTILINEAR && !F.mask_always) {
          float rz[3];
          ray_of(i, rz);
          if (!(rz[2] > 0.0f)) continue;
        }
      }
      ray_of(i, r);
      if (!dev_facet_mask(F, r)) continue;
      float cz = r[2] * F.recip_step;
      if (i == 0 || cz > max_z) {  // facet 0 sets max_z unconditionally (:836-841)
        max_z = cz;
        champion = i;
        best[0] = r[0]; best[1] = r[1]; best[2] = r[2];
      }
    }
    if (champion < 0) {
#pragma unroll
      for (int c = 0; c < NCH; c++) px[c] = 0.0f;
    } else {
      dev_eval_facet<NCH, TS, DEG, GEN, SP != 0>(P, dev_facet_at<SP>(fa, champion), best, px);
    }
    return champion;
  } else if constexpr (MODE == EU_MODE_HDR) {
    // with alpha the colour is de-associated for the weighted sum, alpha is the maximum seen and
    // the result is re-associated (:1527-1547,1597-1620)
    constexpr int NA = (NCH == 2 || NCH == 4) ? NCH - 1 : -1;
    constexpr int NCOL = NA >= 0 ? NA : NCH;
    float qsum = 0.0f, p[NCH];
#pragma unroll
    for (int c = 0; c < NCH; c++) px[c] = 0.0f;
    Located L;
    bool hit = false;
    for (int i = 0; i < P.n_facets; i++) {
      const FacetDev& F = dev_facet_at<SP>(fa, i);
      if constexpr (GEN) {
        float r[3];
        ray_of(i, r);
        dev_eval_facet<NCH, TS, DEG, GEN, SP != 0>(P, F, r, p);
      } else {
        // brackets of one camera position share their geometry: the ray and its window position are the
        // previous facet's (FacetDev::same_geom), only the container differs
        if (i == 0 || !F.same_geom) {
          float r[3];
          ray_of(i, r);
          hit = dev_facet_locate<DEG>(F, P.degree, r, L);
        }
        if (hit) {
          dev_facet_window<NCH, TS, DEG, SP != 0>(F, P.degree, P.wmat, L, p);
        } else {
#pragma unroll
          for (int c = 0; c < NCH; c++) p[c] = 0.0f;
        }
      }
      float grey = p[0];
      if constexpr (NCH >= 3) grey = fmaxf(p[0], fmaxf(p[1], p[2]));
      float q = dev_hdr_quality(grey, F.hdr_optimum, F.hdr_kind);
      if constexpr (NA >= 0) q = p[NA] * q;
      qsum += q;
      if constexpr (NA < 0) {
#pragma unroll
        for (int c = 0; c < NCH; c++) px[c] += p[c] * q;
      } else {
#pragma unroll
        for (int c = 0; c < NCOL; c++) {
          float v = 0.0f;
          if (p[NA] > 0.000001f) v = p[c] / p[NA];
          px[c] += v * q;
        }
        px[NA] = fmaxf(px[NA], p[NA]);
      }
    }
#pragma unroll
    for (int c = 0; c < NCOL; c++) {
      px[c] /= qsum;
      if (!(qsum > 0.0f)) px[c] = 0.0f;
      if constexpr (NA >= 0) px[c] *= px[NA];
    }
    return -1;
  } else {
    // _voronoi_syn_plus: facets the ray hits, sorted by z * recip_step (stable, strictly greater
    // moves up), composited front to back as associated alpha
    const unsigned lane = threadIdx.x & 31u;
    const unsigned gmask = (lane < 16u) ? 0x0000ffffu : 0xffff0000u;
    float zs[EU_MAX_FACETS];
    int ids[EU_MAX_FACETS];
    int cnt = 0, next_best = -1;  // next_best: the last facet any lane of this zimt vector hit
    for (int i = 0; i < P.n_facets; i++) {
      const FacetDev& F = dev_facet_at<SP>(fa, i);
      float r[3];
      ray_of(i, r);
      const bool valid = active && dev_facet_mask(F, r);
      if (__ballot_sync(0xffffffffu, valid) & gmask) next_best = i;
      if (valid) {
        float z = r[2] * F.recip_step;
        int k = cnt++;
        while (k > 0 && z > zs[k - 1]) {
          zs[k] = zs[k - 1];
          ids[k] = ids[k - 1];
          k--;
        }
        zs[k] = z;
        ids[k] = i;
      }
    }
#pragma unroll
    for (int c = 0; c < NCH; c++) px[c] = 0.0f;
    const int top = cnt ? ids[0] : -1;
    float help[NCH];
    bool have_top = false, done = false;
    // shortcut (:1132-1150): all lanes of the vector have next_best in front and are opaque there
    const unsigned not_top = __ballot_sync(0xffffffffu, active && top != next_best) & gmask;
    const bool try_shortcut = next_best >= 0 && not_top == 0u;
    bool opaque = true;
    if (try_shortcut && active) {
      float r[3];
      ray_of(top, r);
      dev_eval_facet<NCH, TS, DEG, GEN, SP != 0>(P, dev_facet_at<SP>(fa, top), r, help);
      have_top = true;
      opaque = help[NCH - 1] >= 1.0f;
    }
    const unsigned not_opaque = __ballot_sync(0xffffffffu, active && !opaque) & gmask;
    if (try_shortcut && not_opaque == 0u) {
      done = true;
      if (active) {
#pragma unroll
        for (int c = 0; c < NCH; c++) px[c] = help[c];
      }
    }
    if (!done) {
      for (int k = 0; k < cnt; k++) {
        if (!(k == 0 && have_top)) {
          float r[3];
          ray_of(ids[k], r);
          dev_eval_facet<NCH, TS, DEG, GEN, SP != 0>(P, dev_facet_at<SP>(fa, ids[k]), r, help);
        }
        if (k == 0) {
#pragma unroll
          for (int c = 0; c < NCH; c++) px[c] = help[c];
        } else {
#pragma unroll
          for (int c = 0; c < NCH; c++) px[c] += (1.0f - px[NCH - 1]) * help[c];
        }
      }
    }
    return top;
  }
}

// GEN: some facet of the job uses the generic stepper (PanoTools translation); kept out of the
// common instantiations because the extra per-facet branch costs ~20 % on multi-facet jobs
// SP: the job shape the kernel is compiled for (plan.h: eu_render_specs; 0 = any)
// Kernels compiled for a job shape are asked to fit six blocks per SM (<= 40 registers): the hdr_merge and voronoi
// shapes are latency-bound at the four to five blocks their 54 / 42 registers allowed (ncu: issue-active 62 %).
template <int NCH, int TS, int MODE, bool TWINE, int DEG, bool PF, bool GEN, int SP = 0>
__global__ void __launch_bounds__(TILE_X* TILE_Y, (SP != 0 && MODE != EU_MODE_SINGLE) ? 6 : 1) k_render(const __grid_constant__ RenderParams P) {
  SpecView<SP> V(P);
  const TargetDev& T = V.trg();
  const FacetDev& f0 = V.f0();
  const FacetDev* __restrict__ fa = P.facets;
  __shared__ __align__(16) float wslot[TILE_Y][NCH == 3 ? TILE_X * 3 : 4];  // dev_store_pixel
  if constexpr (PF && MODE != EU_MODE_SINGLE) {
    __shared__ __align__(16) unsigned char sfa[EU_SMEM_FACETS * sizeof(FacetDev)];
    static_assert(sizeof(FacetDev) % 4 == 0, "FacetDev is copied word by word");
    const int words = P.n_facets * (int)(sizeof(FacetDev) / 4);
    const uint32_t* src = reinterpret_cast<const uint32_t*>(P.facets);
    uint32_t* dst = reinterpret_cast<uint32_t*>(sfa);
    for (int i = threadIdx.y * TILE_X + threadIdx.x; i < words; i += TILE_X * TILE_Y) dst[i] = __ldg(src + i);
    __syncthreads();
    fa = reinterpret_cast<const FacetDev*>(sfa);
  }
  int x = P.col0 + blockIdx.x * TILE_X + threadIdx.x;
  int y = P.row0 + blockIdx.y * TILE_Y + threadIdx.y;
  const bool active = x < P.col1 && y < P.row1;
  if constexpr (MODE != EU_MODE_VORONOI_PLUS) {
    if (!active) return;
  } else {  // all lanes stay for the half-warp votes; idle ones compute on a pixel that exists
    if (y >= P.row1) return;  // whole warp (a warp is one row of the tile)
    x = min(x, P.col1 - 1);
  }
  int xf = first_lane_column(x);
  PixelTerms t;
  {
    float2 c0 = __ldg(P.col_tab + x), r0 = __ldg(P.row_tab + y);
    t.col = ColTerm{c0.x, c0.y};
    t.row = RowTerm{r0.x, r0.y};
    t.first = t.col;
    if (T.projection == EU_CYLINDRICAL && T.normalize) {
      float2 f0 = __ldg(P.col_tab + xf);
      t.first = ColTerm{f0.x, f0.y};
    }
    t.px = t.py = t.pxb = t.pyb = 0.0f;
    if constexpr (GEN) {
      t.px = __ldg(P.planar_raw + x);
      t.py = __ldg(P.planar_raw + 2 * T.width + y);
    }
    t.colb = t.col; t.firstb = t.first; t.rowb = t.row;
    if constexpr (TWINE) {
      float2 c1 = __ldg(P.col_tab + T.width + x), r1 = __ldg(P.row_tab + T.height + y);
      t.colb = ColTerm{c1.x, c1.y};
      t.rowb = RowTerm{r1.x, r1.y};
      t.firstb = t.colb;
      if (T.projection == EU_CYLINDRICAL && T.normalize) {
        float2 f1 = __ldg(P.col_tab + T.width + xf);
        t.firstb = ColTerm{f1.x, f1.y};
      }
      if constexpr (GEN) {
        t.pxb = __ldg(P.planar_raw + T.width + x);
        t.pyb = __ldg(P.planar_raw + 2 * T.width + T.height + y);
      }
    }
  }
  float px[NCH];
  int idx;
  if constexpr (!TWINE) {
    auto ray_of = [&](int i, float r[3]) {
      if constexpr (MODE == EU_MODE_SINGLE) dev_facet_ray<GEN, 0>(T, P.inv, f0, t, y, r, P.cube_tab);
      else dev_facet_ray<GEN, 0>(T, P.inv, dev_facet_at<SP>(fa, i), t, y, r);
    };
    idx = dev_synopsis<NCH, TS, MODE, DEG, GEN, SP>(P, f0, fa, ray_of, active, px);
  } else {
    // deriv_stepper (stepper.h:1606-1694) + twine_t (twining.h:106-263) /
    // synopsis_t (envutil_payload.cc:647-690)
    float acc[NCH], help[NCH];
#pragma unroll
    for (int c = 0; c < NCH; c++) acc[c] = 0.0f;
    idx = -1;
    if constexpr (MODE == EU_MODE_SINGLE) {
      float r00[3], du[3], dv[3];
      dev_facet_ray<GEN, 0>(T, P.inv, f0, t, y, r00, P.cube_tab);
      dev_facet_ray<GEN, 1>(T, P.inv, f0, t, y, du, P.cube_tab);
      dev_facet_ray<GEN, 2>(T, P.inv, f0, t, y, dv, P.cube_tab);
#pragma unroll
      for (int c = 0; c < 3; c++) {
        du[c] = du[c] - r00[c];
        dv[c] = dv[c] - r00[c];
      }
      for (int k = 0; k < P.n_taps; k++) {
        float cx, cy, cw;
        dev_tap<SP != 0>(P, k, cx, cy, cw);
        float r[3];
#pragma unroll
        for (int c = 0; c < 3; c++) r[c] = r00[c] + cx * du[c] + cy * dv[c];
        int id = dev_eval_facet<NCH, TS, DEG, GEN, SP != 0>(P, f0, r, help);
        if (k == 0) idx = id;
#pragma unroll
        for (int c = 0; c < NCH; c++) acc[c] = EU_WIN_MULADD(cw, help[c], acc[c]);
      }
    } else {
      // per-facet ninepacks live in local memory; the taps loop re-reads them
      float np[EU_MAX_FACETS][9];
      for (int i = 0; i < P.n_facets; i++) {
        const FacetDev& F = dev_facet_at<SP>(fa, i);
        float r00[3], r10[3], r01[3];
        dev_facet_ray<GEN, 0>(T, P.inv, F, t, y, r00);
        dev_facet_ray<GEN, 1>(T, P.inv, F, t, y, r10);
        dev_facet_ray<GEN, 2>(T, P.inv, F, t, y, r01);
#pragma unroll
        for (int c = 0; c < 3; c++) {
          np[i][c] = r00[c];
          np[i][3 + c] = r10[c] - r00[c];
          np[i][6 + c] = r01[c] - r00[c];
        }
      }
      for (int k = 0; k < P.n_taps; k++) {
        float cx, cy, cw;
        dev_tap<SP != 0>(P, k, cx, cy, cw);
        auto ray_of = [&](int i, float r[3]) {
#pragma unroll
          for (int c = 0; c < 3; c++) r[c] = np[i][c] + cx * np[i][3 + c] + cy * np[i][6 + c];
        };
        int id = dev_synopsis<NCH, TS, MODE, DEG, GEN, SP>(P, f0, fa, ray_of, active, help);
        if (k == 0) idx = id;
#pragma unroll
        for (int c = 0; c < NCH; c++) acc[c] = EU_WIN_MULADD(cw, help[c], acc[c]);
      }
    }
#pragma unroll
    for (int c = 0; c < NCH; c++) px[c] = acc[c];
  }
  if (!active) return;
  if (T.unbrighten != 1.0f) {  // amplify_type after everything else (envutil_payload.cc:500-511)
    constexpr int NCOL = (NCH == 2 || NCH == 4) ? NCH - 1 : NCH;
#pragma unroll
    for (int c = 0; c < NCOL; c++) px[c] *= T.unbrighten;
  }
  size_t o = (size_t)(y - P.row0) * T.width + x;
  if (P.out) dev_store_pixel<NCH>(P, T, x, y, px, wslot[threadIdx.y]);
  if (P.index_out) P.index_out[o] = idx;
}

// ------------------------------------------------------------------------------------------
// Single-facet render with the block's gather footprint staged in shared memory.
//
// The direct kernel above is bound by the L1 data pipe: a quarter-warp 128-bit gather whose
// texels straddle a 128-byte line costs two wavefronts, and every tap of every pixel goes
// through it. Here a block first locates all its pixels, reduces the bounding box of their
// windows, and pulls exactly those container rows into shared memory with one bulk asynchronous
// copy per row (cp.async.bulk -> mbarrier complete_tx, the TMA engine: no registers, no LSU
// instructions, 16-byte granules). The windows are then read from shared memory, where an
// unaligned run of consecutive texels is conflict-free. Blocks whose footprint does not fit
// (cube-face seams, the +-pi seam, poles) take the direct path; with twining every tap checks
// its own window against the staged box and falls back to HBM individually. Values and their
// order of combination are untouched: the result is bit-identical to the direct kernel.
// ------------------------------------------------------------------------------------------
#define EU_TILE_FLOATS 6144  // 24 KB staged footprint per block
// shared-memory row pitch of the staged tile: the copied width, stepped off multiples of the 32 banks (with such a
// pitch lanes in the same column of different rows collide: 2.3 instead of 1.8 wavefronts per load, simulated and measured)
#define EU_TILE_PITCH(w) (((w) & 31) ? (w) : (w) + 4)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// one row of the footprint: global -> shared, completion counted in bytes on the mbarrier
__device__ __forceinline__ void bulk_row_g2s(float* dst, const float* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <int NCH, int TS, bool TWINE, int DEG, int SP = 0>
__global__ void __launch_bounds__(TILE_X* TILE_Y) k_render_tiled(const __grid_constant__ RenderParams P) {
  static_assert(DEG == 1 || DEG == 3, "tile path is built for the bilinear and cubic evaluators");
  constexpr int ORDER = DEG + 1, H2 = DEG / 2;
  constexpr int NWARP = TILE_X * TILE_Y / 32;
  __shared__ __align__(128) float tile[EU_TILE_FLOATS];
  __shared__ int red[NWARP][4];
  __shared__ __align__(16) float wslot[TILE_Y][NCH == 3 ? TILE_X * 3 : 4];  // dev_store_pixel
  __shared__ int box[4];  // A0 (float offset in the container row), first container row, floats per row, rows
  __shared__ __align__(8) uint64_t mbar;

  SpecView<SP> V(P);
  const TargetDev& T = V.trg();
  const FacetDev& F = V.f0();
  const SourceDev& S = F.src;
  const int tid = threadIdx.y * TILE_X + threadIdx.x;
  const int x = P.col0 + blockIdx.x * TILE_X + threadIdx.x;
  const int y = P.row0 + blockIdx.y * TILE_Y + threadIdx.y;
  const bool inside = x < P.col1 && y < P.row1;
  if (tid == 0) mbar_init(&mbar, 1);

  // ---- phase 1: rays and window origins ------------------------------------------------
  const int xc = inside ? x : 0, yc = inside ? y : P.row0;
  const int xf = first_lane_column(xc);
  float2 c0 = __ldg(P.col_tab + xc), r0 = __ldg(P.row_tab + yc);
  ColTerm col{c0.x, c0.y};
  RowTerm row{r0.x, r0.y};
  ColTerm first = col;
  if (T.projection == EU_CYLINDRICAL && T.normalize) {
    float2 f0 = __ldg(P.col_tab + xf);
    first = ColTerm{f0.x, f0.y};
  }
  float r00[3];
  if (T.projection >= EU_CUBEMAP) dev_stepper_cube_tab(T, P.cube_tab, col, row, r00);
  else dev_stepper(T, F.xx, F.yy, F.zz, col, row, first, yc, r00);
  int face;
  float cx, cy;
  bool hit = dev_facet_coordinate(F, r00, face, cx, cy) && inside;
  Located L = dev_locate(S, DEG, hit ? cx : 0.0f, hit ? cy : 0.0f);
  // window origin in CONTAINER texel coordinates (container rows start 16-byte aligned)
  const int lox = L.ix - H2 + P.src_lx, loy = L.iy - H2 + P.src_ly;
  {
    int mnx = hit ? lox : INT_MAX, mxx = hit ? lox : INT_MIN, mny = hit ? loy : INT_MAX, mxy = hit ? loy : INT_MIN;
    mnx = __reduce_min_sync(0xffffffffu, mnx);
    mxx = __reduce_max_sync(0xffffffffu, mxx);
    mny = __reduce_min_sync(0xffffffffu, mny);
    mxy = __reduce_max_sync(0xffffffffu, mxy);
    if ((tid & 31) == 0) {
      red[tid >> 5][0] = mnx; red[tid >> 5][1] = mxx; red[tid >> 5][2] = mny; red[tid >> 5][3] = mxy;
    }
  }
  __syncthreads();
  if (tid < 32) {
    int l = tid < NWARP ? tid : 0;
    int mnx = __reduce_min_sync(0xffffffffu, red[l][0]);
    int mxx = __reduce_max_sync(0xffffffffu, red[l][1]);
    int mny = __reduce_min_sync(0xffffffffu, red[l][2]);
    int mxy = __reduce_max_sync(0xffffffffu, red[l][3]);
    if (tid == 0) {
      int rows = 0, a0 = 0, wf = 0;
      if (mnx <= mxx) {  // at least one pixel of the block hits the source
        if constexpr (TWINE) {  // sub-rays stray up to half a pixel from the centre ray
          int bw = mxx - mnx + 1, bh = mxy - mny + 1;
          int mx = (bw * 5) / 64 + 2, my = (bh * 5) / 64 + 2;
          mnx -= mx; mxx += mx; mny -= my; mxy += my;
          // keep the box inside the container
          mnx = max(mnx, 0); mny = max(mny, 0);
          mxx = min(mxx, P.src_cw - ORDER); mxy = min(mxy, P.src_ch - ORDER);
        }
        a0 = (mnx * TS) & ~3;  // 16-byte granule within the container row
        wf = (((mxx + ORDER) * TS - a0) + 3) & ~3;
        rows = mxy - mny + ORDER;
        if (rows > TILE_X * TILE_Y || rows * EU_TILE_PITCH(wf) > EU_TILE_FLOATS || wf <= 0 || rows <= 0) rows = 0;
      }
      box[0] = a0; box[1] = mny; box[2] = wf; box[3] = rows;
      if (rows > 0) mbar_expect_tx(&mbar, (uint32_t)(rows * wf) * 4u);
    }
  }
  __syncthreads();
  const int a0 = box[0], by0 = box[1], wcopy = box[2], rows = box[3];
  const int wf = EU_TILE_PITCH(wcopy);  // row pitch in shared memory
  const bool staged = rows > 0;
  if (staged) {
    // rows are dealt round-robin to the warps (the copy is issued from the uniform datapath, so
    // the lanes of one warp take turns)
    const int rid = (tid & 31) * NWARP + (tid >> 5);
    if (rid < rows)
      bulk_row_g2s(tile + rid * wf, P.src_base + (ptrdiff_t)(by0 + rid) * S.stride + a0, (uint32_t)wcopy * 4u, &mbar);
    mbar_wait(&mbar, 0);
  }

  // ---- phase 2: windows ------------------------------------------------------------------
  if (!inside) return;
  float px[NCH];
  if constexpr (!TWINE) {
    if (!hit) {
#pragma unroll
      for (int c = 0; c < NCH; c++) px[c] = 0.0f;
    } else {
      if (staged)
        dev_window_eval<NCH, TS, DEG, true, SP != 0>(tile + (loy - by0) * wf + (lox * TS - a0), wf, DEG, P.wmat, L.fx, L.fy, px);
      else
        dev_window_eval<NCH, TS, DEG, false, SP != 0>(P.src_base + (ptrdiff_t)loy * S.stride + (ptrdiff_t)lox * TS, S.stride, DEG,
                                             P.wmat, L.fx, L.fy, px);
      dev_brighten<NCH>(F, px);
    }
  } else {
    // deriv_stepper (stepper.h:1606-1694) + twine_t (twining.h:106-263)
    float2 c1 = __ldg(P.col_tab + T.width + x), r1 = __ldg(P.row_tab + T.height + y);
    ColTerm colb{c1.x, c1.y};
    RowTerm rowb{r1.x, r1.y};
    ColTerm firstb = colb;
    if (T.projection == EU_CYLINDRICAL && T.normalize) {
      float2 f1 = __ldg(P.col_tab + T.width + xf);
      firstb = ColTerm{f1.x, f1.y};
    }
    float du[3], dv[3], help[NCH];
    if (T.projection >= EU_CUBEMAP) {
      dev_stepper_cube_tab(T, P.cube_tab, colb, row, du);
      dev_stepper_cube_tab(T, P.cube_tab, col, rowb, dv);
    } else {
      dev_stepper(T, F.xx, F.yy, F.zz, colb, row, firstb, y, du);
      dev_stepper(T, F.xx, F.yy, F.zz, col, rowb, first, y, dv);
    }
#pragma unroll
    for (int c = 0; c < 3; c++) {
      du[c] = du[c] - r00[c];
      dv[c] = dv[c] - r00[c];
    }
#pragma unroll
    for (int c = 0; c < NCH; c++) px[c] = 0.0f;
    const int bx1 = a0 + wcopy, by1 = by0 + rows;
    for (int k = 0; k < P.n_taps; k++) {
      float tx, ty, tw;
      dev_tap<SP != 0>(P, k, tx, ty, tw);
      float r[3];
#pragma unroll
      for (int c = 0; c < 3; c++) r[c] = r00[c] + tx * du[c] + ty * dv[c];
      int fc;
      float sx, sy;
      if (!dev_facet_coordinate(F, r, fc, sx, sy)) {
#pragma unroll
        for (int c = 0; c < NCH; c++) help[c] = 0.0f;
      } else {
        Located K = dev_locate(S, DEG, sx, sy);
        const int kx = K.ix - H2 + P.src_lx, ky = K.iy - H2 + P.src_ly;
        const bool in_box = staged && kx * TS >= a0 && (kx + ORDER) * TS <= bx1 && ky >= by0 && ky + ORDER <= by1;
        if (in_box)
          dev_window_eval<NCH, TS, DEG, true, SP != 0>(tile + (ky - by0) * wf + (kx * TS - a0), wf, DEG, P.wmat, K.fx, K.fy, help);
        else
          dev_window_eval<NCH, TS, DEG, false, SP != 0>(P.src_base + (ptrdiff_t)ky * S.stride + (ptrdiff_t)kx * TS, S.stride, DEG,
                                               P.wmat, K.fx, K.fy, help);
        dev_brighten<NCH>(F, help);
      }
#pragma unroll
      for (int c = 0; c < NCH; c++) px[c] = EU_WIN_MULADD(tw, help[c], px[c]);
    }
  }
  if (T.unbrighten != 1.0f) {
    constexpr int NCOL = (NCH == 2 || NCH == 4) ? NCH - 1 : NCH;
#pragma unroll
    for (int c = 0; c < NCOL; c++) px[c] *= T.unbrighten;
  }
  dev_store_pixel<NCH>(P, T, x, y, px, wslot[threadIdx.y]);
}

template <int NCH, int TS, int MODE, bool TWINE>
static void launch_deg(const RenderParams& P, dim3 grid, dim3 block, cudaStream_t st) {
  if constexpr (MODE == EU_MODE_SINGLE) {
    // footprint-staged kernel: needs 16-byte row granules and a pixel output (no index plane).
    // Measured on B200 (profiles/): it wins for the cubic window (16 taps/px: C2 0.60 vs 0.75 ms)
    // and loses for the bilinear one (4 taps/px: C3b 1.56 vs 1.32 ms), where the two block-wide
    // synchronisations cost more than the gathers they replace - so it is used for degree 3 only.
    if (P.use_tiles && P.out && !P.index_out && (P.f0.src.stride & 3) == 0 && !P.any_generic && P.degree == 3) {
      k_render_tiled<NCH, TS, TWINE, 3><<<grid, block, 0, st>>>(P);
      return;
    }
  }
  if (P.any_generic) {  // translation: the generic-stepper build (run-time degree, facets in global memory)
    k_render<NCH, TS, MODE, TWINE, -1, false, true><<<grid, block, 0, st>>>(P);
    return;
  }
  if constexpr (MODE != EU_MODE_SINGLE) {
    if (P.n_facets <= EU_SMEM_FACETS) {
      switch (P.degree) {
        case 1: k_render<NCH, TS, MODE, TWINE, 1, true, false><<<grid, block, 0, st>>>(P); break;
        case 3: k_render<NCH, TS, MODE, TWINE, 3, true, false><<<grid, block, 0, st>>>(P); break;
        default: k_render<NCH, TS, MODE, TWINE, -1, true, false><<<grid, block, 0, st>>>(P); break;
      }
      return;
    }
  }
  switch (P.degree) {
    case 1: k_render<NCH, TS, MODE, TWINE, 1, false, false><<<grid, block, 0, st>>>(P); break;
    case 3: k_render<NCH, TS, MODE, TWINE, 3, false, false><<<grid, block, 0, st>>>(P); break;
    default: k_render<NCH, TS, MODE, TWINE, -1, false, false><<<grid, block, 0, st>>>(P); break;
  }
}

template <int NCH, int TS>
static cudaError_t launch_render(const RenderParams& P, cudaStream_t st) {
  dim3 block(TILE_X, TILE_Y);
  dim3 grid((P.col1 - P.col0 + TILE_X - 1) / TILE_X, (P.row1 - P.row0 + TILE_Y - 1) / TILE_Y);
  bool tw = P.n_taps > 0;
  constexpr bool ALPHA = (NCH == 2 || NCH == 4);
  switch (P.mode) {
    case EU_MODE_SINGLE:
      if (tw) launch_deg<NCH, TS, EU_MODE_SINGLE, true>(P, grid, block, st);
      else launch_deg<NCH, TS, EU_MODE_SINGLE, false>(P, grid, block, st);
      break;
    case EU_MODE_VORONOI:
      if constexpr (!ALPHA) {
        if (tw) launch_deg<NCH, TS, EU_MODE_VORONOI, true>(P, grid, block, st);
        else launch_deg<NCH, TS, EU_MODE_VORONOI, false>(P, grid, block, st);
      } else {
        return cudaErrorInvalidValue;
      }
      break;
    case EU_MODE_VORONOI_PLUS:
      if constexpr (ALPHA) {
        if (tw) launch_deg<NCH, TS, EU_MODE_VORONOI_PLUS, true>(P, grid, block, st);
        else launch_deg<NCH, TS, EU_MODE_VORONOI_PLUS, false>(P, grid, block, st);
      } else {
        return cudaErrorInvalidValue;
      }
      break;
    default:
      if (tw) launch_deg<NCH, TS, EU_MODE_HDR, true>(P, grid, block, st);
      else launch_deg<NCH, TS, EU_MODE_HDR, false>(P, grid, block, st);
  }
  return cudaGetLastError();
}

}  // namespace EU_ARITH_NS
using namespace EU_ARITH_NS;
