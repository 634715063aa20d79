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
#include "eu_device.cuh"
#include "kernels.h"

#define TILE_X 32
#define TILE_Y 8

__device__ __forceinline__ int first_lane_column(int x) {
  int seg0 = (x / EU_SEGMENT) * EU_SEGMENT;
  return seg0 + (x - seg0) % EU_LANES;
}

// one synopsis evaluation (envutil_payload.cc:818-956 voronoi, :1500-1622 hdr_merge) for rays
// produced by `ray_of(i, ray)`; returns the index-plane value
template <int NCH, int TS, int MODE, int DEG, typename RayFn>
__device__ __forceinline__ int dev_synopsis(const RenderParams& P, RayFn ray_of, float px[NCH]) {
  if constexpr (MODE == EU_MODE_SINGLE) {
    float r[3];
    ray_of(0, r);
    return dev_facet_eval<NCH, TS, DEG>(P.f0, P.degree, P.wmat, r, px);
  } else if constexpr (MODE == EU_MODE_VORONOI) {
    int champion = -1;
    float max_z = -FLT_MAX, best[3] = {0.f, 0.f, 0.f};
    for (int i = 0; i < P.n_facets; i++) {
      const FacetDev& F = P.facets[i];
      float r[3];
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
      dev_facet_eval<NCH, TS, DEG>(P.facets[champion], P.degree, P.wmat, best, px);
    }
    return champion;
  } else {
    float qsum = 0.0f, p[NCH];
#pragma unroll
    for (int c = 0; c < NCH; c++) px[c] = 0.0f;
    for (int i = 0; i < P.n_facets; i++) {
      const FacetDev& F = P.facets[i];
      float r[3];
      ray_of(i, r);
      dev_facet_eval<NCH, TS, DEG>(F, P.degree, P.wmat, r, p);
      float grey = p[0];
      if constexpr (NCH >= 3) grey = fmaxf(p[0], fmaxf(p[1], p[2]));
      float q = dev_hdr_quality(grey, F.hdr_optimum, F.hdr_kind);
      qsum += q;
#pragma unroll
      for (int c = 0; c < NCH; c++) px[c] += p[c] * q;
    }
#pragma unroll
    for (int c = 0; c < NCH; c++) {
      px[c] /= qsum;
      if (!(qsum > 0.0f)) px[c] = 0.0f;
    }
    return -1;
  }
}

template <int NCH, int TS, int MODE, bool TWINE, int DEG>
__global__ void __launch_bounds__(TILE_X* TILE_Y) k_render(const __grid_constant__ RenderParams P) {
  const TargetDev& T = P.trg;
  int x = blockIdx.x * TILE_X + threadIdx.x;
  int y = P.row0 + blockIdx.y * TILE_Y + threadIdx.y;
  if (x >= T.width || y >= P.row1) return;
  int xf = first_lane_column(x);
  float2 c0 = __ldg(P.col_tab + x), r0 = __ldg(P.row_tab + y);
  ColTerm col{c0.x, c0.y};
  RowTerm row{r0.x, r0.y};
  ColTerm first = col;
  if (T.projection == EU_CYLINDRICAL && T.normalize) {
    float2 f0 = __ldg(P.col_tab + xf);
    first = ColTerm{f0.x, f0.y};
  }
  float px[NCH];
  int idx;
  if constexpr (!TWINE) {
    auto ray_of = [&](int i, float r[3]) {
      const FacetDev& F = MODE == EU_MODE_SINGLE ? P.f0 : P.facets[i];
      dev_stepper(T, F.xx, F.yy, F.zz, col, row, first, y, r);
    };
    idx = dev_synopsis<NCH, TS, MODE, DEG>(P, ray_of, px);
  } else {
    // deriv_stepper (stepper.h:1606-1694) + twine_t (twining.h:106-263) /
    // synopsis_t (envutil_payload.cc:647-690)
    float2 c1 = __ldg(P.col_tab + T.width + x), r1 = __ldg(P.row_tab + T.height + y);
    ColTerm colb{c1.x, c1.y};
    RowTerm rowb{r1.x, r1.y};
    ColTerm firstb = colb;
    if (T.projection == EU_CYLINDRICAL && T.normalize) {
      float2 f1 = __ldg(P.col_tab + T.width + xf);
      firstb = ColTerm{f1.x, f1.y};
    }
    float acc[NCH], help[NCH];
#pragma unroll
    for (int c = 0; c < NCH; c++) acc[c] = 0.0f;
    idx = -1;
    if constexpr (MODE == EU_MODE_SINGLE) {
      float r00[3], du[3], dv[3];
      dev_stepper(T, P.f0.xx, P.f0.yy, P.f0.zz, col, row, first, y, r00);
      dev_stepper(T, P.f0.xx, P.f0.yy, P.f0.zz, colb, row, firstb, y, du);
      dev_stepper(T, P.f0.xx, P.f0.yy, P.f0.zz, col, rowb, first, y, dv);
#pragma unroll
      for (int c = 0; c < 3; c++) {
        du[c] = du[c] - r00[c];
        dv[c] = dv[c] - r00[c];
      }
      for (int k = 0; k < P.n_taps; k++) {
        float cx = __ldg(P.taps + 3 * k), cy = __ldg(P.taps + 3 * k + 1), cw = __ldg(P.taps + 3 * k + 2);
        float r[3];
#pragma unroll
        for (int c = 0; c < 3; c++) r[c] = r00[c] + cx * du[c] + cy * dv[c];
        int id = dev_facet_eval<NCH, TS, DEG>(P.f0, P.degree, P.wmat, r, help);
        if (k == 0) idx = id;
#pragma unroll
        for (int c = 0; c < NCH; c++) acc[c] += cw * help[c];
      }
    } else {
      // per-facet ninepacks live in local memory; the taps loop re-reads them
      float np[EU_MAX_FACETS][9];
      for (int i = 0; i < P.n_facets; i++) {
        const FacetDev& F = P.facets[i];
        float r00[3], r10[3], r01[3];
        dev_stepper(T, F.xx, F.yy, F.zz, col, row, first, y, r00);
        dev_stepper(T, F.xx, F.yy, F.zz, colb, row, firstb, y, r10);
        dev_stepper(T, F.xx, F.yy, F.zz, col, rowb, first, y, r01);
#pragma unroll
        for (int c = 0; c < 3; c++) {
          np[i][c] = r00[c];
          np[i][3 + c] = r10[c] - r00[c];
          np[i][6 + c] = r01[c] - r00[c];
        }
      }
      for (int k = 0; k < P.n_taps; k++) {
        float cx = __ldg(P.taps + 3 * k), cy = __ldg(P.taps + 3 * k + 1), cw = __ldg(P.taps + 3 * k + 2);
        auto ray_of = [&](int i, float r[3]) {
#pragma unroll
          for (int c = 0; c < 3; c++) r[c] = np[i][c] + cx * np[i][3 + c] + cy * np[i][6 + c];
        };
        int id = dev_synopsis<NCH, TS, MODE, DEG>(P, ray_of, help);
        if (k == 0) idx = id;
#pragma unroll
        for (int c = 0; c < NCH; c++) acc[c] += cw * help[c];
      }
    }
#pragma unroll
    for (int c = 0; c < NCH; c++) px[c] = acc[c];
  }
  size_t o = (size_t)(y - P.row0) * T.width + x;
  if (P.out) {
    float* dst = P.out + o * NCH;
    if constexpr (NCH == 4) {
      *reinterpret_cast<float4*>(dst) = make_float4(px[0], px[1], px[2], px[3]);
    } else {
#pragma unroll
      for (int c = 0; c < NCH; c++) dst[c] = px[c];
    }
  }
  if (P.index_out) P.index_out[o] = idx;
}

template <int NCH, int TS, int MODE, bool TWINE>
static void launch_deg(const RenderParams& P, dim3 grid, dim3 block, cudaStream_t st) {
  switch (P.degree) {
    case 1: k_render<NCH, TS, MODE, TWINE, 1><<<grid, block, 0, st>>>(P); break;
    case 3: k_render<NCH, TS, MODE, TWINE, 3><<<grid, block, 0, st>>>(P); break;
    default: k_render<NCH, TS, MODE, TWINE, -1><<<grid, block, 0, st>>>(P); break;
  }
}

template <int NCH, int TS>
static cudaError_t launch_render(const RenderParams& P, cudaStream_t st) {
  dim3 block(TILE_X, TILE_Y);
  dim3 grid((P.trg.width + TILE_X - 1) / TILE_X, (P.row1 - P.row0 + TILE_Y - 1) / TILE_Y);
  bool tw = P.n_taps > 0;
  switch (P.mode) {
    case EU_MODE_SINGLE:
      if (tw) launch_deg<NCH, TS, EU_MODE_SINGLE, true>(P, grid, block, st);
      else launch_deg<NCH, TS, EU_MODE_SINGLE, false>(P, grid, block, st);
      break;
    case EU_MODE_VORONOI:
      if (tw) launch_deg<NCH, TS, EU_MODE_VORONOI, true>(P, grid, block, st);
      else launch_deg<NCH, TS, EU_MODE_VORONOI, false>(P, grid, block, st);
      break;
    default:
      if (tw) launch_deg<NCH, TS, EU_MODE_HDR, true>(P, grid, block, st);
      else launch_deg<NCH, TS, EU_MODE_HDR, false>(P, grid, block, st);
  }
  return cudaGetLastError();
}
