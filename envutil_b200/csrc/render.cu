// render.cu - the fused render kernel: target ray -> (rotation folded into the stepper basis)
// -> source coordinate -> gates -> b-spline window -> twining / synopsis -> brighten -> store.
// One thread per target pixel, 32x8-pixel tiles so that the threads of a block gather from a
// compact source footprint; no intermediate ray or coordinate buffer ever touches HBM.
//
// Replaces zimt::process + get_t/act_t/put_t of the reference (envutil_payload.cc:425-579,
// zimt/wielding.h:155-463) for all stepper x source x synopsis combinations of `fuse`
// (envutil_payload.cc:1885-2284).
#include "eu_device.cuh"
#include "kernels.h"

#define TILE_X 32
#define TILE_Y 8

// planar coordinates of every column / row, exactly as stepper_base produces them when driven
// by zimt::process: init at the start of each 512-px segment, += delta per 16-px vector
// (stepper.h:324-350, zimt/wielding.h:317-455). out_x[0..W) plain, [W..2W) x-biased stepper;
// out_y[0..H) plain, [H..2H) y-biased stepper.
__global__ void k_planar_tables(TargetDev T, float* __restrict__ out_x, float* __restrict__ out_y) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 2 * T.width) {
    int x = i % T.width;
    float bias = i < T.width ? 0.0f : T.bias_x;
    int seg0 = (x / EU_SEGMENT) * EU_SEGMENT;
    int r = x - seg0, lane = r % EU_LANES, v = r / EU_LANES;
    float ll0 = (float)(2 * lane) + (float)(seg0 * 2 + 1);
    float p = bias + ll0 * T.fx1 + ((float)(2 * T.width) - ll0) * T.fx0;
    for (int k = 0; k < v; k++) p += T.delta;
    out_x[i] = p;
  }
  int j = i - 2 * T.width;
  if (j >= 0 && j < 2 * T.height) {
    int y = j % T.height;
    float bias = j < T.height ? 0.0f : T.bias_y;
    int ll1 = y * 2 + 1;
    out_y[j] = bias + ll1 * T.fy1 + (float)(2 * T.height - ll1) * T.fy0;
  }
}

__device__ __forceinline__ int first_lane_column(int x) {
  int seg0 = (x / EU_SEGMENT) * EU_SEGMENT;
  return seg0 + (x - seg0) % EU_LANES;
}

// one synopsis evaluation (envutil_payload.cc:818-956 voronoi, :1500-1622 hdr_merge) for rays
// produced by `ray_of(i, ray)`; returns the index-plane value
template <int NCH, int MODE, typename RayFn>
__device__ __forceinline__ int dev_synopsis(const RenderParams& P, RayFn ray_of, float px[NCH]) {
  if (MODE == EU_MODE_SINGLE) {
    float r[3];
    ray_of(0, r);
    return dev_facet_eval<NCH>(P.f0, P.degree, P.wmat, r, px);
  } else if (MODE == EU_MODE_VORONOI) {
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
      dev_facet_eval<NCH>(P.facets[champion], P.degree, P.wmat, best, px);
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
      dev_facet_eval<NCH>(F, P.degree, P.wmat, r, p);
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

template <int NCH, int MODE, bool TWINE>
__global__ void __launch_bounds__(TILE_X* TILE_Y) k_render(const __grid_constant__ RenderParams P) {
  const TargetDev& T = P.trg;
  int x = blockIdx.x * TILE_X + threadIdx.x;
  int y = P.row0 + blockIdx.y * TILE_Y + threadIdx.y;
  if (x >= T.width || y >= P.row1) return;
  const float* __restrict__ tab_x = P.planar_x;
  const float* __restrict__ tab_y = P.planar_y;
  float p0x = tab_x[x], p0y = tab_y[y];
  int xf = first_lane_column(x);
  float p0x_first = tab_x[xf];
  float px[NCH];
  int idx;
  if (!TWINE) {
    auto ray_of = [&](int i, float r[3]) {
      const FacetDev& F = MODE == EU_MODE_SINGLE ? P.f0 : P.facets[i];
      dev_stepper(T, F.xx, F.yy, F.zz, p0x, p0y, p0x_first, y, r);
    };
    idx = dev_synopsis<NCH, MODE>(P, ray_of, px);
  } else {
    // deriv_stepper (stepper.h:1606-1694) + twine_t (twining.h:106-263) /
    // synopsis_t (envutil_payload.cc:647-690)
    float p1x = tab_x[T.width + x], p1x_first = tab_x[T.width + xf], p1y = tab_y[T.height + y];
    float acc[NCH], help[NCH];
#pragma unroll
    for (int c = 0; c < NCH; c++) acc[c] = 0.0f;
    idx = -1;
    if (MODE == EU_MODE_SINGLE) {
      float r00[3], du[3], dv[3];
      dev_stepper(T, P.f0.xx, P.f0.yy, P.f0.zz, p0x, p0y, p0x_first, y, r00);
      dev_stepper(T, P.f0.xx, P.f0.yy, P.f0.zz, p1x, p0y, p1x_first, y, du);
      dev_stepper(T, P.f0.xx, P.f0.yy, P.f0.zz, p0x, p1y, p0x_first, y, dv);
#pragma unroll
      for (int c = 0; c < 3; c++) {
        du[c] = du[c] - r00[c];
        dv[c] = dv[c] - r00[c];
      }
      for (int k = 0; k < P.n_taps; k++) {
        float cx = P.taps[3 * k], cy = P.taps[3 * k + 1], cw = P.taps[3 * k + 2];
        float r[3];
#pragma unroll
        for (int c = 0; c < 3; c++) r[c] = r00[c] + cx * du[c] + cy * dv[c];
        int id = dev_facet_eval<NCH>(P.f0, P.degree, P.wmat, r, help);
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
        dev_stepper(T, F.xx, F.yy, F.zz, p0x, p0y, p0x_first, y, r00);
        dev_stepper(T, F.xx, F.yy, F.zz, p1x, p0y, p1x_first, y, r10);
        dev_stepper(T, F.xx, F.yy, F.zz, p0x, p1y, p0x_first, y, r01);
#pragma unroll
        for (int c = 0; c < 3; c++) {
          np[i][c] = r00[c];
          np[i][3 + c] = r10[c] - r00[c];
          np[i][6 + c] = r01[c] - r00[c];
        }
      }
      for (int k = 0; k < P.n_taps; k++) {
        float cx = P.taps[3 * k], cy = P.taps[3 * k + 1], cw = P.taps[3 * k + 2];
        auto ray_of = [&](int i, float r[3]) {
#pragma unroll
          for (int c = 0; c < 3; c++) r[c] = np[i][c] + cx * np[i][3 + c] + cy * np[i][6 + c];
        };
        int id = dev_synopsis<NCH, MODE>(P, ray_of, help);
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

template <int NCH>
static cudaError_t launch_nch(const RenderParams& P, cudaStream_t st) {
  dim3 block(TILE_X, TILE_Y);
  dim3 grid((P.trg.width + TILE_X - 1) / TILE_X, (P.row1 - P.row0 + TILE_Y - 1) / TILE_Y);
  bool tw = P.n_taps > 0;
  switch (P.mode) {
    case EU_MODE_SINGLE:
      if (tw) k_render<NCH, EU_MODE_SINGLE, true><<<grid, block, 0, st>>>(P);
      else k_render<NCH, EU_MODE_SINGLE, false><<<grid, block, 0, st>>>(P);
      break;
    case EU_MODE_VORONOI:
      if (tw) k_render<NCH, EU_MODE_VORONOI, true><<<grid, block, 0, st>>>(P);
      else k_render<NCH, EU_MODE_VORONOI, false><<<grid, block, 0, st>>>(P);
      break;
    default:
      if (tw) k_render<NCH, EU_MODE_HDR, true><<<grid, block, 0, st>>>(P);
      else k_render<NCH, EU_MODE_HDR, false><<<grid, block, 0, st>>>(P);
  }
  return cudaGetLastError();
}

cudaError_t eu_launch_planar_tables(const TargetDev& T, float* d_x, float* d_y, cudaStream_t st) {
  int n = 2 * T.width + 2 * T.height;
  k_planar_tables<<<(n + 255) / 256, 256, 0, st>>>(T, d_x, d_y);
  return cudaGetLastError();
}

cudaError_t eu_launch_render(const RenderParams& P, cudaStream_t st) {
  switch (P.nch) {
    case 1: return launch_nch<1>(P, st);
    case 3: return launch_nch<3>(P, st);
    case 4: return launch_nch<4>(P, st);
    default: return cudaErrorInvalidValue;
  }
}
