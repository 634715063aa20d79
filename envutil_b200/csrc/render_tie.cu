// render_tie.cu - the tie band of a job ("face / winning-facet indices bit-exact except within a
// tie band - |score difference| or |dominant-axis difference| below ~8 ulp - mask those pixels, report their
// count"). Parity tooling of the C ABI (eu_debug_tie_plane), next to eu_debug_planes: per target pixel 1 where
// the discrete choice the render makes there is within `ulps` units in the last place of flipping, i.e.
//   * single cubemap / biatan6 source: the two largest |components| of the ray (ray_to_cubeface's m1/m2/m3
//     comparisons, geometry.h:1178-1289),
//   * panorama synopsis: the two best z * recip_step scores among the facets the ray hits (_voronoi_syn,
//     envutil_payload.cc:818-956).
// A build of the reference with another math library may legitimately choose the other face / facet there.
// Not performance relevant: the general (run-time) device functions are used for every job shape.
#include "render_impl.cuh"

__device__ __forceinline__ bool dev_close(float a, float b, int ulps) {
  float m = fmaxf(fabsf(a), fabsf(b));
  return fabsf(a - b) <= (float)ulps * (m * 1.1920929e-7f);
}

__global__ void __launch_bounds__(TILE_X* TILE_Y) k_tie_plane(const __grid_constant__ RenderParams P, unsigned char* __restrict__ tie,
                                                               int ulps) {
  const TargetDev& T = P.trg;
  const int x = P.col0 + blockIdx.x * TILE_X + threadIdx.x;
  const int y = P.row0 + blockIdx.y * TILE_Y + threadIdx.y;
  if (x >= P.col1 || y >= P.row1) return;
  PixelTerms t;
  {
    float2 c0 = __ldg(P.col_tab + x), r0 = __ldg(P.row_tab + y);
    t.col = ColTerm{c0.x, c0.y};
    t.row = RowTerm{r0.x, r0.y};
    t.first = t.col;
    if (T.projection == EU_CYLINDRICAL && T.normalize) {
      float2 f0 = __ldg(P.col_tab + first_lane_column(x));
      t.first = ColTerm{f0.x, f0.y};
    }
    t.px = __ldg(P.planar_raw + x);
    t.py = __ldg(P.planar_raw + 2 * T.width + y);
    t.colb = t.col; t.firstb = t.first; t.rowb = t.row;
    t.pxb = t.px; t.pyb = t.py;
  }
  bool close = false;
  if (P.mode == EU_MODE_SINGLE) {
    const FacetDev& F = P.f0;
    if (F.kind != EU_SRC_MOUNT) {
      float r[3];
      dev_facet_ray<true, 0>(T, P.inv, F, t, y, r);
      float a = fabsf(r[0]), b = fabsf(r[1]), c = fabsf(r[2]);
      float hi = fmaxf(a, fmaxf(b, c)), mid;
      if (a == hi) mid = fmaxf(b, c); else if (b == hi) mid = fmaxf(a, c); else mid = fmaxf(a, b);
      close = dev_close(hi, mid, ulps);
    }
  } else if (P.mode == EU_MODE_VORONOI || P.mode == EU_MODE_VORONOI_PLUS) {
    float best = -FLT_MAX, second = -FLT_MAX;
    int n = 0;
    for (int i = 0; i < P.n_facets; i++) {
      const FacetDev& F = P.facets[i];
      float r[3];
      dev_facet_ray<true, 0>(T, P.inv, F, t, y, r);
      if (!dev_facet_mask(F, r)) continue;
      float cz = r[2] * F.recip_step;
      n++;
      if (cz > best) { second = best; best = cz; }
      else if (cz > second) second = cz;
    }
    close = n >= 2 && dev_close(best, second, ulps);
  }
  tie[(size_t)(y - P.row0) * T.width + x] = close ? 1 : 0;
}

cudaError_t eu_launch_tie_plane(const RenderParams& P, unsigned char* d_tie, int ulps, cudaStream_t st) {
  dim3 block(TILE_X, TILE_Y);
  dim3 grid((P.col1 - P.col0 + TILE_X - 1) / TILE_X, (P.row1 - P.row0 + TILE_Y - 1) / TILE_Y);
  k_tie_plane<<<grid, block, 0, st>>>(P, d_tie, ulps);
  return cudaGetLastError();
}
