// render_spec_impl.cuh - body of render_spec.cu / render_spec4.cu (EU_SPEC_TS = floats per texel in HBM)
#pragma once
#include "render_impl.cuh"

namespace {
inline bool tiles_ok(const RenderParams& P) {
  return P.use_tiles && P.out && !P.index_out && (P.f0.src.stride & 3) == 0 && P.degree == 3;
}
// TS: floats per texel in HBM (3, or 4 = the padded 16-byte layout: one 128-bit load per tap)
template <int SP, bool TWINE, int TS>
bool single(const RenderParams& P, dim3 grid, dim3 block, cudaStream_t st, bool with_cubic) {
  if (P.degree == 1) {
    k_render<3, TS, EU_MODE_SINGLE, TWINE, 1, false, false, SP><<<grid, block, 0, st>>>(P);
    return true;
  }
  if (P.degree == 3 && with_cubic) {
    if constexpr (!TWINE) {
      if (tiles_ok(P)) {
        k_render_tiled<3, TS, false, 3, SP><<<grid, block, 0, st>>>(P);
        return true;
      }
    }
    k_render<3, TS, EU_MODE_SINGLE, TWINE, 3, false, false, SP><<<grid, block, 0, st>>>(P);
    return true;
  }
  return false;
}
template <int TS>
bool launch_spec(const RenderParams& P, dim3 grid, dim3 block, cudaStream_t st) {
  const bool tw = P.n_taps > 0;
  if (P.mode == EU_MODE_SINGLE) {
    switch (P.spec) {
      case 1: return !tw && single<1, false, TS>(P, grid, block, st, true);
      case 2: return !tw && single<2, false, TS>(P, grid, block, st, false);
      case 3: return !tw && single<3, false, TS>(P, grid, block, st, true);
      case 4: return !tw && single<4, false, TS>(P, grid, block, st, false);
      case 5: return tw ? single<5, true, TS>(P, grid, block, st, false) : single<5, false, TS>(P, grid, block, st, false);
      default: return false;
    }
  }
  if (tw || P.degree != 1 || P.n_facets > EU_SMEM_FACETS) return false;
  if (P.mode == EU_MODE_HDR && P.spec == 6) {
    k_render<3, TS, EU_MODE_HDR, false, 1, true, false, 6><<<grid, block, 0, st>>>(P);
    return true;
  }
  if (P.mode == EU_MODE_VORONOI && P.spec == 7) {
    k_render<3, TS, EU_MODE_VORONOI, false, 1, true, false, 7><<<grid, block, 0, st>>>(P);
    return true;
  }
  return false;
}
}  // namespace

// true: a kernel was launched (check cudaGetLastError); false: no compiled-in shape fits the job
#if EU_SPEC_TS == 4
bool EU_ARITH_FN(eu_launch_render_spec4)(const RenderParams& P, cudaStream_t st) {
#else
bool EU_ARITH_FN(eu_launch_render_spec)(const RenderParams& P, cudaStream_t st) {
#endif
  if (P.spec <= 0 || P.spec >= EU_N_SPECS || P.nch != 3 || P.tstride != EU_SPEC_TS || P.any_generic) return false;
  dim3 block(TILE_X, TILE_Y);
  dim3 grid((P.col1 - P.col0 + TILE_X - 1) / TILE_X, (P.row1 - P.row0 + TILE_Y - 1) / TILE_Y);
  return launch_spec<EU_SPEC_TS>(P, grid, block, st);
}
