// render_spec.cu - render kernels compiled for the commonest job shapes (plan.h: eu_render_specs):
// RGB rasters with 12-byte texels, target projection / source kind / boundary gates fixed at
// compile time. Same device functions, same arithmetic and operation order as the general
// instantiations in render_c*.cu - only the run-time switches are folded away.
#include "render_impl.cuh"

namespace {
inline bool tiles_ok(const RenderParams& P) {
  return P.use_tiles && P.out && !P.index_out && (P.f0.src.stride & 3) == 0 && P.degree == 3;
}
template <int SP, bool TWINE>
bool single(const RenderParams& P, dim3 grid, dim3 block, cudaStream_t st, bool with_cubic) {
  if (P.degree == 1) {
    k_render<3, 3, EU_MODE_SINGLE, TWINE, 1, false, false, SP><<<grid, block, 0, st>>>(P);
    return true;
  }
  if (P.degree == 3 && with_cubic) {
    if constexpr (!TWINE) {
      if (tiles_ok(P)) {
        k_render_tiled<3, 3, false, 3, SP><<<grid, block, 0, st>>>(P);
        return true;
      }
    }
    k_render<3, 3, EU_MODE_SINGLE, TWINE, 3, false, false, SP><<<grid, block, 0, st>>>(P);
    return true;
  }
  return false;
}
}  // namespace

// true: a kernel was launched (check cudaGetLastError); false: no compiled-in shape fits the job
bool EU_ARITH_FN(eu_launch_render_spec)(const RenderParams& P, cudaStream_t st) {
  if (P.spec <= 0 || P.spec >= EU_N_SPECS || P.nch != 3 || P.tstride != 3 || P.any_generic) return false;
  dim3 block(TILE_X, TILE_Y);
  dim3 grid((P.col1 - P.col0 + TILE_X - 1) / TILE_X, (P.row1 - P.row0 + TILE_Y - 1) / TILE_Y);
  const bool tw = P.n_taps > 0;
  if (P.mode == EU_MODE_SINGLE) {
    switch (P.spec) {
      case 1: return !tw && single<1, false>(P, grid, block, st, true);
      case 2: return !tw && single<2, false>(P, grid, block, st, false);
      case 3: return !tw && single<3, false>(P, grid, block, st, true);
      case 4: return !tw && single<4, false>(P, grid, block, st, false);
      case 5: return tw ? single<5, true>(P, grid, block, st, false) : single<5, false>(P, grid, block, st, false);
      default: return false;
    }
  }
  if (tw || P.degree != 1 || P.n_facets > EU_SMEM_FACETS) return false;
  if (P.mode == EU_MODE_HDR && P.spec == 6) {
    k_render<3, 3, EU_MODE_HDR, false, 1, true, false, 6><<<grid, block, 0, st>>>(P);
    return true;
  }
  if (P.mode == EU_MODE_VORONOI && P.spec == 7) {
    k_render<3, 3, EU_MODE_VORONOI, false, 1, true, false, 7><<<grid, block, 0, st>>>(P);
    return true;
  }
  return false;
}
