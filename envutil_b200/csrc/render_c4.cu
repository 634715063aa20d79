// render_c4.cu - k_render instantiations for 4-channel rasters, 4 floats per texel in HBM
#include "render_impl.cuh"
cudaError_t EU_ARITH_FN(eu_launch_render_c4)(const RenderParams& P, cudaStream_t st) { return launch_render<4, 4>(P, st); }
