// render_c1.cu - k_render instantiations for 1-channel rasters, 1 floats per texel in HBM
#include "render_impl.cuh"
cudaError_t EU_ARITH_FN(eu_launch_render_c1)(const RenderParams& P, cudaStream_t st) { return launch_render<1, 1>(P, st); }
