// render_c2.cu - k_render instantiations for 2-channel rasters (grey + alpha), 2 floats per texel in HBM
#include "render_impl.cuh"
cudaError_t EU_ARITH_FN(eu_launch_render_c2)(const RenderParams& P, cudaStream_t st) { return launch_render<2, 2>(P, st); }
