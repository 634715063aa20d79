// render_c3p.cu - k_render instantiations for 3-channel rasters, 4 floats per texel in HBM
#include "render_impl.cuh"
cudaError_t EU_ARITH_FN(eu_launch_render_c3p)(const RenderParams& P, cudaStream_t st) { return launch_render<3, 4>(P, st); }
