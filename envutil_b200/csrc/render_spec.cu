// render_spec.cu - render kernels compiled for the commonest job shapes (plan.h: eu_render_specs):
// RGB rasters with 12- or 16-byte texels, target projection / source kind / boundary gates fixed at
// compile time. Same device functions, same arithmetic and operation order as the general
// instantiations in render_c*.cu - only the run-time switches are folded away.
#define EU_SPEC_TS 3
#include "render_spec_impl.cuh"
