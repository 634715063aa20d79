// render_spec4.cu - the job-shape kernels of render_spec.cu for the padded 16-byte RGB texel layout
#define EU_SPEC_TS 4
#include "render_spec_impl.cuh"
