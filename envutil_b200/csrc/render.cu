// render.cu - per-job stepper tables and the dispatcher over the render translation units.
#include "eu_device.cuh"
#include "kernels.h"

// Per-column / per-row stepper terms (eu_device.cuh), from the planar coordinates exactly as
// stepper_base produces them when driven by zimt::process: init at the start of each 512-px
// segment, += delta per 16-px vector (stepper.h:324-350, zimt/wielding.h:317-455).
// col[0..W) plain, [W..2W) x-biased stepper (deriv_stepper's r10); row[0..H) plain, [H..2H)
// y-biased (r01).
__global__ void k_planar_tables(TargetDev T, float2* __restrict__ col, float2* __restrict__ row,
                                float* __restrict__ raw) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 2 * T.width) {
    int x = i % T.width;
    float bias = i < T.width ? 0.0f : T.bias_x;
    int seg0 = (x / EU_SEGMENT) * EU_SEGMENT;
    int r = x - seg0, lane = r % EU_LANES, v = r / EU_LANES;
    // a cropped output feeds the steppers offset discrete coordinates (zimt/wielding.h:391-407); the
    // segments are those of the crop, the scaling is the whole target's
    float ll0 = (float)(2 * lane) + (float)((seg0 + T.off_x) * 2 + 1);
    float p = bias + ll0 * T.fx1 + ((float)(2 * T.full_w) - ll0) * T.fx0;
    for (int k = 0; k < v; k++) p += T.delta;
    ColTerm c;
    dev_col_term(T, p, c);
    col[i] = make_float2(c.a, c.b);
    raw[i] = p;
  }
  int j = i - 2 * T.width;
  if (j >= 0 && j < 2 * T.height) {
    int y = j % T.height;
    float bias = j < T.height ? 0.0f : T.bias_y;
    int ll1 = (y + T.off_y) * 2 + 1;
    float p = bias + ll1 * T.fy1 + (float)(2 * T.full_h - ll1) * T.fy0;
    RowTerm rt;
    dev_row_term(T, p, y, rt);
    row[j] = make_float2(rt.a, rt.b);
    raw[2 * T.width + j] = p;
  }
}

cudaError_t eu_launch_planar_tables(const TargetDev& T, float2* d_col, float2* d_row, float* d_raw, cudaStream_t st) {
  int n = 2 * T.width + 2 * T.height;
  k_planar_tables<<<(n + 255) / 256, 256, 0, st>>>(T, d_col, d_row, d_raw);
  return cudaGetLastError();
}

cudaError_t eu_launch_render(const RenderParams& P, cudaStream_t st, int* spec_used) {
  if (spec_used) *spec_used = 0;
  const bool fma = P.arith == 1;
  // the general build (any_generic) ignores the compile-time texel stride, any TU of the right
  // channel count serves it
  const bool spec_ran = P.tstride == 4 ? (fma ? eu_launch_render_spec4_fma(P, st) : eu_launch_render_spec4(P, st))
                                       : (fma ? eu_launch_render_spec_fma(P, st) : eu_launch_render_spec(P, st));
  if (spec_ran) {
    if (spec_used) *spec_used = P.spec;
    return cudaGetLastError();
  }
  if (P.nch == 1) return fma ? eu_launch_render_c1_fma(P, st) : eu_launch_render_c1(P, st);
  if (P.nch == 2) return fma ? eu_launch_render_c2_fma(P, st) : eu_launch_render_c2(P, st);
  if (P.nch == 3 && (P.tstride == 3 || P.any_generic)) return fma ? eu_launch_render_c3_fma(P, st) : eu_launch_render_c3(P, st);
  if (P.nch == 3 && P.tstride == 4) return fma ? eu_launch_render_c3p_fma(P, st) : eu_launch_render_c3p(P, st);
  if (P.nch == 4) return fma ? eu_launch_render_c4_fma(P, st) : eu_launch_render_c4(P, st);
  return cudaErrorInvalidValue;
}
