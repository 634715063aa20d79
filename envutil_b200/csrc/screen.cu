// screen.cu - tethered output: float pixels -> one uint32 sRGBA value per pixel, what the reference stores into visor's
// frame buffer when it runs tethered (work(), envutil_payload.cc:524-531: `act + to_screen`).
//
// to_screen_t (envutil_payload.cc:298-413) sends every channel through lut_based_tf (:243-283): a degree-1 b-spline
// over 256 knots of 255 * RGB2sRGB(x) (:221-231) behind a clamp gate (NATURAL boundary, zimt/eval.h:2101-2110),
// evaluated at in * 255 by _eval_linear (zimt/eval.h:1037-1059: sum = c[i] * (1 - t); sum += c[i + 1] * t), truncated
// to an integer and packed A<<24 | B<<16 | G<<8 | R (one channel: grey, opaque; two: grey + alpha; three: opaque).
// Compiled with -fmad=false like everything else: no product is fused into a sum.
#include <stdint.h>

#include "kernels.h"

namespace {

__device__ __forceinline__ uint32_t dev_screen_channel(const float* __restrict__ lut, float in) {
  float c = in * 255.0f;
  if (c < 0.0f) c = 0.0f;
  else if (c > 255.0f) c = 255.0f;
  const float fl = floorf(c), t = c - fl;
  const int i = (int)fl;
  const float wl = 1.0f - t, wr = t;
  float sum = __ldg(lut + i);
  sum *= wl;
  float help = __ldg(lut + i + 1);
  help = help * wr;
  sum += help;
  return (uint32_t)sum;  // cvt.rzi.u32.f32; a NaN pixel gives 0 here (the x86 reference: undefined)
}

// one thread per pixel; the 257-entry table stays in L1
template <int NCH>
__global__ void __launch_bounds__(256) k_to_screen(const float* __restrict__ px, const float* __restrict__ lut, size_t n,
                                                   uint32_t* __restrict__ out) {
  for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
    uint32_t c[NCH];
    if constexpr (NCH == 4) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(px) + k);
      c[0] = dev_screen_channel(lut, v.x); c[1] = dev_screen_channel(lut, v.y);
      c[2] = dev_screen_channel(lut, v.z); c[3] = dev_screen_channel(lut, v.w);
    } else if constexpr (NCH == 2) {
      const float2 v = __ldg(reinterpret_cast<const float2*>(px) + k);
      c[0] = dev_screen_channel(lut, v.x); c[1] = dev_screen_channel(lut, v.y);
    } else {
#pragma unroll
      for (int j = 0; j < NCH; j++) c[j] = dev_screen_channel(lut, __ldg(px + k * NCH + j));
    }
    uint32_t o;
    if constexpr (NCH == 1) o = 0xFF000000u | (c[0] << 16) | (c[0] << 8) | c[0];
    else if constexpr (NCH == 2) o = (c[1] << 24) | (c[0] << 16) | (c[0] << 8) | c[0];
    else if constexpr (NCH == 3) o = 0xFF000000u | (c[2] << 16) | (c[1] << 8) | c[0];
    else o = (c[3] << 24) | (c[2] << 16) | (c[1] << 8) | c[0];
    out[k] = o;
  }
}

}  // namespace

cudaError_t eu_launch_to_screen(const float* px, int nch, size_t n, const float* lut, uint32_t* out, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  const size_t want = (n + 255) / 256;
  const int blocks = (int)(want < (size_t)148 * 16 ? want : (size_t)148 * 16);  // grid-stride beyond 16 blocks per SM
  switch (nch) {
    case 1: k_to_screen<1><<<blocks, 256, 0, st>>>(px, lut, n, out); break;
    case 2: k_to_screen<2><<<blocks, 256, 0, st>>>(px, lut, n, out); break;
    case 3: k_to_screen<3><<<blocks, 256, 0, st>>>(px, lut, n, out); break;
    case 4: k_to_screen<4><<<blocks, 256, 0, st>>>(px, lut, n, out); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}
