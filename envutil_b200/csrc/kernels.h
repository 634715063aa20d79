// kernels.h - launchers of the sm_100a kernels (render.cu, stage.cu), called by api.cu only.
#pragma once
#include <cuda_runtime.h>

#include "plan.h"

// prefilter of one axis (zimt/recursive.h iir_filter): poles, horizons and gain narrowed to
// float exactly as solve_gain_inlined uses them (recursive.h:650-662)
struct IirDev {
  int32_t bc, npoles;
  float pole[EU_MAX_DEGREE / 2 + 1];
  float pole_pow[EU_MAX_DEGREE / 2 + 1];  // float(powl(pole, e)) for the full-loop variants
  int32_t horizon[EU_MAX_DEGREE / 2 + 1];
  float gain;
};

cudaError_t eu_launch_planar_tables(const TargetDev& T, float2* d_col, float2* d_row, float* d_raw, cudaStream_t st);
// per (channels, texel stride) translation units (render_c*.cu)
cudaError_t eu_launch_render_c1(const RenderParams& P, cudaStream_t st);
cudaError_t eu_launch_render_c2(const RenderParams& P, cudaStream_t st);
cudaError_t eu_launch_render_c3(const RenderParams& P, cudaStream_t st);
cudaError_t eu_launch_render_c3p(const RenderParams& P, cudaStream_t st);
cudaError_t eu_launch_render_c4(const RenderParams& P, cudaStream_t st);
// kernels compiled for one job shape (render_spec.cu); false: none fits, use the general ones
bool eu_launch_render_spec(const RenderParams& P, cudaStream_t st);
bool eu_launch_render_spec4(const RenderParams& P, cudaStream_t st);  // 16-byte RGB texels
// the same translation units compiled with -DEU_CONTRACT_WINDOW (RenderParams::arith == 1)
cudaError_t eu_launch_render_c1_fma(const RenderParams& P, cudaStream_t st);
cudaError_t eu_launch_render_c2_fma(const RenderParams& P, cudaStream_t st);
cudaError_t eu_launch_render_c3_fma(const RenderParams& P, cudaStream_t st);
cudaError_t eu_launch_render_c3p_fma(const RenderParams& P, cudaStream_t st);
cudaError_t eu_launch_render_c4_fma(const RenderParams& P, cudaStream_t st);
bool eu_launch_render_spec_fma(const RenderParams& P, cudaStream_t st);
bool eu_launch_render_spec4_fma(const RenderParams& P, cudaStream_t st);
// spec_used (optional): index of the compiled-in job shape whose kernel ran, 0 = a general kernel
cudaError_t eu_launch_render(const RenderParams& P, cudaStream_t st, int* spec_used = nullptr);

// parity tooling (render_tie.cu): 1 where the face / winning-facet choice is within `ulps` of flipping
cudaError_t eu_launch_tie_plane(const RenderParams& P, unsigned char* d_tie, int ulps, cudaStream_t st);

// staging (stage.cu). `core` is texel (0,0) of the core inside the container.
cudaError_t eu_launch_iir_x(float* core, int stride, int nch, int w, int h, const IirDev& f, cudaStream_t st);
cudaError_t eu_launch_iir_y(float* core, int stride, int nch, int w, int h, int n_sections, const IirDev& f,
                            cudaStream_t st);
cudaError_t eu_launch_iir_y_spherical(float* core, int stride, int nch, int w, int h, const IirDev& f,
                                      cudaStream_t st);
// NATURAL brace of one line of n floats: k values before and after, twice the end value minus the
// mirrored one (zimt/brace.h:254-266,299-311)
cudaError_t eu_launch_brace_natural_1d(float* core, int n, int k, cudaStream_t st);
cudaError_t eu_launch_brace(float* core, int stride, int nch, int w, int h, int lx, int rx, int ly, int ry, int bc0,
                            int bc1, int spherical, cudaStream_t st);
cudaError_t eu_launch_cubemap_support(float* ir, int pitch, int nch, int face_px, int section_px, int left, int right,
                                      double refc_md, double model_to_px, int* n_launches, cudaStream_t st);
// the six faces of a cubemap raster in device memory -> the centres of their sections of the IR (one launch)
cudaError_t eu_launch_cubemap_place(const float* src, float* ir, int pitch, int nch, int face_px, int section_px, int left,
                                    cudaStream_t st);
// alpha of masked / cropped facets (stage.cu): feather the 0/1 plane with the 5-tap binomial along x
// then y (REFLECT), then expand the raster to `nch` channels and multiply every channel by it
cudaError_t eu_launch_alpha_apply(const unsigned char* mask, float* tmp_a, float* tmp_b, const float* raw, int native_nch,
                                  float* out, int nch, int w, int h, cudaStream_t st);
// rows of nch-float texels -> 16-byte texels, cw x chh of them, destination rows dst_pitch_texels apart
cudaError_t eu_launch_pad_texels(const float* src, int src_pitch, float* dst, int dst_pitch_texels, int cw, int chh, int nch,
                                 cudaStream_t st);
// tethered output (screen.cu): n pixels of nch interleaved floats -> uint32 sRGBA through the 257-entry table of
// eu_screen_lut (to_screen_t, envutil_payload.cc:298-413)
cudaError_t eu_launch_to_screen(const float* px, int nch, size_t n, const float* lut, uint32_t* out, cudaStream_t st);
