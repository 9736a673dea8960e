// The arithmetic of the visual-prompt preprocessing shared by prep.cu and circle.cu: ATen's bilinear taps (upsample_bilinear2d,
// align_corners=False), the two lerps in ATen's op order, T.ToTensor / T.Normalize per byte (Hybridgl_main.py:115-117, :93, :120).
#pragma once
#include "hgl_common.cuh"

namespace hgl {

struct Taps {
  int i0, d;       // first source index, i1 - i0 (0 or 1)
  float w0, w1;
};

// ATen area_pixel_compute_source_index + compute_source_index_and_lambda (float, align_corners=False)
__device__ __forceinline__ float tap_scale(int in_size, int out_size) { return __fdiv_rn((float)in_size, (float)out_size); }
// `scale` = tap_scale(in_size, out_size): one correctly-rounded division shared by every pixel of an axis
__device__ __forceinline__ Taps make_taps(int dst, int in_size, int out_size, float scale) {
  Taps t;
  if (in_size == out_size) {
    t.i0 = dst; t.d = 0; t.w0 = 1.f; t.w1 = 0.f;
    return t;
  }
  float src = __fmaf_rn(scale, (float)dst + 0.5f, -0.5f);
  src = fmaxf(src, 0.f);
  int i0 = (int)src;
  i0 = min(i0, in_size - 1);
  t.i0 = i0;
  t.d = (i0 < in_size - 1) ? 1 : 0;
  float w1 = __fsub_rn(src, (float)i0);
  w1 = fminf(fmaxf(w1, 0.f), 1.f);
  t.w1 = w1;
  t.w0 = __fsub_rn(1.f, w1);
  return t;
}

__device__ __forceinline__ float bilerp(float a, float b, float c, float d, float wx0, float wx1, float wy0, float wy1) {
  const float top = __fmaf_rn(a, wx0, __fmul_rn(b, wx1));
  const float bot = __fmaf_rn(c, wx0, __fmul_rn(d, wx1));
  return __fmaf_rn(top, wy0, __fmul_rn(bot, wy1));
}

__device__ __forceinline__ Taps make_taps(int dst, int in_size, int out_size) { return make_taps(dst, in_size, out_size, tap_scale(in_size, out_size)); }

static __constant__ float c_in_mean[3] = {0.485f, 0.456f, 0.406f};
static __constant__ float c_in_std[3] = {0.229f, 0.224f, 0.225f};
static __constant__ float c_clip_mean[3] = {0.48145466f, 0.4578275f, 0.40821073f};

__device__ __forceinline__ float to_unit(uint32_t v) { return __fdiv_rn((float)v, 255.f); }                       // T.ToTensor
__device__ __forceinline__ float to_norm(uint32_t v, int c) { return __fdiv_rn(__fsub_rn(to_unit(v), c_in_mean[c]), c_in_std[c]); }  // + T.Normalize

}  // namespace hgl
