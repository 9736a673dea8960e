// Resampling primitives shared by the mask-pass kernels (mask_rows.cu) and the token-space GEM pooling (gem_token.cu):
// ATen's float linspace, the direction ramps of gen_dir_mask (utils.py:135-161) and the tap tables of ATen's antialiased
// bilinear filter (_compute_indices_min_size_weights_aa).
#pragma once
#include "hgl_common.cuh"

namespace hgl {

__device__ __forceinline__ float linspace_at(float a, float b, int n, int i) {  // ATen linspace (float): both-ends evaluation
  if (n <= 1) return a;
  const float step = __fdiv_rn(__fsub_rn(b, a), (float)(n - 1));
  return (i < n / 2) ? __fmaf_rn(step, (float)i, a) : __fmaf_rn(-step, (float)(n - 1 - i), b);
}
// gen_dir_mask utils.py:135-161 (up/down/none are all-ones: the vertical ramps are commented out in the reference)
__device__ __forceinline__ float ramp_at(int dirflag, int x, int W) {
  if (dirflag == HGL_DIR_LEFT) return linspace_at(1.f, 0.f, W, x);
  if (dirflag == HGL_DIR_RIGHT) return linspace_at(0.f, 1.f, W, x);
  if (dirflag == HGL_DIR_MIDDLE) {
    const int h = W / 2;
    return (x < h) ? linspace_at(0.f, 1.f, h, x) : linspace_at(1.f, 0.f, W - h, x - h);
  }
  return 1.f;
}

// ATen _compute_indices_min_size_weights_aa for the triangle (bilinear) filter; one thread per output index.
static __device__ void aa_fill(int i, int in_size, int out_size, int maxk, int* xmin_out, int* xsize_out, float* w) {
  const float scale = __fdiv_rn((float)in_size, (float)out_size);
  float support, invscale;
  if (scale >= 1.f) { support = scale; invscale = __fdiv_rn(1.f, scale); } else { support = 1.f; invscale = 1.f; }
  const float center = (float)((double)scale * ((double)i + 0.5));
  int xmin = (int)((double)__fsub_rn(center, support) + 0.5);
  xmin = max(xmin, 0);
  int xsize = min((int)((double)__fadd_rn(center, support) + 0.5), in_size) - xmin;
  xsize = max(min(xsize, maxk), 0);
  float total = 0.f;
  for (int j = 0; j < xsize; ++j) {
    float t = (float)(((double)__fsub_rn((float)(j + xmin), center) + 0.5) * (double)invscale);
    t = fabsf(t);
    const float wt = (t < 1.f) ? __fsub_rn(1.f, t) : 0.f;
    w[j] = wt;
    total = __fadd_rn(total, wt);
  }
  if (total != 0.f)
    for (int j = 0; j < xsize; ++j) w[j] = __fdiv_rn(w[j], total);
  for (int j = xsize; j < maxk; ++j) w[j] = 0.f;
  *xmin_out = xmin; *xsize_out = xsize;
}


}  // namespace hgl
