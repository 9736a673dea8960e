// Two small kernels around the (plain PyTorch) ViT blocks of the hybrid CLIP forward.
//
// (f1) CLS-row attention under a key bitmap.  The reference masks attention with a bool tensor [N*heads, L+1, L+1] whose only
//      blocked entries are (query 0, patch keys whose soft mask is exactly 0) (model/backbone.py:108-115; nn.MultiheadAttention
//      turns it into an additive float mask of the same size, third_party/modified_CLIP/clip/model.py:220-228).  Every other
//      query row is unmasked, so the forward runs ONE unmasked SDPA over all rows and this kernel recomputes row 0 alone:
//          out[m, h, :] = softmax_j( q[m,0,h,:] . k[m,j,h,:] / sqrt(hd) + bias[m, j] ) @ v[m, :, h, :]
//      straight from the packed projection qkv [M, L1, 3, heads, hd] -- one warp per (proposal, head), a lane per key for the
//      scores (f32), a lane per output channel pair for the weighted sum.
// (a5) CLS head: ln_post(x[:, 0, :]) @ proj (model/backbone.py:254-260, 220-225, 296-306) -- LayerNorm in f32 and the [Dv x De]
//      projection in one launch, optionally accumulated onto `out` (G2L&L2G adds the heads of its two hybrid streams).
#include "hgl_common.cuh"

namespace hgl {

template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void stf(T* p, float v);
template <> __device__ __forceinline__ void stf<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void stf<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

constexpr int kClsWarps = 4;
constexpr int kClsMaxL1 = 1024;      // keys per sequence (577 for ViT-L/14@336)

template <typename T>
__global__ void __launch_bounds__(kClsWarps * 32) cls_attention_kernel(const T* __restrict__ qkv, const float* __restrict__ bias, int M, int L1,
                                                                        int heads, int hd, float scale, T* __restrict__ out) {
  extern __shared__ float cls_sm[];                    // per warp: q [hd] | p [L1]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int task = blockIdx.x * kClsWarps + warp;
  if (task >= M * heads) return;
  const int m = task / heads, h = task - m * heads;
  float* qs = cls_sm + (size_t)warp * (hd + L1);
  float* ps = qs + hd;
  const size_t tok = (size_t)3 * heads * hd;           // elements per token of the packed projection
  const T* base = qkv + (size_t)m * L1 * tok + (size_t)h * hd;
  for (int d = lane; d < hd; d += 32) qs[d] = ldf(base + d);          // q of token 0
  __syncwarp();
  // scores: a lane per key
  float mx = -INFINITY;
  for (int j = lane; j < L1; j += 32) {
    const T* kj = base + (size_t)j * tok + (size_t)heads * hd;
    float s = 0.f;
    for (int d = 0; d < hd; ++d) s = fmaf(qs[d], ldf(kj + d), s);
    s = s * scale + (bias ? __ldg(bias + (size_t)m * L1 + j) : 0.f);
    ps[j] = s;
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j < L1; j += 32) { const float e = __expf(ps[j] - mx); ps[j] = e; sum += e; }
  sum = warp_sum(sum);
  __syncwarp();
  const float inv = 1.f / sum;
  // weighted sum of the values: a lane per output channel
  for (int d = lane; d < hd; d += 32) {
    const T* vd = base + (size_t)2 * heads * hd + d;
    float a = 0.f;
    for (int j = 0; j < L1; ++j) a = fmaf(ps[j], ldf(vd + (size_t)j * tok), a);
    stf(out + ((size_t)m * heads + h) * hd + d, a * inv);
  }
}

constexpr int kHeadRows = 8;
constexpr int kHeadThreads = 256;

template <typename TX, typename TW>
__global__ void __launch_bounds__(kHeadThreads) cls_head_kernel(const TX* __restrict__ x, long long row_stride, const TW* __restrict__ gamma,
                                                                const TW* __restrict__ beta, const TW* __restrict__ proj, int M, int Dv, int De,
                                                                float eps, int accumulate, float* __restrict__ out) {
  extern __shared__ float head_sm[];                   // [kHeadRows][Dv] normalised rows
  const int r0 = blockIdx.x * kHeadRows, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // LayerNorm (f32, biased variance, like F.layer_norm): one warp per row
  for (int r = warp; r < kHeadRows; r += kHeadThreads / 32) {
    float* dst = head_sm + (size_t)r * Dv;
    if (r0 + r >= M) { for (int d = lane; d < Dv; d += 32) dst[d] = 0.f; continue; }
    const TX* xr = x + (size_t)(r0 + r) * row_stride;
    float s = 0.f;
    for (int d = lane; d < Dv; d += 32) { const float v = ldf(xr + d); dst[d] = v; s += v; }
    const float mean = warp_sum(s) / (float)Dv;
    float q = 0.f;
    for (int d = lane; d < Dv; d += 32) { const float c = dst[d] - mean; q += c * c; }
    const float rstd = rsqrtf(warp_sum(q) / (float)Dv + eps);
    for (int d = lane; d < Dv; d += 32) dst[d] = (dst[d] - mean) * rstd * ldf(gamma + d) + ldf(beta + d);
  }
  __syncthreads();
  // projection: a thread per output column, kHeadRows rows at once (proj rows are read coalesced, once per CTA)
  for (int c = tid; c < De; c += kHeadThreads) {
    float acc[kHeadRows];
#pragma unroll
    for (int r = 0; r < kHeadRows; ++r) acc[r] = 0.f;
    for (int k = 0; k < Dv; ++k) {
      const float w = ldf(proj + (size_t)k * De + c);
#pragma unroll
      for (int r = 0; r < kHeadRows; ++r) acc[r] = fmaf(head_sm[(size_t)r * Dv + k], w, acc[r]);
    }
#pragma unroll
    for (int r = 0; r < kHeadRows; ++r) {
      if (r0 + r < M) {
        float* o = out + (size_t)(r0 + r) * De + c;
        *o = accumulate ? *o + acc[r] : acc[r];
      }
    }
  }
}

}  // namespace hgl

extern "C" int hgl_cls_attention(const void* qkv, const float* bias, int M, int L1, int heads, int hd, int dtype, void* out, void* stream) {
  using namespace hgl;
  if (M == 0) return HGL_OK;
  HGL_REQUIRE(qkv && out, "hgl_cls_attention: null pointer");
  HGL_REQUIRE(M > 0 && L1 >= 1 && L1 <= kClsMaxL1 && heads >= 1 && hd >= 1 && hd <= 256, "hgl_cls_attention: bad shape");
  HGL_REQUIRE(dtype == HGL_F32 || dtype == HGL_BF16, "hgl_cls_attention: dtype %d", dtype);
  const size_t smem = (size_t)kClsWarps * (hd + L1) * 4;
  const int blocks = ceil_div(M * heads, kClsWarps);
  const float scale = 1.0f / sqrtf((float)hd);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == HGL_BF16)
    cls_attention_kernel<__nv_bfloat16><<<blocks, kClsWarps * 32, smem, st>>>(reinterpret_cast<const __nv_bfloat16*>(qkv), bias, M, L1, heads, hd, scale,
                                                                               reinterpret_cast<__nv_bfloat16*>(out));
  else
    cls_attention_kernel<float><<<blocks, kClsWarps * 32, smem, st>>>(reinterpret_cast<const float*>(qkv), bias, M, L1, heads, hd, scale,
                                                                       reinterpret_cast<float*>(out));
  return launch_status("hgl_cls_attention");
}

extern "C" int hgl_cls_head(const void* x, int64_t row_stride, const void* gamma, const void* beta, const void* proj, int M, int Dv, int De,
                            double eps, int x_dtype, int w_dtype, int accumulate, float* out, void* stream) {
  using namespace hgl;
  if (M == 0) return HGL_OK;
  HGL_REQUIRE(x && gamma && beta && proj && out, "hgl_cls_head: null pointer");
  HGL_REQUIRE(M > 0 && Dv >= 1 && De >= 1 && row_stride >= Dv, "hgl_cls_head: bad shape");
  HGL_REQUIRE((x_dtype == HGL_F32 || x_dtype == HGL_BF16) && (w_dtype == HGL_F32 || w_dtype == HGL_BF16), "hgl_cls_head: dtype");
  const size_t smem = (size_t)kHeadRows * Dv * 4;
  HGL_REQUIRE(smem <= 200 * 1024, "hgl_cls_head: Dv=%d too wide", Dv);
  const int blocks = ceil_div(M, kHeadRows);
  cudaStream_t st = (cudaStream_t)stream;
  auto go = [&](auto kern, auto xp, auto wp) -> int {
    using TXp = decltype(xp); using TWp = decltype(wp);
    const int rc = ensure_dyn_smem(reinterpret_cast<const void*>(kern), smem, "hgl_cls_head");
    if (rc != HGL_OK) return rc;
    kern<<<blocks, kHeadThreads, smem, st>>>(reinterpret_cast<TXp>(x), (long long)row_stride, reinterpret_cast<TWp>(gamma), reinterpret_cast<TWp>(beta),
                                             reinterpret_cast<TWp>(proj), M, Dv, De, (float)eps, accumulate, out);
    return launch_status("hgl_cls_head");
  };
  if (x_dtype == HGL_F32 && w_dtype == HGL_F32) return go(cls_head_kernel<float, float>, (const float*)nullptr, (const float*)nullptr);
  if (x_dtype == HGL_BF16 && w_dtype == HGL_BF16) return go(cls_head_kernel<__nv_bfloat16, __nv_bfloat16>, (const __nv_bfloat16*)nullptr, (const __nv_bfloat16*)nullptr);
  if (x_dtype == HGL_BF16) return go(cls_head_kernel<__nv_bfloat16, float>, (const __nv_bfloat16*)nullptr, (const float*)nullptr);
  return go(cls_head_kernel<float, __nv_bfloat16>, (const float*)nullptr, (const __nv_bfloat16*)nullptr);
}
