// (a3) CLS-row attention mask, (a4) token masking + stream mix.   ((a2) mask -> patch grid lives in mask_rows.cu)
//
// (a3) replaces CLIPViTFM.make_attn_mask (model/backbone.py:108-115); (a4) replaces the permute/view/mul/cat chains of
// model/backbone.py:214-216, 235-249, 275-291.  Both are pure streaming kernels (16-byte accesses, grid-stride).
#include "hgl_common.cuh"

namespace hgl {

// (a3) full boolean attention mask, written once, 16 bytes per store.  Only row 0 of every (mask, head) slab can be non-zero:
// a 16-byte chunk that lies past row 0 and inside one slab (99.5 % of them) is a plain zero store after ONE 64-bit division.
__global__ void __launch_bounds__(256) attn_mask_kernel(const float* __restrict__ grid, int M, int L, int heads, uint8_t* __restrict__ out) {
  const int L1 = L + 1;
  const size_t per = (size_t)L1 * L1;
  const size_t total = (size_t)M * heads * per;
  const size_t stride = (size_t)gridDim.x * blockDim.x * 16;
  size_t o = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
  size_t mh = o / per, within = o - mh * per;                 // slab and offset inside it: divided once, then advanced by adds
  const size_t dq = stride / per, dr = stride - dq * per;
  for (; o < total; o += stride, mh += dq, within += dr) {
    if (within >= per) { within -= per; ++mh; }
    uint32_t w[4] = {0, 0, 0, 0};
    if (within < (size_t)L1 || within + 16 > per) {          // touches row 0 of this slab or of the next one
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        size_t mh2 = mh, r = within + q;
        while (r >= per) { r -= per; ++mh2; }                 // (tiny grids: a chunk may span several slabs)
        if (o + q < total && r >= 1 && r < (size_t)L1) {
          const int m = (int)(mh2 / heads);
          if (grid[(size_t)m * L + (r - 1)] == 0.f) w[q >> 2] |= 1u << ((q & 3) * 8);
        }
      }
    }
    if (o + 16 <= total) {
      stg_stream(reinterpret_cast<uint4*>(out + o), make_uint4(w[0], w[1], w[2], w[3]));
    } else {
      for (int q = 0; o + q < total; ++q) out[o + q] = (w[q >> 2] >> ((q & 3) * 8)) & 0xff;
    }
  }
}

__global__ void attn_bias_kernel(const float* __restrict__ grid, int M, int L, float* __restrict__ bias) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int L1 = L + 1;
  if (idx >= M * L1) return;
  const int m = idx / L1, k = idx - m * L1;
  float v = 0.f;
  if (k >= 1 && grid[(size_t)m * L + (k - 1)] == 0.f) v = -INFINITY;
  bias[idx] = v;
}

// (a4) out[l,m,:] = a * w(l,m) * src[l,m,:] + b * add[l,m,:]   (8 elements = 16 B (bf16) / 2x16 B (f32) per thread-iteration)
template <bool kBF16>
__global__ void __launch_bounds__(256) token_mask_fuse_kernel(const void* __restrict__ src_, const void* __restrict__ add_, const float* __restrict__ grid,
                                                              float a, float b, int L1, int M, int D, int nld, void* __restrict__ out_) {
  const int L = L1 - 1;
  const int dv = D / 8;                       // vectors of 8 elements per token
  const size_t total = (size_t)L1 * M * dv;
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (size_t)gridDim.x * blockDim.x) {
    const size_t tok = v / dv;                // = l*M + m (LND) or m*L1 + l (NLD)
    int l, m;
    if (nld) { m = (int)(tok / L1); l = (int)(tok - (size_t)m * L1); }
    else { l = (int)(tok / M); m = (int)(tok - (size_t)l * M); }
    float w = a;
    if (grid != nullptr && l > 0) w = a * grid[(size_t)m * L + (l - 1)];
    float x[8], y[8];
    if (kBF16) {
      const uint4 s = reinterpret_cast<const uint4*>(src_)[v];
      const uint32_t sw[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) { x[2 * q] = bf16_bits_to_float(sw[q] & 0xffffu); x[2 * q + 1] = bf16_bits_to_float(sw[q] >> 16); }
      if (add_) {
        const uint4 t = reinterpret_cast<const uint4*>(add_)[v];
        const uint32_t tw[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) { y[2 * q] = bf16_bits_to_float(tw[q] & 0xffffu); y[2 * q + 1] = bf16_bits_to_float(tw[q] >> 16); }
      }
    } else {
      const float4 s0 = reinterpret_cast<const float4*>(src_)[2 * v], s1 = reinterpret_cast<const float4*>(src_)[2 * v + 1];
      x[0] = s0.x; x[1] = s0.y; x[2] = s0.z; x[3] = s0.w; x[4] = s1.x; x[5] = s1.y; x[6] = s1.z; x[7] = s1.w;
      if (add_) {
        const float4 t0 = reinterpret_cast<const float4*>(add_)[2 * v], t1 = reinterpret_cast<const float4*>(add_)[2 * v + 1];
        y[0] = t0.x; y[1] = t0.y; y[2] = t0.z; y[3] = t0.w; y[4] = t1.x; y[5] = t1.y; y[6] = t1.z; y[7] = t1.w;
      }
    }
    float o[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float t = __fmul_rn(x[q], w);
      o[q] = add_ ? __fadd_rn(t, __fmul_rn(y[q], b)) : t;
    }
    if (kBF16) {
      uint4 r;
      r.x = pack_bf16x2(o[0], o[1]); r.y = pack_bf16x2(o[2], o[3]); r.z = pack_bf16x2(o[4], o[5]); r.w = pack_bf16x2(o[6], o[7]);
      reinterpret_cast<uint4*>(out_)[v] = r;
    } else {
      reinterpret_cast<float4*>(out_)[2 * v] = make_float4(o[0], o[1], o[2], o[3]);
      reinterpret_cast<float4*>(out_)[2 * v + 1] = make_float4(o[4], o[5], o[6], o[7]);
    }
  }
}

// (f1) the same fuse with the LayerNorm that follows it (ln_1 of the masked block, clip/model.py:244-257) in the same pass: a WARP
// owns a token row (NLD layout), keeps the fused row in registers, stores it (rounded to the stream dtype, like the unfused path),
// and normalises the STORED values in f32 (CLIP's LayerNorm runs in fp32 whatever the stream dtype, clip/model.py:188-195):
// two-pass mean / variance over the registers, y = (x - mean) * rsqrt(var + eps) * gamma + beta.
constexpr int kLnMaxV = 8;                     // 8-element vectors per lane: D <= 2048
template <bool kBF16>
__global__ void __launch_bounds__(256) token_mask_fuse_ln_kernel(const void* __restrict__ src_, const void* __restrict__ add_,
                                                                 const float* __restrict__ grid, float a, float b,
                                                                 const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                                                 int L1, int M, int D, void* __restrict__ out_x, void* __restrict__ out_ln) {
  const int L = L1 - 1, dv = D / 8, lane = threadIdx.x & 31;
  const size_t rows = (size_t)M * L1;
  const float inv_d = 1.f / (float)D;
  for (size_t row = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5); row < rows; row += (size_t)gridDim.x * 8) {
    const int m = (int)(row / L1), l = (int)(row - (size_t)m * L1);
    float w = a;
    if (grid != nullptr && l > 0) w = a * grid[(size_t)m * L + (l - 1)];
    float o[kLnMaxV][8];
    float sum = 0.f;
#pragma unroll
    for (int u = 0; u < kLnMaxV; ++u) {
      const int vi = lane + 32 * u;
      if (vi >= dv) break;
      const size_t v = row * dv + vi;
      float x[8], y[8];
      if (kBF16) {
        const uint4 s4 = reinterpret_cast<const uint4*>(src_)[v];
        const uint32_t sw[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) { x[2 * q] = bf16_bits_to_float(sw[q] & 0xffffu); x[2 * q + 1] = bf16_bits_to_float(sw[q] >> 16); }
        if (add_) {
          const uint4 t4 = reinterpret_cast<const uint4*>(add_)[v];
          const uint32_t tw[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) { y[2 * q] = bf16_bits_to_float(tw[q] & 0xffffu); y[2 * q + 1] = bf16_bits_to_float(tw[q] >> 16); }
        }
      } else {
        const float4 s0 = reinterpret_cast<const float4*>(src_)[2 * v], s1 = reinterpret_cast<const float4*>(src_)[2 * v + 1];
        x[0] = s0.x; x[1] = s0.y; x[2] = s0.z; x[3] = s0.w; x[4] = s1.x; x[5] = s1.y; x[6] = s1.z; x[7] = s1.w;
        if (add_) {
          const float4 t0 = reinterpret_cast<const float4*>(add_)[2 * v], t1 = reinterpret_cast<const float4*>(add_)[2 * v + 1];
          y[0] = t0.x; y[1] = t0.y; y[2] = t0.z; y[3] = t0.w; y[4] = t1.x; y[5] = t1.y; y[6] = t1.z; y[7] = t1.w;
        }
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float t = __fmul_rn(x[q], w);
        float r = add_ ? __fadd_rn(t, __fmul_rn(y[q], b)) : t;
        if (kBF16) r = bf16_bits_to_float(pack_bf16x2(r, 0.f) & 0xffffu);          // the value the stream stores
        o[u][q] = r;
        sum += r;
      }
      if (out_x) {
        if (kBF16) {
          uint4 r4;
          r4.x = pack_bf16x2(o[u][0], o[u][1]); r4.y = pack_bf16x2(o[u][2], o[u][3]);
          r4.z = pack_bf16x2(o[u][4], o[u][5]); r4.w = pack_bf16x2(o[u][6], o[u][7]);
          reinterpret_cast<uint4*>(out_x)[v] = r4;
        } else {
          reinterpret_cast<float4*>(out_x)[2 * v] = make_float4(o[u][0], o[u][1], o[u][2], o[u][3]);
          reinterpret_cast<float4*>(out_x)[2 * v + 1] = make_float4(o[u][4], o[u][5], o[u][6], o[u][7]);
        }
      }
    }
    const float mean = warp_sum(sum) * inv_d;
    float sq = 0.f;
#pragma unroll
    for (int u = 0; u < kLnMaxV; ++u) {
      if (lane + 32 * u >= dv) break;
#pragma unroll
      for (int q = 0; q < 8; ++q) { const float d = o[u][q] - mean; sq += d * d; }
    }
    const float rstd = rsqrtf(warp_sum(sq) * inv_d + eps);
#pragma unroll
    for (int u = 0; u < kLnMaxV; ++u) {
      const int vi = lane + 32 * u;
      if (vi >= dv) break;
      const size_t v = row * dv + vi;
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * vi), g1 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * vi + 1);
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * vi), b1 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * vi + 1);
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w}, bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      float y[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) y[q] = (o[u][q] - mean) * rstd * gg[q] + bb[q];
      if (kBF16) {
        uint4 r4;
        r4.x = pack_bf16x2(y[0], y[1]); r4.y = pack_bf16x2(y[2], y[3]); r4.z = pack_bf16x2(y[4], y[5]); r4.w = pack_bf16x2(y[6], y[7]);
        reinterpret_cast<uint4*>(out_ln)[v] = r4;
      } else {
        reinterpret_cast<float4*>(out_ln)[2 * v] = make_float4(y[0], y[1], y[2], y[3]);
        reinterpret_cast<float4*>(out_ln)[2 * v + 1] = make_float4(y[4], y[5], y[6], y[7]);
      }
    }
  }
}

}  // namespace hgl

extern "C" int hgl_token_mask_fuse_ln(const void* src, const void* add, const float* grid, float a, float b, const float* gamma,
                                      const float* beta, float eps, int L1, int M, int D, int dtype, void* out_x, void* out_ln, void* stream) {
  using namespace hgl;
  HGL_REQUIRE(src && out_ln && gamma && beta, "hgl_token_mask_fuse_ln: null pointer");
  HGL_REQUIRE(L1 >= 2 && M >= 0 && D >= 8 && D % 8 == 0 && D <= 256 * kLnMaxV, "hgl_token_mask_fuse_ln: bad shape L1=%d M=%d D=%d (D %% 8, D <= %d)",
              L1, M, D, 256 * kLnMaxV);
  HGL_REQUIRE(dtype == HGL_F32 || dtype == HGL_BF16, "hgl_token_mask_fuse_ln: dtype %d", dtype);
  HGL_REQUIRE(((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(out_x) | reinterpret_cast<uintptr_t>(out_ln) |
                reinterpret_cast<uintptr_t>(add) | reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta)) & 15) == 0,
              "hgl_token_mask_fuse_ln: tensors must be 16-byte aligned");
  if (M == 0) return HGL_OK;
  const size_t rows = (size_t)M * L1;
  const int blocks = (int)std::min<size_t>((rows + 7) / 8, (size_t)sm_count() * 8);
  if (dtype == HGL_BF16)
    token_mask_fuse_ln_kernel<true><<<blocks, 256, 0, (cudaStream_t)stream>>>(src, add, grid, a, b, gamma, beta, eps, L1, M, D, out_x, out_ln);
  else
    token_mask_fuse_ln_kernel<false><<<blocks, 256, 0, (cudaStream_t)stream>>>(src, add, grid, a, b, gamma, beta, eps, L1, M, D, out_x, out_ln);
  return launch_status("hgl_token_mask_fuse_ln");
}

extern "C" int hgl_attn_mask(const float* grid, int M, int L, int heads, uint8_t* out, void* stream) {
  using namespace hgl;
  HGL_REQUIRE(grid && out, "hgl_attn_mask: null pointer");
  HGL_REQUIRE(M >= 0 && L >= 1 && heads >= 1, "hgl_attn_mask: bad shape");
  if (M == 0) return HGL_OK;
  HGL_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "hgl_attn_mask: out must be 16-byte aligned");
  const size_t total = (size_t)M * heads * (L + 1) * (L + 1);
  const int blocks = (int)std::min<size_t>(ceil_div64((int64_t)total, 256 * 16), (size_t)sm_count() * 16);
  attn_mask_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(grid, M, L, heads, out);
  return launch_status("hgl_attn_mask");
}

extern "C" int hgl_attn_bias(const float* grid, int M, int L, float* bias, void* stream) {
  using namespace hgl;
  HGL_REQUIRE(grid && bias, "hgl_attn_bias: null pointer");
  HGL_REQUIRE(M >= 0 && L >= 1, "hgl_attn_bias: bad shape");
  if (M == 0) return HGL_OK;
  attn_bias_kernel<<<ceil_div(M * (L + 1), 256), 256, 0, (cudaStream_t)stream>>>(grid, M, L, bias);
  return launch_status("hgl_attn_bias");
}

extern "C" int hgl_token_mask_fuse(const void* src, const void* add, const float* grid, float a, float b, int L1, int M, int D,
                                   int dtype, int layout, void* out, void* stream) {
  using namespace hgl;
  HGL_REQUIRE(src && out, "hgl_token_mask_fuse: null pointer");
  HGL_REQUIRE(L1 >= 2 && M >= 0 && D >= 8 && D % 8 == 0, "hgl_token_mask_fuse: bad shape L1=%d M=%d D=%d (D %% 8)", L1, M, D);
  HGL_REQUIRE(dtype == HGL_F32 || dtype == HGL_BF16, "hgl_token_mask_fuse: dtype %d", dtype);
  HGL_REQUIRE(layout == HGL_LND || layout == HGL_NLD, "hgl_token_mask_fuse: layout %d", layout);
  HGL_REQUIRE(((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(add)) & 15) == 0,
              "hgl_token_mask_fuse: tensors must be 16-byte aligned");
  if (M == 0) return HGL_OK;
  const size_t total = (size_t)L1 * M * (D / 8);
  const int blocks = (int)std::min<size_t>(ceil_div64((int64_t)total, 256), (size_t)sm_count() * 16);
  if (dtype == HGL_BF16)
    token_mask_fuse_kernel<true><<<blocks, 256, 0, (cudaStream_t)stream>>>(src, add, grid, a, b, L1, M, D, layout, out);
  else
    token_mask_fuse_kernel<false><<<blocks, 256, 0, (cudaStream_t)stream>>>(src, add, grid, a, b, L1, M, D, layout, out);
  return launch_status("hgl_token_mask_fuse");
}
