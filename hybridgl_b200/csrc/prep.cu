// (a1) per-mask visual-prompt preprocessing, and the packed-mask format every downstream kernel consumes.
//
// Replaces the per-mask Python loop of the reference (Hybridgl_main.py:92-125; utils.py:292-345):
//   global[n] = Normalize_IN( bilinear_S( where(m_n, img, bg) / 255 ) )
//   local[n]  = bilinear_S( where(m_n, Normalize_IN(img/255), clip_pixel_mean) )
// with the non-antialiased bilinear of T.Resize(..., antialias=None) (ATen upsample_bilinear2d, align_corners=False).
//
// B200 design -- three streaming kernels, all HBM-bound, no shared-memory staging (no data reuse inside a kernel):
//   pack   : the byte masks (M*H*W B, the largest input of the whole path) are read exactly ONCE, 16 B per lane fully
//            coalesced, and squeezed to 1 bit/pixel ([M,H,ceil(W/32)] u32).  Every later kernel (prep, mask grid,
//            heat-map pooling) works on this 8x smaller tensor, which stays L2-resident for typical batches.
//   setup  : per IMAGE (not per mask) the two possible answers of every output pixel -- "all four taps inside the mask"
//            (FG) and "all four outside" (BG) -- for the 6 output planes, already in the output dtype, plus the 24 tap
//            bytes of the pixel for the rare boundary case.  12 planes of S*S per image: L2-resident.
//   main   : a thread owns 4 adjacent output pixels and streams over a chunk of masks: 2-4 word loads of mask bits give
//            all 8 taps; two AND/compare decide FG / BG for the whole group (the overwhelmingly common case) and the
//            6 planes leave as 8-byte (bf16) / 16-byte (f32) streaming stores.  Only pixels whose taps straddle the mask
//            boundary re-evaluate the exact per-tap formula.  No barriers, no smem; chunks are small, so the grid has
//            many waves and no tail.
// Arithmetic follows ATen's CPU kernel op for op (explicit __fmaf_rn/__fmul_rn/__fdiv_rn, no re-contraction),
// so f32 output is bit-identical to the reference on the S=224/336 paths (see oracle/hybridgl_oracle.py header).
#include <stdlib.h>

#include "hgl_common.cuh"

namespace hgl {

constexpr int kPrepPx = 4;        // adjacent output pixels per thread
constexpr int kPrepThreads = 256;

struct Taps {
  int i0, d;       // first source index, i1 - i0 (0 or 1)
  float w0, w1;
};

// ATen area_pixel_compute_source_index + compute_source_index_and_lambda (float, align_corners=False)
__device__ __forceinline__ Taps make_taps(int dst, int in_size, int out_size) {
  Taps t;
  if (in_size == out_size) {
    t.i0 = dst; t.d = 0; t.w0 = 1.f; t.w1 = 0.f;
    return t;
  }
  const float scale = __fdiv_rn((float)in_size, (float)out_size);
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

__constant__ float c_in_mean[3] = {0.485f, 0.456f, 0.406f};
__constant__ float c_in_std[3] = {0.229f, 0.224f, 0.225f};
__constant__ float c_clip_mean[3] = {0.48145466f, 0.4578275f, 0.40821073f};

__device__ __forceinline__ float to_unit(uint32_t v) { return __fdiv_rn((float)v, 255.f); }                       // T.ToTensor
__device__ __forceinline__ float to_norm(uint32_t v, int c) { return __fdiv_rn(__fsub_rn(to_unit(v), c_in_mean[c]), c_in_std[c]); }  // + T.Normalize

// ---------------------------------------------------------------------------------------------------------------------
// pack: byte masks -> bit masks
// ---------------------------------------------------------------------------------------------------------------------
// 4 mask bytes -> 4 bits (byte i non-zero -> bit i): high bit of every non-zero byte lane, gathered by one multiply
__device__ __forceinline__ uint32_t nibble_of(uint32_t v) {
  const uint32_t t = (((v & 0x7f7f7f7fu) + 0x7f7f7f7fu) | v) & 0x80808080u;
  return (t * 0x00204081u) >> 28;
}
__device__ __forceinline__ uint32_t half_of(const uint4 v) {   // 16 mask bytes -> 16 bits
  return nibble_of(v.x) | (nibble_of(v.y) << 4) | (nibble_of(v.z) << 8) | (nibble_of(v.w) << 12);
}

// W % 32 == 0 and 16-byte aligned base: the masks are one flat byte stream, word w = pixels [32w, 32w+32).
// lane loads 16 B (fully coalesced 512 B per warp), neighbours pair up through one shuffle, even lanes store.
__global__ void __launch_bounds__(256) pack_masks_flat_kernel(const uint4* __restrict__ src, size_t n16, uint32_t* __restrict__ bits) {
  constexpr int kU = 4;   // independent 16-byte loads in flight per thread
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (kU - 1) * stride < n16; i += kU * stride) {
    uint4 v[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) v[u] = ldg_stream(src + i + u * stride);
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const uint32_t h = half_of(v[u]);
      const uint32_t o = __shfl_xor_sync(0xffffffffu, h, 1);
      if (!(threadIdx.x & 1)) bits[(i + u * stride) >> 1] = h | (o << 16);
    }
  }
  // tail: n16 is even (W % 32 == 0) and the stride is even, so lane pairs stay together
  for (; i < n16; i += stride) {
    const uint32_t h = half_of(ldg_stream(src + i));
    const uint32_t o = __shfl_xor_sync(__activemask(), h, 1);
    if (!(threadIdx.x & 1)) bits[i >> 1] = h | (o << 16);
  }
}

// general widths: one thread per (mask row, 32-pixel word)
__global__ void __launch_bounds__(256) pack_masks_rows_kernel(const uint8_t* __restrict__ masks, size_t rows_total, int W, uint32_t* __restrict__ bits) {
  const int WW = (W + 31) >> 5;
  const size_t total = rows_total * WW;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const size_t row = t / WW;
    const int wd = (int)(t - row * WW);
    const uint8_t* sp = masks + row * W + 32 * wd;
    const int nv = min(32, W - 32 * wd);
    uint32_t word = 0;
    if (nv == 32 && (reinterpret_cast<uintptr_t>(sp) & 3) == 0) {
      const uint32_t* s4 = reinterpret_cast<const uint32_t*>(sp);
#pragma unroll
      for (int q = 0; q < 8; ++q) word |= nibble_of(ldg_stream32(s4 + q)) << (4 * q);
    } else {
      for (int q = 0; q < nv; ++q) word |= (uint32_t)(sp[q] != 0) << q;
    }
    bits[t] = word;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// setup: per-image FG / BG answer planes + tap bytes
// ---------------------------------------------------------------------------------------------------------------------
// planes: [B][12][S*S] in the output dtype; plane k = 3*kind + channel, kind 0 = FG local, 1 = FG global, 2 = BG local, 3 = BG global
// taps:   [B][6][S*S] u32; words 0-2 image bytes, 3-5 background bytes, byte index inside the 12 = tap*3 + channel
template <bool kBF16>
__global__ void __launch_bounds__(256) prep_setup_kernel(const uint8_t* __restrict__ image, const uint8_t* __restrict__ blur, int H, int W, int S,
                                                         void* __restrict__ planes, uint32_t* __restrict__ taps) {
  const int b = blockIdx.y;
  const int px = blockIdx.x * blockDim.x + threadIdx.x;
  const int SS = S * S;
  if (px >= SS) return;
  const int i = px / S, j = px - i * S;
  const Taps ty = make_taps(i, H, S), tx = make_taps(j, W, S);
  const uint8_t* img = image + (size_t)b * H * W * 3;
  const uint8_t* bg = blur ? blur + (size_t)b * H * W * 3 : nullptr;
  const size_t o00 = ((size_t)ty.i0 * W + tx.i0) * 3, o01 = o00 + tx.d * 3;
  const size_t o10 = o00 + (size_t)ty.d * W * 3, o11 = o10 + tx.d * 3;
  uint32_t iw[3] = {0, 0, 0}, bw[3] = {0, 0, 0};
  float v[12];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const uint32_t vi[4] = {img[o00 + c], img[o01 + c], img[o10 + c], img[o11 + c]};
    uint32_t vb[4] = {0, 0, 0, 0};
    if (bg) { vb[0] = bg[o00 + c]; vb[1] = bg[o01 + c]; vb[2] = bg[o10 + c]; vb[3] = bg[o11 + c]; }
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int k = t * 3 + c;
      iw[k >> 2] |= vi[t] << ((k & 3) * 8);
      bw[k >> 2] |= vb[t] << ((k & 3) * 8);
    }
    const float mean = c_in_mean[c], stdv = c_in_std[c], pm = c_clip_mean[c];
    v[0 + c] = bilerp(to_norm(vi[0], c), to_norm(vi[1], c), to_norm(vi[2], c), to_norm(vi[3], c), tx.w0, tx.w1, ty.w0, ty.w1);
    v[3 + c] = __fdiv_rn(__fsub_rn(bilerp(to_unit(vi[0]), to_unit(vi[1]), to_unit(vi[2]), to_unit(vi[3]), tx.w0, tx.w1, ty.w0, ty.w1), mean), stdv);
    v[6 + c] = bilerp(pm, pm, pm, pm, tx.w0, tx.w1, ty.w0, ty.w1);
    v[9 + c] = __fdiv_rn(__fsub_rn(bilerp(to_unit(vb[0]), to_unit(vb[1]), to_unit(vb[2]), to_unit(vb[3]), tx.w0, tx.w1, ty.w0, ty.w1), mean), stdv);
  }
#pragma unroll
  for (int k = 0; k < 12; ++k) {
    const size_t o = ((size_t)b * 12 + k) * SS + px;
    if (kBF16) reinterpret_cast<__nv_bfloat16*>(planes)[o] = __float2bfloat16_rn(v[k]);
    else reinterpret_cast<float*>(planes)[o] = v[k];
  }
#pragma unroll
  for (int w = 0; w < 3; ++w) {
    taps[((size_t)b * 6 + w) * SS + px] = iw[w];
    taps[((size_t)b * 6 + 3 + w) * SS + px] = bw[w];
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// main: stream the masks
// ---------------------------------------------------------------------------------------------------------------------
struct PrepParams {
  const uint32_t* bits;       // [M,H,WW]
  const int32_t* mask_off;
  const void* planes;         // [B,12,S*S]
  const uint32_t* taps;       // [B,6,S*S]
  void* local_out;
  void* global_out;
  int B, M, H, W, S, WW;
  int chunk;                  // masks per CTA
  int narrow;                 // 1 if the 8 taps of 4 adjacent pixels always fit one 32-bit window
};

// PX adjacent output pixels of one plane, packed in the output dtype: NW 32-bit words (8- or 16-byte stores)
template <bool kBF16, int PX>
struct Pack {
  static constexpr int NW = kBF16 ? PX / 2 : PX;
  static_assert(NW == 2 || NW == 4, "a pack is one 8- or 16-byte store");
  uint32_t w[NW];
  __device__ __forceinline__ void load(const void* planes, size_t elem_off) {
    const uint8_t* ptr = reinterpret_cast<const uint8_t*>(planes) + elem_off * (kBF16 ? 2 : 4);
    if (NW == 2) { const uint2 v = *reinterpret_cast<const uint2*>(ptr); w[0] = v.x; w[1] = v.y; }
    else { const uint4 v = *reinterpret_cast<const uint4*>(ptr); w[0] = v.x; w[1] = v.y; w[NW - 2] = v.z; w[NW - 1] = v.w; }
  }
  // replace pixel q by value v
  __device__ __forceinline__ void put(int q, float v) {
    if (kBF16) {
      const uint32_t h = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v));
      uint32_t& x = w[q >> 1];
      x = (q & 1) ? ((x & 0x0000ffffu) | (h << 16)) : ((x & 0xffff0000u) | h);
    } else {
      w[q % NW] = __float_as_uint(v);
    }
  }
  __device__ __forceinline__ void store(uint8_t* ptr) const {   // ptr: address of the first of the PX pixels
    if (NW == 2) stg_stream(reinterpret_cast<uint2*>(ptr), make_uint2(w[0], w[1]));
    else stg_stream(reinterpret_cast<uint4*>(ptr), make_uint4(w[0], w[1], w[NW - 2], w[NW - 1]));
  }
};

// boundary pixel: exact per-tap evaluation (Hybridgl_main.py:106-121 restricted to the 4 taps of one output pixel).
// Only pixels whose taps straddle the mask outline come here, through the dense fix-up pass at the end of each CTA.
__device__ __forceinline__ void prep_boundary_pixel(const uint32_t* __restrict__ taps, size_t tap0, int SS, uint32_t code,
                                                    float wx0, float wx1, float wy0, float wy1, float* __restrict__ out6) {
  uint32_t iw[3], bw[3];
#pragma unroll
  for (int w = 0; w < 3; ++w) {
    iw[w] = __ldg(taps + tap0 + (size_t)w * SS);
    bw[w] = __ldg(taps + tap0 + (size_t)(3 + w) * SS);
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float gv[4], lv[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int kk = t * 3 + c;
      const uint32_t vi = (iw[kk >> 2] >> ((kk & 3) * 8)) & 0xffu;
      const uint32_t vb = (bw[kk >> 2] >> ((kk & 3) * 8)) & 0xffu;
      const bool in = (code >> t) & 1u;
      gv[t] = to_unit(in ? vi : vb);
      lv[t] = in ? to_norm(vi, c) : c_clip_mean[c];
    }
    out6[3 + c] = __fdiv_rn(__fsub_rn(bilerp(gv[0], gv[1], gv[2], gv[3], wx0, wx1, wy0, wy1), c_in_mean[c]), c_in_std[c]);
    out6[c] = bilerp(lv[0], lv[1], lv[2], lv[3], wx0, wx1, wy0, wy1);
  }
}

constexpr int kPrepQueue = 3072;    // boundary pixels a CTA can defer (entry = k << 16 | local pixel << 4 | tap code)

// One CTA = 256 pixel groups (PX adjacent pixels each) of one image x `chunk` masks.
//   P0  the bit rows the tile touches, for all masks of the chunk, are copied to shared memory with fully independent
//       coalesced loads (one exposed memory latency per CTA instead of one per mask), the FG/BG answers go to registers
//   P1  per mask: window from shared memory, FG/BG decision for the whole group, stores.  No global load in the loop.
//       Pixels on the mask outline get a placeholder and are queued.
//   P2  dense fix-up: one thread per queued pixel evaluates the exact formula and patches the freshly written line (L2 hit)
template <bool kBF16, int PX>
__global__ void __launch_bounds__(kPrepThreads, (kBF16 && PX == 4) ? 4 : 2) prep_main_kernel(const PrepParams p) {
  extern __shared__ __align__(16) uint32_t sm_prep[];
  __shared__ int q_count;
  constexpr int kTilePx = kPrepThreads * PX;
  uint32_t* queue = sm_prep;                       // [kPrepQueue]
  uint32_t* stage = sm_prep + kPrepQueue;          // [chunk][nrows][WW]

  const int H = p.H, W = p.W, S = p.S, WW = p.WW;
  const int SS = S * S, tpr = S / PX;
  const int b = blockIdx.y, tid = threadIdx.x;
  int n_lo = 0, n_hi = p.M;
  if (p.mask_off) { n_lo = p.mask_off[b]; n_hi = p.mask_off[b + 1]; }
  n_lo += blockIdx.z * p.chunk;
  n_hi = min(n_hi, n_lo + p.chunk);
  if (n_lo >= n_hi) return;                                          // uniform for the whole CTA
  const int cnt = n_hi - n_lo;

  const int tile0 = blockIdx.x * kTilePx;                            // first pixel of the tile
  const int row_first = tile0 / S, row_last = min(tile0 + kTilePx - 1, SS - 1) / S;
  const int ylo = make_taps(row_first, H, S).i0;
  const Taps tl = make_taps(row_last, H, S);
  const int nrows = tl.i0 + tl.d - ylo + 1;
  const size_t mask_words = (size_t)H * WW;
  {  // ---- P0
    const int per_mask = nrows * WW;
    const uint32_t* src = p.bits + (size_t)n_lo * mask_words + (size_t)ylo * WW;
    for (int t = tid; t < cnt * per_mask; t += kPrepThreads) {
      const int k = t / per_mask, o = t - k * per_mask;
      stage[t] = __ldg(src + (size_t)k * mask_words + o);
    }
    if (tid == 0) q_count = 0;
  }
  const int pg = blockIdx.x * kPrepThreads + tid;                    // pixel group inside the image
  const bool live = pg < SS / PX;
  const int i = live ? pg / tpr : 0, j0 = live ? (pg - i * tpr) * PX : 0;
  const Taps ty = make_taps(i, H, S);
  const int bx = make_taps(j0, W, S).i0;
  uint32_t tapmask = 0;
  uint64_t xpack = 0;                // per pixel 5 bits (x0 - bx) + 1 bit (dx)
  if (p.narrow) {
#pragma unroll
    for (int q = 0; q < PX; ++q) {
      const Taps tx = make_taps(j0 + q, W, S);
      tapmask |= (1u << (tx.i0 - bx)) | (1u << (tx.i0 + tx.d - bx));
      xpack |= (uint64_t)((tx.i0 - bx) | (tx.d << 5)) << (6 * q);
    }
  }
  const int wi = bx >> 5, sh = bx & 31;
  const int wi1 = min(wi + 1, WW - 1);                               // clamped: bits beyond the row are never selected
  Pack<kBF16, PX> ans[12];
  const size_t px0 = (size_t)i * S + j0;
  if (live) {
#pragma unroll
    for (int k = 0; k < 12; ++k) ans[k].load(p.planes, ((size_t)b * 12 + k) * SS + px0);
  }
  __syncthreads();

  // ---- P1
  constexpr size_t kElem = kBF16 ? 2 : 4;
  const size_t plane_bytes = (size_t)SS * kElem;
  uint8_t* lp = reinterpret_cast<uint8_t*>(p.local_out) + ((size_t)n_lo * 3 * SS + px0) * kElem;
  uint8_t* gp = reinterpret_cast<uint8_t*>(p.global_out) + ((size_t)n_lo * 3 * SS + px0) * kElem;
  const uint32_t* s0 = stage + (ty.i0 - ylo) * WW;                   // tap rows inside the stage of mask 0
  const uint32_t* s1 = s0 + ty.d * WW;
  const int stride = nrows * WW;
  for (int k = 0; live && k < cnt; ++k, s0 += stride, s1 += stride) {
    uint32_t a0 = 0, a1 = 0;
    bool uniform = false, inside = false;
    if (p.narrow) {
      a0 = __funnelshift_r(s0[wi], s0[wi1], sh);
      a1 = __funnelshift_r(s1[wi], s1[wi1], sh);
      const uint32_t t0 = a0 & tapmask, t1 = a1 & tapmask;
      inside = (t0 == tapmask) && (t1 == tapmask);
      uniform = inside || ((t0 | t1) == 0u);
    }
    Pack<kBF16, PX> ol[3], og[3];
    if (uniform) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
#pragma unroll
        for (int w = 0; w < Pack<kBF16, PX>::NW; ++w) {
          ol[c].w[w] = inside ? ans[0 + c].w[w] : ans[6 + c].w[w];
          og[c].w[w] = inside ? ans[3 + c].w[w] : ans[9 + c].w[w];
        }
      }
    } else {
      // mixed group: per-pixel FG/BG merge with bit masks; true boundary pixels keep the BG placeholder and are queued
      uint32_t codes = 0;                 // 4 bits per pixel
#pragma unroll
      for (int q = 0; q < PX; ++q) {
        uint32_t code;
        if (p.narrow) {
          const int oa = (int)(xpack >> (6 * q)) & 31, ob = oa + ((int)(xpack >> (6 * q + 5)) & 1);
          code = ((a0 >> oa) & 1u) | (((a0 >> ob) & 1u) << 1) | (((a1 >> oa) & 1u) << 2) | (((a1 >> ob) & 1u) << 3);
        } else {
          const Taps tx = make_taps(j0 + q, W, S);
          const int xa = tx.i0, xb = tx.i0 + tx.d;
          code = ((s0[xa >> 5] >> (xa & 31)) & 1u) | (((s0[xb >> 5] >> (xb & 31)) & 1u) << 1) |
                 (((s1[xa >> 5] >> (xa & 31)) & 1u) << 2) | (((s1[xb >> 5] >> (xb & 31)) & 1u) << 3);
        }
        codes |= code << (4 * q);
        if (code != 0u && code != 15u) {
          const int slot = atomicAdd(&q_count, 1);
          if (slot < kPrepQueue) queue[slot] = ((uint32_t)k << 16) | ((uint32_t)(tid * PX + q) << 4) | code;
        }
      }
#pragma unroll
      for (int w = 0; w < Pack<kBF16, PX>::NW; ++w) {
        uint32_t m;                        // all-ones in the lanes of pixels that are fully inside
        if (kBF16) m = ((((codes >> (8 * w)) & 15u) == 15u) ? 0x0000ffffu : 0u) | ((((codes >> (8 * w + 4)) & 15u) == 15u) ? 0xffff0000u : 0u);
        else m = (((codes >> (4 * w)) & 15u) == 15u) ? 0xffffffffu : 0u;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          ol[c].w[w] = (ans[0 + c].w[w] & m) | (ans[6 + c].w[w] & ~m);
          og[c].w[w] = (ans[3 + c].w[w] & m) | (ans[9 + c].w[w] & ~m);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      ol[c].store(lp + (size_t)c * plane_bytes);
      og[c].store(gp + (size_t)c * plane_bytes);
    }
    lp += 3 * plane_bytes; gp += 3 * plane_bytes;
  }

  // ---- P2: dense fix-up of the queued boundary pixels
  __syncthreads();
  const int nq_all = q_count;
  const int nq = min(nq_all, kPrepQueue);
  for (int e = tid; e < nq; e += kPrepThreads) {
    const uint32_t ent = queue[e];
    const uint32_t code = ent & 15u;
    const int lpx = (ent >> 4) & 0xfff, k = ent >> 16;
    const int gpx = tile0 + lpx;                       // pixel inside the image
    const int pi = gpx / S, pj = gpx - pi * S;
    const Taps tyy = make_taps(pi, H, S), txx = make_taps(pj, W, S);
    float o6[6];
    prep_boundary_pixel(p.taps, (size_t)b * 6 * SS + gpx, SS, code, txx.w0, txx.w1, tyy.w0, tyy.w1, o6);
    const size_t o = ((size_t)(n_lo + k) * 3) * SS + gpx;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (kBF16) {
        reinterpret_cast<__nv_bfloat16*>(p.local_out)[o + (size_t)c * SS] = __float2bfloat16_rn(o6[c]);
        reinterpret_cast<__nv_bfloat16*>(p.global_out)[o + (size_t)c * SS] = __float2bfloat16_rn(o6[3 + c]);
      } else {
        reinterpret_cast<float*>(p.local_out)[o + (size_t)c * SS] = o6[c];
        reinterpret_cast<float*>(p.global_out)[o + (size_t)c * SS] = o6[3 + c];
      }
    }
  }
  if (nq_all > kPrepQueue) {
    // queue overflow (pathological outlines): rescan the whole tile x chunk, skipping what the queue already covered is
    // not possible, so every boundary pixel is simply re-evaluated in place (idempotent)
    for (int t = tid; t < cnt * kTilePx; t += kPrepThreads) {
      const int k = t / kTilePx, lpx = t - k * kTilePx;
      const int gpx = tile0 + lpx;
      if (gpx >= SS) continue;
      const int pi = gpx / S, pj = gpx - pi * S;
      const Taps tyy = make_taps(pi, H, S), txx = make_taps(pj, W, S);
      const uint32_t* q0 = stage + (size_t)k * stride + (tyy.i0 - ylo) * WW;
      const uint32_t* q1 = q0 + tyy.d * WW;
      const int xa = txx.i0, xb = txx.i0 + txx.d;
      const uint32_t code = ((q0[xa >> 5] >> (xa & 31)) & 1u) | (((q0[xb >> 5] >> (xb & 31)) & 1u) << 1) |
                            (((q1[xa >> 5] >> (xa & 31)) & 1u) << 2) | (((q1[xb >> 5] >> (xb & 31)) & 1u) << 3);
      if (code == 0u || code == 15u) continue;
      float o6[6];
      prep_boundary_pixel(p.taps, (size_t)b * 6 * SS + gpx, SS, code, txx.w0, txx.w1, tyy.w0, tyy.w1, o6);
      const size_t o = ((size_t)(n_lo + k) * 3) * SS + gpx;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        if (kBF16) {
          reinterpret_cast<__nv_bfloat16*>(p.local_out)[o + (size_t)c * SS] = __float2bfloat16_rn(o6[c]);
          reinterpret_cast<__nv_bfloat16*>(p.global_out)[o + (size_t)c * SS] = __float2bfloat16_rn(o6[3 + c]);
        } else {
          reinterpret_cast<float*>(p.local_out)[o + (size_t)c * SS] = o6[c];
          reinterpret_cast<float*>(p.global_out)[o + (size_t)c * SS] = o6[3 + c];
        }
      }
    }
  }
}

struct PrepWs {
  void* planes;
  uint32_t* taps;
  size_t bytes;
};
static PrepWs prep_carve(void* ws, int B, int S, int out_dtype) {
  PrepWs w;
  const size_t SS = (size_t)S * S, elem = out_dtype == HGL_BF16 ? 2 : 4;
  size_t off = 0;
  auto take = [&](size_t n) { size_t o = off; off += (n + 255) & ~size_t(255); return o; };
  uint8_t* base = reinterpret_cast<uint8_t*>(ws);
  w.planes = base + take((size_t)B * 12 * SS * elem);
  w.taps = reinterpret_cast<uint32_t*>(base + take((size_t)B * 6 * SS * 4));
  w.bytes = off;
  return w;
}

}  // namespace hgl

extern "C" int hgl_pack_masks(const uint8_t* masks, int M, int H, int W, uint32_t* bits, void* stream) {
  using namespace hgl;
  if (M == 0) return HGL_OK;
  HGL_REQUIRE(masks && bits, "hgl_pack_masks: null pointer");
  HGL_REQUIRE(M > 0 && H >= 1 && W >= 1, "hgl_pack_masks: bad shape");
  const size_t rows = (size_t)M * H;
  cudaStream_t st = (cudaStream_t)stream;
  if ((W & 31) == 0 && (reinterpret_cast<uintptr_t>(masks) & 15) == 0) {
    const size_t n16 = rows * W / 16;
    const int blocks = (int)std::min<size_t>((n16 + 256 * 4 - 1) / (256 * 4), (size_t)sm_count() * 8);
    pack_masks_flat_kernel<<<std::max(blocks, 1), 256, 0, st>>>(reinterpret_cast<const uint4*>(masks), n16, bits);
  } else {
    const size_t total = rows * ((W + 31) >> 5);
    const int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)sm_count() * 16);
    pack_masks_rows_kernel<<<blocks, 256, 0, st>>>(masks, rows, W, bits);
  }
  return launch_status("hgl_pack_masks");
}

extern "C" int64_t hgl_prep_workspace_bytes(int B, int S, int out_dtype) {
  if (B < 1 || S < 4) return -1;
  return (int64_t)hgl::prep_carve(nullptr, B, S, out_dtype).bytes;
}

extern "C" int hgl_prep(const uint8_t* image, const uint8_t* blur, const uint32_t* bits, const int32_t* mask_off,
                        int B, int M, int max_n, int H, int W, int S, int bg_mode, int out_dtype,
                        void* local_out, void* global_out, void* workspace, void* stream) {
  using namespace hgl;
  if (M == 0 && B >= 1) return HGL_OK;   // nothing to do (empty tensors have null data pointers)
  HGL_REQUIRE(image && bits && local_out && global_out && workspace, "hgl_prep: null pointer");
  HGL_REQUIRE(bg_mode == HGL_BG_BLUR || bg_mode == HGL_BG_BLACK, "hgl_prep: bg_mode %d", bg_mode);
  HGL_REQUIRE(bg_mode != HGL_BG_BLUR || blur, "hgl_prep: blur frame required for HGL_BG_BLUR");
  HGL_REQUIRE(out_dtype == HGL_F32 || out_dtype == HGL_BF16, "hgl_prep: out_dtype %d", out_dtype);
  HGL_REQUIRE(B >= 1 && M >= 0 && H >= 1 && W >= 1 && max_n >= 1, "hgl_prep: bad shape B=%d M=%d H=%d W=%d max_n=%d", B, M, H, W, max_n);
  HGL_REQUIRE(S >= 4 && S % 4 == 0 && S <= 1024, "hgl_prep: S=%d must be a multiple of 4 in [4,1024]", S);
  HGL_REQUIRE(mask_off || B == 1, "hgl_prep: mask_off required when B > 1");
  HGL_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "hgl_prep: workspace must be 256-byte aligned");
  HGL_REQUIRE(((reinterpret_cast<uintptr_t>(local_out) | reinterpret_cast<uintptr_t>(global_out)) & 15) == 0,
              "hgl_prep: outputs must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int SS = S * S;
  PrepWs ws = prep_carve(workspace, B, S, out_dtype);
  const uint8_t* bgp = bg_mode == HGL_BG_BLUR ? blur : nullptr;
  if (out_dtype == HGL_BF16)
    prep_setup_kernel<true><<<dim3(ceil_div(SS, 256), B), 256, 0, st>>>(image, bgp, H, W, S, ws.planes, ws.taps);
  else
    prep_setup_kernel<false><<<dim3(ceil_div(SS, 256), B), 256, 0, st>>>(image, bgp, H, W, S, ws.planes, ws.taps);
  int rc = launch_status("hgl_prep(setup)");
  if (rc != HGL_OK) return rc;

  PrepParams p;
  p.bits = bits; p.mask_off = mask_off; p.planes = ws.planes; p.taps = ws.taps;
  p.local_out = local_out; p.global_out = global_out;
  p.B = B; p.M = M; p.H = H; p.W = W; p.S = S; p.WW = (W + 31) >> 5;
  const double sx = (double)W / (double)S;
  // pixels per thread: 8 (16-byte bf16 stores) when the 16 taps still fit one 32-bit window, else 4
  const int px = (out_dtype == HGL_BF16 && S % 8 == 0 && (int)(7.0 * sx) + 3 <= 31) ? 8 : kPrepPx;
  p.narrow = ((int)((px - 1) * sx) + 3 <= 31) ? 1 : 0;
  // chunk = masks per CTA: bounded by the shared-memory stage of their bit rows, small enough for several waves
  const int gx = ceil_div(SS / px, kPrepThreads);
  const int per_image = (B == 1) ? M : std::min(max_n, M);
  const double sy = (double)H / (double)S;
  const int tile_rows = (kPrepThreads * px + S - 1) / S + 1;                 // output rows a tile can touch
  const int stage_rows = std::min(H, (int)(tile_rows * sy) + 3);            // source rows behind them
  const size_t per_mask_bytes = (size_t)stage_rows * p.WW * 4;
  HGL_REQUIRE(per_mask_bytes <= 96 * 1024, "hgl_prep: frame %dx%d too large for the bit-row stage", H, W);
  int chunk = (int)std::min<size_t>(64, (40 * 1024) / per_mask_bytes);
  chunk = std::max(chunk, 1);
  const long slots = (long)sm_count() * 3;
  while (chunk > 8 && (long)gx * B * ceil_div(per_image, chunk) < 3 * slots) chunk = (chunk * 3) / 4;
  if (const char* ov = getenv("HGL_PREP_CHUNK")) chunk = std::max(1, atoi(ov));   // tuning hook
  chunk = std::min(chunk, std::max(1, (int)((200 * 1024) / per_mask_bytes)));
  p.chunk = chunk;
  dim3 grid(gx, B, ceil_div(per_image, chunk));
  HGL_REQUIRE(grid.z <= 65535 && grid.y <= 65535, "hgl_prep: batch too large for one launch (B=%d, max_n=%d)", B, max_n);
  const size_t smem = (size_t)kPrepQueue * 4 + (size_t)chunk * per_mask_bytes;
  auto launch = [&](auto kern) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<grid, kPrepThreads, smem, st>>>(p);
  };
  if (out_dtype == HGL_BF16) {
    if (px == 8) launch(prep_main_kernel<true, 8>);
    else launch(prep_main_kernel<true, kPrepPx>);
  } else {
    launch(prep_main_kernel<false, kPrepPx>);
  }
  return launch_status("hgl_prep(main)");
}
