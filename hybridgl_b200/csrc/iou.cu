// (a13) IoU accounting  --  Compute_IoU utils.py:365-384, called twice per expression (Hybridgl_main.py:171,230).
//
//   I = |pred & gt|, U = |pred | gt|  (integers, bit-exact requirement);  cum_I += I; cum_U += U
// pred is the selected proposal masks[idx].  Integer sums are order independent, so plain 64-bit atomics give
// results that are identical at any grid / world size; `cum` is the 4-vector the multi-GPU sweep all-reduces.
// HBM-bound on 2 * 2 * E * H * W bytes; 16 bytes per load, SIMD-in-word byte logic + dp4a byte sums.
#include "hgl_common.cuh"

namespace hgl {

constexpr int kIouChunks = 8;   // CTAs per (expression, pick)

__global__ void iou_zero_kernel(int64_t* iu, int E) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < E * 4) iu[i] = 0;
}

__device__ __forceinline__ uint32_t nz_bytes(uint32_t v) {   // 0x01 in every byte lane that is non-zero
  // (v | (v + 0x7f7f7f7f per lane)) high bits: classic has-non-zero-byte trick, lane-safe form
  const uint32_t t = ((v & 0x7f7f7f7fu) + 0x7f7f7f7fu) | v;
  return (t >> 7) & 0x01010101u;
}

__global__ void __launch_bounds__(256) iou_kernel(const uint8_t* __restrict__ masks, const uint8_t* __restrict__ target,
                                                  const int64_t* __restrict__ idx_hybrid, const int64_t* __restrict__ idx_final,
                                                  const int32_t* __restrict__ mask_off, const int32_t* __restrict__ expr_off,
                                                  int B, int M, int E, int H, int W, int64_t* __restrict__ iu, int64_t* __restrict__ cum) {
  const int e = blockIdx.y, pick = blockIdx.z, ch = blockIdx.x;
  int b = 0;
  if (expr_off) { while (b + 1 < B && expr_off[b + 1] <= e) ++b; }
  const int n_lo = mask_off ? mask_off[b] : 0;
  const int64_t sel = pick == 0 ? idx_hybrid[e] : idx_final[e];
  const size_t HW = (size_t)H * W;
  const uint8_t* t = target + (size_t)b * HW;
  const size_t per = ((HW + kIouChunks - 1) / kIouChunks + 15) & ~size_t(15);
  const size_t lo = (size_t)ch * per, hi = min(HW, lo + per);
  int ci = 0, cu = 0;
  if (sel >= 0 && lo < hi) {
    const uint8_t* m = masks + (size_t)(n_lo + sel) * HW;
    const bool vec = ((reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(t)) & 15) == 0;
    size_t i = lo + (size_t)threadIdx.x * 16;
    if (vec) {
      for (; i + 16 <= hi; i += (size_t)blockDim.x * 16) {
        const uint4 a = ldg_stream(reinterpret_cast<const uint4*>(m + i));
        const uint4 g = ldg_stream(reinterpret_cast<const uint4*>(t + i));
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, gw[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t pa = nz_bytes(aw[q]), pg = nz_bytes(gw[q]);
          ci = (int)__dp4a(pa & pg, 0x01010101u, (unsigned)ci);
          cu = (int)__dp4a(pa | pg, 0x01010101u, (unsigned)cu);
        }
      }
    }
    // tail (and the unaligned case): bytes this thread's vector loop did not cover
    size_t tail_lo = vec ? lo + ((hi - lo) / 16) * 16 : lo;
    for (size_t k = tail_lo + threadIdx.x; k < hi; k += blockDim.x) {
      const bool pa = m[k] != 0, pg = t[k] != 0;
      ci += pa && pg; cu += pa || pg;
    }
  } else if (sel < 0 && lo < hi) {   // no proposal selected (empty image): pred is all-false
    for (size_t k = lo + threadIdx.x; k < hi; k += blockDim.x) cu += t[k] != 0;
  }
  ci = warp_sum_i(ci); cu = warp_sum_i(cu);
  __shared__ int red[2][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { red[0][warp] = ci; red[1][warp] = cu; }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long si = 0, su = 0;
    for (int w = 0; w < 8; ++w) { si += red[0][w]; su += red[1][w]; }
    atomicAdd(reinterpret_cast<unsigned long long*>(iu + (size_t)e * 4 + pick * 2), (unsigned long long)si);
    atomicAdd(reinterpret_cast<unsigned long long*>(iu + (size_t)e * 4 + pick * 2 + 1), (unsigned long long)su);
    if (cum) {
      atomicAdd(reinterpret_cast<unsigned long long*>(cum + pick * 2), (unsigned long long)si);
      atomicAdd(reinterpret_cast<unsigned long long*>(cum + pick * 2 + 1), (unsigned long long)su);
    }
  }
}

// Same accounting with the prediction taken from the PACKED masks (hgl_pack_masks / hgl_rle_to_bits): 32 pixels per word
// against 32 target bytes, popcounts instead of byte sums.  One thread per word, rows are the unit (W need not be a
// multiple of 32: bits of the last word beyond W are zero by construction of the packed format).
__device__ __forceinline__ uint32_t nz_nibble(uint32_t v) {   // 4 bytes -> 4 bits (byte i non-zero -> bit i)
  const uint32_t t = (((v & 0x7f7f7f7fu) + 0x7f7f7f7fu) | v) & 0x80808080u;
  return (t * 0x00204081u) >> 28;
}

__global__ void __launch_bounds__(256) iou_bits_kernel(const uint32_t* __restrict__ bits, const uint8_t* __restrict__ target,
                                                       const int64_t* __restrict__ idx_hybrid, const int64_t* __restrict__ idx_final,
                                                       const int32_t* __restrict__ mask_off, const int32_t* __restrict__ expr_off,
                                                       int B, int M, int E, int H, int W, int64_t* __restrict__ iu, int64_t* __restrict__ cum) {
  const int e = blockIdx.y, pick = blockIdx.z, ch = blockIdx.x;
  int b = 0;
  if (expr_off) { while (b + 1 < B && expr_off[b + 1] <= e) ++b; }
  const int n_lo = mask_off ? mask_off[b] : 0;
  const int64_t sel = pick == 0 ? idx_hybrid[e] : idx_final[e];
  const int WW = (W + 31) >> 5;
  const int NWORD = H * WW;
  const uint8_t* t = target + (size_t)b * H * W;
  const uint32_t* p = sel >= 0 ? bits + (size_t)(n_lo + sel) * NWORD : nullptr;
  const int per = (NWORD + kIouChunks - 1) / kIouChunks;
  const int lo = ch * per, hi = min(NWORD, lo + per);
  int ci = 0, cu = 0;
  for (int i = lo + (int)threadIdx.x; i < hi; i += blockDim.x) {
    const int y = i / WW, xw = i - y * WW;
    const uint8_t* tp = t + (size_t)y * W + xw * 32;
    const int n = min(32, W - xw * 32);
    uint32_t g = 0u;
    if (n == 32 && (reinterpret_cast<uintptr_t>(tp) & 15) == 0) {
      const uint4 a = ldg_stream(reinterpret_cast<const uint4*>(tp));
      const uint4 c = ldg_stream(reinterpret_cast<const uint4*>(tp) + 1);
      g = nz_nibble(a.x) | (nz_nibble(a.y) << 4) | (nz_nibble(a.z) << 8) | (nz_nibble(a.w) << 12) | (nz_nibble(c.x) << 16) |
          (nz_nibble(c.y) << 20) | (nz_nibble(c.z) << 24) | (nz_nibble(c.w) << 28);
    } else {
      for (int k = 0; k < n; ++k) g |= (tp[k] != 0 ? 1u : 0u) << k;
    }
    const uint32_t a = p ? p[i] : 0u;
    ci += __popc(a & g);
    cu += __popc(a | g);
  }
  ci = warp_sum_i(ci); cu = warp_sum_i(cu);
  __shared__ int red[2][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { red[0][warp] = ci; red[1][warp] = cu; }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long si = 0, su = 0;
    for (int w = 0; w < 8; ++w) { si += red[0][w]; su += red[1][w]; }
    atomicAdd(reinterpret_cast<unsigned long long*>(iu + (size_t)e * 4 + pick * 2), (unsigned long long)si);
    atomicAdd(reinterpret_cast<unsigned long long*>(iu + (size_t)e * 4 + pick * 2 + 1), (unsigned long long)su);
    if (cum) {
      atomicAdd(reinterpret_cast<unsigned long long*>(cum + pick * 2), (unsigned long long)si);
      atomicAdd(reinterpret_cast<unsigned long long*>(cum + pick * 2 + 1), (unsigned long long)su);
    }
  }
}

}  // namespace hgl

extern "C" int hgl_iou_bits(const uint32_t* bits, const uint8_t* target, const int64_t* idx_hybrid, const int64_t* idx_final,
                            const int32_t* mask_off, const int32_t* expr_off, int B, int M, int E, int H, int W,
                            int64_t* iu, int64_t* cum, void* stream) {
  using namespace hgl;
  HGL_REQUIRE(bits && target && idx_hybrid && idx_final && iu, "hgl_iou_bits: null pointer");
  HGL_REQUIRE(B >= 1 && M >= 0 && E >= 0 && H >= 1 && W >= 1, "hgl_iou_bits: bad shape");
  HGL_REQUIRE((mask_off && expr_off) || B == 1, "hgl_iou_bits: mask_off/expr_off required when B > 1");
  if (E == 0) return HGL_OK;
  cudaStream_t st = (cudaStream_t)stream;
  iou_zero_kernel<<<ceil_div(E * 4, 256), 256, 0, st>>>(iu, E);
  int rc = launch_status("hgl_iou_bits(zero)");
  if (rc != HGL_OK) return rc;
  iou_bits_kernel<<<dim3(kIouChunks, E, 2), 256, 0, st>>>(bits, target, idx_hybrid, idx_final, mask_off, expr_off, B, M, E, H, W, iu, cum);
  return launch_status("hgl_iou_bits");
}

extern "C" int hgl_iou(const uint8_t* masks, const uint8_t* target, const int64_t* idx_hybrid, const int64_t* idx_final,
                       const int32_t* mask_off, const int32_t* expr_off, int B, int M, int E, int H, int W,
                       int64_t* iu, int64_t* cum, void* stream) {
  using namespace hgl;
  HGL_REQUIRE(masks && target && idx_hybrid && idx_final && iu, "hgl_iou: null pointer");
  HGL_REQUIRE(B >= 1 && M >= 0 && E >= 0 && H >= 1 && W >= 1, "hgl_iou: bad shape");
  HGL_REQUIRE((mask_off && expr_off) || B == 1, "hgl_iou: mask_off/expr_off required when B > 1");
  if (E == 0) return HGL_OK;
  cudaStream_t st = (cudaStream_t)stream;
  iou_zero_kernel<<<ceil_div(E * 4, 256), 256, 0, st>>>(iu, E);
  int rc = launch_status("hgl_iou(zero)");
  if (rc != HGL_OK) return rc;
  iou_kernel<<<dim3(kIouChunks, E, 2), 256, 0, st>>>(masks, target, idx_hybrid, idx_final, mask_off, expr_off, B, M, E, H, W, iu, cum);
  return launch_status("hgl_iou");
}
