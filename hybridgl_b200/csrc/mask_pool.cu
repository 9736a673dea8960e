// (b3') token-space mask pooling on the 5th-generation tensor cores, fused with L2 normalisation.
//
//   pooled[n, :] = sum_l  w[n, l] * tokens[b(n)][l, :]          w = soft patch-grid mask of proposal n (hgl_mask_grid,
//   out[n, :]    = pooled[n, :] / || pooled[n, :] ||_2              model/backbone.py:160), tokens = dense patch tokens
// This is the masks x tokens x D contraction of the north star: the token-space form of the reference's per-mask pooling
// loop (Hybridgl_main.py:218-223, SURVEY.md Appendix A-2:  S_in = (M~ . F^) . t ), with cosine scoring's normalisation
// (model/backbone.py:79) folded into the epilogue.  The reference pools in pixel space, one mask at a time, in Python.
//
// B200 design (one CTA = one image x 128 masks, 128 threads):
//   A = w   [128 x Kp]  bf16, K-major : converted from f32 once and kept resident in shared memory (canonical 8x16B cores)
//   B = tok [64 x N]    bf16, MN-major (tokens are [L, D] with D contiguous, so D = MMA-N is the contiguous mode): staged in
//                       a 2-slot ring of 64-token chunks; the slot is released by tcgen05.commit on an mbarrier
//   D = acc [128 x N]   f32 in TMEM (two N-wide buffers so that the next tile's MMAs run under the previous epilogue)
//   tcgen05.mma.cta_group::1.kind::f16 (M=128, N<=256, K=16) issued by ONE thread; tcgen05.ld 32x32b in the epilogue: a
//   thread owns a whole output row, so the sum of squares needs no cross-thread reduction.
// Arithmetic intensity is low (2*N*L*D flop over ~2*(N*L + L*D + N*D) bytes): the kernel is sized to stream,
// not to saturate the tensor pipe; see DESIGN.md section 4 for the ceiling.
#include "hgl_common.cuh"

namespace hgl {

constexpr int kMpThreads = 128;
constexpr int kMpM = 128;       // masks per CTA (UMMA M)
constexpr int kMpKC = 64;       // tokens per staged B chunk (4 MMA k-steps)
constexpr int kMpMaxN = 256;    // widest accumulator buffer (UMMA N)

// ---- tcgen05 / TMEM wrappers (PTX ISA 8.6+, sm_100a) ------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// D[tmem] (+)= A[smem] * B[smem], bf16 x bf16 -> f32, one thread issues for the CTA
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on `bar` when every MMA issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 8 consecutive accumulator columns of this thread's TMEM lane (row)
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// shared-memory matrix descriptor, no swizzle (canonical 8-row x 16-byte core matrices); lbo / sbo in bytes
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) | ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) |
         (1ull << 46);   // descriptor version 1 (sm_100); layout_type 0 = SWIZZLE_NONE; base_offset 0
}

struct PoolParams {
  const float* w;            // [M, L] f32 soft grid masks
  const __nv_bfloat16* tok;  // [B, L, D] bf16
  const int32_t* mask_off;   // [B+1] or null (B == 1)
  int B, M, L, D, Kp, Nw, NT;   // Kp = L rounded up to 64, Nw = accumulator width, NT = D / Nw
  int normalize, out_bf16;
  float* scratch;            // [M, D] f32 un-normalised rows (== out when out is f32)
  void* out;                 // [M, D] f32 | bf16
  uint32_t tmem_cols;
};

__global__ void __launch_bounds__(kMpThreads, 1) mask_pool_kernel(const PoolParams p) {
  extern __shared__ __align__(128) uint8_t smp[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.y;
  int n_lo = 0, n_hi = p.M;
  if (p.mask_off) { n_lo = p.mask_off[b]; n_hi = p.mask_off[b + 1]; }
  n_lo += blockIdx.x * kMpM;
  const int rows = min(kMpM, n_hi - n_lo);
  if (rows <= 0) return;                                        // uniform for the CTA
  const int L = p.L, D = p.D, Kp = p.Kp, Nw = p.Nw;

  // shared-memory carve-up
  uint64_t* bars = reinterpret_cast<uint64_t*>(smp);            // slot_free[2], acc_full[2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smp + 32);
  uint8_t* a_s = smp + 128;                                     // [128 x Kp] bf16: core(rg, kc) at rg * a_sbo + kc * 128
  const uint32_t a_sbo = (uint32_t)(Kp / 8) * 128u;             // stride between 8-row groups
  uint8_t* b_s = a_s + (size_t)kMpM * Kp * 2;                   // 2 slots of [64 x Nw] bf16: core(kb, nc) at kb * b_lbo + nc * 128
  const uint32_t b_lbo = (uint32_t)(Nw / 8) * 128u;             // stride between 8-token blocks
  const uint32_t b_slot = (uint32_t)kMpKC * Nw * 2;

  if (warp == 0) tmem_alloc(tmem_slot, p.tmem_cols);
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(bars + i, 1);
    mbar_fence_init();
  }
  // ---- A: soft masks f32 -> bf16 into the canonical K-major layout (rows beyond the image and k >= L are zero)
  for (int t = tid; t < kMpM * (Kp / 8); t += kMpThreads) {
    const int kc = t / kMpM, r = t - kc * kMpM;                 // consecutive threads -> consecutive rows (16-byte smem stride)
    uint4 o = make_uint4(0, 0, 0, 0);
    if (r < rows) {
      const float* src = p.w + (size_t)(n_lo + r) * L + kc * 8;
      float f[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) f[q] = (kc * 8 + q < L) ? __ldg(src + q) : 0.f;
      o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]); o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
    }
    *reinterpret_cast<uint4*>(a_s + (size_t)(r >> 3) * a_sbo + (size_t)kc * 128 + (r & 7) * 16) = o;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // instruction descriptor: D=f32, A=B=bf16, A K-major, B MN-major, N = Nw, M = 128
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (0u << 15) | (1u << 16) | ((uint32_t)(Nw >> 3) << 17) | ((uint32_t)(kMpM >> 4) << 24);
  const __nv_bfloat16* tok = p.tok + (size_t)b * L * D;
  const int NKC = Kp / kMpKC;
  const int my_row = warp * 32 + lane;                          // TMEM lane = output row of this thread
  const bool row_ok = my_row < rows;
  float* srow = p.scratch + (size_t)(n_lo + my_row) * D;
  float sumsq = 0.f;
  int chunk = 0;                                                // running chunk counter: slot = chunk & 1
  for (int nt = 0; nt < p.NT; ++nt) {
    const int buf = nt & 1;
    for (int kc = 0; kc < NKC; ++kc, ++chunk) {
      const int slot = chunk & 1, use = chunk >> 1;
      if (use > 0) mbar_wait(bars + slot, (uint32_t)((use - 1) & 1));        // the MMAs that read this slot are done
      // ---- B chunk: tokens [kc*64, +64) x columns [nt*Nw, +Nw): 16-byte pieces, lane -> (token % 8, 4 column cores)
      uint8_t* bs = b_s + (size_t)slot * b_slot;
      const int cores = Nw / 8;
      for (int t = tid; t < kMpKC * cores; t += kMpThreads) {
        const int kk = t & 7, rest = t >> 3;
        const int nc = rest % cores, kb = rest / cores;
        const int k = kc * kMpKC + kb * 8 + kk;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (k < L) v = __ldg(reinterpret_cast<const uint4*>(tok + (size_t)k * D + nt * Nw + nc * 8));
        *reinterpret_cast<uint4*>(bs + (size_t)kb * b_lbo + (size_t)nc * 128 + kk * 16) = v;
      }
      proxy_fence_async();                                       // generic-proxy smem writes -> visible to the tensor core (async proxy)
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
        const uint32_t a_addr = smem_u32(a_s) + (uint32_t)(kc * (kMpKC / 8)) * 128u;
        const uint32_t b_addr = smem_u32(bs);
#pragma unroll
        for (int s = 0; s < kMpKC / 16; ++s) {
          const uint64_t ad = smem_desc(a_addr + (uint32_t)s * 256u, 128u, a_sbo);              // 2 K-cores, 128 B apart
          const uint64_t bd = smem_desc(b_addr + (uint32_t)s * 2u * b_lbo, b_lbo, 128u);        // 2 token blocks, b_lbo apart
          umma_bf16(tmem_base + (uint32_t)(buf * Nw), ad, bd, idesc, (kc > 0 || s > 0) ? 1u : 0u);
        }
        umma_commit(bars + slot);                                // slot reusable once these MMAs have read it
        if (kc == NKC - 1) umma_commit(bars + 2 + buf);          // accumulator of tile nt complete
      }
    }
    // ---- epilogue of tile nt: TMEM -> registers -> un-normalised row to scratch, running sum of squares
    mbar_wait(bars + 2 + buf, (uint32_t)((nt >> 1) & 1));
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(buf * Nw);
    for (int c = 0; c < Nw; c += 8) {
      float v[8];
      tmem_ld8(taddr + (uint32_t)c, v);
#pragma unroll
      for (int q = 0; q < 8; ++q) sumsq += v[q] * v[q];
      if (row_ok) {
        float4* dst = reinterpret_cast<float4*>(srow + nt * Nw + c);
        dst[0] = make_float4(v[0], v[1], v[2], v[3]);
        dst[1] = make_float4(v[4], v[5], v[6], v[7]);
      }
    }
    tc_fence_before();
    __syncthreads();                                             // every warp has drained buffer `buf` before it is reused (tile nt + 2)
  }

  // ---- L2 normalisation: the thread re-reads its own row (L1/L2 resident) and writes the final dtype
  if (row_ok && (p.normalize || p.out_bf16)) {
    const float inv = p.normalize ? __frcp_rn(sqrtf(sumsq)) : 1.f;
    for (int c = 0; c < D; c += 8) {
      const float4 x0 = *reinterpret_cast<const float4*>(srow + c), x1 = *reinterpret_cast<const float4*>(srow + c + 4);
      const float v[8] = {x0.x * inv, x0.y * inv, x0.z * inv, x0.w * inv, x1.x * inv, x1.y * inv, x1.z * inv, x1.w * inv};
      if (p.out_bf16) {
        uint4 o;
        o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]); o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
        *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)(n_lo + my_row) * D + c) = o;
      } else {
        float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + (size_t)(n_lo + my_row) * D + c);
        dst[0] = make_float4(v[0], v[1], v[2], v[3]);
        dst[1] = make_float4(v[4], v[5], v[6], v[7]);
      }
    }
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, p.tmem_cols);
}

}  // namespace hgl

extern "C" int64_t hgl_mask_pool_workspace_bytes(int M, int D, int out_dtype) {
  if (M < 0 || D < 1) return -1;
  return out_dtype == HGL_BF16 ? (int64_t)M * D * 4 + 256 : 256;   // f32 rows before normalisation when the output is bf16
}

extern "C" int hgl_mask_pool(const float* weights, const void* tokens, const int32_t* mask_off, int B, int M, int max_n, int L, int D,
                             int normalize, int out_dtype, void* out, void* workspace, void* stream) {
  using namespace hgl;
  if (M == 0) return HGL_OK;
  HGL_REQUIRE(weights && tokens && out, "hgl_mask_pool: null pointer");
  HGL_REQUIRE(B >= 1 && M >= 0 && max_n >= 1 && L >= 1, "hgl_mask_pool: bad shape");
  HGL_REQUIRE(mask_off || B == 1, "hgl_mask_pool: mask_off required when B > 1");
  HGL_REQUIRE(out_dtype == HGL_F32 || out_dtype == HGL_BF16, "hgl_mask_pool: out_dtype %d", out_dtype);
  HGL_REQUIRE(D >= 16 && D % 16 == 0, "hgl_mask_pool: D=%d must be a multiple of 16", D);
  HGL_REQUIRE(((reinterpret_cast<uintptr_t>(tokens) | reinterpret_cast<uintptr_t>(out)) & 15) == 0, "hgl_mask_pool: tokens / out must be 16-byte aligned");
  HGL_REQUIRE(out_dtype == HGL_F32 || (workspace && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0), "hgl_mask_pool: bf16 output needs a 16-byte aligned workspace");
  HGL_REQUIRE(B <= 65535, "hgl_mask_pool: B=%d too large for one launch", B);
  PoolParams p;
  p.w = weights; p.tok = reinterpret_cast<const __nv_bfloat16*>(tokens); p.mask_off = mask_off;
  p.B = B; p.M = M; p.L = L; p.D = D;
  p.Kp = ceil_div(L, kMpKC) * kMpKC;
  int nw = std::min(D, kMpMaxN);
  while (D % nw != 0 || nw % 16 != 0) nw -= 16;                   // widest accumulator that tiles D
  p.Nw = nw; p.NT = D / nw;
  p.normalize = normalize ? 1 : 0; p.out_bf16 = out_dtype == HGL_BF16;
  p.out = out;
  p.scratch = p.out_bf16 ? reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255)) : reinterpret_cast<float*>(out);
  uint32_t cols = 32;
  while (cols < (uint32_t)(2 * nw)) cols *= 2;
  p.tmem_cols = cols;
  const size_t smem = 128 + (size_t)kMpM * p.Kp * 2 + 2 * (size_t)kMpKC * nw * 2;
  HGL_REQUIRE(smem <= 227 * 1024, "hgl_mask_pool: L=%d needs %zu B of shared memory (limit 227 KB)", L, smem);
  cudaError_t e = cudaFuncSetAttribute(mask_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("hgl_mask_pool: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return HGL_ECUDA; }
  const int per_image = (B == 1) ? M : std::min(max_n, M);
  dim3 grid(ceil_div(per_image, kMpM), B);
  mask_pool_kernel<<<grid, kMpThreads, smem, (cudaStream_t)stream>>>(p);
  return launch_status("hgl_mask_pool");
}
