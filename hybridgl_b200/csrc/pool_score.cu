// (b3') + (a6)-(a9),(a12): token-space mask pooling on the 5th-generation tensor cores, L2 normalisation, cosine scoring
// against the text ensemble, spatial-relationship re-ranking and the per-expression argmax -- ONE kernel, one launch.
//
//   pooled[n, :] = sum_l  w[n, l] * tokens[b(n)][l, :]        w = soft patch-grid mask of proposal n (model/backbone.py:160)
//   f^[n, :]     = pooled[n, :] / || pooled[n, :] ||_2         (model/backbone.py:79)
//   s[e, n]      = scale * f^[n] . t^[e]                       (model/backbone.py:80-85, Hybridgl_main.py:153-166)
//   ... soft-max / top-k / relation_boxes / blend / argmax     (Hybridgl_main.py:168-196, 225-227; select_tail.cuh)
// The masks x tokens x D contraction is the token-space form of the reference's per-mask pooling loop
// (Hybridgl_main.py:218-223).  The pooled rows never leave the SM: the dot products with the text
// vectors and the row norms are taken straight from the TMEM accumulators (f32), so the scores carry no bf16 rounding of the
// features; writing the normalised rows out is optional (hgl_mask_pool is the same kernel with no expressions).
//
// B200 design
//   grid = (NT column tiles, B images), thread-block cluster = the NT column tiles of one image (NT <= 8).
//   A = w    [128 x 64] per stage, bf16 K-major, no swizzle (8x16B core matrices).  The f32 soft masks are split into
//            hi = bf16(w) and lo = bf16(w - hi) and BOTH are multiplied (two MMAs per k-step): w is reproduced to 2^-17
//            relative, so the result equals the f32 x bf16 product to ~1e-5 -- flops are free here, precision is not.
//   B = tok  [64 tokens x Nw] per stage, bf16 MN-major (tokens are [L, D] with D contiguous), staged by TMA
//            (cp.async.bulk.tensor.3d over a [B, L, D] tensor map, 64x64 boxes, 128-byte swizzle; rows >= L are zero-filled
//            by the TMA unit) on a full/free mbarrier ring.
//   D = acc  [128 x Nw] f32 per row tile in TMEM (up to 4 row tiles = 512 proposals per image share the B stream).
//   tcgen05.mma.cta_group::1.kind::f16 (M=128, N=Nw, K=16) issued by ONE thread; tcgen05.commit releases ring slots.
//   Epilogue: tcgen05.ld 32x32b -- a thread owns one proposal row -- accumulates |row|^2 and the dots with <= 4 text / 4
//   negative vectors (this CTA's column slice, in shared memory); partial sums meet in the cluster's rank-0 CTA through
//   distributed shared memory (mapa + ld.shared::cluster), which finishes the scores and runs the selection tail.
// Arithmetic intensity is low (2*N*L*D flop over ~2*(N*L + L*D) bytes): the kernel is a latency chain
// (TMA -> MMA -> TMEM -> DSMEM -> tail), sized to run wide (B*NT CTAs) rather than to saturate the tensor pipe.
#include <cuda.h>   // CUtensorMap and its enums only; the encoder is fetched through cudaGetDriverEntryPoint (no libcuda link)

#include <algorithm>

#include "hgl_common.cuh"
#include "select_tail.cuh"

namespace hgl {

constexpr int kPsThreads = 256;   // 8 warps: all stage A; thread 0 issues TMA + MMA; warps w and w+4 share TMEM lanes 32*(w%4)..
constexpr int kPsM = 128;         // proposals per row tile (UMMA M)
constexpr int kPsKC = 64;         // tokens per ring stage (4 MMA k-steps)
constexpr int kPsER = 4;          // expressions per scoring round
constexpr int kPsMaxTiles = 4;    // row tiles per image (512 proposals)
constexpr int kPsMaxStages = 8;
constexpr uint32_t kPsAHalf = kPsM * kPsKC * 2;   // one bf16 [128 x 64] operand block (hi or lo)
constexpr uint32_t kPsBBox = kPsKC * 64 * 2;      // one TMA box: 64 tokens x 64 columns bf16
constexpr int kPsPiece = 64;      // output columns staged through shared memory per round of the feature pass
// shared-memory map (offsets from the 1024-byte aligned base)
constexpr uint32_t kOffFull = 0, kOffFree = 64, kOffAcc = 128, kOffTmem = 192, kOffTnp = 256, kOffTn = 320, kOffText = 512;
constexpr uint32_t kOffRing = kOffText + 2 * kPsER * 256 * 4;          // text slice: [2*kPsER][Nw <= 256] f32 -> ring at 8704 ...
constexpr uint32_t kRingBase = (kOffRing + 1023) & ~1023u;             // ... rounded to the swizzle atom: 9216
// epilogue scratch, aliased onto the ring once every MMA has completed
constexpr uint32_t kEpPart = 0;                                                    // [tiles][1 + 2*kPsER][128] f32
constexpr uint32_t kEpSc = kEpPart + kPsMaxTiles * (1 + 2 * kPsER) * kPsM * 4;     // [2][kPsER][tiles*128] f32
constexpr uint32_t kEpPicks = kEpSc + 2 * kPsER * kPsMaxTiles * kPsM * 4;          // [8 warps][9] int
constexpr uint32_t kEpInv = kEpPicks + 512;                                        // [tiles*128] f32
constexpr uint32_t kEpStage = kEpInv + kPsMaxTiles * kPsM * 4;                     // [128][kPsPiece + 1] f32; also the half-1 partials
constexpr uint32_t kEpEnd = kEpStage + kPsM * (kPsPiece + 1) * 4;

// ---- tcgen05 / TMEM / TMA wrappers (PTX ISA 8.6+, sm_100a) ------------------------------------------------------------
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
// 16 consecutive accumulator columns of this thread's TMEM lane (row)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                 "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// shared-memory matrix descriptor (sm_100 version 1); layout: 0 = no swizzle (8-row x 16-byte core matrices), 2 = 128-byte swizzle
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) | ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) |
         (1ull << 46) | ((uint64_t)layout << 61);
}
// one box of a 3-D tiled tensor map -> shared memory (TMA; SASS UTMALDG), completes on `bar`
__device__ __forceinline__ void tma_load_3d(uint32_t dst_smem, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst_smem),
               "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
               : "memory");
}

struct PoolScoreParams {
  const float* w;              // [M, L] f32 soft grid masks
  const int32_t* mask_off;     // [B+1] or null (B == 1)
  const int32_t* expr_off;     // [B+1] or null (B == 1)
  int B, M, E, L, D, max_n;
  int Nw, NT, NKC, stages, boxes;   // columns per CTA, CTAs per cluster, token chunks, ring depth, TMA boxes per stage (Nw / 64)
  uint32_t stage_bytes, tmem_cols;
  // optional pooled rows
  void* out; int out_bf16, normalize;
  // scoring (E == 0: pooling only)
  const float* sent; const float* noun; const float* others; const int32_t* other_off;
  float scale, r, one_minus_r;
  TailArgs tail;
};

__global__ void __launch_bounds__(kPsThreads, 1) pool_score_kernel(const __grid_constant__ CUtensorMap tmap, const PoolScoreParams p) {
  extern __shared__ uint8_t ps_smem_raw[];
  uint8_t* smem = ps_smem_raw + ((1024u - (smem_u32(ps_smem_raw) & 1023u)) & 1023u);      // swizzle atoms need 1024-byte alignment
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();                         // == blockIdx.x: the cluster spans the grid's x extent
  const int b = blockIdx.y;
  int n_lo = 0, n_hi = p.M, e_lo = 0, e_hi = p.E;
  if (p.mask_off) { n_lo = p.mask_off[b]; n_hi = p.mask_off[b + 1]; }
  if (p.expr_off) { e_lo = p.expr_off[b]; e_hi = p.expr_off[b + 1]; }
  const int n = min(max(n_hi - n_lo, 0), p.max_n);                 // uniform for the whole cluster
  const int EB = max(e_hi - e_lo, 0);
  const int L = p.L, D = p.D, Nw = p.Nw, S = p.stages, NKC = p.NKC;
  const int col0 = (int)rank * Nw;

  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + kOffFull);
  uint64_t* bar_free = reinterpret_cast<uint64_t*>(smem + kOffFree);
  uint64_t* bar_acc = reinterpret_cast<uint64_t*>(smem + kOffAcc);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffTmem);
  float* tnp = reinterpret_cast<float*>(smem + kOffTnp);          // [2*kPsER] this CTA's partial |text|^2 (its column slice)
  float* tn_s = reinterpret_cast<float*>(smem + kOffTn);          // [2*kPsER] rank 0: text norms
  float* text = reinterpret_cast<float*>(smem + kOffText);        // [2*kPsER][Nw]: rows 0..3 text ensemble, 4..7 negatives
  uint8_t* ring = smem + kRingBase;
  float* part = reinterpret_cast<float*>(ring + kEpPart);
  float* sc = reinterpret_cast<float*>(ring + kEpSc);
  int* picks = reinterpret_cast<int*>(ring + kEpPicks);
  float* inv_s = reinterpret_cast<float*>(ring + kEpInv);
  float* stage = reinterpret_cast<float*>(ring + kEpStage);

  if (n == 0) {
    // an image without proposals: nothing to pool; the selection tail still defines its outputs (-1 picks), like hgl_score_select
    if (rank == 0) {
      float* s0 = reinterpret_cast<float*>(smem + kOffText);
      for (int e = e_lo + warp; e < e_hi; e += kPsThreads / 32) select_tail_warp(p.tail, e, 0, n_lo, s0, s0, picks + warp * 9, lane);
    }
    return;
  }
  const int tiles = (n + kPsM - 1) / kPsM;
  const int total = tiles * NKC;

  if (tid == 0) {
    for (int i = 0; i < kPsMaxStages; ++i) { mbar_init(bar_full + i, 1); mbar_init(bar_free + i, 1); }
    mbar_init(bar_acc, 1);
    mbar_fence_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
  }
  if (warp == 0) tmem_alloc(tmem_slot, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // B stage of iteration `it` (row tile it / NKC, token chunk it % NKC): Nw/64 boxes of 64 tokens x 64 columns
  const CUtensorMap* tmap_ptr = &tmap;
  auto issue_b = [&](int it) {
    const int slot = it % S, kc = it % NKC;
    mbar_expect_tx(bar_full + slot, (uint32_t)p.boxes * kPsBBox);
    const uint32_t dst = smem_u32(ring + (size_t)slot * p.stage_bytes + 2 * kPsAHalf);
    for (int i = 0; i < p.boxes; ++i) tma_load_3d(dst + (uint32_t)i * kPsBBox, tmap_ptr, col0 + 64 * i, kc * kPsKC, b, bar_full + slot);
  };
  if (tid == 0)
    for (int it = 0; it < min(S, total); ++it) issue_b(it);

  // (a6) text side of one scoring round, this CTA's column slice: warp w -> expression w % 4, ensemble (w < 4) or negatives
  // (Hybridgl_main.py:153-164): text = r*sent + (1-r)*noun; neg = mean_k others[k] (zeros if none)
  auto text_round = [&](int rd) {
    const int j = warp & 3, kind = warp >> 2;
    const int e = e_lo + rd * kPsER + j;
    if (e < e_hi) {
      const int k0 = p.other_off[e], k1 = p.other_off[e + 1];
      float acc = 0.f;
      for (int c = lane; c < Nw; c += 32) {
        const int col = col0 + c;
        float v = 0.f;
        if (col < D) {
          if (kind == 0) {
            v = __fadd_rn(__fmul_rn(p.r, __ldg(p.sent + (size_t)e * D + col)), __fmul_rn(p.one_minus_r, __ldg(p.noun + (size_t)e * D + col)));
          } else {
            for (int k = k0; k < k1; ++k) v = __fadd_rn(v, __ldg(p.others + (size_t)k * D + col));
            if (k1 > k0) v = __fdiv_rn(v, (float)(k1 - k0));
          }
        }
        text[(kind * kPsER + j) * Nw + c] = v;
        acc += v * v;
      }
      acc = warp_sum(acc);
      if (lane == 0) tnp[kind * kPsER + j] = acc;
    }
  };
  if (EB > 0) text_round(0);

  // ---- A: soft masks f32 -> (hi, lo) bf16 in the canonical K-major layout, one 64-token chunk per iteration.  A warp
  //      converts 8 rows x 32 columns per item: lane -> (row % 8, 8-float column group): 32-byte global sectors in, one
  //      contiguous 512-byte run of 16-byte core rows out (bank-conflict-free).  Rows beyond the image and k >= L are zero.
  const bool w_vec = (L & 3) == 0 && (reinterpret_cast<uintptr_t>(p.w) & 15) == 0;
  auto convert_a = [&](int tile, int kc, int slot) {
    uint8_t* a_hi = ring + (size_t)slot * p.stage_bytes;
    constexpr int kItems = kPsM * (kPsKC / 8) / kPsThreads;      // 4
    float4 f[kItems][2];
#pragma unroll
    for (int it = 0; it < kItems; ++it) {
      const int wi = warp + 8 * it;
      const int r = (wi >> 1) * 8 + (lane & 7), c8 = (wi & 1) * 4 + (lane >> 3);
      const int row = tile * kPsM + r, k = kc * kPsKC + c8 * 8;
      f[it][0] = make_float4(0.f, 0.f, 0.f, 0.f); f[it][1] = f[it][0];
      if (row < n && k < L) {
        const float* src = p.w + (size_t)(n_lo + row) * L + k;
        if (w_vec && k + 8 <= L) {
          f[it][0] = __ldg(reinterpret_cast<const float4*>(src)); f[it][1] = __ldg(reinterpret_cast<const float4*>(src) + 1);
        } else {
          float v[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) v[q] = (k + q < L) ? __ldg(src + q) : 0.f;
          f[it][0] = make_float4(v[0], v[1], v[2], v[3]); f[it][1] = make_float4(v[4], v[5], v[6], v[7]);
        }
      }
    }
#pragma unroll
    for (int it = 0; it < kItems; ++it) {
      const int wi = warp + 8 * it;
      const int r = (wi >> 1) * 8 + (lane & 7), c8 = (wi & 1) * 4 + (lane >> 3);
      const float v[8] = {f[it][0].x, f[it][0].y, f[it][0].z, f[it][0].w, f[it][1].x, f[it][1].y, f[it][1].z, f[it][1].w};
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const __nv_bfloat16 h0 = __float2bfloat16_rn(v[2 * q]), h1 = __float2bfloat16_rn(v[2 * q + 1]);
        hi[q] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
        lo[q] = pack_bf16x2(v[2 * q] - __bfloat162float(h0), v[2 * q + 1] - __bfloat162float(h1));
      }
      const size_t off = (size_t)(r >> 3) * 1024 + (size_t)c8 * 128 + (r & 7) * 16;
      *reinterpret_cast<uint4*>(a_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(a_hi + kPsAHalf + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
  };

  // instruction descriptor: D = f32, A = B = bf16, A K-major, B MN-major, N = Nw, M = 128
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (0u << 15) | (1u << 16) | ((uint32_t)(Nw >> 3) << 17) | ((uint32_t)(kPsM >> 4) << 24);
  for (int it = 0; it < total; ++it) {
    const int slot = it % S, kc = it % NKC, tile = it / NKC;
    if (tid == 0) {                                              // keep S-1 B stages in flight ahead of the MMAs
      const int nxt = it + S - 1;
      if (nxt >= S && nxt < total) {
        mbar_wait(bar_free + (nxt % S), (uint32_t)((nxt / S - 1) & 1));
        issue_b(nxt);
      }
    }
    if (it >= S) mbar_wait(bar_free + slot, (uint32_t)((it / S - 1) & 1));     // the MMAs that read this slot have completed
    convert_a(tile, kc, slot);
    proxy_fence_async();                                         // generic-proxy smem writes -> visible to the tensor core (async proxy)
    __syncthreads();
    if (tid == 0) {
      mbar_wait(bar_full + slot, (uint32_t)((it / S) & 1));      // TMA has landed this stage's tokens
      tc_fence_after();
      const uint32_t a_addr = smem_u32(ring + (size_t)slot * p.stage_bytes);
      const uint32_t b_addr = a_addr + 2 * kPsAHalf;
      const int ksteps = min(kPsKC / 16, (L - kc * kPsKC + 15) / 16);
      for (int s2 = 0; s2 < ksteps; ++s2) {
        // A: two K-cores 128 B apart, 8-row groups 1024 B apart.  B: 128-byte swizzle, MN-major: 8-token groups (SBO) 1024 B
        // apart, 64-column groups (LBO) one box apart; a k-step is 16 tokens = 2048 B.
        const uint64_t bd = smem_desc(b_addr + (uint32_t)s2 * 2048u, kPsBBox, 1024u, 2u);
        const uint64_t ah = smem_desc(a_addr + (uint32_t)s2 * 256u, 128u, 1024u, 0u);
        const uint64_t al = smem_desc(a_addr + kPsAHalf + (uint32_t)s2 * 256u, 128u, 1024u, 0u);
        umma_bf16(tmem_base + (uint32_t)(tile * Nw), ah, bd, idesc, (kc > 0 || s2 > 0) ? 1u : 0u);
        umma_bf16(tmem_base + (uint32_t)(tile * Nw), al, bd, idesc, 1u);
      }
      umma_commit(bar_free + slot);
      if (it == total - 1) umma_commit(bar_acc);
    }
  }

  // ---- epilogue.  Warp w reads TMEM lanes 32*(w%4)..+31 (= proposal rows); warps 0-3 take the even 16-column groups,
  //      warps 4-7 the odd ones.
  mbar_wait(bar_acc, 0u);
  tc_fence_after();
  const int half = warp >> 2;
  const int my_row = (warp & 3) * 32 + lane;
  const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  constexpr int kNP = 1 + 2 * kPsER;                             // partial sums per row: |row|^2, 4 text dots, 4 negative dots
  float* part1 = stage;                                          // half-1 partials [tiles][kNP][128] (<= 18 KB of the stage buffer)
  const int rounds = max(1, (EB + kPsER - 1) / kPsER);
  const int rows_pad = tiles * kPsM;
  for (int rd = 0; rd < rounds; ++rd) {
    const int ne = min(kPsER, EB - rd * kPsER);                  // <= 0 when the launch only pools
    if (rd > 0) { text_round(rd); __syncthreads(); }
    for (int tile = 0; tile < tiles; ++tile) {
      float ss = 0.f, dt[kPsER], dn[kPsER];
#pragma unroll
      for (int j = 0; j < kPsER; ++j) { dt[j] = 0.f; dn[j] = 0.f; }
      for (int c = half * 16; c < Nw; c += 32) {
        float v[16];
        tmem_ld16(lane_addr + (uint32_t)(tile * Nw + c), v);
#pragma unroll
        for (int q = 0; q < 16; ++q) ss += v[q] * v[q];
#pragma unroll
        for (int j = 0; j < kPsER; ++j) {
          if (j < ne) {
            const float4* t4 = reinterpret_cast<const float4*>(text + j * Nw + c);
            const float4* n4 = reinterpret_cast<const float4*>(text + (kPsER + j) * Nw + c);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 a = t4[q], g = n4[q];
              dt[j] += v[4 * q] * a.x + v[4 * q + 1] * a.y + v[4 * q + 2] * a.z + v[4 * q + 3] * a.w;
              dn[j] += v[4 * q] * g.x + v[4 * q + 1] * g.y + v[4 * q + 2] * g.z + v[4 * q + 3] * g.w;
            }
          }
        }
      }
      float* dst = (half ? part1 : part) + (size_t)tile * kNP * kPsM + my_row;
      dst[0] = ss;
#pragma unroll
      for (int j = 0; j < kPsER; ++j) { dst[(1 + j) * kPsM] = dt[j]; dst[(1 + kPsER + j) * kPsM] = dn[j]; }
    }
    __syncthreads();
    for (int i = tid; i < tiles * kNP * kPsM; i += kPsThreads) part[i] += part1[i];
    cluster_sync_all();                                          // #1: every CTA's partial sums are visible cluster-wide

    if (rank == 0 && ne > 0) {                                   // text norms: sum of the column slices
      if (tid < 2 * kPsER) {
        float t = 0.f;
        for (int c = 0; c < p.NT; ++c) t += ld_peer_f32(tnp + tid, (uint32_t)c);
        tn_s[tid] = sqrtf(t);
      }
      __syncthreads();
    }
    if (tid < kPsM && (rank == 0 || (rd == 0 && p.out != nullptr))) {
      for (int tile = 0; tile < tiles; ++tile) {
        const float* src = part + (size_t)tile * kNP * kPsM + tid;
        float ss = 0.f;
        for (int c = 0; c < p.NT; ++c) ss += ld_peer_f32(src, (uint32_t)c);
        const float fnorm = sqrtf(ss);
        if (rd == 0) inv_s[tile * kPsM + tid] = p.normalize ? __frcp_rn(fnorm) : 1.f;
        if (rank == 0 && ne > 0) {
          const int grow = tile * kPsM + tid;
#pragma unroll
          for (int j = 0; j < kPsER; ++j) {
            if (j < ne) {
              float a = 0.f, g = 0.f;
              for (int c = 0; c < p.NT; ++c) {
                a += ld_peer_f32(src + (1 + j) * kPsM, (uint32_t)c);
                g += ld_peer_f32(src + (1 + kPsER + j) * kPsM, (uint32_t)c);
              }
              // scale * (f/|f|) . (t/|t|); a zero 'neg' vector gives 0/0 = NaN exactly like the reference (App. B-5)
              const float s_pos = p.scale * __fdiv_rn(__fdiv_rn(a, fnorm), tn_s[j]);
              const float s_neg = p.scale * __fdiv_rn(__fdiv_rn(g, fnorm), tn_s[kPsER + j]);
              sc[j * rows_pad + grow] = s_pos;
              sc[(kPsER + j) * rows_pad + grow] = s_neg;
              if (grow < n) p.tail.score_clip[(size_t)(e_lo + rd * kPsER + j) * p.max_n + grow] = s_pos;
            }
          }
        }
      }
    }
    cluster_sync_all();                                          // #2: nobody reads a peer's shared memory past this point
    if (rank == 0 && ne > 0) {
      if (warp < ne) select_tail_warp(p.tail, e_lo + rd * kPsER + warp, n, n_lo, sc + warp * rows_pad, sc + (kPsER + warp) * rows_pad, picks + warp * 9, lane);
      __syncthreads();
    }
  }

  // ---- optional: the pooled (normalised) rows themselves, staged 64 columns at a time through shared memory
  if (p.out != nullptr) {
    constexpr int kPitch = kPsPiece + 1;
    __syncthreads();
    for (int tile = 0; tile < tiles; ++tile) {
      const int rows = min(kPsM, n - tile * kPsM);
      const float inv = inv_s[tile * kPsM + my_row];
      for (int c0 = 0; c0 < Nw; c0 += kPsPiece) {
        for (int c = half * 16; c < kPsPiece; c += 32) {
          float v[16];
          tmem_ld16(lane_addr + (uint32_t)(tile * Nw + c0 + c), v);
#pragma unroll
          for (int q = 0; q < 16; ++q) stage[my_row * kPitch + c + q] = v[q] * inv;
        }
        __syncthreads();
        const int pw = min(kPsPiece, D - (col0 + c0));             // valid columns of this piece (multiple of 8)
        for (int t = tid; t < rows * (pw / 4); t += kPsThreads) {
          const int r = t / (pw / 4), q4 = t - r * (pw / 4);
          const float* sp = stage + r * kPitch + q4 * 4;
          const size_t o = (size_t)(n_lo + tile * kPsM + r) * D + col0 + c0 + q4 * 4;
          if (p.out_bf16) {
            uint2 w2;
            w2.x = pack_bf16x2(sp[0], sp[1]); w2.y = pack_bf16x2(sp[2], sp[3]);
            *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(p.out) + o) = w2;
          } else {
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + o) = make_float4(sp[0], sp[1], sp[2], sp[3]);
          }
        }
        __syncthreads();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, p.tmem_cols);
}

// ---- host side ----------------------------------------------------------------------------------------------------------
typedef CUresult (*TensorMapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TensorMapEncodeFn tensor_map_encoder() {
  static TensorMapEncodeFn fn = []() -> TensorMapEncodeFn {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
    return reinterpret_cast<TensorMapEncodeFn>(sym);
  }();
  return fn;
}

static int pool_score_launch(PoolScoreParams& p, const void* tokens, cudaStream_t st, const char* what) {
  const int D = p.D, L = p.L;
  // columns per CTA: whole 64-column swizzle atoms, at most 8 CTAs per (portable) cluster
  int nw = 64;
  while (nw < 256 && ceil_div(D, nw) > 8) nw *= 2;
  p.Nw = nw; p.NT = ceil_div(D, nw); p.boxes = nw / 64;
  HGL_REQUIRE(p.NT <= 8, "%s: D=%d needs %d column tiles (> 8 CTAs per cluster)", what, D, p.NT);
  const int tiles = ceil_div(std::max(1, std::min(p.max_n, p.M)), kPsM);
  HGL_REQUIRE(tiles <= kPsMaxTiles && tiles * nw <= 512, "%s: max_n=%d with D=%d exceeds the TMEM accumulator budget (%d proposals per image)", what,
              p.max_n, D, std::min(kPsMaxTiles, 512 / nw) * kPsM);
  uint32_t cols = 32;
  while (cols < (uint32_t)(tiles * nw)) cols *= 2;
  p.tmem_cols = cols;
  p.NKC = ceil_div(L, kPsKC);
  p.stage_bytes = 2 * kPsAHalf + (uint32_t)p.boxes * kPsBBox;
  const size_t budget = 227 * 1024 - 1024 - kRingBase;
  int stages = (int)std::min<size_t>(std::min(p.NKC * tiles, 4), budget / p.stage_bytes);
  while ((size_t)stages * p.stage_bytes < kEpEnd) ++stages;      // the epilogue scratch lives in the ring
  HGL_REQUIRE(stages >= 1 && (size_t)stages * p.stage_bytes <= budget && stages <= kPsMaxStages, "%s: shared-memory budget", what);
  p.stages = stages;
  const size_t smem = 1024 + kRingBase + (size_t)stages * p.stage_bytes;

  TensorMapEncodeFn enc = tensor_map_encoder();
  if (!enc) { set_error("%s: cuTensorMapEncodeTiled is not available from this driver", what); return HGL_ECUDA; }
  CUtensorMap tmap;
  const cuuint64_t gdim[3] = {(cuuint64_t)D, (cuuint64_t)L, (cuuint64_t)p.B};
  const cuuint64_t gstride[2] = {(cuuint64_t)D * 2, (cuuint64_t)L * D * 2};
  const cuuint32_t box[3] = {64, (cuuint32_t)kPsKC, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(tokens), gdim, gstride, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) { set_error("%s: cuTensorMapEncodeTiled failed (CUresult %d; D=%d L=%d B=%d)", what, (int)cr, D, L, p.B); return HGL_ECUDA; }

  {
    const int rc_s = ensure_dyn_smem(reinterpret_cast<const void*>(pool_score_kernel), smem, what);
    if (rc_s != HGL_OK) return rc_s;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(p.NT, p.B, 1);
  cfg.blockDim = dim3(kPsThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)p.NT; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, pool_score_kernel, tmap, p);
  if (e != cudaSuccess) { set_error("%s: cudaLaunchKernelEx: %s", what, cudaGetErrorString(e)); return HGL_ECUDA; }
  return launch_status(what);
}

}  // namespace hgl

extern "C" int64_t hgl_mask_pool_workspace_bytes(int M, int D, int out_dtype) {
  if (M < 0 || D < 1) return -1;
  (void)out_dtype;
  return 0;         // accumulators stay in TMEM until normalised: no global scratch
}

extern "C" int hgl_mask_pool(const float* weights, const void* tokens, const int32_t* mask_off, int B, int M, int max_n, int L, int D,
                             int normalize, int out_dtype, void* out, void* workspace, void* stream) {
  using namespace hgl;
  (void)workspace;
  if (M == 0) return HGL_OK;
  HGL_REQUIRE(weights && tokens && out, "hgl_mask_pool: null pointer");
  HGL_REQUIRE(B >= 1 && M >= 0 && max_n >= 1 && L >= 1, "hgl_mask_pool: bad shape");
  HGL_REQUIRE(mask_off || B == 1, "hgl_mask_pool: mask_off required when B > 1");
  HGL_REQUIRE(out_dtype == HGL_F32 || out_dtype == HGL_BF16, "hgl_mask_pool: out_dtype %d", out_dtype);
  HGL_REQUIRE(D >= 8 && D % 8 == 0, "hgl_mask_pool: D=%d must be a multiple of 8", D);
  HGL_REQUIRE(((reinterpret_cast<uintptr_t>(tokens) | reinterpret_cast<uintptr_t>(out)) & 15) == 0, "hgl_mask_pool: tokens / out must be 16-byte aligned");
  HGL_REQUIRE(B <= 65535, "hgl_mask_pool: B=%d too large for one launch", B);
  PoolScoreParams p = {};
  p.w = weights; p.mask_off = mask_off; p.expr_off = nullptr;
  p.B = B; p.M = M; p.E = 0; p.L = L; p.D = D; p.max_n = (B == 1) ? std::max(M, 1) : max_n;
  p.out = out; p.out_bf16 = out_dtype == HGL_BF16; p.normalize = normalize ? 1 : 0;
  return pool_score_launch(p, tokens, (cudaStream_t)stream, "hgl_mask_pool");
}

extern "C" int hgl_pool_score_select(const float* weights, const void* tokens, const int32_t* mask_off, const int32_t* expr_off,
                                     int B, int M, int E, int max_n, int L, int D,
                                     const float* sent, const float* noun, const float* others, const int32_t* other_off,
                                     const int64_t* boxes, const int32_t* relaflag, const float* score_gem,
                                     double logit_scale_exp, double r, double alpha, void* features_out, int out_dtype,
                                     float* score_clip, int64_t* idx_hybrid, int64_t* idx_final, int32_t* top_idx, float* blended,
                                     void* stream) {
  using namespace hgl;
  HGL_REQUIRE(weights && tokens && sent && noun && other_off && boxes && relaflag && score_clip && idx_hybrid && idx_final && top_idx && blended,
              "hgl_pool_score_select: null pointer");
  HGL_REQUIRE(B >= 1 && M >= 0 && E >= 0 && max_n >= 1 && L >= 1, "hgl_pool_score_select: bad shape");
  HGL_REQUIRE((mask_off && expr_off) || B == 1, "hgl_pool_score_select: mask_off / expr_off required when B > 1");
  HGL_REQUIRE(out_dtype == HGL_F32 || out_dtype == HGL_BF16, "hgl_pool_score_select: out_dtype %d", out_dtype);
  HGL_REQUIRE(D >= 8 && D % 8 == 0, "hgl_pool_score_select: D=%d must be a multiple of 8", D);
  HGL_REQUIRE(((reinterpret_cast<uintptr_t>(tokens) | reinterpret_cast<uintptr_t>(features_out)) & 15) == 0,
              "hgl_pool_score_select: tokens / features_out must be 16-byte aligned");
  HGL_REQUIRE(B <= 65535, "hgl_pool_score_select: B=%d too large for one launch", B);
  if (E == 0 && features_out == nullptr) return HGL_OK;
  PoolScoreParams p = {};
  p.w = weights; p.mask_off = mask_off; p.expr_off = expr_off;
  p.B = B; p.M = M; p.E = E; p.L = L; p.D = D; p.max_n = max_n;
  p.out = features_out; p.out_bf16 = out_dtype == HGL_BF16; p.normalize = 1;
  p.sent = sent; p.noun = noun; p.others = others; p.other_off = other_off;
  p.scale = (float)logit_scale_exp; p.r = (float)r; p.one_minus_r = (float)(1.0 - r);
  p.tail.boxes = boxes; p.tail.relaflag = relaflag; p.tail.other_off = other_off; p.tail.score_gem = score_gem;
  p.tail.alpha = (float)alpha; p.tail.one_minus_alpha = (float)(1.0 - alpha); p.tail.max_n = max_n;
  p.tail.score_clip = score_clip; p.tail.idx_hybrid = idx_hybrid; p.tail.idx_final = idx_final; p.tail.top_idx = top_idx; p.tail.blended = blended;
  return pool_score_launch(p, tokens, (cudaStream_t)stream, "hgl_pool_score_select");
}
