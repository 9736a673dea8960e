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
//   negative vectors (this CTA's column slice, in shared memory); every CTA PUSHES its partial sums into the rank-0 CTA's
//   shared memory (mapa + st.shared::cluster, one barrier.cluster), which finishes the scores and runs the selection tail.
// Arithmetic intensity is low (2*N*L*D flop over ~2*(N*L + L*D) bytes): the kernel is a latency chain
// (TMA -> MMA -> TMEM -> DSMEM -> tail), sized to run wide (B*NT CTAs) rather than to saturate the tensor pipe.
#include <cuda.h>   // CUtensorMap and its enums only; the encoder is fetched through cudaGetDriverEntryPoint (no libcuda link)

#include <algorithm>

#include "hgl_common.cuh"
#include "select_tail.cuh"

namespace hgl {

constexpr int kPsThreads = 512;   // 16 warps (4 per scheduler: a lone warp issues one instruction every ~13 cycles, the kernel is issue-latency
                                  // bound): all stage A; thread 0 issues TMA + MMA; warps w, w+4, w+8, w+12 share TMEM lanes 32*(w%4)..
constexpr int kPsM = 128;         // proposals per row tile (UMMA M)
constexpr int kPsKC = 64;         // tokens per ring stage (4 MMA k-steps)
constexpr int kPsER = 4;          // expressions per scoring round (== kPsThreads / kPsM: the reduction maps a thread to (row, expression))
constexpr int kPsMaxTiles = 4;    // row tiles per image (512 proposals)
constexpr int kPsMaxStages = 8;
static_assert(kPsThreads / kPsM == kPsER, "rank-0 reduction: one thread per (row, expression of the round)");
constexpr uint32_t kPsAHalf = kPsM * kPsKC * 2;   // one bf16 [128 x 64] operand block (hi or lo)
constexpr uint32_t kPsBBox = kPsKC * 64 * 2;      // one TMA box: 64 tokens x 64 columns bf16
constexpr int kPsPiece = 64;      // output columns staged through shared memory per round of the feature pass
// shared-memory map (offsets from the 1024-byte aligned base).  Fixed part; the rest (PoolScoreParams::off_*) depends on Nw / NT / tiles:
//   text   [2*kPsER][Nw] f32           this CTA's column slice of the text ensemble / negatives
//   boxes  [tiles*128][4] int64        rank 0: the image's boxes            } prefetched at kernel start: the selection tail
//   sg     [kPsER][tiles*128] f32      rank 0: score_gem rows of the round  } then never waits for global memory
//   gather [NT][kNP][128] f32          rank 0: partial sums pushed by every CTA of the cluster (NOT in the ring: peers push
//   tn     [NT][2*kPsER] f32                    while rank 0's MMAs may still be reading its ring)
//   ring   S stages of (A hi | A lo | B boxes), 1024-byte aligned
constexpr uint32_t kOffFull = 0, kOffFree = 64, kOffAcc = 128, kOffTmem = 136, kOffAFull = 192, kOffTnp = 256, kOffMeta = 320, kOffText = 512;
constexpr int kNP = 1 + 2 * kPsER;                                     // partial sums per row: |row|^2, 4 text dots, 4 negative dots
// epilogue scratch, aliased onto the CTA's OWN ring once every one of its MMAs has completed
constexpr uint32_t kEpPart1 = 0;                                       // [3][kNP][128] f32: partial sums of the column quarters 1..3
constexpr uint32_t kEpSc = kEpPart1 + 3 * kNP * kPsM * 4;              // rank 0: [2*kPsER][tiles*128] f32 raw scores
constexpr uint32_t kEpPicks = kEpSc + 2 * kPsER * kPsMaxTiles * kPsM * 4;   // rank 0: [kPsER][kTailPicks] int
constexpr uint32_t kEpInv = kEpPicks + 256;                            // rank 0: [tiles*128] f32 |row|
constexpr uint32_t kEpStage = kEpInv + kPsMaxTiles * kPsM * 4;         // [128][kPsPiece + 1] f32 (feature output only)
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
// the same with the accumulate flag known at compile time and the descriptors' invariant high words kept apart (the issuing
// thread is alone in its warp: every instruction it does not execute is ~13 cycles off the critical path)
template <bool kAcc>
__device__ __forceinline__ void umma_bf16_lohi(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(tmem_d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "n"(kAcc ? 1 : 0)
      : "memory");
}
// arrive on `bar` when every MMA issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 consecutive accumulator columns of this thread's TMEM lane (row); NO wait: pair with tmem_ld_wait()
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                 "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
                 "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
                 "=r"(r[30]), "=r"(r[31])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
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

// Phase trace for tuning builds (-DHGL_TUNING, never the shipped library): thread 0 of every CTA stamps clock64() at the
// phase boundaries; read back with hgl_debug_pool_score_trace (profiles/trace_pool_score.py).
#ifdef HGL_TUNING
__device__ long long g_ps_trace[16 * 4096];
#define PS_TRACE(k) do { if (threadIdx.x == 0 && blockIdx.y * gridDim.x + blockIdx.x < 4096) \
    g_ps_trace[(blockIdx.y * gridDim.x + blockIdx.x) * 16 + (k)] = clock64(); } while (0)
#else
#define PS_TRACE(k) do { } while (0)
#endif

struct PoolScoreParams {
  const float* w;              // [M, L] f32 soft grid masks
  const int32_t* mask_off;     // [B+1] or null (B == 1)
  const int32_t* expr_off;     // [B+1] or null (B == 1)
  int B, M, E, L, D, max_n;
  int Nw, NT, NKC, stages, boxes;   // columns per CTA, CTAs per cluster, token chunks, ring depth, TMA boxes per stage (Nw / 64)
  uint32_t stage_bytes, tmem_cols;
  uint32_t off_boxes, off_sg, off_gather, off_tn, off_ring;   // shared-memory map (bytes from the aligned base)
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
  const int L = p.L, D = p.D, Nw = p.Nw, S = p.stages, NKC = p.NKC, NT = p.NT;
  const int col0 = (int)rank * Nw;

  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + kOffFull);
  uint64_t* bar_free = reinterpret_cast<uint64_t*>(smem + kOffFree);
  uint64_t* bar_acc = reinterpret_cast<uint64_t*>(smem + kOffAcc);
  uint64_t* bar_afull = reinterpret_cast<uint64_t*>(smem + kOffAFull);   // [kPsMaxStages] A stage converted: one arrival per warp
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kOffTmem);
  float* tnp = reinterpret_cast<float*>(smem + kOffTnp);          // [2*kPsER] this CTA's partial |text|^2 (its column slice)
  float* text = reinterpret_cast<float*>(smem + kOffText);        // [2*kPsER][Nw]: rows 0..3 text ensemble, 4..7 negatives
  int64_t* box_s = reinterpret_cast<int64_t*>(smem + p.off_boxes);  // rank 0: boxes of the image
  float* sg_s = reinterpret_cast<float*>(smem + p.off_sg);          // rank 0: score_gem rows of the round's expressions
  int* meta_s = reinterpret_cast<int*>(smem + kOffMeta);            // rank 0: (n_other, relaflag) of the round's expressions
  float* gather = reinterpret_cast<float*>(smem + p.off_gather);
  float* gather_tn = reinterpret_cast<float*>(smem + p.off_tn);
  uint8_t* ring = smem + p.off_ring;
  float* part1 = reinterpret_cast<float*>(ring + kEpPart1);
  float* sc = reinterpret_cast<float*>(ring + kEpSc);
  int* picks = reinterpret_cast<int*>(ring + kEpPicks);
  float* inv_s = reinterpret_cast<float*>(ring + kEpInv);
  float* stage = reinterpret_cast<float*>(ring + kEpStage);

  if (n == 0) {
    // an image without proposals: nothing to pool; the selection tail still defines its outputs (-1 picks), like hgl_score_select
    if (rank == 0)
      for (int e0 = e_lo; e0 < e_hi; e0 += kPsER) {
        select_tail_block(p.tail, e0, min(kPsER, e_hi - e0), 0, n_lo, sc, kPsM, picks, nullptr, nullptr, nullptr, warp, lane);
        __syncthreads();
      }
    return;
  }
  const int tiles = (n + kPsM - 1) / kPsM;
  const int total = tiles * NKC;
  PS_TRACE(0);

  if (tid == 0) {
    for (int i = 0; i < kPsMaxStages; ++i) { mbar_init(bar_full + i, 1); mbar_init(bar_free + i, 1); mbar_init(bar_afull + i, kPsThreads / 32); }
    mbar_init(bar_acc, 1);
    mbar_fence_init();
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
  }
  if (warp == 0) tmem_alloc(tmem_slot, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  cluster_arrive();                                              // "this CTA runs": waited for before the first push into rank 0
  bool peers_started = false;
  const uint32_t tmem_base = *tmem_slot;
  PS_TRACE(1);

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

  // ---- A: soft masks f32 -> (hi, lo) bf16 in the canonical K-major layout, one 64-token chunk per iteration.  A warp
  //      converts 8 rows x 32 columns per item: lane -> (row % 8, 8-float column group): 32-byte global sectors in, one
  //      contiguous 512-byte run of 16-byte core rows out (bank-conflict-free).  Rows beyond the image and k >= L are zero.
  //      The loads of chunk it+1 are issued before chunk it is converted, so only the first chunk's latency is exposed.
  constexpr int kItems = kPsM * (kPsKC / 8) / kPsThreads;      // 2
  const bool w_vec = (L & 3) == 0 && (reinterpret_cast<uintptr_t>(p.w) & 15) == 0;
  // per-thread invariants of the two items: source pointer, row, first column, offset inside an operand block
  const float* a_src[kItems];
  int a_row[kItems], a_col[kItems];
  uint32_t a_off[kItems];
#pragma unroll
  for (int q = 0; q < kItems; ++q) {
    const int wi = warp + (kPsThreads / 32) * q;
    const int r = (wi >> 1) * 8 + (lane & 7), c8 = (wi & 1) * 4 + (lane >> 3);
    a_row[q] = r; a_col[q] = c8 * 8;
    a_src[q] = p.w + (size_t)(n_lo + r) * L + c8 * 8;
    a_off[q] = (uint32_t)(r >> 3) * 1024u + (uint32_t)c8 * 128u + (uint32_t)(r & 7) * 16u;
  }
  auto load_a = [&](int it, float4 (&f)[kItems][2]) {
    const int kc = it % NKC, tile = it / NKC;
    const size_t shift = (size_t)tile * kPsM * L + (size_t)kc * kPsKC;
#pragma unroll
    for (int q = 0; q < kItems; ++q) {
      const int row = tile * kPsM + a_row[q], k = kc * kPsKC + a_col[q];
      f[q][0] = make_float4(0.f, 0.f, 0.f, 0.f); f[q][1] = f[q][0];
      if (row < n && k < L) {
        const float* src = a_src[q] + shift;
        if (w_vec && k + 8 <= L) {
          f[q][0] = __ldg(reinterpret_cast<const float4*>(src)); f[q][1] = __ldg(reinterpret_cast<const float4*>(src) + 1);
        } else {
          float v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) v[u] = (k + u < L) ? __ldg(src + u) : 0.f;
          f[q][0] = make_float4(v[0], v[1], v[2], v[3]); f[q][1] = make_float4(v[4], v[5], v[6], v[7]);
        }
      }
    }
  };
  // hi = bf16(w), lo = bf16(w - hi), two values per cvt.rn.bf16x2.f32
  auto store_a = [&](int slot, const float4 (&f)[kItems][2]) {
    uint8_t* a_hi = ring + (size_t)slot * p.stage_bytes;
#pragma unroll
    for (int q = 0; q < kItems; ++q) {
      const float v[8] = {f[q][0].x, f[q][0].y, f[q][0].z, f[q][0].w, f[q][1].x, f[q][1].y, f[q][1].z, f[q][1].w};
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        hi[u] = pack_bf16x2(v[2 * u], v[2 * u + 1]);
        lo[u] = pack_bf16x2(v[2 * u] - __uint_as_float(hi[u] << 16), v[2 * u + 1] - __uint_as_float(hi[u] & 0xffff0000u));
      }
      *reinterpret_cast<uint4*>(a_hi + a_off[q]) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(a_hi + kPsAHalf + a_off[q]) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
  };
  // Every global load this kernel depends on is issued up front and in parallel (the kernel is a latency chain: a dependent
  // DRAM round trip costs 1-2 us while the prep kernel saturates HBM beside it): the first four A chunks into registers ...
  float4 f0[kItems][2], f1[kItems][2], f2[kItems][2], f3[kItems][2];
  auto load_set = [&](int it) {
    switch (it & 3) { case 0: load_a(it, f0); break; case 1: load_a(it, f1); break; case 2: load_a(it, f2); break; default: load_a(it, f3); break; }
  };
  auto store_set = [&](int it, int slot) {
    switch (it & 3) { case 0: store_a(slot, f0); break; case 1: store_a(slot, f1); break; case 2: store_a(slot, f2); break; default: store_a(slot, f3); break; }
  };
  const int rows_pad = ((n + kPsM - 1) / kPsM) * kPsM;
  for (int it = 0; it < min(4, total); ++it) load_set(it);
  // ... and, on rank 0, what the selection tail will need: the image's boxes, the round's score_gem rows, (n_other, relaflag).
  // Loads go to registers here and are stored to shared memory only after the text side below has issued ITS loads: a store
  // right behind its load would stall the thread for a full memory round trip per statement.
  constexpr int kPF = kPsMaxTiles * kPsM * 4 / kPsThreads;       // 8 elements per thread cover 512 boxes / 4 x 512 score_gem values
  int64_t pf_box[kPF];
  float pf_sg[kPF];
  int pf_meta[2] = {0, 0};
  auto tail_loads = [&](int rd, bool with_boxes) {
    const int ne_r = min(kPsER, EB - rd * kPsER);
    if (tid < ne_r) {
      const int e = e_lo + rd * kPsER + tid;
      pf_meta[0] = p.tail.other_off[e + 1] - p.tail.other_off[e];
      pf_meta[1] = p.tail.relaflag[e];
    }
#pragma unroll
    for (int u = 0; u < kPF; ++u) {
      const int i = tid + u * kPsThreads;
      pf_box[u] = (with_boxes && i < n * 4) ? __ldg(p.tail.boxes + (size_t)n_lo * 4 + i) : 0;
      pf_sg[u] = 0.f;
      if (p.tail.score_gem != nullptr && i < ne_r * n) {
        const int j = i / n, c = i - j * n;
        pf_sg[u] = __ldg(p.tail.score_gem + (size_t)(e_lo + rd * kPsER + j) * p.max_n + c);
      }
    }
  };
  auto tail_stores = [&](int rd, bool with_boxes) {
    const int ne_r = min(kPsER, EB - rd * kPsER);
    if (tid < ne_r) { meta_s[2 * tid] = pf_meta[0]; meta_s[2 * tid + 1] = pf_meta[1]; }
#pragma unroll
    for (int u = 0; u < kPF; ++u) {
      const int i = tid + u * kPsThreads;
      if (with_boxes && i < n * 4) box_s[i] = pf_box[u];
      if (p.tail.score_gem != nullptr && i < ne_r * n) { const int j = i / n, c = i - j * n; sg_s[j * rows_pad + c] = pf_sg[u]; }
    }
  };
  const bool tail_here = rank == 0 && EB > 0;
  if (tail_here) tail_loads(0, true);

  // (a6) text side of one scoring round, this CTA's column slice: warp w -> expression w % 4, ensemble (w < 4) or negatives
  // (Hybridgl_main.py:153-164): text = r*sent + (1-r)*noun; neg = mean_k others[k] (zeros if none)
  auto text_round = [&](int rd) {
    const int j = warp & 3, kind = (warp >> 2) & 1;
    const int e = e_lo + rd * kPsER + j;
    if (e < e_hi && warp < 8) {
      constexpr int kCU = 8;                                     // columns per lane: Nw / 32 <= 8
      float v[kCU];
      if (kind == 0) {                                           // all loads first, then the arithmetic
        float a[kCU], g[kCU];
#pragma unroll
        for (int ci = 0; ci < kCU; ++ci) {
          const int c = lane + 32 * ci, col = col0 + c;
          const bool ok = c < Nw && col < D;
          a[ci] = ok ? __ldg(p.sent + (size_t)e * D + col) : 0.f;
          g[ci] = ok ? __ldg(p.noun + (size_t)e * D + col) : 0.f;
        }
#pragma unroll
        for (int ci = 0; ci < kCU; ++ci) v[ci] = __fadd_rn(__fmul_rn(p.r, a[ci]), __fmul_rn(p.one_minus_r, g[ci]));
      } else {
        const int k0 = p.other_off[e], k1 = p.other_off[e + 1];
#pragma unroll
        for (int ci = 0; ci < kCU; ++ci) v[ci] = 0.f;
        for (int kb = k0; kb < k1; kb += 4) {                    // four negatives' loads in flight, added in order
          float o[4][kCU];
#pragma unroll
          for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int ci = 0; ci < kCU; ++ci) {
              const int c = lane + 32 * ci, col = col0 + c;
              o[u][ci] = (kb + u < k1 && c < Nw && col < D) ? __ldg(p.others + (size_t)(kb + u) * D + col) : 0.f;
            }
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (kb + u < k1) {
#pragma unroll
              for (int ci = 0; ci < kCU; ++ci) v[ci] = __fadd_rn(v[ci], o[u][ci]);
            }
        }
        if (k1 > k0) {
#pragma unroll
          for (int ci = 0; ci < kCU; ++ci) v[ci] = __fdiv_rn(v[ci], (float)(k1 - k0));
        }
      }
      float acc = 0.f;
#pragma unroll
      for (int ci = 0; ci < kCU; ++ci) {
        const int c = lane + 32 * ci;
        if (c < Nw) { text[(kind * kPsER + j) * Nw + c] = v[ci]; acc += v[ci] * v[ci]; }
      }
      acc = warp_sum(acc);
      if (lane == 0) tnp[kind * kPsER + j] = acc;
    }
  };
  if (EB > 0) text_round(0);
  if (tail_here) tail_stores(0, true);
  PS_TRACE(2);

  // instruction descriptor: D = f32, A = B = bf16, A K-major, B MN-major, N = Nw, M = 128
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (0u << 15) | (1u << 16) | ((uint32_t)(Nw >> 3) << 17) | ((uint32_t)(kPsM >> 4) << 24);
  // descriptor high words (invariant): A: no swizzle, SBO = 1024 (8-row groups), LBO = 128 (the two K-cores of a k-step);
  // B: 128-byte swizzle, MN-major: SBO = 1024 (8-token groups), LBO = one box (64-column groups); a k-step is 16 tokens = 2048 B
  const uint64_t a_tmpl = smem_desc(0u, 128u, 1024u, 0u), b_tmpl = smem_desc(0u, kPsBBox, 1024u, 2u);
  const uint32_t a_hi32 = (uint32_t)(a_tmpl >> 32), b_hi32 = (uint32_t)(b_tmpl >> 32);
  const uint32_t a_lo_tmpl = (uint32_t)a_tmpl, b_lo_tmpl = (uint32_t)b_tmpl;
  // No CTA-wide barrier in this loop: a warp that has converted its share of a stage arrives on the stage's `afull` mbarrier
  // and moves on; thread 0 trails behind, waits for (afull, full) of a stage and issues its MMAs.
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
    store_set(it, slot);
    if (it + 4 < total) load_set(it + 4);                        // refill the register set just drained
    proxy_fence_async();                                         // generic-proxy smem writes -> visible to the tensor core (async proxy)
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_afull + slot);
    if (it < 4) PS_TRACE(3 + it);
    if (tid == 0) {
      mbar_wait(bar_afull + slot, (uint32_t)((it / S) & 1));     // all eight warps have converted their share of this stage
      mbar_wait(bar_full + slot, (uint32_t)((it / S) & 1));      // TMA has landed this stage's tokens
      tc_fence_after();
      const uint32_t a_addr = smem_u32(ring + (size_t)slot * p.stage_bytes);
      // (14-bit start-address field: in a cluster the shared-window address carries the CTA rank in its upper bits)
      const uint32_t ah = a_lo_tmpl | ((a_addr >> 4) & 0x3fffu), al = a_lo_tmpl | (((a_addr + kPsAHalf) >> 4) & 0x3fffu);
      const uint32_t bd = b_lo_tmpl | (((a_addr + 2 * kPsAHalf) >> 4) & 0x3fffu);
      const uint32_t dcol = tmem_base + (uint32_t)(tile * Nw);
      const int ksteps = min(kPsKC / 16, (L - kc * kPsKC + 15) / 16);
      if (kc == 0) umma_bf16_lohi<false>(dcol, ah, a_hi32, bd, b_hi32, idesc);
      else umma_bf16_lohi<true>(dcol, ah, a_hi32, bd, b_hi32, idesc);
      umma_bf16_lohi<true>(dcol, al, a_hi32, bd, b_hi32, idesc);
#pragma unroll
      for (int s2 = 1; s2 < kPsKC / 16; ++s2) {
        if (s2 < ksteps) {
          umma_bf16_lohi<true>(dcol, ah + 16u * s2, a_hi32, bd + 128u * s2, b_hi32, idesc);
          umma_bf16_lohi<true>(dcol, al + 16u * s2, a_hi32, bd + 128u * s2, b_hi32, idesc);
        }
      }
      umma_commit(bar_free + slot);
      if (it == total - 1) umma_commit(bar_acc);
    }
  }

  // ---- epilogue.  Warp w reads TMEM lanes 32*(w%4)..+31 (= proposal rows); warps w, w+4, w+8, w+12 share the row's
  //      32-column groups.  Every CTA PUSHES its rows' partial sums into rank 0's shared memory (st.shared::cluster:
  //      fire and forget, no remote-load latency); one cluster barrier later rank 0 adds the NT slices up.
  PS_TRACE(7);
  mbar_wait(bar_acc, 0u);
  tc_fence_after();
  __syncthreads();      // everything before the epilogue (text rows, ring contents) is also ordered by a CTA barrier, not only through the
                        // mbarrier chain stores -> bar_afull -> MMA -> tcgen05.commit -> bar_acc (which compute-sanitizer cannot follow)
  PS_TRACE(8);
  const int quarter = warp >> 2;                                 // which 32-column groups of a row this thread reads
  const int my_row = (warp & 3) * 32 + lane;
  const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  const int rounds = max(1, (EB + kPsER - 1) / kPsER);
  const bool want_out = p.out != nullptr;
  for (int rd = 0; rd < rounds; ++rd) {
    const int ne = min(kPsER, EB - rd * kPsER);                  // <= 0 when the launch only pools
    if (rd > 0) {
      if (tail_here) tail_loads(rd, false);
      text_round(rd);
      if (tail_here) tail_stores(rd, false);
      __syncthreads();
    }
    for (int tile = 0; tile < tiles; ++tile) {
      float ss = 0.f, dt[kPsER], dn[kPsER];
#pragma unroll
      for (int j = 0; j < kPsER; ++j) { dt[j] = 0.f; dn[j] = 0.f; }
      // this thread's columns: the 32-column groups quarter, quarter + 4, ... of the row
      for (int c = quarter * 32; c < Nw; c += 128) {
        uint32_t r[32];
        tmem_ld32_nowait(lane_addr + (uint32_t)(tile * Nw + c), r);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 32; ++q) { const float x = __uint_as_float(r[q]); ss += x * x; }
#pragma unroll
        for (int j = 0; j < kPsER; ++j) {
          if (j < ne) {
            const float4* t4 = reinterpret_cast<const float4*>(text + j * Nw + c);
            const float4* n4 = reinterpret_cast<const float4*>(text + (kPsER + j) * Nw + c);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 a = t4[q], g = n4[q];
              const float x0 = __uint_as_float(r[4 * q]), x1 = __uint_as_float(r[4 * q + 1]);
              const float x2 = __uint_as_float(r[4 * q + 2]), x3 = __uint_as_float(r[4 * q + 3]);
              dt[j] += x0 * a.x + x1 * a.y + x2 * a.z + x3 * a.w;
              dn[j] += x0 * g.x + x1 * g.y + x2 * g.z + x3 * g.w;
            }
          }
        }
      }
      if (!peers_started) { cluster_wait(); peers_started = true; }   // uniform: once per CTA, before its first remote store
      if (quarter) {
        float* pq = part1 + (size_t)(quarter - 1) * kNP * kPsM + my_row;
        pq[0] = ss;
#pragma unroll
        for (int j = 0; j < kPsER; ++j) { pq[(1 + j) * kPsM] = dt[j]; pq[(1 + kPsER + j) * kPsM] = dn[j]; }
        if (tile == 0 && ne > 0 && tid >= 128 && tid < 128 + 2 * kPsER)
          st_peer_f32(gather_tn + rank * 2 * kPsER + (tid - 128), 0u, tnp[tid - 128]);
      }
      __syncthreads();
      if (!quarter) {
        float* dst = gather + (size_t)rank * kNP * kPsM + my_row;       // slice `rank` of rank 0's gather buffer
        const float* p1 = part1 + my_row;
        constexpr int kQ = kNP * kPsM;
        if (rd == 0) st_peer_f32(dst, 0u, ss + p1[0] + p1[kQ] + p1[2 * kQ]);
#pragma unroll
        for (int j = 0; j < kPsER; ++j) {
          if (j < ne) {
            const int kt = (1 + j) * kPsM, kn = (1 + kPsER + j) * kPsM;
            st_peer_f32(dst + kt, 0u, dt[j] + p1[kt] + p1[kQ + kt] + p1[2 * kQ + kt]);
            st_peer_f32(dst + kn, 0u, dn[j] + p1[kn] + p1[kQ + kn] + p1[2 * kQ + kn]);
          }
        }
      }
      PS_TRACE(9);
      cluster_sync_all();                                        // every CTA's partial sums have landed in rank 0
      PS_TRACE(10);
      if (rank == 0) {
        // thread (row, j): the row's scores against expression j of the round -- all 16 warps share the reduction
        const int row = tid & (kPsM - 1), j = tid >> 7;          // kPsThreads / kPsM == kPsER
        const float* src = gather + row;
        const int grow = tile * kPsM + row;
        float ss = 0.f, a = 0.f, g = 0.f, ta = 0.f, tg = 0.f;
#pragma unroll 4
        for (int c = 0; c < NT; ++c) {
          if (rd == 0) ss += src[(size_t)c * kNP * kPsM];
          if (j < ne) {
            a += src[((size_t)c * kNP + 1 + j) * kPsM];
            g += src[((size_t)c * kNP + 1 + kPsER + j) * kPsM];
            ta += gather_tn[c * 2 * kPsER + j];
            tg += gather_tn[c * 2 * kPsER + kPsER + j];
          }
        }
        const float fnorm = (rd == 0) ? sqrtf(ss) : inv_s[grow];
        if (rd == 0 && j == 0) inv_s[grow] = fnorm;              // |row|; turned into 1/|row| (or 1) by the feature pass
        if (j < ne) {
          // scale * (f/|f|) . (t/|t|); a zero 'neg' vector gives 0/0 = NaN exactly like the reference (App. B-5)
          const float s_pos = p.scale * __fdiv_rn(__fdiv_rn(a, fnorm), sqrtf(ta));
          const float s_neg = p.scale * __fdiv_rn(__fdiv_rn(g, fnorm), sqrtf(tg));
          sc[j * rows_pad + grow] = s_pos;
          sc[(kPsER + j) * rows_pad + grow] = s_neg;
          if (grow < n) p.tail.score_clip[(size_t)(e_lo + rd * kPsER + j) * p.max_n + grow] = s_pos;
        }
      }
      // another exchange follows (next tile / round) or the peers still need the row norms: rank 0 must be done reading first
      if (tile + 1 < tiles || rd + 1 < rounds || want_out) cluster_sync_all();
    }
    PS_TRACE(11);
    if (rank == 0 && ne > 0) {
      __syncthreads();                                           // the scores written by warps 0-3 above are read by every tail warp
      select_tail_block(p.tail, e_lo + rd * kPsER, ne, n, n_lo, sc, rows_pad, picks, box_s, p.tail.score_gem ? sg_s : nullptr, meta_s, warp, lane);
      __syncthreads();
    }
  }

  // ---- optional: the pooled (normalised) rows themselves, staged 64 columns at a time through shared memory
  if (want_out) {
    constexpr int kPitch = kPsPiece + 1;
    // the row norms live in rank 0 (the last cluster barrier above ordered its writes): every CTA fetches its rows' norms
    __syncthreads();
    float my_inv[kPsMaxTiles];
#pragma unroll
    for (int tile = 0; tile < kPsMaxTiles; ++tile)
      my_inv[tile] = (tile < tiles && p.normalize) ? __frcp_rn(ld_peer_f32(inv_s + tile * kPsM + my_row, 0u)) : 1.f;
    for (int tile = 0; tile < tiles; ++tile) {
      const int rows = min(kPsM, n - tile * kPsM);
      float inv = 1.f;
#pragma unroll
      for (int q = 0; q < kPsMaxTiles; ++q) if (q == tile) inv = my_inv[q];
      for (int c0 = 0; c0 < Nw; c0 += kPsPiece) {
        for (int c = quarter * 16; c < kPsPiece; c += 64) {
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
    cluster_sync_all();                                          // rank 0 stays until every peer has read the norms
  }
  PS_TRACE(12);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, p.tmem_cols);
  PS_TRACE(13);
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
  // (128 columns per CTA: for D = 512 a cluster of 4 -- sixteen images' clusters of 8 do not all fit the GPCs at once)
  int nw = D <= 64 ? 64 : 128;
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
  p.off_boxes = kOffText + (uint32_t)(2 * kPsER * nw * 4);
  p.off_sg = p.off_boxes + (uint32_t)(tiles * kPsM * 32);
  p.off_gather = p.off_sg + (uint32_t)(kPsER * tiles * kPsM * 4);
  p.off_tn = p.off_gather + (uint32_t)(p.NT * kNP * kPsM * 4);
  p.off_ring = (p.off_tn + (uint32_t)(p.NT * 2 * kPsER * 4) + 1023u) & ~1023u;
  const size_t budget = 227 * 1024 - 1024 - p.off_ring;
  int stages = (int)std::min<size_t>(std::min(p.NKC * tiles, 4), budget / p.stage_bytes);
  while ((size_t)stages * p.stage_bytes < kEpEnd) ++stages;      // the epilogue scratch lives in the ring
  HGL_REQUIRE(stages >= 1 && (size_t)stages * p.stage_bytes <= budget && stages <= kPsMaxStages, "%s: shared-memory budget", what);
  p.stages = stages;
  const size_t smem = 1024 + p.off_ring + (size_t)stages * p.stage_bytes;

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
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)p.NT; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1] = priority_attr(st);
  cfg.attrs = attr; cfg.numAttrs = 2;
  cudaError_t e = cudaLaunchKernelEx(&cfg, pool_score_kernel, tmap, p);
  if (e != cudaSuccess) { set_error("%s: cudaLaunchKernelEx: %s", what, cudaGetErrorString(e)); return HGL_ECUDA; }
  return launch_status(what);
}

}  // namespace hgl

#ifdef HGL_TUNING
extern "C" HGL_API int hgl_debug_pool_score_trace(long long* host_out, int n_ctas) {
  return cudaMemcpyFromSymbol(host_out, hgl::g_ps_trace, sizeof(long long) * 16 * (size_t)std::min(n_ctas, 4096)) == cudaSuccess ? 0 : -2;
}
#endif

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
