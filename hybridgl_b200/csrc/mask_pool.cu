// (b3') token-space mask pooling on the 5th-generation tensor cores, fused with L2 normalisation.
//
//   pooled[n, :] = sum_l  w[n, l] * tokens[b(n)][l, :]          w = soft patch-grid mask of proposal n (hgl_mask_grid,
//   out[n, :]    = pooled[n, :] / || pooled[n, :] ||_2              model/backbone.py:160), tokens = dense patch tokens
// This is the masks x tokens x D contraction of the north star: the token-space form of the reference's per-mask pooling
// loop (Hybridgl_main.py:218-223, SURVEY.md Appendix A-2:  S_in = (M~ . F^) . t ), with cosine scoring's normalisation
// (model/backbone.py:79) folded into the epilogue.  The reference pools in pixel space, one mask at a time, in Python.
//
// B200 design (one CTA = one image x 128 masks x one column tile; the column tiles of a row block form a thread-block cluster):
//   A = w   [128 x Kp]  bf16, K-major : converted from f32 once and kept resident in shared memory (canonical 8x16B cores)
//   B = tok [64 x N]    bf16, MN-major (tokens are [L, D] with D contiguous, so D = MMA-N is the contiguous mode): staged in
//                       a 2-slot ring of 64-token chunks; the slot is released by tcgen05.commit on an mbarrier
//   D = acc [128 x N]   f32 in TMEM; it stays there until it is normalised
//   tcgen05.mma.cta_group::1.kind::f16 (M=128, N<=256, K=16) issued by ONE thread; tcgen05.ld 32x32b in the epilogue: a
//   thread owns a whole output row, so the sum of squares needs no cross-thread reduction inside the CTA; the partial sums
//   of the cluster's column tiles are exchanged through distributed shared memory (mapa + ld.shared::cluster).
// Arithmetic intensity is low (2*N*L*D flop over ~2*(N*L + L*D + N*D) bytes): the kernel is sized to stream,
// not to saturate the tensor pipe; see DESIGN.md section 4 for the ceiling.
#include "hgl_common.cuh"

namespace hgl {

constexpr int kMpThreads = 256;   // warps 0-3 own the TMEM lanes in the epilogue; all 8 warps stage operands and store rows
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
  int B, M, L, D, Kp, Nw, NT;   // Kp = L rounded up to 64, Nw = columns per CTA (accumulator width), NT = D / Nw = cluster size
  int normalize, out_bf16;
  int stages;                // B ring depth (2..4)
  void* out;                 // [M, D] f32 | bf16
  uint32_t tmem_cols;
};

// cluster helpers (the NT CTAs of a cluster own the NT column tiles of the same 128 rows)
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float ld_peer_f32(const float* local_smem_ptr, uint32_t cta_rank) {
  uint32_t raddr;
  float v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(local_smem_ptr)), "r"(cta_rank));
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(raddr) : "memory");
  return v;
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

constexpr int kMpPiece = 64;      // output columns staged through shared memory per epilogue round

// 16-byte asynchronous global -> shared copy (LDGSTS) and its group bookkeeping
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// grid = (m-tiles of 128 masks, images, NT column tiles), cluster = (1, 1, NT): the CTAs of a cluster share the rows and
// exchange their partial sums of squares through distributed shared memory, so every accumulator stays in TMEM until it
// is normalised -- one pass, no round trip through global memory.
__global__ void __launch_bounds__(kMpThreads, 1) mask_pool_kernel(const PoolParams p) {
  extern __shared__ __align__(128) uint8_t smp[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.y, nt = blockIdx.z;                    // nt == rank of this CTA in its cluster
  int n_lo = 0, n_hi = p.M;
  if (p.mask_off) { n_lo = p.mask_off[b]; n_hi = p.mask_off[b + 1]; }
  n_lo += blockIdx.x * kMpM;
  const int rows = min(kMpM, n_hi - n_lo);
  if (rows <= 0) return;                                        // uniform for the whole cluster
  const int L = p.L, D = p.D, Kp = p.Kp, Nw = p.Nw;

  // shared-memory carve-up
  uint64_t* bars = reinterpret_cast<uint64_t*>(smp);            // slot_free[4], acc_full
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smp + 64);
  float* ssq = reinterpret_cast<float*>(smp + 128);             // [2][128] partial sums of squares per row (two column halves)
  uint8_t* a_s = smp + 128 + 1024;                              // [128 x Kp] bf16: core(rg, kc) at rg * a_sbo + kc * 128
  const uint32_t a_sbo = (uint32_t)(Kp / 8) * 128u;             // stride between 8-row groups
  uint8_t* b_s = a_s + (size_t)kMpM * Kp * 2;                   // ring of [64 x Nw] bf16 slots: core(kb, nc) at kb * b_lbo + nc * 128
  const uint32_t b_lbo = (uint32_t)(Nw / 8) * 128u;             // stride between 8-token blocks
  const uint32_t b_slot = (uint32_t)kMpKC * Nw * 2;
  const int S = p.stages;

  if (warp == 0) tmem_alloc(tmem_slot, p.tmem_cols);
  if (tid == 0) {
    for (int i = 0; i < 5; ++i) mbar_init(bars + i, 1);
    mbar_fence_init();
  }
  const __nv_bfloat16* tok = p.tok + (size_t)b * L * D + (size_t)nt * Nw;
  const int NKC = Kp / kMpKC;
  const int cores = Nw / 8;
  // B chunk kc: tokens [kc*64, +64) x this CTA's Nw columns as 16-byte asynchronous copies straight into the canonical
  // MN-major layout; lane -> (token % 8, 4 column cores): a warp reads 8 tokens x 64 contiguous bytes (whole sectors) and
  // lands on 32 distinct 16-byte bank groups.  Tokens beyond L are zero-filled.
  auto load_chunk = [&](int kc) {
    uint8_t* bs = b_s + (size_t)(kc % S) * b_slot;
    for (int t = tid; t < kMpKC * cores; t += kMpThreads) {
      const int kk = t & 7, rest = t >> 3;
      const int nc = rest % cores, kl = (rest / cores) * 8 + kk;
      const int k = kc * kMpKC + kl;
      uint8_t* dst = bs + (size_t)(kl >> 3) * b_lbo + (size_t)nc * 128 + (kl & 7) * 16;
      if (k < L) cp_async16(dst, tok + (size_t)k * D + nc * 8);
      else *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
    }
  };
  for (int kc = 0; kc < S - 1; ++kc) {                          // prologue: S-1 chunks in flight before A is even converted
    if (kc < NKC) load_chunk(kc);
    cp_async_commit();
  }
  // ---- A: soft masks f32 -> bf16 into the canonical K-major layout (rows beyond the image and k >= L are zero), one
  //      64-column chunk at a time so that only the first chunk's load latency is exposed: chunk kc+1 is converted while the
  //      MMAs of chunk kc run.  A lane reads one full 32-byte sector (8 floats of its row) per item and writes one 16-byte
  //      core row (conflict-free); the 4 items of a thread are loaded before any is converted.
  const bool w_vec = (L & 3) == 0 && (reinterpret_cast<uintptr_t>(p.w) & 15) == 0;
  auto convert_a = [&](int kc) {
    constexpr int kItems = kMpM * (kMpKC / 8) / kMpThreads;      // 4
    float4 f[kItems][2];
#pragma unroll
    for (int it = 0; it < kItems; ++it) {
      const int t = tid + it * kMpThreads;
      const int c8 = kc * (kMpKC / 8) + t / kMpM, r = t % kMpM;  // consecutive threads -> consecutive rows
      f[it][0] = make_float4(0.f, 0.f, 0.f, 0.f); f[it][1] = f[it][0];
      if (r < rows && c8 * 8 < L) {
        const float* src = p.w + (size_t)(n_lo + r) * L + c8 * 8;
        if (w_vec && c8 * 8 + 8 <= L) {
          f[it][0] = __ldg(reinterpret_cast<const float4*>(src)); f[it][1] = __ldg(reinterpret_cast<const float4*>(src) + 1);
        } else {
          float v[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) v[q] = (c8 * 8 + q < L) ? __ldg(src + q) : 0.f;
          f[it][0] = make_float4(v[0], v[1], v[2], v[3]); f[it][1] = make_float4(v[4], v[5], v[6], v[7]);
        }
      }
    }
#pragma unroll
    for (int it = 0; it < kItems; ++it) {
      const int t = tid + it * kMpThreads;
      const int c8 = kc * (kMpKC / 8) + t / kMpM, r = t % kMpM;
      uint4 o;
      o.x = pack_bf16x2(f[it][0].x, f[it][0].y); o.y = pack_bf16x2(f[it][0].z, f[it][0].w);
      o.z = pack_bf16x2(f[it][1].x, f[it][1].y); o.w = pack_bf16x2(f[it][1].z, f[it][1].w);
      *reinterpret_cast<uint4*>(a_s + (size_t)(r >> 3) * a_sbo + (size_t)c8 * 128 + (r & 7) * 16) = o;
    }
  };
  convert_a(0);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // instruction descriptor: D=f32, A=B=bf16, A K-major, B MN-major, N = Nw, M = 128
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (0u << 15) | (1u << 16) | ((uint32_t)(Nw >> 3) << 17) | ((uint32_t)(kMpM >> 4) << 24);
  for (int kc = 0; kc < NKC; ++kc) {
    // keep S-1 chunks in flight: chunk kc+S-1 goes into the slot chunk kc-1 used, once its MMAs have read it
    const int nxt = kc + S - 1;
    if (nxt < NKC) {
      const int use = nxt / S;
      if (use > 0) mbar_wait(bars + (nxt % S), (uint32_t)((use - 1) & 1));
      load_chunk(nxt);
    }
    cp_async_commit();
    if (S == 4) cp_async_wait<3>(); else if (S == 3) cp_async_wait<2>(); else cp_async_wait<1>();   // chunk kc has landed (this thread's copies)
    proxy_fence_async();                                         // generic-proxy smem writes -> visible to the tensor core (async proxy)
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t a_addr = smem_u32(a_s) + (uint32_t)(kc * (kMpKC / 8)) * 128u;
      const uint32_t b_addr = smem_u32(b_s + (size_t)(kc % S) * b_slot);
#pragma unroll
      for (int s2 = 0; s2 < kMpKC / 16; ++s2) {
        const uint64_t ad = smem_desc(a_addr + (uint32_t)s2 * 256u, 128u, a_sbo);             // 2 K-cores, 128 B apart
        const uint64_t bd = smem_desc(b_addr + (uint32_t)s2 * 2u * b_lbo, b_lbo, 128u);       // 2 token blocks, b_lbo apart
        umma_bf16(tmem_base, ad, bd, idesc, (kc > 0 || s2 > 0) ? 1u : 0u);
      }
      umma_commit(bars + (kc % S));                              // slot reusable once these MMAs have read it
      if (kc == NKC - 1) umma_commit(bars + 4);                  // accumulator complete
    }
    if (kc + 1 < NKC) convert_a(kc + 1);                         // under the MMAs of chunk kc; fenced + barriered next iteration
  }

  // ---- epilogue.  Warp w reads TMEM lanes 32*(w%4)..+31 (= output rows); warps 0-3 take the low half of the columns,
  //      warps 4-7 the high half.  Pass A: sum of squares of the row over this CTA's columns.
  mbar_wait(bars + 4, 0u);
  tc_fence_after();
  const int half = warp >> 2;
  const int my_row = (warp & 3) * 32 + lane;
  const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  float inv = 1.f;
  if (p.normalize) {
    float s1 = 0.f;
    for (int c = half * 16; c < Nw; c += 32) {                   // 16-column groups interleaved between the two halves
      float v[16];
      tmem_ld16(taddr + (uint32_t)c, v);
#pragma unroll
      for (int q = 0; q < 16; ++q) s1 += v[q] * v[q];
    }
    ssq[half * kMpM + my_row] = s1;
    if (p.NT > 1) cluster_sync_all(); else __syncthreads();
    float tot = 0.f;
    for (int c = 0; c < p.NT; ++c)
      tot += (p.NT > 1) ? ld_peer_f32(ssq + my_row, (uint32_t)c) + ld_peer_f32(ssq + kMpM + my_row, (uint32_t)c) : ssq[my_row] + ssq[kMpM + my_row];
    inv = __frcp_rn(sqrtf(tot));
  }
  // pass B: scale, convert, stage 64 columns at a time through shared memory (the B ring is idle: every MMA has completed)
  float* stage = reinterpret_cast<float*>(b_s);                  // [128][kMpPiece + 1] f32
  constexpr int kPitch = kMpPiece + 1;
  const size_t out_col0 = (size_t)nt * Nw;
  for (int c0 = 0; c0 < Nw; c0 += kMpPiece) {
    const int pw = min(kMpPiece, Nw - c0);
    for (int c = half * 16; c < pw; c += 32) {
      float v[16];
      tmem_ld16(taddr + (uint32_t)(c0 + c), v);
#pragma unroll
      for (int q = 0; q < 16; ++q) stage[my_row * kPitch + c + q] = v[q] * inv;
    }
    __syncthreads();
    // rows x pw columns -> global, consecutive threads along a row
    for (int t = tid; t < rows * (pw / 4); t += kMpThreads) {
      const int r = t / (pw / 4), q4 = t - r * (pw / 4);
      const float* sp = stage + r * kPitch + q4 * 4;
      const size_t o = (size_t)(n_lo + r) * D + out_col0 + c0 + q4 * 4;
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
  tc_fence_before();
  if (p.NT > 1) cluster_sync_all(); else __syncthreads();        // peers may still be reading ssq[]; TMEM reads are complete
  if (warp == 0) tmem_dealloc(tmem_base, p.tmem_cols);
}

}  // namespace hgl

extern "C" int64_t hgl_mask_pool_workspace_bytes(int M, int D, int out_dtype) {
  if (M < 0 || D < 1) return -1;
  (void)out_dtype;
  return 256;       // accumulators stay in TMEM until normalised: no global scratch (kept in the ABI for future tilings)
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
  HGL_REQUIRE(D >= 16 && D % 16 == 0, "hgl_mask_pool: D=%d must be a multiple of 16", D);
  HGL_REQUIRE(((reinterpret_cast<uintptr_t>(tokens) | reinterpret_cast<uintptr_t>(out)) & 15) == 0, "hgl_mask_pool: tokens / out must be 16-byte aligned");
  HGL_REQUIRE(B <= 65535, "hgl_mask_pool: B=%d too large for one launch", B);
  PoolParams p;
  p.w = weights; p.tok = reinterpret_cast<const __nv_bfloat16*>(tokens); p.mask_off = mask_off;
  p.B = B; p.M = M; p.L = L; p.D = D;
  p.Kp = ceil_div(L, kMpKC) * kMpKC;
  // columns per CTA: the widest multiple of 16 (<= 256) that tiles D with at most 8 CTAs per cluster
  int nw = std::min(D, kMpMaxN);
  while (D % nw != 0 || nw % 16 != 0) nw -= 16;
  p.Nw = nw; p.NT = D / nw;
  HGL_REQUIRE(p.NT <= 8, "hgl_mask_pool: D=%d needs %d column tiles (> 8 CTAs per cluster)", D, p.NT);
  p.normalize = normalize ? 1 : 0; p.out_bf16 = out_dtype == HGL_BF16;
  p.out = out;
  uint32_t cols = 32;
  while (cols < (uint32_t)nw) cols *= 2;
  p.tmem_cols = cols;
  const size_t slot = (size_t)kMpKC * nw * 2;
  const size_t stage = (size_t)kMpM * (kMpPiece + 1) * 4;                       // the output stage lives in the B ring
  const size_t fixed = 128 + 1024 + (size_t)kMpM * p.Kp * 2;
  int stages = 4;
  while (stages > 2 && fixed + std::max((size_t)stages * slot, stage) > 227 * 1024) --stages;
  p.stages = stages;
  const size_t smem = fixed + std::max((size_t)stages * slot, stage);
  HGL_REQUIRE(smem <= 227 * 1024, "hgl_mask_pool: L=%d needs %zu B of shared memory (limit 227 KB)", L, smem);
  cudaError_t e = cudaFuncSetAttribute(mask_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("hgl_mask_pool: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return HGL_ECUDA; }
  const int per_image = (B == 1) ? M : std::min(max_n, M);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ceil_div(per_image, kMpM), B, p.NT);
  cfg.blockDim = dim3(kMpThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = (unsigned)p.NT;
  cfg.attrs = attr; cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, mask_pool_kernel, p);
  if (e != cudaSuccess) { set_error("hgl_mask_pool: cudaLaunchKernelEx: %s", cudaGetErrorString(e)); return HGL_ECUDA; }
  return launch_status("hgl_mask_pool");
}
