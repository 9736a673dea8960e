// (a6)-(a9),(a12) cosine scoring + spatial-relationship re-ranking + per-expression argmax for SUPPLIED features
// (the hybrid CLIP features of CLIPViTFM.forward), one launch per batch.  When the features are pooled from dense tokens the
// same scoring runs inside the tensor-core kernel instead (pool_score.cu).
//
// Replaces, per expression, ~40 tiny torch kernels and 9-18 device->host syncs of the reference:
//   text ensemble / negatives        Hybridgl_main.py:153-166
//   calculate_score                   model/backbone.py:74-87
//   argmax, softmax, top-k            Hybridgl_main.py:168-183
//   relation_boxes double loop        Hybridgl_main.py:185-196, utils.py:240-268
//   blend with score_gem, argmax      Hybridgl_main.py:225-227     (select_tail.cuh)
//
// Grid = (CS, images) with the CS (<= 8) CTAs of an image forming a thread-block cluster.  Every CTA builds the text vectors
// of (up to 4) expressions in shared memory (r*sent + (1-r)*noun, mean of the negatives, their norms), then each WARP owns
// mask rows of the CTA's slice of the feature matrix: 8-byte (bf16) / 16-byte (f32) coalesced loads, 1 + 2*4 running dot
// products (|f|^2, f.text_j, f.neg_j), shuffle-tree reductions.  The scores are stored straight into the rank-0 CTA's shared
// memory (st.shared::cluster); after one cluster barrier rank 0 runs the selection tail, two warps per expression.  One
// kernel, no workspace, no memset, no atomics.  HBM traffic per image ~ n*De*b + small (SURVEY 8(d) row c).
#include <algorithm>

#include "hgl_common.cuh"
#include "select_tail.cuh"

namespace hgl {

constexpr int kEG = 4;            // expressions per round
constexpr int kScoreThreads = 256;
constexpr int kScoreWarps = kScoreThreads / 32;

struct ScoreParams {
  const void* feat; int feat_bf16;
  const float* sent; const float* noun; const float* others; const int32_t* other_off;
  const int32_t* mask_off; const int32_t* expr_off;
  int B, M, E, De, max_n, CS;
  float scale, r, one_minus_r;
  TailArgs tail;
};

__global__ void __launch_bounds__(kScoreThreads) score_select_kernel(const ScoreParams p) {
  extern __shared__ __align__(16) float sm[];
  const int De = p.De, max_n = p.max_n;
  const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  int n_lo = 0, n_hi = p.M, e_lo = 0, e_hi = p.E;
  if (p.mask_off) { n_lo = p.mask_off[b]; n_hi = p.mask_off[b + 1]; }
  if (p.expr_off) { e_lo = p.expr_off[b]; e_hi = p.expr_off[b + 1]; }
  const int n = min(max(n_hi - n_lo, 0), max_n);
  const int rpc = (n + p.CS - 1) / p.CS;                          // mask rows per CTA of the cluster
  const int row_lo = (int)rank * rpc, row_hi = min(n, row_lo + rpc);

  float* txt = sm;                                  // [2*kEG][De]: rows 0..3 text ensemble, 4..7 negatives
  float* tnorm = txt + 2 * kEG * De;                // [2*kEG]
  float* sc = tnorm + 2 * kEG;                      // [2*kEG][max_n]  rank 0: score_clip / score_clip_Neg rows of the round
  int* picks = reinterpret_cast<int*>(sc + 2 * kEG * max_n);   // [kEG][kTailPicks]
  int64_t* box_s = reinterpret_cast<int64_t*>(picks + kEG * kTailPicks);   // rank 0: the image's boxes [n][4] (8-byte aligned: even float count before)
  float* sg_s = reinterpret_cast<float*>(box_s + (size_t)max_n * 4);       // rank 0: score_gem rows of the round [kEG][max_n]
  int* meta_s = reinterpret_cast<int*>(sg_s + kEG * max_n);                // rank 0: (n_other, relaflag) of the round [kEG][2]
  // what the selection tail needs is fetched while the scores are still being computed (the tail is a dependent chain:
  // every global round trip it skips is 1-2 us)
  if (rank == 0 && e_hi > e_lo)
    for (int i = tid; i < n * 4; i += kScoreThreads) box_s[i] = __ldg(p.tail.boxes + (size_t)n_lo * 4 + i);
  auto prefetch_tail = [&](int eg, int ne) {
    if (tid < ne) {
      meta_s[2 * tid] = p.tail.other_off[eg + tid + 1] - p.tail.other_off[eg + tid];
      meta_s[2 * tid + 1] = p.tail.relaflag[eg + tid];
    }
    if (p.tail.score_gem != nullptr)
      for (int i = tid; i < ne * n; i += kScoreThreads) {
        const int j = i / n, c = i - j * n;
        sg_s[j * max_n + c] = __ldg(p.tail.score_gem + (size_t)(eg + j) * max_n + c);
      }
  };

  cluster_arrive();                                               // "this CTA runs": waited for before the first store into rank 0
  bool peers_started = false;
  for (int eg = e_lo; eg < e_hi; eg += kEG) {
    const int ne = min(kEG, e_hi - eg);
    if (rank == 0) prefetch_tail(eg, ne);
    // ---- (a6) text side, Hybridgl_main.py:153-164: warp w -> expression w % 4, ensemble (w < 4) or mean of the negatives;
    //      the loads of 8 columns per lane are issued together
    {
      const int j = warp & 3, kind = warp >> 2;
      const int e = eg + j;
      if (j < ne) {
        const int k0 = p.other_off[e], k1 = p.other_off[e + 1];
        float acc = 0.f;
        constexpr int kCU = 8;
        for (int d0 = 0; d0 < De; d0 += 32 * kCU) {
          float v[kCU];
          if (kind == 0) {
            float a[kCU], g[kCU];
#pragma unroll
            for (int ci = 0; ci < kCU; ++ci) {
              const int d = d0 + lane + 32 * ci;
              a[ci] = d < De ? __ldg(p.sent + (size_t)e * De + d) : 0.f;
              g[ci] = d < De ? __ldg(p.noun + (size_t)e * De + d) : 0.f;
            }
#pragma unroll
            for (int ci = 0; ci < kCU; ++ci) v[ci] = __fadd_rn(__fmul_rn(p.r, a[ci]), __fmul_rn(p.one_minus_r, g[ci]));
          } else {
#pragma unroll
            for (int ci = 0; ci < kCU; ++ci) v[ci] = 0.f;
            for (int kb = k0; kb < k1; kb += 2) {
              float o[2][kCU];
#pragma unroll
              for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int ci = 0; ci < kCU; ++ci) {
                  const int d = d0 + lane + 32 * ci;
                  o[u][ci] = (kb + u < k1 && d < De) ? __ldg(p.others + (size_t)(kb + u) * De + d) : 0.f;
                }
#pragma unroll
              for (int u = 0; u < 2; ++u)
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
#pragma unroll
          for (int ci = 0; ci < kCU; ++ci) {
            const int d = d0 + lane + 32 * ci;
            if (d < De) { txt[(kind * kEG + j) * De + d] = v[ci]; acc += v[ci] * v[ci]; }
          }
        }
        acc = warp_sum(acc);
        if (lane == 0) tnorm[kind * kEG + j] = sqrtf(acc);
      }
    }
    __syncthreads();
    if (!peers_started) { cluster_wait(); peers_started = true; }
    // ---- cosine scores of this CTA's mask rows (model/backbone.py:79-85); lane owns features 4*lane + 128*i
    for (int m = row_lo + warp; m < row_hi; m += kScoreWarps) {
      const size_t row = (size_t)(n_lo + m) * De;
      float ff = 0.f, dt[kEG], dn[kEG];
#pragma unroll
      for (int j = 0; j < kEG; ++j) { dt[j] = 0.f; dn[j] = 0.f; }
#pragma unroll 4
      for (int d0 = lane * 4; d0 < De; d0 += 128) {
        float f[4];
        if (p.feat_bf16) {
          const uint2 u = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(p.feat) + row + d0));
          f[0] = bf16_bits_to_float(u.x & 0xffffu); f[1] = bf16_bits_to_float(u.x >> 16);
          f[2] = bf16_bits_to_float(u.y & 0xffffu); f[3] = bf16_bits_to_float(u.y >> 16);
        } else {
          const float4 u = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.feat) + row + d0));
          f[0] = u.x; f[1] = u.y; f[2] = u.z; f[3] = u.w;
        }
        ff += f[0] * f[0] + f[1] * f[1] + f[2] * f[2] + f[3] * f[3];
#pragma unroll
        for (int j = 0; j < kEG; ++j) {
          if (j < ne) {
            const float4 t4 = *reinterpret_cast<const float4*>(txt + j * De + d0);
            const float4 n4 = *reinterpret_cast<const float4*>(txt + (kEG + j) * De + d0);
            dt[j] += f[0] * t4.x + f[1] * t4.y + f[2] * t4.z + f[3] * t4.w;
            dn[j] += f[0] * n4.x + f[1] * n4.y + f[2] * n4.z + f[3] * n4.w;
          }
        }
      }
      ff = warp_sum(ff);
      const float fnorm = sqrtf(ff);
#pragma unroll
      for (int j = 0; j < kEG; ++j) {
        if (j < ne) {
          const float a = warp_sum(dt[j]), c = warp_sum(dn[j]);
          if (lane == 0) {
            // scale * (f/|f|) . (t/|t|); a zero 'neg' vector gives 0/0 = NaN exactly like the reference (App. B-5)
            const float s_pos = p.scale * __fdiv_rn(__fdiv_rn(a, fnorm), tnorm[j]);
            const float s_neg = p.scale * __fdiv_rn(__fdiv_rn(c, fnorm), tnorm[kEG + j]);
            st_peer_f32(sc + j * max_n + m, 0u, s_pos);
            st_peer_f32(sc + (kEG + j) * max_n + m, 0u, s_neg);
            p.tail.score_clip[(size_t)(eg + j) * max_n + m] = s_pos;
          }
        }
      }
    }
    cluster_sync_all();                                          // every CTA's scores have landed in rank 0
    if (rank == 0) select_tail_block(p.tail, eg, ne, n, n_lo, sc, max_n, picks, box_s, p.tail.score_gem ? sg_s : nullptr, meta_s, warp, lane);
    if (eg + kEG < e_hi) cluster_sync_all();                     // next round overwrites txt (own) and sc (rank 0)
  }
}

}  // namespace hgl

extern "C" int64_t hgl_score_select_workspace_bytes(int B, int E, int max_n) {
  if (B < 1 || E < 0 || max_n < 1) return -1;
  return 0;      // scores travel through distributed shared memory: no global scratch (the argument stays in the ABI)
}

extern "C" int hgl_score_select(const void* feat, int feat_dtype, const float* sent, const float* noun, const float* others,
                                const int32_t* other_off, const int64_t* boxes, const int32_t* relaflag, const float* score_gem,
                                const int32_t* mask_off, const int32_t* expr_off, int B, int M, int E, int De, int max_n,
                                double logit_scale_exp, double r, double alpha,
                                float* score_clip, int64_t* idx_hybrid, int64_t* idx_final, int32_t* top_idx, float* blended,
                                void* workspace, void* stream) {
  using namespace hgl;
  (void)workspace;
  HGL_REQUIRE(feat && sent && noun && other_off && boxes && relaflag && score_clip && idx_hybrid && idx_final && top_idx && blended,
              "hgl_score_select: null pointer");
  HGL_REQUIRE(feat_dtype == HGL_F32 || feat_dtype == HGL_BF16, "hgl_score_select: feat_dtype %d", feat_dtype);
  HGL_REQUIRE(B >= 1 && M >= 0 && E >= 0 && max_n >= 1, "hgl_score_select: bad shape");
  HGL_REQUIRE(De >= 8 && De % 8 == 0 && De <= 4096, "hgl_score_select: De=%d must be a multiple of 8 in [8,4096]", De);
  HGL_REQUIRE((mask_off && expr_off) || B == 1, "hgl_score_select: mask_off/expr_off required when B > 1");
  HGL_REQUIRE((reinterpret_cast<uintptr_t>(feat) & 15) == 0, "hgl_score_select: feat must be 16-byte aligned");
  HGL_REQUIRE(B <= 65535, "hgl_score_select: B=%d too large for one launch", B);
  if (E == 0) return HGL_OK;
  ScoreParams p = {};
  p.feat = feat; p.feat_bf16 = (feat_dtype == HGL_BF16);
  p.sent = sent; p.noun = noun; p.others = others; p.other_off = other_off;
  p.mask_off = mask_off; p.expr_off = expr_off;
  p.B = B; p.M = M; p.E = E; p.De = De; p.max_n = max_n;
  p.scale = (float)logit_scale_exp; p.r = (float)r; p.one_minus_r = (float)(1.0 - r);
  p.tail.boxes = boxes; p.tail.relaflag = relaflag; p.tail.other_off = other_off; p.tail.score_gem = score_gem;
  p.tail.alpha = (float)alpha; p.tail.one_minus_alpha = (float)(1.0 - alpha); p.tail.max_n = max_n;
  p.tail.score_clip = score_clip; p.tail.idx_hybrid = idx_hybrid; p.tail.idx_final = idx_final; p.tail.top_idx = top_idx; p.tail.blended = blended;
  const int per_image = (B == 1) ? std::min(M, max_n) : std::min(max_n, M);
  p.CS = std::max(1, std::min(8, ceil_div(std::max(per_image, 1), 16)));       // >= 2 rows per warp before another CTA pays off
  const size_t smem = ((size_t)2 * kEG * De + 2 * kEG + (size_t)2 * kEG * max_n) * 4 + (size_t)kEG * kTailPicks * 4 + (size_t)max_n * 32 + (size_t)kEG * max_n * 4 + 2 * kEG * 4 + 16;
  HGL_REQUIRE(smem <= 227 * 1024, "hgl_score_select: De=%d max_n=%d needs %zu B of shared memory", De, max_n, smem);
  {
    const int rc_s = ensure_dyn_smem(reinterpret_cast<const void*>(score_select_kernel), smem, "hgl_score_select");
    if (rc_s != HGL_OK) return rc_s;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(p.CS, B, 1);
  cfg.blockDim = dim3(kScoreThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)p.CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1] = priority_attr((cudaStream_t)stream);
  cfg.attrs = attr; cfg.numAttrs = 2;
  cudaError_t e = cudaLaunchKernelEx(&cfg, score_select_kernel, p);
  if (e != cudaSuccess) { set_error("hgl_score_select: cudaLaunchKernelEx: %s", cudaGetErrorString(e)); return HGL_ECUDA; }
  return launch_status("hgl_score_select");
}
