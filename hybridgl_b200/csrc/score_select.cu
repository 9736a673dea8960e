// (a6)-(a9),(a12) cosine scoring + spatial-relationship re-ranking + per-expression argmax, one launch per batch.
//
// Replaces, per expression, ~40 tiny torch kernels and 9-18 device->host syncs of the reference:
//   text ensemble / negatives        Hybridgl_main.py:153-166
//   calculate_score                   model/backbone.py:74-87
//   argmax, softmax, top-k            Hybridgl_main.py:168-183
//   relation_boxes double loop        Hybridgl_main.py:185-196, utils.py:240-268
//   blend with score_gem, argmax      Hybridgl_main.py:225-227
//
// Grid = (chunks of 8 masks, images).  A CTA builds the text vectors of its image's expressions in shared memory
// (r*sent + (1-r)*noun, mean of the negatives), then every WARP owns one mask row of the feature matrix: 8-byte (bf16) /
// 16-byte (f32) coalesced loads, 1 + 2*4 running dot products (|f|^2, f.text_j, f.neg_j), shuffle-tree reductions, scores
// written to global memory.  The LAST CTA of an image to finish (atomic ticket after a __threadfence) runs the selection
// tail -- soft-max over masks, top-3 / top-6, 3x6 box relations, blend with score_gem, argmax -- one warp per expression,
// so the whole of (c) is still one kernel ending in the per-expression argmax, but the feature stream is spread over
// B*ceil(n/8) CTAs instead of B.  HBM traffic per image ~ n*De*b + small (SURVEY 8(d) row c).
#include "hgl_common.cuh"

namespace hgl {

constexpr int kEG = 4;            // expressions per pass over the features
constexpr int kScoreThreads = 256;
constexpr int kScoreWarps = kScoreThreads / 32;   // = mask rows per CTA

struct ScoreParams {
  const void* feat; int feat_bf16;
  const float* sent; const float* noun; const float* others; const int32_t* other_off;
  const int64_t* boxes; const int32_t* relaflag; const float* score_gem;
  const int32_t* mask_off; const int32_t* expr_off;
  int B, M, E, De, max_n;
  float scale, r, one_minus_r, alpha, one_minus_alpha;
  float* score_clip; int64_t* idx_hybrid; int64_t* idx_final; int32_t* top_idx; float* blended;
  float* score_neg;       // workspace [E, max_n]
  int32_t* tickets;       // workspace [B], zero at launch
  float* txt;             // workspace [E, De]  r*sent + (1-r)*noun
  float* neg;             // workspace [E, De]  mean of the 'a photo of <other noun>' embeddings (zeros if none)
  float* tnorm;           // workspace [E, 2]   |txt|, |neg|
};

// (a6) text side, Hybridgl_main.py:153-164: one CTA per expression, done once instead of once per feature chunk
__global__ void __launch_bounds__(128) score_text_kernel(const ScoreParams p) {
  const int e = blockIdx.x, tid = threadIdx.x, De = p.De;
  const int k0 = p.other_off[e], k1 = p.other_off[e + 1];
  float st = 0.f, sn = 0.f;
  for (int d = tid; d < De; d += blockDim.x) {
    const float t = __fadd_rn(__fmul_rn(p.r, p.sent[(size_t)e * De + d]), __fmul_rn(p.one_minus_r, p.noun[(size_t)e * De + d]));
    float a = 0.f;
    for (int k = k0; k < k1; ++k) a = __fadd_rn(a, p.others[(size_t)k * De + d]);
    if (k1 > k0) a = __fdiv_rn(a, (float)(k1 - k0));
    p.txt[(size_t)e * De + d] = t;
    p.neg[(size_t)e * De + d] = a;
    st += t * t; sn += a * a;
  }
  __shared__ float red[2][4];
  st = warp_sum(st); sn = warp_sum(sn);
  if ((tid & 31) == 0) { red[0][tid >> 5] = st; red[1][tid >> 5] = sn; }
  __syncthreads();
  if (tid == 0) {
    p.tnorm[2 * e] = sqrtf(red[0][0] + red[0][1] + red[0][2] + red[0][3]);
    p.tnorm[2 * e + 1] = sqrtf(red[1][0] + red[1][1] + red[1][2] + red[1][3]);
  }
}

// relation_boxes utils.py:240-268 (boxes XYWH int64; torch promotes to float32 for the divisions)
__device__ float relation(const int64_t* bi, const int64_t* bj, float si, float sj, int rel) {
  switch (rel) {
    case HGL_REL_LEFT: return si * sj * (((float)bi[0] + (float)bi[2] / 2.f) < ((float)bj[0] + (float)bj[2] / 2.f) ? 1.f : 0.f);
    case HGL_REL_RIGHT: return si * sj * (((float)bi[0] + (float)bi[2] / 2.f) > ((float)bj[0] + (float)bj[2] / 2.f) ? 1.f : 0.f);
    case HGL_REL_UP: return si * sj * (((float)bi[1] + (float)bi[3] / 2.f) < ((float)bj[1] + (float)bj[3] / 2.f) ? 1.f : 0.f);
    case HGL_REL_DOWN: return si * sj * (((float)bi[1] + (float)bi[3] / 2.f) > ((float)bj[1] + (float)bj[3] / 2.f) ? 1.f : 0.f);
    case HGL_REL_BIG: return si * sj * ((bi[2] * bi[3]) > (bj[2] * bj[3]) ? 1.f : 0.f);
    case HGL_REL_SMALL: return si * sj * ((bi[2] * bi[3]) < (bj[2] * bj[3]) ? 1.f : 0.f);
    case HGL_REL_WITHIN: {
      const int64_t x1 = max(bi[0], bj[0]);
      const int64_t x2 = max(x1, min(bi[0] + bi[2], bj[0] + bj[2]));
      const int64_t y1 = max(bi[1], bj[1]);
      const int64_t y2 = max(y1, min(bi[1] + bi[3], bj[1] + bj[3]));
      return __fdiv_rn(__fmul_rn(__fmul_rn(__fmul_rn(si, sj), (float)(x2 - x1)), (float)(y2 - y1)), (float)(bi[2] * bi[3]));
    }
    default: return si;   // 'none' and unknown words
  }
}

// torch.argmax / topk ordering: larger wins, NaN counts as the largest, lower index wins ties
__device__ __forceinline__ bool better(float v, int i, float bv, int bi) {
  if (bi < 0) return true;
  const bool vn = isnan(v), bn = isnan(bv);
  if (vn != bn) return vn;
  if (!vn && v != bv) return v > bv;
  return i < bi;
}
__device__ __forceinline__ void warp_argbest(float& v, int& i) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, i, o);
    if (oi >= 0 && better(ov, oi, v, i)) { v = ov; i = oi; }
  }
}

// soft-max over n values in shared memory (in place), one warp; torch.nn.Softmax(0) on [n,1]
__device__ void warp_softmax(float* x, int n, int lane) {
  float mx = -INFINITY;
  for (int i = lane; i < n; i += 32) mx = fmaxf(mx, x[i]);
  mx = warp_max(mx);
  float s = 0.f;
  for (int i = lane; i < n; i += 32) { const float e = expf(x[i] - mx); x[i] = e; s += e; }
  s = warp_sum(s);
  for (int i = lane; i < n; i += 32) x[i] = __fdiv_rn(x[i], s);
  __syncwarp();
}

// indices of the k largest entries (descending), one warp; `out` in shared memory
__device__ void warp_topk(const float* x, int n, int k, int* out, int lane) {
  for (int t = 0; t < k; ++t) {
    float bv = 0.f; int bi = -1;
    for (int i = lane; i < n; i += 32) {
      bool taken = false;
      for (int u = 0; u < t; ++u) taken |= (out[u] == i);
      if (!taken && better(x[i], i, bv, bi)) { bv = x[i]; bi = i; }
    }
    warp_argbest(bv, bi);
    if (lane == 0) out[t] = bi;
    __syncwarp();
  }
}

__global__ void __launch_bounds__(kScoreThreads) score_select_kernel(const ScoreParams p) {
  extern __shared__ __align__(16) float sm[];
  const int De = p.De, max_n = p.max_n;
  __shared__ int s_last;

  const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int n_lo = 0, n_hi = p.M, e_lo = 0, e_hi = p.E;
  if (p.mask_off) { n_lo = p.mask_off[b]; n_hi = p.mask_off[b + 1]; }
  if (p.expr_off) { e_lo = p.expr_off[b]; e_hi = p.expr_off[b + 1]; }
  const int n = min(n_hi - n_lo, max_n);
  const int nchunks = max(1, (n + kScoreWarps - 1) / kScoreWarps);
  if ((int)blockIdx.x >= nchunks) return;                       // grid.x is sized for the largest image
  const int m = blockIdx.x * kScoreWarps + warp;               // this warp's mask row (image-local)

  // ---- cosine scores of this warp's mask row (model/backbone.py:79-85); lane owns features 4*lane + 128*i.
  //      The text vectors come from score_text_kernel (L2 / L1 resident, shared by every CTA of the image).
  if (m < n) {
    const size_t row = (size_t)(n_lo + m) * De;
    for (int eg = e_lo; eg < e_hi; eg += kEG) {
      const int ne = min(kEG, e_hi - eg);
      float ff = 0.f, dt[kEG], dn[kEG];
#pragma unroll
      for (int j = 0; j < kEG; ++j) { dt[j] = 0.f; dn[j] = 0.f; }
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
            const float4 t4 = __ldg(reinterpret_cast<const float4*>(p.txt + (size_t)(eg + j) * De + d0));
            const float4 n4 = __ldg(reinterpret_cast<const float4*>(p.neg + (size_t)(eg + j) * De + d0));
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
            p.score_clip[(size_t)(eg + j) * max_n + m] = p.scale * __fdiv_rn(__fdiv_rn(a, fnorm), p.tnorm[2 * (eg + j)]);
            p.score_neg[(size_t)(eg + j) * max_n + m] = p.scale * __fdiv_rn(__fdiv_rn(c, fnorm), p.tnorm[2 * (eg + j) + 1]);
          }
        }
      }
    }
  }

  // ---- ticket: the last CTA of this image runs the selection tail
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(p.tickets + b, 1) == nchunks - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();

  float* sc = sm;                                   // [kScoreWarps][max_n]  score_clip, later soft-max p
  float* sn = sc + kScoreWarps * max_n;             // [kScoreWarps][max_n]  score_clip_Neg, later soft-max q
  int* picks = reinterpret_cast<int*>(sn + kScoreWarps * max_n);   // [kScoreWarps][3 + 6]
  for (int e = e_lo + warp; e < e_hi; e += kScoreWarps) {        // one warp per expression, warp-private shared rows
    float* s = sc + warp * max_n;
    float* q = sn + warp * max_n;
    int* top = picks + warp * 9;
    int* topn = top + 3;
    for (int i = lane; i < n; i += 32) {
      s[i] = __ldcg(p.score_clip + (size_t)e * max_n + i);
      q[i] = __ldcg(p.score_neg + (size_t)e * max_n + i);
    }
    for (int i = n + lane; i < max_n; i += 32) p.score_clip[(size_t)e * max_n + i] = 0.f;
    __syncwarp();
    float bv = 0.f; int bi = -1;                                   // :168 argmax
    for (int i = lane; i < n; i += 32) if (better(s[i], i, bv, bi)) { bv = s[i]; bi = i; }
    warp_argbest(bv, bi);
    const int n_other = p.other_off[e + 1] - p.other_off[e];
    warp_softmax(s, n, lane);                                       // :173
    const int k1 = min(3, n), k2 = min(6, n);                       // :178-181
    warp_topk(s, n, k1, top, lane);                                 // :182
    if (n_other > 0) { warp_softmax(q, n, lane); warp_topk(q, n, k2, topn, lane); }   // :174,:183
    __syncwarp();
    // relation sums (:185-193), lanes 0..k1-1, sequential fp32 accumulation over j like the reference
    float T = 0.f;
    const int rel = p.relaflag[e];
    if (lane < k1) {
      const int ti = top[lane];
      const int64_t* bi4 = p.boxes + (size_t)(n_lo + ti) * 4;
      const int cntj = (n_other == 0) ? k1 : k2;
      for (int u = 0; u < cntj; ++u) {
        const int tj = (n_other == 0) ? top[u] : topn[u];
        const float sj = (n_other == 0) ? s[tj] : q[tj];
        T = __fadd_rn(T, relation(bi4, p.boxes + (size_t)(n_lo + tj) * 4, s[ti], sj, rel));
      }
    }
    // softmax over the k1 values (:196)
    float mx = (lane < k1) ? T : -INFINITY;
    mx = warp_max(mx);
    float ex = (lane < k1) ? expf(T - mx) : 0.f;
    const float sum = warp_sum(ex);
    float Tn = __fdiv_rn(ex, sum);
    if (p.score_gem != nullptr && lane < k1)                        // :225-226
      Tn = __fadd_rn(__fmul_rn(Tn, p.one_minus_alpha), __fmul_rn(p.alpha, p.score_gem[(size_t)e * max_n + top[lane]]));
    float fv = Tn; int fi = (lane < k1) ? lane : -1;                // :227
    warp_argbest(fv, fi);
    if (lane < 3) {
      p.top_idx[(size_t)e * 3 + lane] = (lane < k1) ? top[lane] : -1;
      p.blended[(size_t)e * 3 + lane] = (lane < k1) ? Tn : 0.f;
    }
    if (lane == 0) {
      p.idx_hybrid[e] = bi;
      p.idx_final[e] = (fi >= 0) ? top[fi] : -1;
    }
    __syncwarp();
  }
}

}  // namespace hgl

extern "C" int64_t hgl_score_select_workspace_bytes(int B, int E, int max_n) {
  if (B < 1 || E < 0 || max_n < 1) return -1;
  // negative scores [E,max_n] | tickets [B] | txt, neg [E,De] each (De <= 4096 assumed for sizing) | norms [E,2]
  return (int64_t)(((size_t)E * max_n * 4 + 255) & ~size_t(255)) + (int64_t)(((size_t)B * 4 + 255) & ~size_t(255)) +
         2 * (int64_t)(((size_t)E * 4096 * 4 + 255) & ~size_t(255)) + (int64_t)(((size_t)E * 8 + 255) & ~size_t(255));
}

extern "C" int hgl_score_select(const void* feat, int feat_dtype, const float* sent, const float* noun, const float* others,
                                const int32_t* other_off, const int64_t* boxes, const int32_t* relaflag, const float* score_gem,
                                const int32_t* mask_off, const int32_t* expr_off, int B, int M, int E, int De, int max_n,
                                double logit_scale_exp, double r, double alpha,
                                float* score_clip, int64_t* idx_hybrid, int64_t* idx_final, int32_t* top_idx, float* blended,
                                void* workspace, void* stream) {
  using namespace hgl;
  HGL_REQUIRE(feat && sent && noun && other_off && boxes && relaflag && score_clip && idx_hybrid && idx_final && top_idx && blended && workspace,
              "hgl_score_select: null pointer");
  HGL_REQUIRE(feat_dtype == HGL_F32 || feat_dtype == HGL_BF16, "hgl_score_select: feat_dtype %d", feat_dtype);
  HGL_REQUIRE(B >= 1 && M >= 0 && E >= 0 && max_n >= 1, "hgl_score_select: bad shape");
  HGL_REQUIRE(De >= 8 && De % 8 == 0 && De <= 4096, "hgl_score_select: De=%d must be a multiple of 8 in [8,4096]", De);
  HGL_REQUIRE((mask_off && expr_off) || B == 1, "hgl_score_select: mask_off/expr_off required when B > 1");
  HGL_REQUIRE((reinterpret_cast<uintptr_t>(feat) & 15) == 0, "hgl_score_select: feat must be 16-byte aligned");
  HGL_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "hgl_score_select: workspace must be 256-byte aligned");
  HGL_REQUIRE(B <= 65535, "hgl_score_select: B=%d too large for one launch", B);
  if (E == 0) return HGL_OK;
  cudaStream_t st = (cudaStream_t)stream;
  ScoreParams p;
  p.feat = feat; p.feat_bf16 = (feat_dtype == HGL_BF16);
  p.sent = sent; p.noun = noun; p.others = others; p.other_off = other_off;
  p.boxes = boxes; p.relaflag = relaflag; p.score_gem = score_gem;
  p.mask_off = mask_off; p.expr_off = expr_off;
  p.B = B; p.M = M; p.E = E; p.De = De; p.max_n = max_n;
  p.scale = (float)logit_scale_exp; p.r = (float)r; p.one_minus_r = (float)(1.0 - r);
  p.alpha = (float)alpha; p.one_minus_alpha = (float)(1.0 - alpha);
  p.score_clip = score_clip; p.idx_hybrid = idx_hybrid; p.idx_final = idx_final; p.top_idx = top_idx; p.blended = blended;
  const size_t neg_bytes = ((size_t)E * max_n * 4 + 255) & ~size_t(255);
  p.score_neg = reinterpret_cast<float*>(workspace);
  uint8_t* wsb = reinterpret_cast<uint8_t*>(workspace);
  p.tickets = reinterpret_cast<int32_t*>(wsb + neg_bytes);
  const size_t tick_bytes = ((size_t)B * 4 + 255) & ~size_t(255), vec_bytes = ((size_t)E * 4096 * 4 + 255) & ~size_t(255);
  p.txt = reinterpret_cast<float*>(wsb + neg_bytes + tick_bytes);
  p.neg = reinterpret_cast<float*>(wsb + neg_bytes + tick_bytes + vec_bytes);
  p.tnorm = reinterpret_cast<float*>(wsb + neg_bytes + tick_bytes + 2 * vec_bytes);
  cudaError_t e = cudaMemsetAsync(p.tickets, 0, (size_t)B * 4, st);
  if (e != cudaSuccess) { set_error("hgl_score_select: cudaMemsetAsync: %s", cudaGetErrorString(e)); return HGL_ECUDA; }
  score_text_kernel<<<E, 128, 0, st>>>(p);
  int rc = launch_status("hgl_score_select(text)");
  if (rc != HGL_OK) return rc;
  const size_t smem = ((size_t)2 * kScoreWarps * max_n) * 4 + (size_t)kScoreWarps * 9 * 4 + 16;
  HGL_REQUIRE(smem <= 200 * 1024, "hgl_score_select: De=%d max_n=%d needs %zu B of shared memory", De, max_n, smem);
  e = cudaFuncSetAttribute(score_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("hgl_score_select: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return HGL_ECUDA; }
  const int per_image = (B == 1) ? M : std::min(max_n, M);
  dim3 grid(std::max(1, ceil_div(per_image, kScoreWarps)), B);
  score_select_kernel<<<grid, kScoreThreads, smem, st>>>(p);
  return launch_status("hgl_score_select");
}
