// (a6)-(a9),(a12) cosine scoring + spatial-relationship re-ranking + per-expression argmax, one launch per batch.
//
// Replaces, per expression, ~40 tiny torch kernels and 9-18 device->host syncs of the reference:
//   text ensemble / negatives        Hybridgl_main.py:153-166
//   calculate_score                   model/backbone.py:74-87
//   argmax, softmax, top-k            Hybridgl_main.py:168-183
//   relation_boxes double loop        Hybridgl_main.py:185-196, utils.py:240-268
//   blend with score_gem, argmax      Hybridgl_main.py:225-227
//
// One CTA per image.  Features ([n, De], bf16 or f32) are streamed once per group of 4 expressions: a warp owns a
// mask row, keeps 1 + 2*4 running dot products (|f|^2, f.text_j, f.neg_j) and finishes with shuffle trees.  The
// selection tail (soft-max over masks, top-3 / top-6, 3x6 box relations, blend, argmax) runs one warp per
// expression entirely in registers / shared memory.  HBM traffic per image ~ n*De*b + small (SURVEY 8(d) row c).
#include "hgl_common.cuh"

namespace hgl {

constexpr int kEG = 4;            // expressions per pass over the features
constexpr int kScoreThreads = 256;

struct ScoreParams {
  const void* feat; int feat_bf16;
  const float* sent; const float* noun; const float* others; const int32_t* other_off;
  const int64_t* boxes; const int32_t* relaflag; const float* score_gem;
  const int32_t* mask_off; const int32_t* expr_off;
  int B, M, E, De, max_n;
  float scale, r, one_minus_r, alpha, one_minus_alpha;
  float* score_clip; int64_t* idx_hybrid; int64_t* idx_final; int32_t* top_idx; float* blended;
};

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
  float* txt = sm;                        // [kEG][De]  r*sent + (1-r)*noun
  float* neg = txt + kEG * De;            // [kEG][De]  mean of 'a photo of <other noun>' embeddings
  float* sc = neg + kEG * De;             // [kEG][max_n]  score_clip, later soft-max p
  float* sn = sc + kEG * max_n;           // [kEG][max_n]  score_clip_Neg, later soft-max q
  float* tnorm = sn + kEG * max_n;        // [2*kEG] 1/|text|, 1/|neg|
  int* picks = reinterpret_cast<int*>(tnorm + 2 * kEG);   // [kEG][3 + 6]

  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
  int n_lo = 0, n_hi = p.M, e_lo = 0, e_hi = p.E;
  if (p.mask_off) { n_lo = p.mask_off[b]; n_hi = p.mask_off[b + 1]; }
  if (p.expr_off) { e_lo = p.expr_off[b]; e_hi = p.expr_off[b + 1]; }
  const int n = min(n_hi - n_lo, max_n);

  for (int eg = e_lo; eg < e_hi; eg += kEG) {
    const int ne = min(kEG, e_hi - eg);
    // ---- text side (Hybridgl_main.py:153-164)
    for (int t = tid; t < ne * De; t += blockDim.x) {
      const int j = t / De, d = t - j * De;
      const int e = eg + j;
      txt[t] = __fadd_rn(__fmul_rn(p.r, p.sent[(size_t)e * De + d]), __fmul_rn(p.one_minus_r, p.noun[(size_t)e * De + d]));
      const int k0 = p.other_off[e], k1 = p.other_off[e + 1];
      float a = 0.f;
      for (int k = k0; k < k1; ++k) a = __fadd_rn(a, p.others[(size_t)k * De + d]);
      if (k1 > k0) a = __fdiv_rn(a, (float)(k1 - k0));
      neg[t] = a;
    }
    __syncthreads();
    for (int v = warp; v < 2 * ne; v += nwarp) {   // norms of the 2*ne text vectors
      const float* x = (v < ne) ? txt + v * De : neg + (v - ne) * De;
      float s = 0.f;
      for (int d = lane; d < De; d += 32) s += x[d] * x[d];
      s = warp_sum(s);
      if (lane == 0) tnorm[(v < ne) ? v : kEG + (v - ne)] = sqrtf(s);
    }
    __syncthreads();

    // ---- cosine scores: warp per mask row (model/backbone.py:79-85)
    for (int m = warp; m < n; m += nwarp) {
      float ff = 0.f, dt[kEG], dn[kEG];
#pragma unroll
      for (int j = 0; j < kEG; ++j) { dt[j] = 0.f; dn[j] = 0.f; }
      const size_t row = (size_t)(n_lo + m) * De;
      for (int d0 = lane * 8; d0 < De; d0 += 256) {   // 8 features per lane per step (16 B bf16 / 32 B f32)
        float f[8];
        if (p.feat_bf16) {
          const uint4 u = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.feat) + row + d0);
          const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) { f[2 * q] = bf16_bits_to_float(w[q] & 0xffffu); f[2 * q + 1] = bf16_bits_to_float(w[q] >> 16); }
        } else {
          const float4 u0 = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.feat) + row + d0);
          const float4 u1 = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.feat) + row + d0 + 4);
          f[0] = u0.x; f[1] = u0.y; f[2] = u0.z; f[3] = u0.w; f[4] = u1.x; f[5] = u1.y; f[6] = u1.z; f[7] = u1.w;
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          ff += f[q] * f[q];
#pragma unroll
          for (int j = 0; j < kEG; ++j) {
            if (j < ne) { dt[j] += f[q] * txt[j * De + d0 + q]; dn[j] += f[q] * neg[j * De + d0 + q]; }
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
            sc[j * max_n + m] = p.scale * __fdiv_rn(__fdiv_rn(a, fnorm), tnorm[j]);
            sn[j * max_n + m] = p.scale * __fdiv_rn(__fdiv_rn(c, fnorm), tnorm[kEG + j]);
          }
        }
      }
    }
    __syncthreads();

    // ---- selection tail: one warp per expression
    for (int j = warp; j < ne; j += nwarp) {
      const int e = eg + j;
      float* s = sc + j * max_n;
      float* q = sn + j * max_n;
      int* top = picks + j * 9;
      int* topn = top + 3;
      for (int i = lane; i < n; i += 32) p.score_clip[(size_t)e * max_n + i] = s[i];
      for (int i = n + lane; i < max_n; i += 32) p.score_clip[(size_t)e * max_n + i] = 0.f;
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
    }
    __syncthreads();
  }
}

}  // namespace hgl

extern "C" int hgl_score_select(const void* feat, int feat_dtype, const float* sent, const float* noun, const float* others,
                                const int32_t* other_off, const int64_t* boxes, const int32_t* relaflag, const float* score_gem,
                                const int32_t* mask_off, const int32_t* expr_off, int B, int M, int E, int De, int max_n,
                                double logit_scale_exp, double r, double alpha,
                                float* score_clip, int64_t* idx_hybrid, int64_t* idx_final, int32_t* top_idx, float* blended,
                                void* stream) {
  using namespace hgl;
  HGL_REQUIRE(feat && sent && noun && other_off && boxes && relaflag && score_clip && idx_hybrid && idx_final && top_idx && blended,
              "hgl_score_select: null pointer");
  HGL_REQUIRE(feat_dtype == HGL_F32 || feat_dtype == HGL_BF16, "hgl_score_select: feat_dtype %d", feat_dtype);
  HGL_REQUIRE(B >= 1 && M >= 0 && E >= 0 && max_n >= 1, "hgl_score_select: bad shape");
  HGL_REQUIRE(De >= 8 && De % 8 == 0, "hgl_score_select: De=%d must be a multiple of 8", De);
  HGL_REQUIRE((mask_off && expr_off) || B == 1, "hgl_score_select: mask_off/expr_off required when B > 1");
  HGL_REQUIRE((reinterpret_cast<uintptr_t>(feat) & 15) == 0, "hgl_score_select: feat must be 16-byte aligned");
  if (E == 0) return HGL_OK;
  ScoreParams p;
  p.feat = feat; p.feat_bf16 = (feat_dtype == HGL_BF16);
  p.sent = sent; p.noun = noun; p.others = others; p.other_off = other_off;
  p.boxes = boxes; p.relaflag = relaflag; p.score_gem = score_gem;
  p.mask_off = mask_off; p.expr_off = expr_off;
  p.B = B; p.M = M; p.E = E; p.De = De; p.max_n = max_n;
  p.scale = (float)logit_scale_exp; p.r = (float)r; p.one_minus_r = (float)(1.0 - r);
  p.alpha = (float)alpha; p.one_minus_alpha = (float)(1.0 - alpha);
  p.score_clip = score_clip; p.idx_hybrid = idx_hybrid; p.idx_final = idx_final; p.top_idx = top_idx; p.blended = blended;
  const size_t smem = ((size_t)2 * kEG * De + (size_t)2 * kEG * max_n + 2 * kEG) * 4 + (size_t)kEG * 9 * 4;
  HGL_REQUIRE(smem <= 200 * 1024, "hgl_score_select: De=%d max_n=%d needs %zu B of shared memory", De, max_n, smem);
  cudaError_t e = cudaFuncSetAttribute(score_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("hgl_score_select: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return HGL_ECUDA; }
  score_select_kernel<<<B, kScoreThreads, smem, (cudaStream_t)stream>>>(p);
  return launch_status("hgl_score_select");
}
