// Selection tail shared by the two scoring kernels (score_select.cu: features supplied; pool_score.cu: features pooled on the
// tensor cores in the same kernel): everything the reference does per expression AFTER the cosine scores exist.
//   argmax                              Hybridgl_main.py:168
//   softmax over masks, top-3 / top-6   Hybridgl_main.py:173-183
//   relation_boxes double loop, softmax Hybridgl_main.py:185-196, utils.py:240-268
//   blend with score_gem, argmax        Hybridgl_main.py:225-227
// One warp per expression; `s` / `q` are warp-private shared-memory rows holding the raw positive / negative scores.
#pragma once
#include "hgl_common.cuh"

namespace hgl {

struct TailArgs {
  const int64_t* boxes;        // [M,4] XYWH
  const int32_t* relaflag;     // [E]
  const int32_t* other_off;    // [E+1]
  const float* score_gem;      // [E,max_n] or null
  float alpha, one_minus_alpha;
  int max_n;
  float* score_clip;           // [E,max_n]: the tail zero-fills columns n..max_n-1
  int64_t* idx_hybrid; int64_t* idx_final; int32_t* top_idx; float* blended;
};

// relation_boxes utils.py:240-268 (boxes XYWH int64; torch promotes to float32 for the divisions)
__device__ __forceinline__ float relation(const int64_t* bi, const int64_t* bj, float si, float sj, int rel) {
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
__device__ __forceinline__ void warp_softmax(float* x, int n, int lane) {
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
__device__ __forceinline__ void warp_topk(const float* x, int n, int k, int* out, int lane) {
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

// One warp, one expression e of an image whose n masks are rows n_lo.. of the [M,...] tensors.  s[0..n) / q[0..n): raw
// score_clip / score_clip_Neg of the expression (overwritten); picks: 9 ints of warp-private shared memory.
__device__ __forceinline__ void select_tail_warp(const TailArgs& t, int e, int n, int n_lo, float* s, float* q, int* picks, int lane) {
  int* top = picks;
  int* topn = picks + 3;
  for (int i = n + lane; i < t.max_n; i += 32) t.score_clip[(size_t)e * t.max_n + i] = 0.f;
  float bv = 0.f; int bi = -1;                                   // :168 argmax
  for (int i = lane; i < n; i += 32) if (better(s[i], i, bv, bi)) { bv = s[i]; bi = i; }
  warp_argbest(bv, bi);
  const int n_other = t.other_off[e + 1] - t.other_off[e];
  warp_softmax(s, n, lane);                                       // :173
  const int k1 = min(3, n), k2 = min(6, n);                       // :178-181
  warp_topk(s, n, k1, top, lane);                                 // :182
  if (n_other > 0) { warp_softmax(q, n, lane); warp_topk(q, n, k2, topn, lane); }   // :174,:183
  __syncwarp();
  // relation sums (:185-193), lanes 0..k1-1, sequential fp32 accumulation over j like the reference
  float T = 0.f;
  const int rel = t.relaflag[e];
  if (lane < k1) {
    const int ti = top[lane];
    const int64_t* bi4 = t.boxes + (size_t)(n_lo + ti) * 4;
    const int cntj = (n_other == 0) ? k1 : k2;
    for (int u = 0; u < cntj; ++u) {
      const int tj = (n_other == 0) ? top[u] : topn[u];
      const float sj = (n_other == 0) ? s[tj] : q[tj];
      T = __fadd_rn(T, relation(bi4, t.boxes + (size_t)(n_lo + tj) * 4, s[ti], sj, rel));
    }
  }
  // softmax over the k1 values (:196)
  float mx = (lane < k1) ? T : -INFINITY;
  mx = warp_max(mx);
  float ex = (lane < k1) ? expf(T - mx) : 0.f;
  const float sum = warp_sum(ex);
  float Tn = __fdiv_rn(ex, sum);
  if (t.score_gem != nullptr && lane < k1)                        // :225-226
    Tn = __fadd_rn(__fmul_rn(Tn, t.one_minus_alpha), __fmul_rn(t.alpha, t.score_gem[(size_t)e * t.max_n + top[lane]]));
  float fv = Tn; int fi = (lane < k1) ? lane : -1;                // :227
  warp_argbest(fv, fi);
  if (lane < 3) {
    t.top_idx[(size_t)e * 3 + lane] = (lane < k1) ? top[lane] : -1;
    t.blended[(size_t)e * 3 + lane] = (lane < k1) ? Tn : 0.f;
  }
  if (lane == 0) {
    t.idx_hybrid[e] = bi;
    t.idx_final[e] = (fi >= 0) ? top[fi] : -1;
  }
  __syncwarp();
}

// ---- thread-block cluster helpers ---------------------------------------------------------------------------------------
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t map_peer(const void* local_smem_ptr, uint32_t cta_rank) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(local_smem_ptr)), "r"(cta_rank));
  return raddr;
}
__device__ __forceinline__ float ld_peer_f32(const float* local_smem_ptr, uint32_t cta_rank) {
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(map_peer(local_smem_ptr, cta_rank)) : "memory");
  return v;
}
__device__ __forceinline__ void st_peer_f32(float* local_smem_ptr, uint32_t cta_rank, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(map_peer(local_smem_ptr, cta_rank)), "f"(v) : "memory");
}

}  // namespace hgl
