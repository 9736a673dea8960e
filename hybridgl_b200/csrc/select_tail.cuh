// Selection tail shared by the two scoring kernels (score_select.cu: features supplied; pool_score.cu: features pooled on the
// tensor cores in the same kernel): everything the reference does per expression AFTER the cosine scores exist.
//   argmax                              Hybridgl_main.py:168
//   softmax over masks, top-3 / top-6   Hybridgl_main.py:173-183
//   relation_boxes double loop, softmax Hybridgl_main.py:185-196, utils.py:240-268
//   blend with score_gem, argmax        Hybridgl_main.py:225-227
// A CTA serves up to 4 expressions at a time: two warps per expression (positive / negative side), see select_tail_block.
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

// torch.argmax / topk ordering: larger wins, NaN counts as the largest, lower index wins ties.  Scores are mapped to
// order-preserving 32-bit keys (NaN -> the largest key, -0 -> +0), so that a warp-wide selection is two redux.sync
// instructions (max of the keys, then min of the candidate indices) instead of a ten-shuffle compare tree.  Key 0 never
// occurs for a real value and means "nothing".
__device__ __forceinline__ uint32_t order_key(float v) {
  if (isnan(v)) return 0xffffffffu;
  const uint32_t u = __float_as_uint(v + 0.0f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_to_float(uint32_t k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }

// index of the best entry of x[0..n) that is not in skip[0..nskip) (-1 if there is none); whole warp
__device__ __forceinline__ int warp_select_best(const float* x, int n, const int* skip, int nskip, int lane) {
  uint32_t bk = 0u;
  int bi = 0x7fffffff;
  for (int i = lane; i < n; i += 32) {
    bool taken = false;
    for (int u = 0; u < nskip; ++u) taken |= (skip[u] == i);
    const uint32_t k = taken ? 0u : order_key(x[i]);
    if (k > bk) { bk = k; bi = i; }                      // ascending i: the lane keeps its lowest index among equals
  }
  const uint32_t kmax = __reduce_max_sync(0xffffffffu, bk);
  if (kmax == 0u) return -1;
  const int cand = (bk == kmax) ? bi : 0x7fffffff;
  return __reduce_min_sync(0xffffffffu, cand);
}

// soft-max over n values in shared memory (in place), one warp; torch.nn.Softmax(0) on [n,1]
__device__ __forceinline__ void warp_softmax(float* x, int n, int lane) {
  uint32_t mk = 0u;
  for (int i = lane; i < n; i += 32) mk = max(mk, order_key(x[i]));
  mk = __reduce_max_sync(0xffffffffu, mk);
  // a NaN anywhere makes every output NaN (max is NaN in torch); reproduce that through the arithmetic below
  const float mx = (mk == 0xffffffffu) ? __uint_as_float(0x7fc00000u) : (mk == 0u ? -INFINITY : key_to_float(mk));
  float s = 0.f;
  for (int i = lane; i < n; i += 32) { const float e = expf(x[i] - mx); x[i] = e; s += e; }
  s = warp_sum(s);
  for (int i = lane; i < n; i += 32) x[i] = __fdiv_rn(x[i], s);
  __syncwarp();
}

// indices of the k largest entries (descending), one warp; `out` in shared memory
__device__ __forceinline__ void warp_topk(const float* x, int n, int k, int* out, int lane) {
  for (int t = 0; t < k; ++t) {
    const int bi = warp_select_best(x, n, out, t, lane);
    if (lane == 0) out[t] = bi;
    __syncwarp();
  }
}

constexpr int kTailPicks = 12;     // ints of shared memory per expression: top3 | top6 | argmax | -, -
constexpr int kTailNPL = 4;        // register fast path: up to 4 scores per lane (n <= 128)

// ---- register-resident fast path (n <= 128): a lone warp pays ~13 cycles per issued instruction, so the tail is written as
//      straight-line code over <= 4 scores per lane instead of loops over shared memory ----------------------------------
// soft-max of the warp's n scores (lane holds x[lane + 32*q]); returns the probabilities in v[], writes them to xs[] too
__device__ __forceinline__ void reg_softmax(float (&v)[kTailNPL], float* xs, int n, int lane) {
  uint32_t mk = 0u;
#pragma unroll
  for (int q = 0; q < kTailNPL; ++q) if (lane + 32 * q < n) mk = max(mk, order_key(v[q]));
  mk = __reduce_max_sync(0xffffffffu, mk);
  const float mx = (mk == 0xffffffffu) ? __uint_as_float(0x7fc00000u) : (mk == 0u ? -INFINITY : key_to_float(mk));
  float s = 0.f;
#pragma unroll
  for (int q = 0; q < kTailNPL; ++q) {
    v[q] = (lane + 32 * q < n) ? expf(v[q] - mx) : 0.f;
    s += v[q];
  }
  s = warp_sum(s);
#pragma unroll
  for (int q = 0; q < kTailNPL; ++q) {
    v[q] = __fdiv_rn(v[q], s);
    if (lane + 32 * q < n) xs[lane + 32 * q] = v[q];
  }
}
// the k best of the warp's scores, descending (torch.topk order); out[] in shared memory, written by lane 0
__device__ __forceinline__ void reg_topk(const float (&v)[kTailNPL], int n, int k, int* out, int lane) {
  uint32_t kk[kTailNPL];
#pragma unroll
  for (int q = 0; q < kTailNPL; ++q) kk[q] = (lane + 32 * q < n) ? order_key(v[q]) : 0u;
  for (int t = 0; t < k; ++t) {
    uint32_t lm = kk[0];
#pragma unroll
    for (int q = 1; q < kTailNPL; ++q) lm = max(lm, kk[q]);
    const uint32_t km = __reduce_max_sync(0xffffffffu, lm);
    int cand = 0x7fffffff;
#pragma unroll
    for (int q = kTailNPL - 1; q >= 0; --q) if (kk[q] == km && km != 0u) cand = lane + 32 * q;     // lowest index of the lane last
    const int idx = __reduce_min_sync(0xffffffffu, cand);
#pragma unroll
    for (int q = 0; q < kTailNPL; ++q) if (idx == lane + 32 * q) kk[q] = 0u;
    if (lane == 0) out[t] = (km == 0u) ? -1 : idx;
  }
  __syncwarp();
}

// The selection tail of up to 4 expressions (e_first .. e_first + ne - 1) of one image, run by a WHOLE CTA of 8 warps:
//   phase 1  warp j      : argmax of the raw scores, soft-max, top-3              (positive side of expression j)
//            warp 4 + j  : soft-max and top-6 of the negative scores             (only when the expression has other nouns)
//   phase 2  warp j      : relation sums over the picked boxes, soft-max, blend with score_gem, final argmax, outputs
// sc: [8][stride] raw scores in shared memory, rows 0..3 positive / 4..7 negative (overwritten); picks: [4][kTailPicks] ints.
// Optional shared-memory copies the caller prefetched while it was busy elsewhere (null -> read from global memory here; the
// tail is a chain of dependent steps, so every global round trip it can skip is ~1-2 us):
//   box_s  the image's boxes [n][4];   sg_s  score_gem rows of the 4 expressions [4][stride];   meta_s  [4][2] = (n_other, relaflag)
// Contains one __syncthreads(): every thread of the CTA must call it.
__device__ __forceinline__ void select_tail_block(const TailArgs& t, int e_first, int ne, int n, int n_lo, float* sc, int stride, int* picks,
                                                  const int64_t* box_s, const float* sg_s, const int* meta_s, int warp, int lane) {
  const int j = warp & 3, side = warp >> 2;
  const int k1 = min(3, n), k2 = min(6, n);                       // :178-181
  const bool fast = n <= 32 * kTailNPL;
  if (j < ne && warp < 8) {
    const int e = e_first + j;
    int* pk = picks + j * kTailPicks;
    if (side == 0) {
      float* s = sc + j * stride;
      for (int i = n + lane; i < t.max_n; i += 32) t.score_clip[(size_t)e * t.max_n + i] = 0.f;
      if (fast) {
        float v[kTailNPL];
#pragma unroll
        for (int q = 0; q < kTailNPL; ++q) v[q] = (lane + 32 * q < n) ? s[lane + 32 * q] : 0.f;
        reg_topk(v, n, min(1, n), pk + 9, lane);                  // :168 argmax of the raw scores
        reg_softmax(v, s, n, lane);                               // :173
        reg_topk(v, n, k1, pk, lane);                             // :182
      } else {
        const int bi = warp_select_best(s, n, nullptr, 0, lane);
        if (lane == 0) pk[9] = bi;
        warp_softmax(s, n, lane);
        warp_topk(s, n, k1, pk, lane);
      }
    } else if ((meta_s ? meta_s[2 * j] : t.other_off[e + 1] - t.other_off[e]) > 0) {
      float* q_ = sc + (4 + j) * stride;
      if (fast) {
        float v[kTailNPL];
#pragma unroll
        for (int q = 0; q < kTailNPL; ++q) v[q] = (lane + 32 * q < n) ? q_[lane + 32 * q] : 0.f;
        reg_softmax(v, q_, n, lane);                              // :174
        reg_topk(v, n, k2, pk + 3, lane);                         // :183
      } else {
        warp_softmax(q_, n, lane);
        warp_topk(q_, n, k2, pk + 3, lane);
      }
    }
  }
  __syncthreads();
  if (side == 0 && j < ne && warp < 8) {
    const int e = e_first + j;
    const int* top = picks + j * kTailPicks;
    const int* topn = top + 3;
    const float* s = sc + j * stride;
    const float* q = sc + (4 + j) * stride;
    const int n_other = meta_s ? meta_s[2 * j] : t.other_off[e + 1] - t.other_off[e];
    // relation sums (:185-193), lanes 0..k1-1, sequential fp32 accumulation over j like the reference
    float T = 0.f;
    const int rel = meta_s ? meta_s[2 * j + 1] : t.relaflag[e];
    if (lane < k1) {
      const int ti = top[lane];
      const int64_t* bsrc = box_s ? box_s : t.boxes + (size_t)n_lo * 4;
      const int64_t* bi4 = bsrc + (size_t)ti * 4;
      const int cntj = (n_other == 0) ? k1 : k2;
      const float si = s[ti];
      if (rel == HGL_REL_NONE || rel > HGL_REL_WITHIN || rel < 0) {
        for (int u = 0; u < cntj; ++u) T = __fadd_rn(T, si);      // 'none' and unknown words: utils.py:267
      } else {
        for (int u = 0; u < cntj; ++u) {
          const int tj = (n_other == 0) ? top[u] : topn[u];
          const float sj = (n_other == 0) ? s[tj] : q[tj];
          T = __fadd_rn(T, relation(bi4, bsrc + (size_t)tj * 4, si, sj, rel));
        }
      }
    }
    // softmax over the k1 values (:196)
    float mx = (lane < k1) ? T : -INFINITY;
    mx = warp_max(mx);
    float ex = (lane < k1) ? expf(T - mx) : 0.f;
    const float sum = warp_sum(ex);
    float Tn = __fdiv_rn(ex, sum);
    if (t.score_gem != nullptr && lane < k1)                      // :225-226
      Tn = __fadd_rn(__fmul_rn(Tn, t.one_minus_alpha),
                     __fmul_rn(t.alpha, sg_s ? sg_s[j * stride + top[lane]] : t.score_gem[(size_t)e * t.max_n + top[lane]]));
    uint32_t fk = (lane < k1) ? order_key(Tn) : 0u;               // :227 argmax over the k1 blended values
    const uint32_t fmax = __reduce_max_sync(0xffffffffu, fk);
    const int fi = (fmax == 0u) ? -1 : __reduce_min_sync(0xffffffffu, (fk == fmax) ? lane : 0x7fffffff);
    if (lane < 3) {
      t.top_idx[(size_t)e * 3 + lane] = (lane < k1) ? top[lane] : -1;
      t.blended[(size_t)e * 3 + lane] = (lane < k1) ? Tn : 0.f;
    }
    if (lane == 0) {
      t.idx_hybrid[e] = (n > 0) ? top[9] : -1;
      t.idx_final[e] = (fi >= 0) ? top[fi] : -1;
    }
  }
}

// ---- thread-block cluster helpers ---------------------------------------------------------------------------------------
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Split form.  Every kernel that writes into a peer's shared memory arrives right after its own set-up and waits just before its
// first remote access: a peer CTA must have STARTED before its shared memory may be touched (compute-sanitizer racecheck: "block
// that might not have entered yet"), and by then every CTA has long arrived, so the wait costs nothing.
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
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
