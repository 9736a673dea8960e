// (f3) GEM heat-map pooling in TOKEN SPACE -- Hybridgl_main.py:200-223 without the H x W heat-map.
//
// The reference resizes the raw GEM map h [hh, hw] (28 x 37: one value per GEM patch token, h = F^ . t) to the frame,
//   A = T.Resize((H,W), antialias=True)(h) = Uy h Ux^T,
// conditions it (A' = minmax(A) * ramp, A'' = A' / mean(A')) and pools it inside / outside every mask.  Because the resize is
// linear,
//   S_in[n] = sum_p m_n(p) A''(p) = kk * ( G_n . h  -  mn * sum(G_n) ),     G_n[i,j] = sum_{y,x} m_n(y,x) ramp(x) Uy[y,i] Ux[x,j]
// i.e. a masks x tokens contraction against the mask resampled onto the TOKEN grid with the adjoint of the up-sampler (ramp
// folded into the horizontal weights); the affine terms of the conditioning enter as per-expression scalars:
//   mn = min_p A(p),   kk = H*W / (C . h - mn * sum(C)),  C = G of the all-ones mask,   S_tot = sum_p A'' = H*W.
// (With h = F^ t this is (M~ F^) . t -- the same contraction hgl_mask_pool runs on the tensor cores for the CLIP tokens; here K =
// hh*hw and the right-hand side is <= a few vectors per image, so it stays on the CUDA cores.)
// Nothing frame-sized is written: no [E,H,W+1] prefix tables (59 MB at the bench shape), no gathers from them.
//
// Kernels (all tiny except the mask pass):
//   gem_tables_kernel   adjoint tap tables: for token column j, prefix sums over its <= ~2W/hw fine columns of ramp_k(x) Ux[x,j]
//                       (4 ramp kinds), for token row i the weights Uy[y,i]; per-image set of ramp kinds; resets of scratch
//   gem_minmax_kernel   min / max of A over the frame (A evaluated on the fly in ATen's tap order, never stored)
//   gem_consts_kernel   mn, kk per expression
//   gem_rows_kernel     the pass over the packed masks, run based like mask_rows_kernel: a warp owns a band of rows of one
//                       mask, a lane a bit row; a run [s,e) adds table differences to the token columns it touches, the
//                       rows of a 32-row block are folded into the <= few token rows they feed, and the band's G is dotted
//                       with the image's raw maps.  The last band of a mask adds the bands up in order.
#include <algorithm>

#include "hgl_common.cuh"
#include "resample.cuh"

namespace hgl {

constexpr int kGtThreads = 256;
constexpr int kGtBands = 8;
constexpr int kGtKinds = 4;          // ramps: 0 = ones (none / up / down), 1 = left, 2 = right, 3 = middle
constexpr int kGtMaxRows = 12;       // token rows a band of frame rows can touch

__device__ __forceinline__ int ramp_kind(int dirflag) {
  return dirflag == HGL_DIR_LEFT ? 1 : dirflag == HGL_DIR_RIGHT ? 2 : dirflag == HGL_DIR_MIDDLE ? 3 : 0;
}
__device__ __forceinline__ int kind_dirflag(int kind) {
  return kind == 1 ? HGL_DIR_LEFT : kind == 2 ? HGL_DIR_RIGHT : kind == 3 ? HGL_DIR_MIDDLE : HGL_DIR_NONE;
}
__device__ __forceinline__ int f2ord(float f) { const int b = __float_as_int(f); return b >= 0 ? b : b ^ 0x7fffffff; }
__device__ __forceinline__ float ord2f(int o) { return __int_as_float(o >= 0 ? o : o ^ 0x7fffffff); }

struct GemWs {
  float* px;        // [kGtKinds][hw][KX + 1]  prefix sums of ramp_k(x) * Ux[x, j] over x in [lox[j], hix[j])
  int* lox; int* hix;     // [hw]
  float* wy;        // [hh][KY]  Uy[y, i] for y in [loy[i], hiy[i])
  int* loy; int* hiy;     // [hh]
  float* consts;    // [E][4]  mn, kk, -, -
  int* mm;          // [E][2]  ordered-int min / max of the resized map
  int* kinds;       // [B]     bit k set: some expression of the image uses ramp kind k
  int* tickets;     // [M + 1] zero at launch; entry M: task counter
  int* pcnt;        // [M][kGtBands]
  float* pheat;     // [E][max_n][kGtBands][2]  (G . h, sum G) per band
  int KX, KY;
  size_t bytes;
};

static GemWs gem_carve(void* ws, int B, int M, int E, int H, int W, int hh, int hw, int max_n) {
  GemWs g;
  g.KX = 2 * ((W + hw - 1) / hw) + 6;
  g.KY = 2 * ((H + hh - 1) / hh) + 6;
  size_t off = 0;
  auto take = [&](size_t n) { size_t o = off; off += (n + 255) & ~size_t(255); return o; };
  uint8_t* base = reinterpret_cast<uint8_t*>(ws);
  g.px = reinterpret_cast<float*>(base + take((size_t)kGtKinds * hw * (g.KX + 1) * 4));
  g.lox = reinterpret_cast<int*>(base + take((size_t)hw * 4)); g.hix = reinterpret_cast<int*>(base + take((size_t)hw * 4));
  g.wy = reinterpret_cast<float*>(base + take((size_t)hh * g.KY * 4));
  g.loy = reinterpret_cast<int*>(base + take((size_t)hh * 4)); g.hiy = reinterpret_cast<int*>(base + take((size_t)hh * 4));
  g.consts = reinterpret_cast<float*>(base + take((size_t)E * 4 * 4));
  g.mm = reinterpret_cast<int*>(base + take((size_t)E * 2 * 4));
  g.kinds = reinterpret_cast<int*>(base + take((size_t)B * 4));
  g.tickets = reinterpret_cast<int*>(base + take((size_t)(M + 1) * 4));
  g.pcnt = reinterpret_cast<int*>(base + take((size_t)M * kGtBands * 4));
  g.pheat = reinterpret_cast<float*>(base + take((size_t)E * max_n * kGtBands * 2 * 4));
  g.bytes = off;
  return g;
}

// CTA c < hw: token column c;  hw <= c < hw + hh: token row c - hw;  c == hw + hh: resets + per-image ramp kinds.
// A thread evaluates the up-sampler's taps of ONE fine coordinate near the token (ATen's arithmetic, aa_fill) and keeps the
// weight that falls on the token; thread 0 turns them into prefix sums (double accumulation: tiny edge taps survive).
__global__ void __launch_bounds__(128) gem_tables_kernel(GemWs g, int B, int M, int E, int H, int W, int hh, int hw,
                                                         const int32_t* __restrict__ expr_off, const int32_t* __restrict__ dirflag) {
  __shared__ float wgt[128];
  __shared__ int s_lo, s_hi;
  const int c = blockIdx.x, tid = threadIdx.x;
  if (c == hw + hh) {
    for (int e = tid; e < E; e += blockDim.x) { g.mm[2 * e] = 0x7fffffff; g.mm[2 * e + 1] = (int)0x80000000; }
    for (int m = tid; m <= M; m += blockDim.x) g.tickets[m] = 0;
    for (int b = tid; b < B; b += blockDim.x) {
      const int e0 = expr_off ? expr_off[b] : 0, e1 = expr_off ? expr_off[b + 1] : E;
      int k = 0;
      for (int e = e0; e < e1; ++e) k |= 1 << ramp_kind(dirflag[e]);
      g.kinds[b] = k;
    }
    return;
  }
  const bool col = c < hw;
  const int t_idx = col ? c : c - hw;                  // token column / row
  const int fine = col ? W : H, coarse = col ? hw : hh;
  const int K = col ? g.KX : g.KY;
  // fine coordinates that can touch the token: centre (t + 0.5) * fine / coarse, support one coarse cell to each side (+ slack)
  const float inv = (float)fine / (float)coarse;
  const int start = max(0, (int)floorf(((float)t_idx - 1.0f) * inv) - 2);
  const int xx = start + tid;
  float wv = 0.f;
  if (tid < K + 8 && xx < fine) {
    int xm, xs;
    float w3[3];
    aa_fill(xx, coarse, fine, 3, &xm, &xs, w3);
    if (t_idx >= xm && t_idx < xm + xs) wv = w3[t_idx - xm];
  }
  wgt[tid] = wv;
  __syncthreads();
  if (tid == 0) {
    int lo = -1, hi = -1;
    for (int t = 0; t < 128; ++t)
      if (wgt[t] != 0.f) { if (lo < 0) lo = t; hi = t + 1; }
    if (lo < 0) { lo = 0; hi = 0; }
    hi = min(hi, lo + K);                             // (never binding: K bounds the support with slack)
    s_lo = lo; s_hi = hi;
    if (col) { g.lox[t_idx] = start + lo; g.hix[t_idx] = start + hi; }
    else { g.loy[t_idx] = start + lo; g.hiy[t_idx] = start + hi; }
  }
  __syncthreads();
  const int lo = s_lo, hi = s_hi;
  if (col) {
    if (tid < kGtKinds) {                              // one thread per ramp kind: prefix over the column's support
      float* dst = g.px + ((size_t)tid * hw + t_idx) * (g.KX + 1);
      double run = 0.0;
      dst[0] = 0.f;
      for (int t = lo; t < hi; ++t) {
        run += (double)wgt[t] * (double)ramp_at(kind_dirflag(tid), start + t, W);
        dst[t - lo + 1] = (float)run;
      }
      for (int t = hi - lo + 1; t <= g.KX; ++t) dst[t] = (float)run;
    }
  } else {
    for (int t = tid; t < g.KY; t += blockDim.x) g.wy[(size_t)t_idx * g.KY + t] = (lo + t < hi) ? wgt[lo + t] : 0.f;
  }
}

// min / max of the resized map A = Uy h Ux^T over the frame (Hybridgl_main.py:201, 204): evaluated exactly as ATen does
// (horizontal pass of the raw rows behind the CTA's 32 frame rows into shared memory, then the vertical taps), never stored.
constexpr int kMmRows = 32;
__global__ void __launch_bounds__(256) gem_minmax_kernel(const float* __restrict__ heat, int H, int W, int hh, int hw, int nr_max, int* __restrict__ mm) {
  extern __shared__ float mm_sm[];                     // hrow [nr_max][W]
  const int e = blockIdx.y, tid = threadIdx.x;
  const int y_first = blockIdx.x * kMmRows, y_last = min(H, y_first + kMmRows) - 1;
  int ry0, ys, ym;
  float wtmp[3];
  aa_fill(y_first, hh, H, 3, &ry0, &ys, wtmp);
  aa_fill(y_last, hh, H, 3, &ym, &ys, wtmp);
  const int nr = min(ym + ys - ry0, nr_max);
  for (int x = tid; x < W; x += blockDim.x) {
    int xm, xs;
    float wx[3];
    aa_fill(x, hw, W, 3, &xm, &xs, wx);
    const float* A0 = heat + ((size_t)e * hh + ry0) * hw + xm;
    for (int r = 0; r < nr; ++r) {
      float t = 0.f;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx)
        if (kx < xs) t = __fadd_rn(t, __fmul_rn(__ldg(A0 + (size_t)r * hw + kx), wx[kx]));
      mm_sm[r * W + x] = t;
    }
  }
  __syncthreads();
  float lo = INFINITY, hi = -INFINITY;
  for (int y = y_first + (tid >> 5); y <= y_last; y += 8) {
    int ymin, ysize;
    float wy[3];
    aa_fill(y, hh, H, 3, &ymin, &ysize, wy);
    const float* hr = mm_sm + (ymin - ry0) * W;
    for (int x = tid & 31; x < W; x += 32) {
      float a = 0.f;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
        if (ky < ysize) a = __fadd_rn(a, __fmul_rn(hr[ky * W + x], wy[ky]));
      lo = fminf(lo, a); hi = fmaxf(hi, a);
    }
  }
  lo = warp_min(lo); hi = warp_max(hi);
  if ((tid & 31) == 0) { atomicMin(mm + 2 * e, f2ord(lo)); atomicMax(mm + 2 * e + 1, f2ord(hi)); }
}

// per expression: mn = min A, kk = H*W / (C . h - mn * sum C) with C[i,j] = (sum_y Uy[y,i]) * (sum_x ramp(x) Ux[x,j])
__global__ void __launch_bounds__(128) gem_consts_kernel(GemWs g, const float* __restrict__ heat, const int32_t* __restrict__ dirflag, int E, int H,
                                                         int W, int hh, int hw) {
  const int e = blockIdx.x, tid = threadIdx.x;
  const int kind = ramp_kind(dirflag[e]);
  double ch = 0.0, cs = 0.0;
  for (int t = tid; t < hh * hw; t += blockDim.x) {
    const int i = t / hw, j = t - i * hw;
    double cy = 0.0;
    for (int k = 0; k < g.KY; ++k) cy += (double)g.wy[(size_t)i * g.KY + k];
    const double cx = (double)g.px[((size_t)kind * hw + j) * (g.KX + 1) + g.KX];
    ch += cy * cx * (double)heat[(size_t)e * hh * hw + t];
    cs += cy * cx;
  }
  __shared__ double red[2][4];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { ch += __shfl_xor_sync(0xffffffffu, ch, o); cs += __shfl_xor_sync(0xffffffffu, cs, o); }
  if ((tid & 31) == 0) { red[0][tid >> 5] = ch; red[1][tid >> 5] = cs; }
  __syncthreads();
  if (tid == 0) {
    ch = red[0][0] + red[0][1] + red[0][2] + red[0][3];
    cs = red[1][0] + red[1][1] + red[1][2] + red[1][3];
    const float mn = ord2f(g.mm[2 * e]);
    float* o = g.consts + (size_t)e * 4;
    o[0] = mn;
    o[1] = (float)((double)H * (double)W / (ch - (double)mn * cs));
    o[2] = ord2f(g.mm[2 * e + 1]);
    o[3] = 0.f;
  }
}

struct GemRowsParams {
  const uint32_t* bits; int M, H, W, WW;
  const int32_t* mask_off; const int32_t* expr_off; const int32_t* dirflag; const float* black;
  const float* heat; int hh, hw; int B, E, max_n;
  float* score_gem;
  GemWs g;
  int gpitch;                    // hw rounded up to odd (bank-friendly row pitch of the per-lane row buffers)
};

__global__ void __launch_bounds__(kGtThreads, 2) gem_rows_kernel(const GemRowsParams p) {
  extern __shared__ __align__(16) uint8_t gt_sm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int H = p.H, W = p.W, WW = p.WW, hw = p.hw, hh = p.hh, KX = p.g.KX, KY = p.g.KY, gp = p.gpitch;
  // shared: px tables [4][hw][KX+1] | lox, hix [hw] | wy [hh][KY] | loy, hiy [hh] | per warp: hbuf [32][gp], pg [kGtMaxRows][hw]
  float* px = reinterpret_cast<float*>(gt_sm);
  int* lox = reinterpret_cast<int*>(px + (size_t)kGtKinds * hw * (KX + 1));
  int* hix = lox + hw;
  float* wy = reinterpret_cast<float*>(hix + hw);
  int* loy = reinterpret_cast<int*>(wy + (size_t)hh * KY);
  int* hiy = loy + hh;
  float* warp_base = reinterpret_cast<float*>(hiy + hh) + (size_t)warp * (32 * gp + kGtMaxRows * hw);
  float* hbuf = warp_base;
  float* pg = warp_base + 32 * gp;
  for (int t = tid; t < kGtKinds * hw * (KX + 1); t += kGtThreads) px[t] = p.g.px[t];
  for (int t = tid; t < hw; t += kGtThreads) { lox[t] = p.g.lox[t]; hix[t] = p.g.hix[t]; }
  for (int t = tid; t < hh * KY; t += kGtThreads) wy[t] = p.g.wy[t];
  for (int t = tid; t < hh; t += kGtThreads) { loy[t] = p.g.loy[t]; hiy[t] = p.g.hiy[t]; }
  __syncthreads();

  const int band_rows = (H + kGtBands - 1) / kGtBands;
  const size_t mask_words = (size_t)H * WW;
  const float inv_sx = (float)hw / (float)W;
  const float hwf = (float)((size_t)H * W);
  const int total_tasks = p.M * kGtBands;
  for (;;) {
    int task = 0;
    if (lane == 0) task = atomicAdd(p.g.tickets + p.M, 1);
    task = __shfl_sync(0xffffffffu, task, 0);
    if (task >= total_tasks) break;
    const int m = task / kGtBands, band = task - m * kGtBands;
    const int y_begin = band * band_rows, y_end = min(H, y_begin + band_rows);
    int b = 0, n_lo = 0;
    if (p.mask_off) {
      int lo = 0, hi = p.B - 1;
      while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (p.mask_off[mid] <= m) lo = mid; else hi = mid - 1; }
      b = lo; n_lo = p.mask_off[b];
    }
    const int e_lo = p.expr_off ? p.expr_off[b] : 0, e_hi = p.expr_off ? p.expr_off[b + 1] : p.E;
    const int kinds = p.g.kinds[b];
    const bool in_row = (m - n_lo) < p.max_n;
    // token rows this band can touch
    int i_lo = 0;
    while (i_lo < hh && hiy[i_lo] <= y_begin) ++i_lo;
    int i_hi = i_lo;
    while (i_hi + 1 < hh && loy[i_hi + 1] < y_end) ++i_hi;
    const int ni = min(i_hi - i_lo + 1, kGtMaxRows);
    const uint32_t* mb = p.bits + (size_t)m * mask_words;
    int cnt = 0;
    bool first = true;
    for (int kind = 0; kind < kGtKinds; ++kind) {
      if (!((kinds >> kind) & 1) && !(first && kind == kGtKinds - 1)) continue;   // (an image without expressions still counts its pixels)
      const bool pool = (kinds >> kind) & 1;
      const float* pk = px + (size_t)kind * hw * (KX + 1);
      for (int t = lane; t < ni * hw; t += 32) pg[t] = 0.f;
      for (int y0 = y_begin; y0 < y_end; y0 += 32) {
        const int y = y0 + lane;
        const bool have = y < y_end;
        const uint32_t* rowp = mb + (size_t)(have ? y : y_begin) * WW;
        float* hr = hbuf + lane * gp;
        int c_row = 0, bx_lo = hw, bx_hi = -1;
        if (have) {
          uint32_t carry = 0;
          int xs = 0;
          bool zeroed = false;
          for (int w = 0; w <= WW; ++w) {
            const uint32_t v = (w < WW) ? __ldg(rowp + w) : 0u;
            if (w < WW && v == (0u - carry)) { c_row += __popc(v); continue; }      // all outside after a 0 / all inside after a 1
            c_row += __popc(v);
            uint32_t t = v ^ ((v << 1) | carry);
            carry = v >> 31;
            while (t) {
              const int bp = __ffs(t) - 1;
              t &= t - 1;
              const int x = 32 * w + bp;
              if ((v >> bp) & 1u) { xs = x; continue; }
              const int s0 = xs, e0 = min(x, W);                      // pixels [s0, e0) of row y are inside the mask
              if (!pool) continue;
              if (!zeroed) { for (int j = 0; j < hw; ++j) hr[j] = 0.f; zeroed = true; }
              int j = max(0, min(hw - 1, (int)((float)s0 * inv_sx - 1.5f) - 1));
              while (j < hw && hix[j] <= s0) ++j;
              bx_lo = min(bx_lo, j);
              for (; j < hw && lox[j] < e0; ++j) {
                const float* ps = pk + (size_t)j * (KX + 1) - lox[j];
                const int a0 = min(max(s0, lox[j]), hix[j]), a1 = min(max(e0, lox[j]), hix[j]);
                hr[j] += ps[a1] - ps[a0];
              }
              bx_hi = max(bx_hi, j - 1);
            }
          }
        }
        if (first) cnt += c_row;
        if (!pool) continue;
        // fold the block's rows into the token rows they feed (ascending frame row)
        const bool nonempty = bx_hi >= 0;
        const uint32_t rows_mask = __ballot_sync(0xffffffffu, nonempty);
        if (rows_mask != 0u) {
          int gx0 = bx_lo, gx1 = bx_hi;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            gx0 = min(gx0, __shfl_xor_sync(0xffffffffu, gx0, o));
            gx1 = max(gx1, __shfl_xor_sync(0xffffffffu, gx1, o));
          }
          const int nbx = gx1 - gx0 + 1;
          // lanes whose row holds no run in [gx0, gx1] at some column still read zeros: clear the columns they did not touch
          if (nonempty) { for (int j = gx0; j < bx_lo; ++j) hr[j] = 0.f; for (int j = bx_hi + 1; j <= gx1; ++j) hr[j] = 0.f; }
          __syncwarp();
          for (int t = lane; t < ni * nbx; t += 32) {
            const int il = t / nbx, j = gx0 + t - il * nbx, i = i_lo + il;
            const int k_lo = max(0, loy[i] - y0), k_hi = min(31, hiy[i] - 1 - y0);
            const float* wp = wy + (size_t)i * KY + (y0 - loy[i]);
            float a = pg[il * hw + j];
            for (int r = k_lo; r <= k_hi; ++r)
              if ((rows_mask >> r) & 1u) a = fmaf(hbuf[r * gp + j], wp[r], a);
            pg[il * hw + j] = a;
          }
          __syncwarp();
        }
      }
      if (pool) {
        // this band's share of G . h for every expression of the image that uses this ramp, and of sum(G)
        float sr = 0.f;
        for (int t = lane; t < ni * hw; t += 32) sr += pg[t];
        sr = warp_sum(sr);
        for (int e = e_lo; e < e_hi; ++e) {
          if (ramp_kind(p.dirflag[e]) != kind) continue;
          const float* he = p.heat + ((size_t)e * hh + i_lo) * hw;
          float sc = 0.f;
          for (int t = lane; t < ni * hw; t += 32) sc = fmaf(pg[t], __ldg(he + t), sc);
          sc = warp_sum(sc);
          if (lane == 0 && in_row) {
            float* o = p.g.pheat + (((size_t)e * p.max_n + (m - n_lo)) * kGtBands + band) * 2;
            o[0] = sc; o[1] = sr;
          }
        }
      }
      first = false;
    }
    cnt = warp_sum_i(cnt);
    if (lane == 0) p.g.pcnt[m * kGtBands + band] = cnt;
    // ---- ticket: the last band of mask m adds the bands up in band order
    __threadfence();
    __syncwarp();
    int last = 0;
    if (lane == 0) last = (atomicAdd(p.g.tickets + m, 1) == kGtBands - 1);
    last = __shfl_sync(0xffffffffu, last, 0);
    if (!last || !in_row) continue;
    __threadfence();
    int area_m = 0;
#pragma unroll
    for (int q = 0; q < kGtBands; ++q) area_m += __ldcg(p.g.pcnt + m * kGtBands + q);
    for (int e = e_lo + lane; e < e_hi; e += 32) {
      const float* o = p.g.pheat + ((size_t)e * p.max_n + (m - n_lo)) * kGtBands * 2;
      float sc = 0.f, sr = 0.f;
#pragma unroll
      for (int q = 0; q < kGtBands; ++q) { sc += __ldcg(o + 2 * q); sr += __ldcg(o + 2 * q + 1); }
      const float mn = p.g.consts[e * 4], kk = p.g.consts[e * 4 + 1];
      const float s_in = kk * (sc - mn * sr);
      const float bl = p.black[e];
      p.score_gem[(size_t)e * p.max_n + (m - n_lo)] = (2.f - bl) * s_in / (float)area_m - bl * (hwf - s_in) / (hwf - (float)area_m);
    }
  }
}

}  // namespace hgl

extern "C" int64_t hgl_gem_token_workspace_bytes(int B, int M, int E, int H, int W, int hh, int hw, int max_n) {
  if (B < 1 || M < 0 || E < 0 || H < 1 || W < 1 || hh < 1 || hw < 1 || max_n < 1) return -1;
  return (int64_t)hgl::gem_carve(nullptr, B, M, E, H, W, hh, hw, max_n).bytes + 256;
}

extern "C" int hgl_gem_token_pool(const uint32_t* bits, const int32_t* mask_off, int B, int M, int H, int W,
                                  const float* heat_raw, int hh, int hw, const int32_t* expr_off, const int32_t* dirflag, const float* black, int E,
                                  int max_n, float* score_gem, void* workspace, void* stream) {
  using namespace hgl;
  HGL_REQUIRE(B >= 1 && M >= 0 && E >= 0 && H >= 1 && W >= 1 && max_n >= 1, "hgl_gem_token_pool: bad shape");
  if (E == 0 || M == 0) return HGL_OK;
  HGL_REQUIRE(bits && heat_raw && dirflag && black && score_gem && workspace, "hgl_gem_token_pool: null pointer");
  HGL_REQUIRE((mask_off && expr_off) || B == 1, "hgl_gem_token_pool: mask_off / expr_off required when B > 1");
  HGL_REQUIRE(hh >= 1 && hw >= 1 && hh <= H && hw <= W && hw <= 128 && hh <= 128, "hgl_gem_token_pool: the raw map %dx%d must be coarser than the frame", hh, hw);
  HGL_REQUIRE(E <= 65535 && (W + 31) / 32 <= 63, "hgl_gem_token_pool: too many expressions / frame too wide");
  cudaStream_t st = (cudaStream_t)stream;
  GemWs g = gem_carve(reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255)), B, M, E, H, W, hh, hw, max_n);
  HGL_REQUIRE(g.KX + 8 <= 128 && g.KY + 8 <= 128, "hgl_gem_token_pool: frame / raw map ratio too large (%d, %d taps)", g.KX, g.KY);
  const int band_rows = ceil_div(H, kGtBands);
  HGL_REQUIRE(band_rows / (H / hh) + 4 <= kGtMaxRows, "hgl_gem_token_pool: a band of %d rows spans too many token rows", band_rows);
  cudaError_t e = cudaMemsetAsync(score_gem, 0, (size_t)E * max_n * 4, st);      // rows of images with fewer than max_n masks
  if (e != cudaSuccess) { set_error("hgl_gem_token_pool: cudaMemsetAsync: %s", cudaGetErrorString(e)); return HGL_ECUDA; }
  gem_tables_kernel<<<hw + hh + 1, 128, 0, st>>>(g, B, M, E, H, W, hh, hw, expr_off, dirflag);
  int rc = launch_status("hgl_gem_token_pool(tables)");
  if (rc != HGL_OK) return rc;
  {
    const int nr_max = (int)((double)kMmRows * hh / H) + 4;
    const size_t smem = (size_t)nr_max * W * 4;
    rc = ensure_dyn_smem(reinterpret_cast<const void*>(gem_minmax_kernel), smem, "hgl_gem_token_pool(minmax)");
    if (rc != HGL_OK) return rc;
    gem_minmax_kernel<<<dim3(ceil_div(H, kMmRows), E), 256, smem, st>>>(heat_raw, H, W, hh, hw, nr_max, g.mm);
    rc = launch_status("hgl_gem_token_pool(minmax)");
    if (rc != HGL_OK) return rc;
  }
  gem_consts_kernel<<<E, 128, 0, st>>>(g, heat_raw, dirflag, E, H, W, hh, hw);
  rc = launch_status("hgl_gem_token_pool(consts)");
  if (rc != HGL_OK) return rc;
  GemRowsParams p;
  p.bits = bits; p.M = M; p.H = H; p.W = W; p.WW = (W + 31) >> 5;
  p.mask_off = mask_off; p.expr_off = expr_off; p.dirflag = dirflag; p.black = black;
  p.heat = heat_raw; p.hh = hh; p.hw = hw; p.B = B; p.E = E; p.max_n = max_n; p.score_gem = score_gem; p.g = g;
  p.gpitch = hw | 1;
  const int warps = kGtThreads / 32;
  const size_t smem = ((size_t)kGtKinds * hw * (g.KX + 1) + 2 * hw + (size_t)hh * g.KY + 2 * hh) * 4 +
                      (size_t)warps * (32 * p.gpitch + kGtMaxRows * hw) * 4;
  HGL_REQUIRE(smem <= 110 * 1024, "hgl_gem_token_pool: raw map %dx%d needs %zu B of shared memory", hh, hw, smem);
  rc = ensure_dyn_smem(reinterpret_cast<const void*>(gem_rows_kernel), smem, "hgl_gem_token_pool");
  if (rc != HGL_OK) return rc;
  const int ctas = std::min(ceil_div(M * kGtBands, warps), sm_count() * 2);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(kGtThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1] = {priority_attr(st)};
  cfg.attrs = attr; cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, gem_rows_kernel, p);
  if (e != cudaSuccess) { set_error("hgl_gem_token_pool: cudaLaunchKernelEx: %s", cudaGetErrorString(e)); return HGL_ECUDA; }
  return launch_status("hgl_gem_token_pool");
}
