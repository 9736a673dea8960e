// (a2) mask -> patch grid (antialiased)  and  (a10)+(a11) heat-map conditioning + per-mask pooling, in ONE pass over the
// packed masks.
//
//   (a2)   model/backbone.py:160   TF.resize(pred_masks.float(), (g,g)) under torchvision >= 0.17: ATen
//          _upsample_bilinear2d_aa = separable triangle filter, horizontal pass then vertical pass.
//   (a10)  Hybridgl_main.py:204-209, utils.py:135-161   A' = minmax(A) * ramp(dirflag);  A'' = A' / mean(A')
//   (a11)  Hybridgl_main.py:211-223   score_gem[n] = (2-black) * sum(A''*m_n)/|m_n|  -  black * sum(A''*(1-m_n)) / |1-m_n|
// The reference loops over masks in Python (~8 full-frame kernels + a D2H sync per mask) and resamples every mask pixel.
//
// B200 design -- everything is expressed on the RUNS of a packed bit row (SAM proposals are blobs: one or two runs per
// row), so a row costs O(#runs) instead of O(W):
//   * horizontal filter sum of bin gx over a run [s,e)   =  P_gx(e) - P_gx(s)      P_gx = prefix sum of the filter taps
//   * heat-map sum over the same run                     =  C_e[y][e] - C_e[y][s]  C_e  = row prefix sum of A*ramp
//   * area                                               =  popcount
// heat_prefix_kernel builds C_e (and the row min / max / sum that give min-max and mean) in one coalesced pass over the
// heat-maps; heat_consts_kernel folds them into (min, 1/(range*mean), sum A'') per expression; mask_rows_kernel then
// streams the packed masks: a WARP owns a band of rows of one mask and one LANE owns a bit row (16-byte loads straight
// from global memory, no staging, no CTA barrier in the hot loop), 32 rows at a time.  While the words of a row are in
// registers the lane only records which words contain a 0<->1 transition; the run logic then touches just those (two per
// row for a blob).  After every 32-row block the warp folds the rows' horizontal sums into the <= 4 vertical bins they
// feed (ascending row order).  Bands are independent tasks; the LAST band of a mask to finish (atomic ticket after a
// __threadfence) adds the four partial results in a fixed order, so results are bit-reproducible from run to run.
#include <stdlib.h>

#include "hgl_common.cuh"
#include "resample.cuh"

namespace hgl {

constexpr int kMaxG = 32;          // grid side limit (14 for ViT-B/16, 24 for ViT-L/14@336)
constexpr int kRowsThreads = 256;  // one thread per bit row of a 256-row tile
constexpr int kEB = 4;             // expressions pooled per pass over a mask
constexpr int kPrefWarps = 8;
constexpr int kPrefRows = 32;       // heat-map rows per CTA of heat_prefix_kernel

// ---- shared with heat conditioning -----------------------------------------------------------------------------------
// gen_dir_mask as a tensor (the drop-in form; the pooling path never materialises it)
__global__ void dir_mask_kernel(int dirflag, int H, int W, float* __restrict__ out) {
  const size_t total = (size_t)H * W;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
    out[i] = ramp_at(dirflag, (int)(i % W), W);
}

// The same table built by a whole warp: the weights (double-precision index arithmetic, the expensive part) are evaluated one
// per lane; the normalising total stays a sequential float sum in tap order (one lane), so every value is bit-identical to
// aa_fill.  `w` is shared memory.
__device__ void aa_fill_warp(int i, int in_size, int out_size, int maxk, int* xmin_out, int* xsize_out, float* w, int lane) {
  const float scale = __fdiv_rn((float)in_size, (float)out_size);
  float support, invscale;
  if (scale >= 1.f) { support = scale; invscale = __fdiv_rn(1.f, scale); } else { support = 1.f; invscale = 1.f; }
  const float center = (float)((double)scale * ((double)i + 0.5));
  int xmin = (int)((double)__fsub_rn(center, support) + 0.5);
  xmin = max(xmin, 0);
  int xsize = min((int)((double)__fadd_rn(center, support) + 0.5), in_size) - xmin;
  xsize = max(min(xsize, maxk), 0);
  for (int j = lane; j < xsize; j += 32) {
    float t = (float)(((double)__fsub_rn((float)(j + xmin), center) + 0.5) * (double)invscale);
    t = fabsf(t);
    w[j] = (t < 1.f) ? __fsub_rn(1.f, t) : 0.f;
  }
  __syncwarp();
  float total = 0.f;
  if (lane == 0)
    for (int j = 0; j < xsize; ++j) total = __fadd_rn(total, w[j]);
  total = __shfl_sync(0xffffffffu, total, 0);
  for (int j = lane; j < maxk; j += 32) {
    if (j >= xsize) w[j] = 0.f;
    else if (total != 0.f) w[j] = __fdiv_rn(w[j], total);
  }
  __syncwarp();
  if (lane == 0) { *xmin_out = xmin; *xsize_out = xsize; }
}

// The tap tables of the mask pass, built ONCE per launch instead of once per CTA: block t < g = vertical table t, block g + i =
// horizontal table i plus the double-precision prefix sums of its taps.  The block written at `out` has exactly the layout of the
// mask pass's shared-memory header (ymin | ysize | xlo | xhi as int[kMaxG] each, then wy, wx, psum at the given byte offsets), so
// a CTA of the pass copies it in with 16-byte loads.
__global__ void __launch_bounds__(32) aa_tables_kernel(int H, int W, int g, int maxky, int maxkx, int off_wy, int off_wx, int off_psum,
                                                        uint8_t* __restrict__ out) {
  extern __shared__ float aa_w[];
  __shared__ int lo_s, size_s;
  const int t = blockIdx.x, lane = threadIdx.x;
  int* ints = reinterpret_cast<int*>(out);
  if (t < g) {
    aa_fill_warp(t, H, g, maxky, &lo_s, &size_s, aa_w, lane);
    __syncwarp();
    float* wy = reinterpret_cast<float*>(out + off_wy) + (size_t)t * maxky;
    for (int j = lane; j < maxky; j += 32) wy[j] = aa_w[j];
    if (lane == 0) { ints[t] = lo_s; ints[kMaxG + t] = size_s; }
  } else {
    const int i = t - g;
    aa_fill_warp(i, W, g, maxkx, &lo_s, &size_s, aa_w, lane);
    __syncwarp();
    float* wx = reinterpret_cast<float*>(out + off_wx) + (size_t)i * maxkx;
    for (int j = lane; j < maxkx; j += 32) wx[j] = aa_w[j];
    if (lane == 0) {
      double* ps = reinterpret_cast<double*>(out + off_psum) + (size_t)i * (maxkx + 1);
      double run = 0.0;
      ps[0] = 0.0;
      for (int k = 0; k < maxkx; ++k) { run += (double)aa_w[k]; ps[k + 1] = run; }
      ints[2 * kMaxG + i] = lo_s; ints[3 * kMaxG + i] = lo_s + size_s;
    }
  }
}

static int aa_maxk(int in_size, int out_size) {
  const float scale = (float)in_size / (float)out_size;
  const float support = scale >= 1.f ? scale : 1.f;
  return (int)ceilf(support) * 2 + 1;
}

// ---- heat-map prefix tables -------------------------------------------------------------------------------------------
struct HeatWs {            // workspace carve-up (byte offsets 256-aligned)
  float* cr;               // [E][H][Wp]   cr[y][x] = sum_{x' < x} A[y][x'] * ramp(x'),  x in [0, W]
  float* rp;               // [E][Wp]      rp[x]    = sum_{x' < x} ramp(x')
  float* rowstat;          // [E][H][4]    min A, max A, cr[y][W], -
  float* consts;           // [E][4]       min A, kk = 1/(range*mean A'), sum A'' , -
  int Wp;
  size_t bytes;
};

static HeatWs heat_carve(void* ws, int E, int H, int W) {
  HeatWs h;
  h.Wp = (W + 1 + 3) & ~3;
  size_t off = 0;
  auto take = [&](size_t n) { size_t o = off; off += (n + 255) & ~size_t(255); return o; };
  uint8_t* base = reinterpret_cast<uint8_t*>(ws);
  h.cr = reinterpret_cast<float*>(base + take((size_t)E * H * h.Wp * 4));
  h.rp = reinterpret_cast<float*>(base + take((size_t)E * h.Wp * 4));
  h.rowstat = reinterpret_cast<float*>(base + take((size_t)E * H * 4 * 4));
  h.consts = reinterpret_cast<float*>(base + take((size_t)E * 4 * 4));
  h.bytes = off;
  return h;
}

// One warp per (expression, row), 128 pixels per round: a lane owns 4 adjacent pixels (one 16-byte load, one 16-byte store),
// scans them in registers, the 32 lane totals go through one shuffle scan, and a running carry links the rounds.  Row 0's
// warp also emits the ramp prefix.  A CTA covers kPrefRows consecutive rows (kPrefRows / kPrefWarps per warp), so the
// per-CTA tables (ramp row, and for kLR the horizontal pass below) are built once per 32 rows.
//
// kLR: `heat` is the RAW GEM map [E,hh,hw] and the frame-sized map of Hybridgl_main.py:201, T.Resize((H,W), antialias=True),
// is evaluated on the fly (ATen _upsample_bilinear2d_aa, separable: horizontal pass, then vertical pass; for an up-sampler
// the triangle filter has <= 3 taps per axis).  The 32 output rows of a CTA only see a handful of raw rows (kPrefRows*hh/H + 3),
// so the CTA first runs the HORIZONTAL pass for those raw rows into shared memory (exactly ATen's intermediate tensor,
// same tap order), and a pixel is then three shared-memory reads and the vertical taps.  The raw map (a few KB per
// expression) is read through L1; the frame-sized heat-map never exists in HBM.
// kOut (with kLR): write the resized map itself ([E,H,W] in `cr`, pitch Wp = W) instead of the tables -- the up-sampling fast
// path of hgl_heat_resize_aa (same per-CTA horizontal pass, same tap order as the table form and as ATen).
template <bool kLR, bool kOut = false>
__global__ void __launch_bounds__(kPrefWarps * 32) heat_prefix_kernel(const float* __restrict__ heat, const int32_t* __restrict__ dirflag,
                                                                      int H, int W, int Wp, int hh, int hw, int nr_max, float* __restrict__ cr,
                                                                      float* __restrict__ rp, float* __restrict__ rowstat) {
  extern __shared__ __align__(16) float ramp[];      // [W rounded up to 128]  (+ kLR: hrow [nr_max][W128])
  const int e = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int W128 = (W + 127) & ~127;
  const int dir = kOut ? 0 : dirflag[e];
  const int y_first = blockIdx.x * kPrefRows, y_last = min(H, y_first + kPrefRows) - 1;
  float* hrow = ramp + W128;                              // kLR: horizontal pass of raw rows [ry0, ry0 + nr)
  __shared__ int s_ymin[kPrefRows], s_ysize[kPrefRows];   // kLR: vertical taps of the CTA's rows, built once (not once per lane)
  __shared__ float s_wy[kPrefRows][3];
  if (kLR && (int)threadIdx.x <= y_last - y_first)
    aa_fill(y_first + threadIdx.x, hh, H, 3, &s_ymin[threadIdx.x], &s_ysize[threadIdx.x], s_wy[threadIdx.x]);
  int ry0 = 0, nr = 0;
  if (kLR) {
    int ym, ys;
    float wtmp[3];
    aa_fill(y_first, hh, H, 3, &ry0, &ys, wtmp);
    aa_fill(y_last, hh, H, 3, &ym, &ys, wtmp);
    nr = min(ym + ys - ry0, nr_max);
  }
  for (int x = threadIdx.x; x < W128; x += blockDim.x) {
    ramp[x] = (x < W) ? ramp_at(dir, x, W) : 0.f;
    if (kLR) {
      int xm = 0, xs = 0;
      float wx[3] = {0.f, 0.f, 0.f};
      if (x < W) aa_fill(x, hw, W, 3, &xm, &xs, wx);
      const float* A0 = heat + ((size_t)e * hh + ry0) * hw + xm;
      for (int r = 0; r < nr; ++r) {
        float t = 0.f;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx)
          if (kx < xs) t = __fadd_rn(t, __fmul_rn(__ldg(A0 + (size_t)r * hw + kx), wx[kx]));
        hrow[r * W128 + x] = t;
      }
    }
  }
  __syncthreads();
  for (int y = y_first + warp; y <= y_last; y += kPrefWarps) {      // whole warps; no block-level barrier below
    const float* A = heat + ((size_t)e * H + y) * W;                // !kLR only
    const bool vec = !kLR && (W & 3) == 0 && (reinterpret_cast<uintptr_t>(heat) & 15) == 0;
    int ymin = 0, ysize = 0;
    float wy[3] = {0.f, 0.f, 0.f};
    if (kLR) {
      const int r = y - y_first;
      ymin = s_ymin[r]; ysize = s_ysize[r]; wy[0] = s_wy[r][0]; wy[1] = s_wy[r][1]; wy[2] = s_wy[r][2];
    }
    const float* hr = hrow + (ymin - ry0) * W128;
    if (kLR && kOut) {                                     // the resized row itself: [E,H,W], scalar stores (any W / alignment)
      float* orow = cr + ((size_t)e * H + y) * W;
      for (int x0 = 0; x0 < W; x0 += 128) {
        const int x = x0 + 4 * lane;
        float h4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          if (ky < ysize) {
            const float4 t4 = *reinterpret_cast<const float4*>(hr + ky * W128 + x);
            h4[0] = __fadd_rn(h4[0], __fmul_rn(t4.x, wy[ky])); h4[1] = __fadd_rn(h4[1], __fmul_rn(t4.y, wy[ky]));
            h4[2] = __fadd_rn(h4[2], __fmul_rn(t4.z, wy[ky])); h4[3] = __fadd_rn(h4[3], __fmul_rn(t4.w, wy[ky]));
          }
        }
        if (x + 3 < W && ((reinterpret_cast<uintptr_t>(orow + x) & 15) == 0)) *reinterpret_cast<float4*>(orow + x) = make_float4(h4[0], h4[1], h4[2], h4[3]);
        else {
#pragma unroll
          for (int q = 0; q < 4; ++q) if (x + q < W) orow[x + q] = h4[q];
        }
      }
      continue;
    }
    for (int pass = (y == 0 ? 0 : 1); pass < 2; ++pass) {   // pass 0 (row 0 only): the ramp itself; pass 1: A * ramp
      float* dst = (pass == 0) ? rp + (size_t)e * Wp : cr + ((size_t)e * H + y) * Wp;
      float mn = INFINITY, mx = -INFINITY, carry = 0.f;
      for (int x0 = 0; x0 < W; x0 += 128) {
        const int x = x0 + 4 * lane;
        const float4 r4 = *reinterpret_cast<const float4*>(ramp + x);
        float a[4] = {0.f, 0.f, 0.f, 0.f};
        if (pass == 0) { a[0] = r4.x; a[1] = r4.y; a[2] = r4.z; a[3] = r4.w; }
        else {
          float h4[4] = {0.f, 0.f, 0.f, 0.f};
          if (kLR) {                                         // vertical pass: o = sum_ky hrow[ymin+ky][x] * wy[ky], ascending ky
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
              if (ky < ysize) {
                const float4 t4 = *reinterpret_cast<const float4*>(hr + ky * W128 + x);
                h4[0] = __fadd_rn(h4[0], __fmul_rn(t4.x, wy[ky])); h4[1] = __fadd_rn(h4[1], __fmul_rn(t4.y, wy[ky]));
                h4[2] = __fadd_rn(h4[2], __fmul_rn(t4.z, wy[ky])); h4[3] = __fadd_rn(h4[3], __fmul_rn(t4.w, wy[ky]));
              }
            }
          } else if (vec && x + 3 < W) { const float4 v = __ldg(reinterpret_cast<const float4*>(A + x)); h4[0] = v.x; h4[1] = v.y; h4[2] = v.z; h4[3] = v.w; }
          else {
#pragma unroll
            for (int q = 0; q < 4; ++q) if (x + q < W) h4[q] = __ldg(A + x + q);
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) if (x + q < W) { mn = fminf(mn, h4[q]); mx = fmaxf(mx, h4[q]); }
          a[0] = __fmul_rn(h4[0], r4.x); a[1] = __fmul_rn(h4[1], r4.y); a[2] = __fmul_rn(h4[2], r4.z); a[3] = __fmul_rn(h4[3], r4.w);
        }
        const float p0 = a[0], p1 = __fadd_rn(p0, a[1]), p2 = __fadd_rn(p1, a[2]), p3 = __fadd_rn(p2, a[3]);
        float incl = p3;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const float nb = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl = __fadd_rn(incl, nb);
        }
        float excl = __shfl_up_sync(0xffffffffu, incl, 1);            // sum of the lanes before this one (exact, no subtraction)
        if (lane == 0) excl = 0.f;
        const float base = __fadd_rn(carry, excl);                     // exclusive prefix of this lane's first pixel
        const float4 o4 = make_float4(base, __fadd_rn(base, p0), __fadd_rn(base, p1), __fadd_rn(base, p2));
        if (x + 3 < Wp) *reinterpret_cast<float4*>(dst + x) = o4;      // entries beyond W inside the padded row are never read
        else {
          const float ov[4] = {o4.x, o4.y, o4.z, o4.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) if (x + q < Wp) dst[x + q] = ov[q];
        }
        carry = __fadd_rn(carry, __shfl_sync(0xffffffffu, incl, 31));
      }
      if (lane == 0 && (W & 127) == 0) dst[W] = carry;              // otherwise the lane that owns index W wrote it above
      if (pass == 1) {
        mn = warp_min(mn); mx = warp_max(mx);
        if (lane == 0) {
          float* o = rowstat + ((size_t)e * H + y) * 4;
          o[0] = mn; o[1] = mx; o[2] = carry; o[3] = 0.f;
        }
      }
      __syncwarp();
    }
  }
}

// per expression: min, kk = 1 / (range * mean(A')), sum(A'')   (Hybridgl_main.py:204,209) from the row statistics
__global__ void __launch_bounds__(128) heat_consts_kernel(const float* __restrict__ rowstat, const float* __restrict__ rp, int H, int W, int Wp,
                                                          float* __restrict__ consts) {
  const int e = blockIdx.x, tid = threadIdx.x;
  float lo = INFINITY, hi = -INFINITY;
  double s1 = 0.0;
  for (int y = tid; y < H; y += blockDim.x) {
    const float* o = rowstat + ((size_t)e * H + y) * 4;
    lo = fminf(lo, o[0]); hi = fmaxf(hi, o[1]); s1 += (double)o[2];
  }
  __shared__ float slo[4], shi[4];
  __shared__ double ss[4];
  lo = warp_min(lo); hi = warp_max(hi);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s1 += __shfl_xor_sync(0xffffffffu, s1, o);
  if ((tid & 31) == 0) { slo[tid >> 5] = lo; shi[tid >> 5] = hi; ss[tid >> 5] = s1; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < 4; ++w) { lo = fminf(lo, slo[w]); hi = fmaxf(hi, shi[w]); s1 += ss[w]; }
    const double s0 = (double)H * (double)rp[(size_t)e * Wp + W];     // sum of the ramp over the frame
    const double range = (double)hi - (double)lo;
    const double sum_ap = (s1 - (double)lo * s0) / range;               // sum of A' = minmax(A) * ramp
    const double mean = sum_ap / ((double)H * (double)W);
    const double kk = 1.0 / (range * mean);
    float* o = consts + (size_t)e * 4;
    o[0] = lo; o[1] = (float)kk; o[2] = (float)(kk * (s1 - (double)lo * s0)); o[3] = 0.f;
  }
}

// T.Resize((H,W), antialias=True) of the raw GEM maps as a tensor (Hybridgl_main.py:201; the drop-in form, any scale):
// one thread per output pixel, separable triangle filter with ATen's index / weight arithmetic, horizontal pass then vertical.
struct AaTap { float scale, support, invscale, center, total; int xmin, xsize; };
__device__ __forceinline__ float aa_raw_weight(const AaTap& t, int j) {
  const float v = fabsf((float)(((double)__fsub_rn((float)(j + t.xmin), t.center) + 0.5) * (double)t.invscale));
  return (v < 1.f) ? __fsub_rn(1.f, v) : 0.f;
}
__device__ __forceinline__ AaTap aa_tap(int i, int in_size, int out_size) {
  AaTap t;
  t.scale = __fdiv_rn((float)in_size, (float)out_size);
  if (t.scale >= 1.f) { t.support = t.scale; t.invscale = __fdiv_rn(1.f, t.scale); } else { t.support = 1.f; t.invscale = 1.f; }
  t.center = (float)((double)t.scale * ((double)i + 0.5));
  t.xmin = max((int)((double)__fsub_rn(t.center, t.support) + 0.5), 0);
  t.xsize = max(min((int)((double)__fadd_rn(t.center, t.support) + 0.5), in_size) - t.xmin, 0);
  t.total = 0.f;
  for (int j = 0; j < t.xsize; ++j) t.total = __fadd_rn(t.total, aa_raw_weight(t, j));
  return t;
}
__device__ __forceinline__ float aa_weight(const AaTap& t, int j) {
  const float w = aa_raw_weight(t, j);
  return (t.total != 0.f) ? __fdiv_rn(w, t.total) : w;
}
__global__ void __launch_bounds__(256) heat_resize_aa_kernel(const float* __restrict__ src, int E, int hh, int hw, int H, int W,
                                                             float* __restrict__ out) {
  const size_t total = (size_t)E * H * W;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % W), y = (int)((i / W) % H), e = (int)(i / ((size_t)W * H));
    const AaTap tx = aa_tap(x, hw, W), ty = aa_tap(y, hh, H);
    const float* s = src + (size_t)e * hh * hw;
    float o = 0.f;
    for (int ky = 0; ky < ty.xsize; ++ky) {
      const float* r = s + (size_t)(ty.xmin + ky) * hw + tx.xmin;
      float t = 0.f;
      for (int kx = 0; kx < tx.xsize; ++kx) t = __fadd_rn(t, __fmul_rn(__ldg(r + kx), aa_weight(tx, kx)));
      o = __fadd_rn(o, __fmul_rn(t, aa_weight(ty, ky)));
    }
    out[i] = o;
  }
}

// ---- the pass over the packed masks ----------------------------------------------------------------------------------
constexpr int kBands = 8;                              // row bands (= independent warp tasks) per mask
#ifndef HGL_ROWS_FRONT
#define HGL_ROWS_FRONT 1
#endif
#ifndef HGL_ROWS_GBATCH
#define HGL_ROWS_GBATCH 1
#endif
constexpr int kRowVecChunk = 5;                        // 16-byte loads in flight per lane while a bit row is scanned (5 = a 640-pixel row)
constexpr int kRowWordChunk = 13;                      // same for rows that are not 16-byte multiples (13 + 12 words = an 800-pixel row)

constexpr size_t kAaTabBytes = 224 * 1024;             // room for the tap-table block (never larger than the pass's shared memory)

struct RowsScratch {         // global scratch of one launch (byte offsets 256-aligned)
  uint8_t* aatab;            // [kAaTabBytes]  tap tables of the antialiased resize (aa_tables_kernel), grid passes only
  int32_t* tickets;          // [M + 1]        zero at launch; entry M is the task counter of the dynamic scheduler
  int32_t* pcnt;             // [M][kBands]    pixels per band
  float* pgrid;              // [M][kBands][g*g]
  float* pheat;              // [E][max_n][kBands][2]   (sum cr, sum rp) per band
  size_t bytes;
};
static RowsScratch rows_carve(void* ws, int M, int g, int E, int max_n) {
  RowsScratch r;
  size_t off = 0;
  auto take = [&](size_t n) { size_t o = off; off += (n + 255) & ~size_t(255); return o; };
  uint8_t* base = reinterpret_cast<uint8_t*>(ws);
  r.aatab = base + take(g > 0 ? kAaTabBytes : 0);
  r.tickets = reinterpret_cast<int32_t*>(base + take((size_t)(M + 1) * 4));
  r.pcnt = reinterpret_cast<int32_t*>(base + take((size_t)M * kBands * 4));
  r.pgrid = reinterpret_cast<float*>(base + take((size_t)M * kBands * g * g * 4));
  r.pheat = reinterpret_cast<float*>(base + take((size_t)E * max_n * kBands * 2 * 4));
  r.bytes = off;
  return r;
}

struct RowsParams {
  const uint32_t* bits;          // [M,H,WW]
  int M, H, W, WW;
  // grid part
  int g, maxky, maxkx;
  float* grid;                   // [M,g,g]
  int32_t* area;                 // [M] or null
  // heat part
  const int32_t* mask_off; const int32_t* expr_off;
  int B, E, max_n, Wp;
  const float* cr; const float* rp; const float* consts; const float* black;
  float* score_gem;              // [E,max_n]
  RowsScratch sc;
  // shared-memory carve-up (byte offsets)
  int off_wy, off_wx, off_psum, off_hbuf, off_pgrid;
};

template <bool kGrid, bool kHeat>
__global__ void __launch_bounds__(kRowsThreads, 4) mask_rows_kernel(const RowsParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  int* ymin = reinterpret_cast<int*>(smem);
  int* ysize = ymin + kMaxG;
  int* xlo = ysize + kMaxG;
  int* xhi = xlo + kMaxG;
  float* wy = reinterpret_cast<float*>(smem + p.off_wy);        // [g][maxky]
  double* psum = reinterpret_cast<double*>(smem + p.off_psum);  // [g][maxkx+1] prefix sums of wx (double: tiny edge taps survive)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int H = p.H, W = p.W, WW = p.WW, g = p.g, gg = g * g;
  const int gp = g | 1;
  float* hbuf = reinterpret_cast<float*>(smem + p.off_hbuf) + (size_t)warp * 32 * gp;     // [32][gp] horizontal sums of this warp's 32 rows
  float* pgrid = reinterpret_cast<float*>(smem + p.off_pgrid) + (size_t)warp * gg;        // [g][g] this band's share of the grid
  const int pk = p.maxkx + 1;
  float inv_sx = 0.f;
  if (kGrid) {                                                     // the tap tables come ready-made from aa_tables_kernel
    const uint4* src = reinterpret_cast<const uint4*>(p.sc.aatab);
    uint4* dst = reinterpret_cast<uint4*>(smem);
    for (int t = tid; t < (p.off_hbuf >> 4); t += kRowsThreads) dst[t] = __ldg(src + t);
    inv_sx = (float)g / (float)W;
    __syncthreads();
  }

  const int band_rows = (H + kBands - 1) / kBands;
  const size_t mask_words = (size_t)H * WW;
  const bool vec = (WW & 3) == 0 && (reinterpret_cast<uintptr_t>(p.bits) & 15) == 0;
  const float hw = (float)((size_t)H * W);
  const int total_tasks = p.M * kBands;
  // dynamic scheduling: masks differ a lot in how many rows they touch, so warps pull (mask, band) tasks from a counter
  // (the ticket of the NEXT task is drawn before the current one is processed, so its round trip is never waited for)
  int ticket = 0;
  if (lane == 0) ticket = atomicAdd(p.sc.tickets + p.M, 1);
  for (;;) {
    const int task = __shfl_sync(0xffffffffu, ticket, 0);
    if (task >= total_tasks) break;
    if (lane == 0) ticket = atomicAdd(p.sc.tickets + p.M, 1);
    const int m = task / kBands, band = task - m * kBands;
    const int y_begin = band * band_rows, y_end = min(H, y_begin + band_rows);
    int b = 0, n_lo = 0, e_lo = 0, e_hi = 0;
    if (kHeat) {
      if (p.mask_off) {                                  // image of mask m: last b with mask_off[b] <= m
        int lo = 0, hi = p.B - 1;
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (p.mask_off[mid] <= m) lo = mid; else hi = mid - 1; }
        b = lo; n_lo = p.mask_off[b];
      }
      e_lo = p.expr_off ? p.expr_off[b] : 0;
      e_hi = p.expr_off ? p.expr_off[b + 1] : p.E;
    }
    const uint32_t* mb = p.bits + (size_t)m * mask_words;
    int cnt = 0;
    if (kGrid)
      for (int t = lane; t < gg; t += 32) pgrid[t] = 0.f;
    bool first = true;
    for (int eg = e_lo; first || eg < e_hi; eg += kEB) {     // groups of kEB expressions; more than one pass is rare
      const int ne = kHeat ? max(0, min(kEB, e_hi - eg)) : 0;
      const bool do_grid = kGrid && first;
      float acc[kEB], accr[kEB];
#pragma unroll
      for (int j = 0; j < kEB; ++j) { acc[j] = 0.f; accr[j] = 0.f; }
      for (int y0 = y_begin; y0 < y_end; y0 += 32) {
        const int y = y0 + lane;
        const bool have = y < y_end;
        const uint32_t* rowp = mb + (size_t)(have ? y : y_begin) * WW;
        // ---- phase A: the row's words pass through registers once: pixel count + bitmap of words with a transition
        uint64_t im = 0;
        int c_row = 0;
        if (have) {
          uint32_t prev = 0;                                          // last bit of the previous word
          auto look = [&](uint32_t v, int w) {
            c_row += __popc(v);
            if (v != (0u - prev)) im |= 1ull << w;                    // neither "all outside" after a 0 nor "all inside" after a 1
            prev = v >> 31;
          };
#if HGL_ROWS_FRONT
          // every load of a chunk is issued before the first word is looked at: one memory round trip per chunk (a 640-pixel
          // row is one chunk) instead of one per 16 bytes -- the round trips, not the bytes, are what this pass waits for
          if (vec) {
            const uint4* r4 = reinterpret_cast<const uint4*>(rowp);
            const int nu = WW >> 2;
            for (int u0 = 0; u0 < nu; u0 += kRowVecChunk) {
              uint4 q[kRowVecChunk];
#pragma unroll
              for (int k = 0; k < kRowVecChunk; ++k) q[k] = (u0 + k < nu) ? __ldg(r4 + u0 + k) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
              for (int k = 0; k < kRowVecChunk; ++k) {
                const int u = u0 + k;
                if (u >= nu) break;
                const uint4 v = q[k];
                // 128 pixels on one side of the outline (most of a frame for a blob): nothing to record
                if ((v.x | v.y | v.z | v.w) == 0u && prev == 0u) continue;
                if ((v.x & v.y & v.z & v.w) == 0xffffffffu && prev == 1u) { c_row += 128; continue; }
                look(v.x, 4 * u); look(v.y, 4 * u + 1); look(v.z, 4 * u + 2); look(v.w, 4 * u + 3);
              }
            }
          } else {
            for (int w0 = 0; w0 < WW; w0 += kRowWordChunk) {
              uint32_t q[kRowWordChunk];
#pragma unroll
              for (int k = 0; k < kRowWordChunk; ++k) q[k] = (w0 + k < WW) ? __ldg(rowp + w0 + k) : 0u;
#pragma unroll
              for (int k = 0; k < kRowWordChunk; ++k) {
                if (w0 + k >= WW) break;
                look(q[k], w0 + k);
              }
            }
          }
#else
          if (vec) {
            const uint4* r4 = reinterpret_cast<const uint4*>(rowp);
            for (int u = 0; u < (WW >> 2); ++u) {
              const uint4 v = __ldg(r4 + u);
              // 128 pixels on one side of the outline (most of a frame for a blob): nothing to record
              if ((v.x | v.y | v.z | v.w) == 0u && prev == 0u) continue;
              if ((v.x & v.y & v.z & v.w) == 0xffffffffu && prev == 1u) { c_row += 128; continue; }
              look(v.x, 4 * u); look(v.y, 4 * u + 1); look(v.z, 4 * u + 2); look(v.w, 4 * u + 3);
            }
          } else {
            for (int w = 0; w < WW; ++w) look(__ldg(rowp + w), w);
          }
#endif
          if (prev) im |= 1ull << WW;                                 // run open at the frame edge: closed by the virtual zero word WW
        }
        if (first) cnt += c_row;
        const bool nonempty = c_row != 0;
        float* hr = hbuf + lane * gp;
        int bx_lo = g, bx_hi = -1;
        if (do_grid && nonempty)
          for (int i = 0; i < g; ++i) hr[i] = 0.f;
        // ---- phase B: only the words that hold a transition (two per row for a blob)
        if (im != 0ull) {
          const float* crow[kEB];
          const float* rrow[kEB];
#pragma unroll
          for (int j = 0; j < kEB; ++j) {
            const int ej = eg + min(j, max(ne - 1, 0));
            crow[j] = kHeat ? p.cr + ((size_t)ej * H + y) * p.Wp : nullptr;
            rrow[j] = kHeat ? p.rp + (size_t)ej * p.Wp : nullptr;
          }
          uint32_t carry = 0;
          int xs = 0;
#pragma unroll 1
          while (im) {
            const int w = __ffsll((long long)im) - 1;
            im &= im - 1;
            const uint32_t v = (w < WW) ? __ldg(rowp + w) : 0u;       // L1 hit: loaded a moment ago
            uint32_t t = v ^ ((v << 1) | carry);
            carry = v >> 31;
            while (t) {
              const int bp = __ffs(t) - 1;
              t &= t - 1;
              const int x = 32 * w + bp;
              if ((v >> bp) & 1u) { xs = x; continue; }
              const int s0 = xs, e0 = min(x, W);                      // pixels [s0, e0) of row y are inside the mask
              if (do_grid) {
                int i = max(0, min(g - 1, (int)((float)s0 * inv_sx - 1.5f) - 1));
                while (i < g && xhi[i] <= s0) ++i;
                bx_lo = min(bx_lo, i);
                for (; i < g && xlo[i] < e0; ++i) {
                  const double* ps = psum + i * pk - xlo[i];
                  const int a0 = min(max(s0, xlo[i]), xhi[i]), a1 = min(max(e0, xlo[i]), xhi[i]);
                  hr[i] += (float)(ps[a1] - ps[a0]);
                }
                bx_hi = max(bx_hi, i - 1);
              }
#if HGL_ROWS_GBATCH
              if (kHeat && ne > 0) {                                  // all table reads of the run are in flight before the first add
                float c1[kEB], c0[kEB], r1[kEB], r0[kEB];
#pragma unroll
                for (int j = 0; j < kEB; ++j) {                       // predicated loads, no branches between them
                  const bool on = j < ne;
                  c1[j] = on ? __ldg(crow[j] + e0) : 0.f; c0[j] = on ? __ldg(crow[j] + s0) : 0.f;
                  r1[j] = on ? __ldg(rrow[j] + e0) : 0.f; r0[j] = on ? __ldg(rrow[j] + s0) : 0.f;
                }
#pragma unroll
                for (int j = 0; j < kEB; ++j) {
                  if (j < ne) { acc[j] += c1[j] - c0[j]; accr[j] += r1[j] - r0[j]; }
                }
              }
#else
              if (kHeat) {
#pragma unroll
                for (int j = 0; j < kEB; ++j) {
                  if (j < ne) {
                    acc[j] += __ldg(crow[j] + e0) - __ldg(crow[j] + s0);
                    accr[j] += __ldg(rrow[j] + e0) - __ldg(rrow[j] + s0);
                  }
                }
              }
#endif
            }
          }
        }
        if (do_grid) {
          // fold the block's rows into the vertical bins they feed, ascending row order:
          //   pgrid[gy][gx] = pgrid[gy][gx] + h[y][gx] * wy[gy][y - ymin[gy]]     (separate multiply and add, like the reference)
          const uint32_t rows_mask = __ballot_sync(0xffffffffu, nonempty);
          if (rows_mask != 0u) {
            int gx0 = bx_lo, gx1 = bx_hi;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              gx0 = min(gx0, __shfl_xor_sync(0xffffffffu, gx0, o));
              gx1 = max(gx1, __shfl_xor_sync(0xffffffffu, gx1, o));
            }
            const int r_first = __ffs(rows_mask) - 1, r_last = 31 - __clz(rows_mask);
            const int ya = y0 + r_first, yb = y0 + r_last;           // first / last non-empty row of the block
            int gy_lo = 0;
            while (gy_lo < g && ymin[gy_lo] + ysize[gy_lo] <= ya) ++gy_lo;
            int gy_hi = gy_lo;
            while (gy_hi + 1 < g && ymin[gy_hi + 1] <= yb) ++gy_hi;
            const int nbx = gx1 - gx0 + 1, nel = (gy_hi - gy_lo + 1) * nbx;
            __syncwarp();
            for (int t = lane; t < nel; t += 32) {
              const int gyl = t / nbx, gx = gx0 + t - gyl * nbx, gy = gy_lo + gyl;
              const int y_lo = ymin[gy], y_sz = ysize[gy];
              const float* wp = wy + gy * p.maxky + (y0 - y_lo);
              const float* hp = hbuf + gx;
              const int k_lo = max(r_first, y_lo - y0), k_hi = min(r_last, y_lo + y_sz - 1 - y0);
              float a = pgrid[gy * g + gx];
#pragma unroll 4
              for (int r = k_lo; r <= k_hi; ++r)
                if ((rows_mask >> r) & 1u) a = __fadd_rn(a, __fmul_rn(hp[r * gp], wp[r]));
              pgrid[gy * g + gx] = a;
            }
            __syncwarp();
          }
        }
      }
      if (kHeat && ne > 0) {                                          // this band's pooled sums of the group
#pragma unroll
        for (int j = 0; j < kEB; ++j) {
          const float sc = warp_sum(acc[j]), sr = warp_sum(accr[j]);
          if (lane == 0 && j < ne && (m - n_lo) < p.max_n) {        // a mask beyond the row stride is dropped, never written past its row
            float* o = p.sc.pheat + (((size_t)(eg + j) * p.max_n + (m - n_lo)) * kBands + band) * 2;
            o[0] = sc; o[1] = sr;
          }
        }
      }
      first = false;
    }
    cnt = warp_sum_i(cnt);
    if (lane == 0) p.sc.pcnt[m * kBands + band] = cnt;
    if (kGrid) {
      __syncwarp();
      float* o = p.sc.pgrid + ((size_t)m * kBands + band) * gg;
      for (int t = lane; t < gg; t += 32) o[t] = pgrid[t];
    }
    // ---- ticket: the last band of mask m combines the four partial results in band order
    __threadfence();
    __syncwarp();                                                     // every lane's partial results are fenced before the ticket
    int last = 0;
    if (lane == 0) last = (atomicAdd(p.sc.tickets + m, 1) == kBands - 1);
    last = __shfl_sync(0xffffffffu, last, 0);
    if (!last) continue;
    __threadfence();
    int area_m = 0;
#pragma unroll
    for (int q = 0; q < kBands; ++q) area_m += __ldcg(p.sc.pcnt + m * kBands + q);
    if (lane == 0 && p.area) p.area[m] = area_m;
    if (kGrid) {
      const float* pg = p.sc.pgrid + (size_t)m * kBands * gg;
      for (int t = lane; t < gg; t += 32) {
        float a = __ldcg(pg + t);
#pragma unroll
        for (int q = 1; q < kBands; ++q) a = __fadd_rn(a, __ldcg(pg + q * gg + t));
        p.grid[(size_t)m * gg + t] = a;
      }
    }
    if (kHeat && (m - n_lo) < p.max_n) {
      for (int e = e_lo + lane; e < e_hi; e += 32) {
        const float* o = p.sc.pheat + ((size_t)e * p.max_n + (m - n_lo)) * kBands * 2;
        float sc = 0.f, sr = 0.f;
#pragma unroll
        for (int q = 0; q < kBands; ++q) { sc += __ldcg(o + 2 * q); sr += __ldcg(o + 2 * q + 1); }
        const float mn = p.consts[e * 4], kk = p.consts[e * 4 + 1], s_tot = p.consts[e * 4 + 2];
        const float s_in = kk * (sc - mn * sr);
        const float bl = p.black[e];
        p.score_gem[(size_t)e * p.max_n + (m - n_lo)] = (2.f - bl) * s_in / (float)area_m - bl * (s_tot - s_in) / (hw - (float)area_m);
      }
    }
  }
}

// ---- legacy 4-tap path and small helpers ------------------------------------------------------------------------------
// non-antialiased: 4-tap bilinear sample of the packed mask at the g x g grid positions (ATen upsample_bilinear2d;
// what the reference's pinned torchvision 0.15.2 computes for TF.resize on tensors)
__global__ void mask_grid_noaa_kernel(const uint32_t* __restrict__ bits, int M, int H, int W, int g, float* __restrict__ grid) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= M * g * g) return;
  const int gx = idx % g, gy = (idx / g) % g, m = idx / (g * g);
  const int WW = (W + 31) >> 5;
  auto taps = [](int dst, int in_size, int out_size, int& i0, int& d, float& w0, float& w1) {
    if (in_size == out_size) { i0 = dst; d = 0; w0 = 1.f; w1 = 0.f; return; }
    const float scale = __fdiv_rn((float)in_size, (float)out_size);
    float src = fmaxf(__fmaf_rn(scale, (float)dst + 0.5f, -0.5f), 0.f);
    i0 = min((int)src, in_size - 1);
    d = (i0 < in_size - 1) ? 1 : 0;
    w1 = fminf(fmaxf(__fsub_rn(src, (float)i0), 0.f), 1.f);
    w0 = __fsub_rn(1.f, w1);
  };
  int y0, dy, x0, dx; float wy0, wy1, wx0, wx1;
  taps(gy, H, g, y0, dy, wy0, wy1);
  taps(gx, W, g, x0, dx, wx0, wx1);
  const uint32_t* b = bits + (size_t)m * H * WW;
  auto bit = [&](int y, int x) { return ((b[(size_t)y * WW + (x >> 5)] >> (x & 31)) & 1u) ? 1.f : 0.f; };
  const float a = bit(y0, x0), bb = bit(y0, x0 + dx), c = bit(y0 + dy, x0), d = bit(y0 + dy, x0 + dx);
  const float top = __fmaf_rn(a, wx0, __fmul_rn(bb, wx1)), bot = __fmaf_rn(c, wx0, __fmul_rn(d, wx1));
  grid[idx] = __fmaf_rn(top, wy0, __fmul_rn(bot, wy1));
}

// pixel count per mask from the packed words (used when antialias == 0 and the caller still wants areas)
__global__ void __launch_bounds__(256) mask_area_kernel(const uint32_t* __restrict__ bits, int M, size_t words, int32_t* __restrict__ area) {
  __shared__ int red[8];
  for (int m = blockIdx.x; m < M; m += gridDim.x) {
    const uint32_t* b = bits + (size_t)m * words;
    int s = 0;
    for (size_t i = threadIdx.x; i < words; i += blockDim.x) s += __popc(b[i]);
    s = warp_sum_i(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < 8; ++w) t += red[w]; area[m] = t; }
    __syncthreads();
  }
}

// Launch the pass.  want_grid / want_heat select the template instance.
static int launch_rows(RowsParams p, bool want_grid, bool want_heat, void* scratch, cudaStream_t st) {
  HGL_REQUIRE(p.WW <= 63, "mask rows pass: W=%d wider than 2016", p.W);
  p.sc = rows_carve(scratch, p.M, want_grid ? p.g : 0, want_heat ? p.E : 0, want_heat ? p.max_n : 0);
  // one memset: the tickets, and (grid passes) the tap-table block in front of them, whose unused slots and alignment gaps the CTAs
  // copy along with the tables -- zero rather than uninitialised (compute-sanitizer initcheck)
  uint8_t* z0 = want_grid ? p.sc.aatab : reinterpret_cast<uint8_t*>(p.sc.tickets);
  cudaError_t e = cudaMemsetAsync(z0, 0, (size_t)(reinterpret_cast<uint8_t*>(p.sc.tickets) - z0) + (size_t)(p.M + 1) * 4, st);
  if (e != cudaSuccess) { set_error("mask rows pass: cudaMemsetAsync: %s", cudaGetErrorString(e)); return HGL_ECUDA; }
  size_t off = (size_t)4 * kMaxG * 4;
  auto take = [&](size_t n, size_t align) { off = (off + align - 1) & ~(align - 1); size_t o = off; off += n; return (int)o; };
  const int warps = kRowsThreads / 32;
  p.off_wy = take(want_grid ? (size_t)p.g * p.maxky * 4 : 0, 16);
  p.off_wx = take(want_grid ? (size_t)p.g * p.maxkx * 4 : 0, 16);
  p.off_psum = take(want_grid ? (size_t)p.g * (p.maxkx + 1) * 8 : 0, 16);
  p.off_hbuf = take(want_grid ? (size_t)warps * 32 * (p.g | 1) * 4 : 0, 16);
  p.off_pgrid = take(want_grid ? (size_t)warps * p.g * p.g * 4 : 0, 16);
  const size_t smem = off;
  HGL_REQUIRE(smem <= 220 * 1024, "mask rows pass: frame %dx%d (g=%d) needs %zu B of shared memory", p.H, p.W, p.g, smem);
  if (want_grid) {
    const size_t tab_smem = (size_t)std::max(p.maxky, p.maxkx) * 4;
    aa_tables_kernel<<<2 * p.g, 32, tab_smem, st>>>(p.H, p.W, p.g, p.maxky, p.maxkx, p.off_wy, p.off_wx, p.off_psum, p.sc.aatab);
    const int rc1 = launch_status("mask rows pass (tap tables)");
    if (rc1 != HGL_OK) return rc1;
  }
  int per_sm = (int)std::max<size_t>(1, std::min<size_t>(6, (200 * 1024) / smem));
  per_sm = std::max(1, std::min(per_sm, tuning_int("HGL_ROWS_CTAS_PER_SM", per_sm)));
  const int ctas = std::min(ceil_div(p.M * kBands, warps), sm_count() * per_sm);
  auto go = [&](auto kern) -> int {
    const int rc2 = ensure_dyn_smem(reinterpret_cast<const void*>(kern), smem, "mask rows pass");
    if (rc2 != HGL_OK) return rc2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(kRowsThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1] = {priority_attr(st)};
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaError_t e3 = cudaLaunchKernelEx(&cfg, kern, p);
    if (e3 != cudaSuccess) { set_error("mask rows pass: cudaLaunchKernelEx: %s", cudaGetErrorString(e3)); return HGL_ECUDA; }
    return launch_status("mask rows pass");
  };
  if (want_grid && want_heat) return go(mask_rows_kernel<true, true>);
  if (want_grid) return go(mask_rows_kernel<true, false>);
  return go(mask_rows_kernel<false, true>);
}

// hh > 0: `heat` is the raw map [E,hh,hw] (resized on the fly when both axes are up-sampled, else materialised in `full`)
static int launch_heat_tables(const float* heat, int hh, int hw, float* full, const int32_t* dirflag, int E, int H, int W, const HeatWs& ws,
                              cudaStream_t st) {
  bool lr = hh > 0;
  if (lr && (aa_maxk(hh, H) > 3 || aa_maxk(hw, W) > 3)) {          // a down-sampling axis: wide filters, resize as a tensor first
    HGL_REQUIRE(full, "hgl_heat_pool: internal: no room for the resized heat-map");
    const int blocks = (int)std::min<size_t>(((size_t)E * H * W + 255) / 256, (size_t)sm_count() * 16);
    heat_resize_aa_kernel<<<blocks, 256, 0, st>>>(heat, E, hh, hw, H, W, full);
    int rc = launch_status("hgl_heat_pool(resize)");
    if (rc != HGL_OK) return rc;
    heat = full; lr = false;
  }
  const size_t w128 = (size_t)((W + 127) & ~127);
  const int nr_max = lr ? (int)((double)kPrefRows * hh / H) + 4 : 0;       // raw rows behind the kPrefRows output rows of a CTA
  const size_t smem = w128 * 4 * (1 + (size_t)nr_max);
  HGL_REQUIRE(smem <= 200 * 1024, "hgl_heat_pool: W=%d too wide", W);
  auto kern = lr ? heat_prefix_kernel<true> : heat_prefix_kernel<false>;
  int rc0 = ensure_dyn_smem(reinterpret_cast<const void*>(kern), smem, "hgl_heat_pool(prefix)");
  if (rc0 != HGL_OK) return rc0;
  kern<<<dim3(ceil_div(H, kPrefRows), E), kPrefWarps * 32, smem, st>>>(heat, dirflag, H, W, ws.Wp, hh, hw, nr_max, ws.cr, ws.rp, ws.rowstat);
  int rc = launch_status("hgl_heat_pool(prefix)");
  if (rc != HGL_OK) return rc;
  heat_consts_kernel<<<E, 128, 0, st>>>(ws.rowstat, ws.rp, H, W, ws.Wp, ws.consts);
  return launch_status("hgl_heat_pool(consts)");
}

}  // namespace hgl

extern "C" int hgl_dir_mask(int dirflag, int H, int W, float* out, void* stream) {
  using namespace hgl;
  HGL_REQUIRE(out, "hgl_dir_mask: null pointer");
  HGL_REQUIRE(H >= 1 && W >= 1, "hgl_dir_mask: bad shape");
  const int blocks = (int)std::min<size_t>(((size_t)H * W + 255) / 256, (size_t)sm_count() * 8);
  dir_mask_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(dirflag, H, W, out);
  return launch_status("hgl_dir_mask");
}

extern "C" int64_t hgl_mask_grid_workspace_bytes(int M, int g) {
  if (M < 0 || g < 1 || g > hgl::kMaxG) return -1;
  return (int64_t)hgl::rows_carve(nullptr, M, g, 0, 0).bytes + 256;
}

extern "C" int hgl_mask_grid(const uint32_t* bits, int M, int H, int W, int g, int antialias, float* grid, int32_t* area,
                             void* workspace, void* stream) {
  using namespace hgl;
  if (M == 0) return HGL_OK;
  HGL_REQUIRE(bits && grid, "hgl_mask_grid: null pointer");
  HGL_REQUIRE(M >= 0 && H >= 1 && W >= 1 && g >= 1 && g <= kMaxG, "hgl_mask_grid: bad shape M=%d H=%d W=%d g=%d", M, H, W, g);
  cudaStream_t st = (cudaStream_t)stream;
  const int WW = (W + 31) >> 5;
  if (!antialias) {
    const int total = M * g * g;
    mask_grid_noaa_kernel<<<ceil_div(total, 256), 256, 0, st>>>(bits, M, H, W, g, grid);
    int rc = launch_status("hgl_mask_grid(noaa)");
    if (rc != HGL_OK) return rc;
    if (area) {
      mask_area_kernel<<<std::min(M, sm_count() * 8), 256, 0, st>>>(bits, M, (size_t)H * WW, area);
      return launch_status("hgl_mask_grid(area)");
    }
    return HGL_OK;
  }
  HGL_REQUIRE(H >= g && W >= g, "hgl_mask_grid: antialiased path is a down-sampler (H=%d W=%d g=%d)", H, W, g);
  HGL_REQUIRE(workspace && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "hgl_mask_grid: 16-byte aligned workspace required");
  RowsParams p = {};
  p.bits = bits; p.M = M; p.H = H; p.W = W; p.WW = WW;
  p.g = g; p.maxky = aa_maxk(H, g); p.maxkx = aa_maxk(W, g); p.grid = grid; p.area = area;
  void* scratch = reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
  return launch_rows(p, true, false, scratch, st);
}

extern "C" int64_t hgl_grid_heat_pool_workspace_bytes(int B, int M, int E, int H, int W, int g, int max_n) {
  using namespace hgl;
  if (B < 1 || M < 0 || E < 0 || H < 1 || W < 1 || max_n < 0 || g < 0 || g > kMaxG) return -1;
  return (int64_t)heat_carve(nullptr, E, H, W).bytes + (int64_t)rows_carve(nullptr, M, g, E, max_n).bytes + 256;
}

extern "C" int64_t hgl_heat_pool_workspace_bytes(int B, int M, int E, int H, int W, int max_n) {
  return hgl_grid_heat_pool_workspace_bytes(B, M, E, H, W, 0, max_n);
}

// bytes of the frame-sized heat-maps that the raw-map entry point must materialise (only when an axis is down-sampled)
static size_t lr_full_bytes(int E, int H, int W, int hh, int hw) {
  if (hh <= 0 || (hgl::aa_maxk(hh, H) <= 3 && hgl::aa_maxk(hw, W) <= 3)) return 0;
  return ((size_t)E * H * W * 4 + 255) & ~size_t(255);
}

// workspace carve-up shared by the table pass and the mask pass (hh > 0: heat is the raw map [E,hh,hw])
static int hgl_heat_carve(void* workspace, int E, int H, int W, int hh, int hw, hgl::HeatWs* ws, float** full, void** scratch) {
  using namespace hgl;
  HGL_REQUIRE(workspace, "hgl_heat_pool: null pointer");
  HGL_REQUIRE(E >= 0 && H >= 1 && W >= 1, "hgl_heat_pool: bad shape");
  HGL_REQUIRE(hh >= 0 && hw >= 0 && hw <= 65535 && (hh > 0) == (hw > 0), "hgl_heat_pool: bad raw heat-map shape %dx%d", hh, hw);
  HGL_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "hgl_heat_pool: workspace must be 16-byte aligned");
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
  *ws = heat_carve(base, E, H, W);
  const size_t fb = lr_full_bytes(E, H, W, hh, hw);
  *full = fb ? reinterpret_cast<float*>(base + ws->bytes) : nullptr;
  *scratch = base + ws->bytes + fb;
  return HGL_OK;
}

// the half of the pooling entry points that needs the heat-maps only (not the masks): prefix tables + per-expression constants
extern "C" int hgl_heat_tables(const float* heat, int hh, int hw, const int32_t* dirflag, int E, int H, int W, void* workspace, void* stream) {
  using namespace hgl;
  if (E == 0) return HGL_OK;
  HGL_REQUIRE(heat && dirflag, "hgl_heat_tables: null pointer");
  HeatWs ws;
  float* full = nullptr;
  void* scratch = nullptr;
  int rc = hgl_heat_carve(workspace, E, H, W, hh, hw, &ws, &full, &scratch);
  if (rc != HGL_OK) return rc;
  return launch_heat_tables(heat, hh, hw, full, dirflag, E, H, W, ws, (cudaStream_t)stream);
}

// argument checks of the mask pass + zeroing of its output rows; tables = 1: also build the tables here (one-call entry points)
static int hgl_heat_common(const float* heat, int hh, int hw, const int32_t* expr_off, const int32_t* dirflag, const float* black,
                           const uint32_t* bits, const int32_t* mask_off, int B, int M, int E, int H, int W, int max_n, float* score_gem,
                           void* workspace, hgl::HeatWs* ws, void** scratch, int tables, cudaStream_t st) {
  using namespace hgl;
  HGL_REQUIRE(black && bits && score_gem && workspace, "hgl_heat_pool: null pointer");
  HGL_REQUIRE(!tables || (heat && dirflag), "hgl_heat_pool: null pointer");
  HGL_REQUIRE(B >= 1 && M >= 0 && E >= 0 && H >= 1 && W >= 1 && max_n >= 1, "hgl_heat_pool: bad shape");
  HGL_REQUIRE((mask_off && expr_off) || B == 1, "hgl_heat_pool: mask_off/expr_off required when B > 1");
  float* full = nullptr;
  int rc = hgl_heat_carve(workspace, E, H, W, hh, hw, ws, &full, scratch);
  if (rc != HGL_OK) return rc;
  cudaError_t e = cudaMemsetAsync(score_gem, 0, (size_t)E * max_n * 4, st);    // rows of images with fewer than max_n masks
  if (e != cudaSuccess) { set_error("hgl_heat_pool: cudaMemsetAsync: %s", cudaGetErrorString(e)); return HGL_ECUDA; }
  return tables ? launch_heat_tables(heat, hh, hw, full, dirflag, E, H, W, *ws, st) : HGL_OK;
}

static void hgl_fill_heat(hgl::RowsParams& p, const hgl::HeatWs& ws, const float* black, const int32_t* mask_off, const int32_t* expr_off,
                          int B, int E, int max_n, float* score_gem) {
  p.mask_off = mask_off; p.expr_off = expr_off; p.B = B; p.E = E; p.max_n = max_n; p.Wp = ws.Wp;
  p.cr = ws.cr; p.rp = ws.rp; p.consts = ws.consts; p.black = black; p.score_gem = score_gem;
}

extern "C" int hgl_heat_pool(const float* heat, const int32_t* expr_off, const int32_t* dirflag, const float* black,
                             const uint32_t* bits, const int32_t* mask_off, int B, int M, int E, int H, int W,
                             int max_n, float* score_gem, void* workspace, void* stream) {
  using namespace hgl;
  if (E == 0 || M == 0) return HGL_OK;
  cudaStream_t st = (cudaStream_t)stream;
  HeatWs ws;
  void* scratch = nullptr;
  int rc = hgl_heat_common(heat, 0, 0, expr_off, dirflag, black, bits, mask_off, B, M, E, H, W, max_n, score_gem, workspace, &ws, &scratch, 1, st);
  if (rc != HGL_OK) return rc;
  RowsParams p = {};
  p.bits = bits; p.M = M; p.H = H; p.W = W; p.WW = (W + 31) >> 5;
  hgl_fill_heat(p, ws, black, mask_off, expr_off, B, E, max_n, score_gem);
  return launch_rows(p, false, true, scratch, st);
}

extern "C" int hgl_heat_resize_aa(const float* heat_raw, int E, int hh, int hw, int H, int W, float* out, void* stream) {
  using namespace hgl;
  if (E == 0) return HGL_OK;
  HGL_REQUIRE(heat_raw && out, "hgl_heat_resize_aa: null pointer");
  HGL_REQUIRE(E > 0 && hh >= 1 && hw >= 1 && H >= 1 && W >= 1, "hgl_heat_resize_aa: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  if (aa_maxk(hh, H) <= 3 && aa_maxk(hw, W) <= 3 && E <= 65535) {
    // both axes up-sampled (the GEM case, 28x37 -> frame): a CTA runs the horizontal pass for the few raw rows behind its 32
    // output rows once (shared memory), every pixel is then three shared-memory reads and the vertical taps
    const size_t w128 = (size_t)((W + 127) & ~127);
    const int nr_max = (int)((double)kPrefRows * hh / H) + 4;
    const size_t smem = w128 * 4 * (1 + (size_t)nr_max);
    if (smem <= 200 * 1024) {
      auto kern = heat_prefix_kernel<true, true>;
      const int rc0 = ensure_dyn_smem(reinterpret_cast<const void*>(kern), smem, "hgl_heat_resize_aa");
      if (rc0 != HGL_OK) return rc0;
      kern<<<dim3(ceil_div(H, kPrefRows), E), kPrefWarps * 32, smem, st>>>(heat_raw, nullptr, H, W, W, hh, hw, nr_max, out, nullptr, nullptr);
      return launch_status("hgl_heat_resize_aa");
    }
  }
  const int blocks = (int)std::min<size_t>(((size_t)E * H * W + 255) / 256, (size_t)sm_count() * 16);
  heat_resize_aa_kernel<<<blocks, 256, 0, st>>>(heat_raw, E, hh, hw, H, W, out);
  return launch_status("hgl_heat_resize_aa");
}

extern "C" int64_t hgl_grid_heat_pool_raw_workspace_bytes(int B, int M, int E, int H, int W, int g, int max_n, int hh, int hw) {
  const int64_t base = hgl_grid_heat_pool_workspace_bytes(B, M, E, H, W, g, max_n);
  if (base < 0 || hh < 1 || hw < 1) return -1;
  return base + (int64_t)lr_full_bytes(E, H, W, hh, hw);
}

// the body of hgl_grid_heat_pool / hgl_grid_heat_pool_raw / hgl_grid_heat_pool_rows (hh == 0: heat is frame-sized;
// tables == 0: hgl_heat_tables already ran on this workspace)
static int grid_heat_pool_impl(const uint32_t* bits, const int32_t* mask_off, int B, int M, int H, int W, int g, float* grid, int32_t* area,
                               const float* heat, int hh, int hw, const int32_t* expr_off, const int32_t* dirflag, const float* black, int E,
                               int max_n, float* score_gem, void* workspace, int tables, void* stream) {
  using namespace hgl;
  if (M == 0) return HGL_OK;
  if (E == 0) return hgl_mask_grid(bits, M, H, W, g, 1, grid, area, workspace, stream);
  HGL_REQUIRE(grid, "hgl_grid_heat_pool: null pointer");
  HGL_REQUIRE(g >= 1 && g <= kMaxG && H >= g && W >= g, "hgl_grid_heat_pool: bad grid side g=%d for %dx%d", g, H, W);
  cudaStream_t st = (cudaStream_t)stream;
  HeatWs ws;
  void* scratch = nullptr;
  int rc = hgl_heat_common(heat, hh, hw, expr_off, dirflag, black, bits, mask_off, B, M, E, H, W, max_n, score_gem, workspace, &ws, &scratch,
                           tables, st);
  if (rc != HGL_OK) return rc;
  RowsParams p = {};
  p.bits = bits; p.M = M; p.H = H; p.W = W; p.WW = (W + 31) >> 5;
  p.g = g; p.maxky = aa_maxk(H, g); p.maxkx = aa_maxk(W, g); p.grid = grid; p.area = area;
  hgl_fill_heat(p, ws, black, mask_off, expr_off, B, E, max_n, score_gem);
  return launch_rows(p, true, true, scratch, st);
}

// the mask half after hgl_heat_tables(heat, hh, hw, ...) on the same workspace (hh = hw = 0 for frame-sized maps)
extern "C" int hgl_grid_heat_pool_rows(const uint32_t* bits, const int32_t* mask_off, int B, int M, int H, int W, int g,
                                       float* grid, int32_t* area, int hh, int hw, const int32_t* expr_off, const float* black, int E,
                                       int max_n, float* score_gem, void* workspace, void* stream) {
  return grid_heat_pool_impl(bits, mask_off, B, M, H, W, g, grid, area, nullptr, hh, hw, expr_off, nullptr, black, E, max_n, score_gem, workspace,
                             0, stream);
}

extern "C" int hgl_grid_heat_pool(const uint32_t* bits, const int32_t* mask_off, int B, int M, int H, int W, int g,
                                  float* grid, int32_t* area,
                                  const float* heat, const int32_t* expr_off, const int32_t* dirflag, const float* black, int E,
                                  int max_n, float* score_gem, void* workspace, void* stream) {
  return grid_heat_pool_impl(bits, mask_off, B, M, H, W, g, grid, area, heat, 0, 0, expr_off, dirflag, black, E, max_n, score_gem, workspace, 1, stream);
}

extern "C" int hgl_grid_heat_pool_raw(const uint32_t* bits, const int32_t* mask_off, int B, int M, int H, int W, int g,
                                      float* grid, int32_t* area,
                                      const float* heat_raw, int hh, int hw, const int32_t* expr_off, const int32_t* dirflag,
                                      const float* black, int E, int max_n, float* score_gem, void* workspace, void* stream) {
  using namespace hgl;
  HGL_REQUIRE(hh >= 1 && hw >= 1, "hgl_grid_heat_pool_raw: bad raw heat-map shape %dx%d", hh, hw);
  return grid_heat_pool_impl(bits, mask_off, B, M, H, W, g, grid, area, heat_raw, hh, hw, expr_off, dirflag, black, E, max_n, score_gem, workspace,
                             1, stream);
}
