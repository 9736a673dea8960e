// (a2) mask -> patch grid, (a3) CLS-row attention mask, (a4) token masking + stream mix.
//
// (a2) replaces TF.resize(pred_masks.float(), (g,g)) (model/backbone.py:160).  Two semantics exist in the wild
// (SURVEY.md Appendix B-1): torchvision >= 0.17 antialiases tensors (ATen _upsample_bilinear2d_aa: separable
// triangle filter of half-width `scale`), the reference's pinned 0.15.2 does not (4-tap bilinear).  Both are here.
//
// Antialiased kernel (HBM-bound, reads every mask byte exactly once, 128-byte coalesced per warp per row):
//   thread = 4 adjacent columns, walks down the rows; every source row feeds at most 3 vertical bins, so the
//   vertical pass is 3 predicated FMAs per pixel into registers; when a bin retires its row V[gy][0..W) goes
//   to shared memory and one warp per output column does the short horizontal dot product.  The per-mask
//   pixel count (area) falls out of the same pass.  Weights are built in shared memory with the exact float32
//   (and float64-intermediate) expression sequence of ATen so the zero pattern of the grid is identical.
#include "hgl_common.cuh"

namespace hgl {

constexpr int kMaxG = 32;        // grid side limit (14 for ViT-B/16, 24 for ViT-L/14@336)

// ATen _compute_indices_min_size_weights_aa for the triangle (bilinear) filter; one thread per output index.
__device__ void aa_fill(int i, int in_size, int out_size, int maxk, int* xmin_out, int* xsize_out, float* w) {
  const float scale = __fdiv_rn((float)in_size, (float)out_size);
  float support, invscale;
  if (scale >= 1.f) { support = scale; invscale = __fdiv_rn(1.f, scale); } else { support = 1.f; invscale = 1.f; }
  const float center = (float)((double)scale * ((double)i + 0.5));
  int xmin = (int)((double)__fsub_rn(center, support) + 0.5);
  xmin = max(xmin, 0);
  int xsize = min((int)((double)__fadd_rn(center, support) + 0.5), in_size) - xmin;
  xsize = max(min(xsize, maxk), 0);
  float total = 0.f;
  for (int j = 0; j < xsize; ++j) {
    float t = (float)(((double)__fsub_rn((float)(j + xmin), center) + 0.5) * (double)invscale);
    t = fabsf(t);
    const float wt = (t < 1.f) ? __fsub_rn(1.f, t) : 0.f;
    w[j] = wt;
    total = __fadd_rn(total, wt);
  }
  if (total != 0.f)
    for (int j = 0; j < xsize; ++j) w[j] = __fdiv_rn(w[j], total);
  for (int j = xsize; j < maxk; ++j) w[j] = 0.f;
  *xmin_out = xmin; *xsize_out = xsize;
}

static int aa_maxk(int in_size, int out_size) {
  const float scale = (float)in_size / (float)out_size;
  const float support = scale >= 1.f ? scale : 1.f;
  return (int)ceilf(support) * 2 + 1;
}

// Antialiased down-sample of PACKED masks.  Persistent CTAs (4 warps) loop over masks; warp w owns the row band
// [w*H/4, (w+1)*H/4).  Lane = horizontal bin gx.  Per row: lane l fetches word l of the bit row (one coalesced load),
// an all-zero row is skipped after one ballot; otherwise every lane walks the <= 4 words its filter support overlaps:
// an all-ones overlap adds a precomputed partial sum, anything else iterates the set bits.  The horizontal result is
// folded into <= 3 running vertical bins (register accumulators, retired in row order into shared memory).
struct GridTables {
  int* ymin; int* ysize; int* xmin; int* xsize;   // [kMaxG]
  float* wy; float* wx;                            // [g][maxky], [g][maxkx]
  int* rowbin; float* roww;                        // [H], [H][3]
  float* fullsum;                                  // [g][kMaxWordsPerBin]
};
constexpr int kMaxWordsPerBin = 8;
constexpr int kGridWarps = 4;

__global__ void __launch_bounds__(kGridWarps * 32) mask_grid_aa_bits_kernel(const uint32_t* __restrict__ bits, int M, int H, int W, int g,
                                                                            int maxky, int maxkx, int psum_off, float* __restrict__ grid,
                                                                            int32_t* __restrict__ area) {
  extern __shared__ __align__(16) uint8_t smem[];
  int* ymin = reinterpret_cast<int*>(smem);
  int* ysize = ymin + kMaxG;
  int* xmin = ysize + kMaxG;
  int* xsize = xmin + kMaxG;
  float* wy = reinterpret_cast<float*>(xsize + kMaxG);
  float* wx = wy + g * maxky;
  int* rowbin = reinterpret_cast<int*>(wx + g * maxkx);     // first vertical bin fed by source row y
  float* roww = reinterpret_cast<float*>(rowbin + H);       // [H][3] weights into bins rowbin[y]+{0,1,2}
  double* psum = reinterpret_cast<double*>(smem + psum_off);  // [g][maxkx+1] prefix sums of wx (double: tiny edge taps survive)
  float* part = reinterpret_cast<float*>(psum + g * (maxkx + 1));   // [kGridWarps][g][g] per-warp partial grids
  int* ared = reinterpret_cast<int*>(part + kGridWarps * g * g);    // [kGridWarps]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int WW = (W + 31) >> 5;
  if (tid < g) aa_fill(tid, H, g, maxky, &ymin[tid], &ysize[tid], wy + tid * maxky);
  else if (tid >= 32 && tid < 32 + g) { const int i = tid - 32; aa_fill(i, W, g, maxkx, &xmin[i], &xsize[i], wx + i * maxkx); }
  __syncthreads();
  for (int y = tid; y < H; y += blockDim.x) {
    int first = -1;
    float w3[3] = {0.f, 0.f, 0.f};
    for (int bb = 0; bb < g; ++bb) {
      const int k = y - ymin[bb];
      if (k >= 0 && k < ysize[bb]) {
        if (first < 0) first = bb;
        const int slot = bb - first;
        if (slot < 3) w3[slot] = wy[bb * maxky + k];
      }
    }
    rowbin[y] = first < 0 ? g : first;
    roww[3 * y + 0] = w3[0]; roww[3 * y + 1] = w3[1]; roww[3 * y + 2] = w3[2];
  }
  if (tid < g) {
    double run = 0.0;
    psum[tid * (maxkx + 1)] = 0.0;
    for (int k = 0; k < maxkx; ++k) { run += (double)wx[tid * maxkx + k]; psum[tid * (maxkx + 1) + k + 1] = run; }
  }
  __syncthreads();

  // per-lane constants of bin gx = lane
  const bool has_bin = lane < g;
  const int x_lo = has_bin ? xmin[lane] : 0, x_hi = has_bin ? xmin[lane] + xsize[lane] : 0;
  const double* myps = psum + (has_bin ? lane : 0) * (maxkx + 1);
  // prefix sum of this bin's taps at absolute pixel x, clamped to the support: P(x) = sum_{x' < x} wx[x']
  auto P = [&](int x) -> double { return myps[min(max(x, x_lo), x_hi) - x_lo]; };

  for (int m = blockIdx.x; m < M; m += gridDim.x) {
    for (int t = tid; t < kGridWarps * g * g; t += blockDim.x) part[t] = 0.f;
    __syncthreads();
    const uint32_t* mb = bits + (size_t)m * H * WW;
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;
    int cur = (warp < H) ? min(rowbin[warp], g) : g;   // first live vertical bin; acc slot s belongs to bin cur+s
    int cnt = 0;
    float* mypart = part + warp * g * g;
    constexpr int kAhead = 4;
    // rows are interleaved across the warps (row y -> warp y % kGridWarps): every warp sees the same share of the mask
    for (int yb = warp; yb < H; yb += kAhead * kGridWarps) {
      uint32_t v0[kAhead], v1[kAhead];
#pragma unroll
      for (int u = 0; u < kAhead; ++u) {
        const int y = yb + u * kGridWarps;
        v0[u] = (y < H && lane < WW) ? __ldg(mb + (size_t)y * WW + lane) : 0u;
        v1[u] = (y < H && lane + 32 < WW) ? __ldg(mb + (size_t)y * WW + lane + 32) : 0u;
      }
#pragma unroll
      for (int u = 0; u < kAhead; ++u) {
        const int y = yb + u * kGridWarps;
        if (y >= H) break;
        const int fb = min(rowbin[y], g);
        while (cur < fb) {   // bin `cur` gets no more rows from this warp (warp-uniform)
          if (has_bin && cur < g) mypart[cur * g + lane] = acc0;
          acc0 = acc1; acc1 = acc2; acc2 = 0.f;
          ++cur;
        }
        cnt += __popc(v0[u]) + __popc(v1[u]);
        // transitions of the row, as bit positions where the value changes (0->1 run start, 1->0 run end); words 0..31
        // live in v0 (lane = word), words 32..63 in v1; pixels beyond W are zero, so every run is closed inside the row
        // or by the virtual word WW (all zero).
        const uint32_t c0 = __shfl_up_sync(0xffffffffu, v0[u], 1) >> 31;               // last bit of the previous word
        const uint32_t top0 = __shfl_sync(0xffffffffu, v0[u], 31) >> 31;              // last bit of word 31
        const uint32_t c1 = __shfl_up_sync(0xffffffffu, v1[u], 1) >> 31;
        const uint32_t t0 = v0[u] ^ ((v0[u] << 1) | (lane == 0 ? 0u : c0));
        const uint32_t t1 = v1[u] ^ ((v1[u] << 1) | (lane == 0 ? top0 : c1));
        uint32_t has0 = __ballot_sync(0xffffffffu, t0 != 0u), has1 = __ballot_sync(0xffffffffu, t1 != 0u);
        if ((has0 | has1) == 0u) continue;                                             // no run in this row
        double hs = 0.0;
        while (has0) {
          const int src = __ffs(has0) - 1; has0 &= has0 - 1;
          uint32_t t = __shfl_sync(0xffffffffu, t0, src);
          const uint32_t w = __shfl_sync(0xffffffffu, v0[u], src);
          while (t) {
            const int bp = __ffs(t) - 1; t &= t - 1;
            const double pv = P(32 * src + bp);
            hs += ((w >> bp) & 1u) ? -pv : pv;                                         // start: -P(s), end: +P(e)
          }
        }
        while (has1) {
          const int src = __ffs(has1) - 1; has1 &= has1 - 1;
          uint32_t t = __shfl_sync(0xffffffffu, t1, src);
          const uint32_t w = __shfl_sync(0xffffffffu, v1[u], src);
          while (t) {
            const int bp = __ffs(t) - 1; t &= t - 1;
            const double pv = P(32 * (32 + src) + bp);
            hs += ((w >> bp) & 1u) ? -pv : pv;
          }
        }
        const float hsum = (float)hs;
        acc0 += roww[3 * y] * hsum; acc1 += roww[3 * y + 1] * hsum; acc2 += roww[3 * y + 2] * hsum;
      }
    }
    // flush the bins still open at the end
    if (has_bin) {
      if (cur < g) mypart[cur * g + lane] = acc0;
      if (cur + 1 < g) mypart[(cur + 1) * g + lane] = acc1;
      if (cur + 2 < g) mypart[(cur + 2) * g + lane] = acc2;
    }
    cnt = warp_sum_i(cnt);
    if (lane == 0) ared[warp] = cnt;
    __syncthreads();
    for (int t = tid; t < g * g; t += blockDim.x) {
      float sum = 0.f;
#pragma unroll
      for (int w = 0; w < kGridWarps; ++w) sum += part[w * g * g + t];
      grid[(size_t)m * g * g + t] = sum;
    }
    if (tid == 0 && area != nullptr) {
      int sum = 0;
      for (int w = 0; w < kGridWarps; ++w) sum += ared[w];
      area[m] = sum;
    }
    __syncthreads();
  }
}

// non-antialiased: 4-tap bilinear sample of the packed mask at the g x g grid positions (ATen upsample_bilinear2d)
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

// (a3) full boolean attention mask, written once, 16 bytes per store
__global__ void attn_mask_kernel(const float* __restrict__ grid, int M, int L, int heads, uint8_t* __restrict__ out) {
  const int L1 = L + 1;
  const size_t per = (size_t)L1 * L1;
  const size_t total = (size_t)M * heads * per;
  const size_t stride = (size_t)gridDim.x * blockDim.x * 16;
  for (size_t o = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 16; o < total; o += stride) {
    uint32_t w[4] = {0, 0, 0, 0};
    const size_t mh = o / per;
    const size_t within = o - mh * per;
    if (within < (size_t)L1 || (within + 15) / per != 0 || true) {
      // only bytes that fall in row 0 of some (m,h) slab can be non-zero
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const size_t oo = o + q;
        if (oo >= total) break;
        const size_t mh2 = oo / per;
        const size_t r = oo - mh2 * per;
        if (r >= 1 && r < (size_t)L1) {
          const int m = (int)(mh2 / heads);
          if (grid[(size_t)m * L + (r - 1)] == 0.f) w[q >> 2] |= 1u << ((q & 3) * 8);
        }
      }
    }
    if (o + 16 <= total) {
      *reinterpret_cast<uint4*>(out + o) = make_uint4(w[0], w[1], w[2], w[3]);
    } else {
      for (int q = 0; o + q < total; ++q) out[o + q] = (w[q >> 2] >> ((q & 3) * 8)) & 0xff;
    }
  }
}

__global__ void attn_bias_kernel(const float* __restrict__ grid, int M, int L, float* __restrict__ bias) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int L1 = L + 1;
  if (idx >= M * L1) return;
  const int m = idx / L1, k = idx - m * L1;
  float v = 0.f;
  if (k >= 1 && grid[(size_t)m * L + (k - 1)] == 0.f) v = -INFINITY;
  bias[idx] = v;
}

// (a4) out[l,m,:] = a * w(l,m) * src[l,m,:] + b * add[l,m,:]   (8 elements = 16 B (bf16) / 2x16 B (f32) per thread-iteration)
template <bool kBF16>
__global__ void __launch_bounds__(256) token_mask_fuse_kernel(const void* __restrict__ src_, const void* __restrict__ add_, const float* __restrict__ grid,
                                                              float a, float b, int L1, int M, int D, int nld, void* __restrict__ out_) {
  const int L = L1 - 1;
  const int dv = D / 8;                       // vectors of 8 elements per token
  const size_t total = (size_t)L1 * M * dv;
  for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < total; v += (size_t)gridDim.x * blockDim.x) {
    const size_t tok = v / dv;                // = l*M + m (LND) or m*L1 + l (NLD)
    int l, m;
    if (nld) { m = (int)(tok / L1); l = (int)(tok - (size_t)m * L1); }
    else { l = (int)(tok / M); m = (int)(tok - (size_t)l * M); }
    float w = a;
    if (grid != nullptr && l > 0) w = a * grid[(size_t)m * L + (l - 1)];
    float x[8], y[8];
    if (kBF16) {
      const uint4 s = reinterpret_cast<const uint4*>(src_)[v];
      const uint32_t sw[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) { x[2 * q] = bf16_bits_to_float(sw[q] & 0xffffu); x[2 * q + 1] = bf16_bits_to_float(sw[q] >> 16); }
      if (add_) {
        const uint4 t = reinterpret_cast<const uint4*>(add_)[v];
        const uint32_t tw[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) { y[2 * q] = bf16_bits_to_float(tw[q] & 0xffffu); y[2 * q + 1] = bf16_bits_to_float(tw[q] >> 16); }
      }
    } else {
      const float4 s0 = reinterpret_cast<const float4*>(src_)[2 * v], s1 = reinterpret_cast<const float4*>(src_)[2 * v + 1];
      x[0] = s0.x; x[1] = s0.y; x[2] = s0.z; x[3] = s0.w; x[4] = s1.x; x[5] = s1.y; x[6] = s1.z; x[7] = s1.w;
      if (add_) {
        const float4 t0 = reinterpret_cast<const float4*>(add_)[2 * v], t1 = reinterpret_cast<const float4*>(add_)[2 * v + 1];
        y[0] = t0.x; y[1] = t0.y; y[2] = t0.z; y[3] = t0.w; y[4] = t1.x; y[5] = t1.y; y[6] = t1.z; y[7] = t1.w;
      }
    }
    float o[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float t = __fmul_rn(x[q], w);
      o[q] = add_ ? __fadd_rn(t, __fmul_rn(y[q], b)) : t;
    }
    if (kBF16) {
      uint4 r;
      r.x = pack_bf16x2(o[0], o[1]); r.y = pack_bf16x2(o[2], o[3]); r.z = pack_bf16x2(o[4], o[5]); r.w = pack_bf16x2(o[6], o[7]);
      reinterpret_cast<uint4*>(out_)[v] = r;
    } else {
      reinterpret_cast<float4*>(out_)[2 * v] = make_float4(o[0], o[1], o[2], o[3]);
      reinterpret_cast<float4*>(out_)[2 * v + 1] = make_float4(o[4], o[5], o[6], o[7]);
    }
  }
}

}  // namespace hgl

extern "C" int hgl_mask_grid(const uint32_t* bits, int M, int H, int W, int g, int antialias, float* grid, int32_t* area, void* stream) {
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
      mask_area_kernel<<<min(M, sm_count() * 8), 256, 0, st>>>(bits, M, (size_t)H * WW, area);
      return launch_status("hgl_mask_grid(area)");
    }
    return HGL_OK;
  }
  HGL_REQUIRE(H >= g && W >= g, "hgl_mask_grid: antialiased path is a down-sampler (H=%d W=%d g=%d)", H, W, g);
  HGL_REQUIRE(WW <= 63, "hgl_mask_grid: W=%d wider than 2016", W);
  const int maxky = aa_maxk(H, g), maxkx = aa_maxk(W, g);
  size_t psum_off = (size_t)4 * kMaxG * 4 + (size_t)g * (maxky + maxkx) * 4 + (size_t)H * 4 + (size_t)3 * H * 4;
  psum_off = (psum_off + 15) & ~size_t(15);
  const size_t smem = psum_off + (size_t)g * (maxkx + 1) * 8 + (size_t)kGridWarps * g * g * 4 + kGridWarps * 4 + 64;
  HGL_REQUIRE(smem <= 200 * 1024, "hgl_mask_grid: frame %dx%d too large for the tap tables (%zu B)", H, W, smem);
  cudaError_t e = cudaFuncSetAttribute(mask_grid_aa_bits_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("hgl_mask_grid: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return HGL_ECUDA; }
  const int ctas = min(M, sm_count() * 8);
  mask_grid_aa_bits_kernel<<<ctas, kGridWarps * 32, smem, st>>>(bits, M, H, W, g, maxky, maxkx, (int)psum_off, grid, area);
  return launch_status("hgl_mask_grid(aa)");
}

extern "C" int hgl_attn_mask(const float* grid, int M, int L, int heads, uint8_t* out, void* stream) {
  using namespace hgl;
  HGL_REQUIRE(grid && out, "hgl_attn_mask: null pointer");
  HGL_REQUIRE(M >= 0 && L >= 1 && heads >= 1, "hgl_attn_mask: bad shape");
  if (M == 0) return HGL_OK;
  HGL_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "hgl_attn_mask: out must be 16-byte aligned");
  const size_t total = (size_t)M * heads * (L + 1) * (L + 1);
  const int blocks = (int)std::min<size_t>(ceil_div64((int64_t)total, 256 * 16), (size_t)sm_count() * 16);
  attn_mask_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(grid, M, L, heads, out);
  return launch_status("hgl_attn_mask");
}

extern "C" int hgl_attn_bias(const float* grid, int M, int L, float* bias, void* stream) {
  using namespace hgl;
  HGL_REQUIRE(grid && bias, "hgl_attn_bias: null pointer");
  HGL_REQUIRE(M >= 0 && L >= 1, "hgl_attn_bias: bad shape");
  if (M == 0) return HGL_OK;
  attn_bias_kernel<<<ceil_div(M * (L + 1), 256), 256, 0, (cudaStream_t)stream>>>(grid, M, L, bias);
  return launch_status("hgl_attn_bias");
}

extern "C" int hgl_token_mask_fuse(const void* src, const void* add, const float* grid, float a, float b, int L1, int M, int D,
                                   int dtype, int layout, void* out, void* stream) {
  using namespace hgl;
  HGL_REQUIRE(src && out, "hgl_token_mask_fuse: null pointer");
  HGL_REQUIRE(L1 >= 2 && M >= 0 && D >= 8 && D % 8 == 0, "hgl_token_mask_fuse: bad shape L1=%d M=%d D=%d (D %% 8)", L1, M, D);
  HGL_REQUIRE(dtype == HGL_F32 || dtype == HGL_BF16, "hgl_token_mask_fuse: dtype %d", dtype);
  HGL_REQUIRE(layout == HGL_LND || layout == HGL_NLD, "hgl_token_mask_fuse: layout %d", layout);
  HGL_REQUIRE(((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(add)) & 15) == 0,
              "hgl_token_mask_fuse: tensors must be 16-byte aligned");
  if (M == 0) return HGL_OK;
  const size_t total = (size_t)L1 * M * (D / 8);
  const int blocks = (int)std::min<size_t>(ceil_div64((int64_t)total, 256), (size_t)sm_count() * 16);
  if (dtype == HGL_BF16)
    token_mask_fuse_kernel<true><<<blocks, 256, 0, (cudaStream_t)stream>>>(src, add, grid, a, b, L1, M, D, layout, out);
  else
    token_mask_fuse_kernel<false><<<blocks, 256, 0, (cudaStream_t)stream>>>(src, add, grid, a, b, L1, M, D, layout, out);
  return launch_status("hgl_token_mask_fuse");
}
