// (a10)+(a11) heat-map conditioning and per-mask pooling  --  Hybridgl_main.py:204-223, utils.py:135-161.
//
//   A'  = (A - min A) / (max A - min A) * ramp(dirflag)            :204-207
//   A'' = A' / mean(A')                                            :209
//   score_gem[n] = (2-black) * sum(A''*m_n)/|m_n|  -  black * sum(A''*(1-m_n)) / |1-m_n|     :218-223
// The reference loops over masks in Python (~8 full-frame kernels + one D2H sync per mask).  Here (SURVEY.md
// Appendix A-2): S_in[e,n] = masks[n,:] . A''_e, area[n] = |m_n|, S_tot[e] = sum A''_e, then a closed form.
//
// B200 design: works on the PACKED masks (hgl_pack_masks), 8x fewer bytes than the byte masks and mostly zero words.
//   pass 1  heat_stats : per expression min / max / sum(A*ramp)            (E*H*W*4 bytes, tiny)
//   pass 2  heat_pool  : a CTA owns (image, band of RB rows): the conditioned heat values of up to EB expressions of
//                        that band sit in shared memory; each warp streams masks: 32 bit-words per coalesced load, one
//                        ballot finds the non-zero words, and for each of those all 32 lanes add their pixel's heat
//                        value when their bit is set (conflict-free LDS, lane-private accumulators, ONE shuffle tree per
//                        (mask, band)).  Zero words -- most of a mask -- cost nothing.  Partials are written per band and
//                        combined in a fixed order (deterministic, no float atomics).
//   pass 3  finalize   : sum the band partials in order, closed form -> score_gem[E, max_n]
#include "hgl_common.cuh"

namespace hgl {

constexpr int kStatChunks = 32;
constexpr int kPoolWarps = 8;
constexpr int kPoolRows = 4;      // rows per band

__device__ __forceinline__ float linspace_at(float a, float b, int n, int i) {  // ATen linspace (float): both-ends evaluation
  if (n <= 1) return a;
  const float step = __fdiv_rn(__fsub_rn(b, a), (float)(n - 1));
  return (i < n / 2) ? __fmaf_rn(step, (float)i, a) : __fmaf_rn(-step, (float)(n - 1 - i), b);
}
// gen_dir_mask utils.py:135-161 (up/down/none are all-ones: the vertical ramps are commented out in the reference)
__device__ __forceinline__ float ramp_at(int dirflag, int x, int W) {
  if (dirflag == HGL_DIR_LEFT) return linspace_at(1.f, 0.f, W, x);
  if (dirflag == HGL_DIR_RIGHT) return linspace_at(0.f, 1.f, W, x);
  if (dirflag == HGL_DIR_MIDDLE) {
    const int h = W / 2;
    return (x < h) ? linspace_at(0.f, 1.f, h, x) : linspace_at(1.f, 0.f, W - h, x - h);
  }
  return 1.f;
}

// gen_dir_mask utils.py:135-161 as a tensor (the drop-in form; the pooling kernels evaluate the ramp on the fly)
__global__ void dir_mask_kernel(int dirflag, int H, int W, float* __restrict__ out) {
  const size_t total = (size_t)H * W;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
    out[i] = ramp_at(dirflag, (int)(i % W), W);
}

struct HeatWs {            // workspace carve-up (all offsets in bytes, 16-aligned)
  float* stats;            // [E][kStatChunks][4]  min, max, sum(A*ramp), sum(ramp)
  float* part_sin;         // [tiles][E][max_n]
  float* part_tot;         // [tiles][E]
  int32_t* part_area;      // [tiles][M]
  int tiles;
  size_t bytes;
};

static HeatWs carve(void* ws, int M, int E, int H, int W, int max_n) {
  HeatWs h;
  const int tiles = ceil_div(H, kPoolRows);   // bands
  size_t off = 0;
  auto take = [&](size_t n) { size_t o = off; off += (n + 15) & ~size_t(15); return o; };
  uint8_t* base = reinterpret_cast<uint8_t*>(ws);
  h.stats = reinterpret_cast<float*>(base + take((size_t)E * kStatChunks * 4 * 4));
  h.part_sin = reinterpret_cast<float*>(base + take((size_t)tiles * E * max_n * 4));
  h.part_tot = reinterpret_cast<float*>(base + take((size_t)tiles * E * 4));
  h.part_area = reinterpret_cast<int32_t*>(base + take((size_t)tiles * M * 4));
  h.tiles = tiles;
  h.bytes = off;
  return h;
}

__global__ void __launch_bounds__(256) heat_stats_kernel(const float* __restrict__ heat, const int32_t* __restrict__ dirflag, int H, int W,
                                                         float* __restrict__ stats) {
  // chunk = a band of rows; a thread walks columns x = tid, tid+256, ... so the ramp is evaluated once per column
  const int e = blockIdx.y, ch = blockIdx.x;
  const int rows_per = (H + kStatChunks - 1) / kStatChunks;
  const int r_lo = min(H, ch * rows_per), r_hi = min(H, r_lo + rows_per);   // trailing chunks may be empty
  const float* A = heat + (size_t)e * H * W;
  const int dir = dirflag[e];
  float mn = INFINITY, mx = -INFINITY, s1 = 0.f, s0 = 0.f;
  for (int x = threadIdx.x; x < W; x += blockDim.x) {
    const float rp = ramp_at(dir, x, W);
    float cs = 0.f;
    for (int r = r_lo; r < r_hi; ++r) {
      const float a = A[(size_t)r * W + x];
      mn = fminf(mn, a); mx = fmaxf(mx, a);
      cs += a;
    }
    s1 += cs * rp; s0 += rp * (float)(r_hi - r_lo);
  }
  __shared__ float red[4][8];
  mn = warp_min(mn); mx = warp_max(mx); s1 = warp_sum(s1); s0 = warp_sum(s0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { red[0][warp] = mn; red[1][warp] = mx; red[2][warp] = s1; red[3][warp] = s0; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) { mn = fminf(mn, red[0][w]); mx = fmaxf(mx, red[1][w]); s1 += red[2][w]; s0 += red[3][w]; }
    float* o = stats + ((size_t)e * kStatChunks + ch) * 4;
    o[0] = mn; o[1] = mx; o[2] = s1; o[3] = s0;
  }
}

// per-expression conditioning constants from the chunk statistics: A'' = (A - mn) * kk * ramp
__device__ __forceinline__ void heat_consts(const float* stats, int e, size_t HW, float& mn, float& kk) {
  float lo = INFINITY, hi = -INFINITY;
  double s1 = 0.0, s0 = 0.0;
  for (int c = 0; c < kStatChunks; ++c) {
    const float* o = stats + ((size_t)e * kStatChunks + c) * 4;
    lo = fminf(lo, o[0]); hi = fmaxf(hi, o[1]); s1 += (double)o[2]; s0 += (double)o[3];
  }
  const double range = (double)hi - (double)lo;
  const double mean = (s1 - (double)lo * s0) / range / (double)HW;   // mean of A'
  mn = lo;
  kk = (float)(1.0 / (range * mean));
}

template <int EB>
__global__ void __launch_bounds__(kPoolWarps * 32) heat_pool_kernel(const float* __restrict__ heat, const int32_t* __restrict__ expr_off,
                                                                    const int32_t* __restrict__ dirflag, const uint32_t* __restrict__ bits,
                                                                    const int32_t* __restrict__ mask_off, int M, int E, int H, int W, int max_n,
                                                                    HeatWs ws) {
  extern __shared__ __align__(16) float smf[];
  // layout: cond [EB][bpx] conditioned heat of the band (zero beyond W) | wsum [EB][bwp] sum over each 32-pixel word | ramp [EB][W32]
  const size_t HW = (size_t)H * W;
  const int WW = (W + 31) >> 5, W32 = WW * 32;
  const int band = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int y0 = band * kPoolRows, rows = min(kPoolRows, H - y0);
  const int bw = rows * WW;                          // words of one mask inside the band
  const int bwp = kPoolRows * WW;                    // padded
  const int bpx = bwp * 32;                          // padded pixels per expression plane
  float* cond = smf;
  float* wsum = cond + EB * bpx;
  float* ramp = wsum + EB * bwp;
  __shared__ float tot_s[kPoolWarps][8];
  int n_lo = 0, n_hi = M, e_lo = 0, e_hi = E;
  if (mask_off) { n_lo = mask_off[b]; n_hi = mask_off[b + 1]; }
  if (expr_off) { e_lo = expr_off[b]; e_hi = expr_off[b + 1]; }
  const int n_cnt = n_hi - n_lo;
  constexpr int kMaxLoads = 4;                       // band words per lane (bw <= 128)
  const int nloads = (bw + 31) >> 5;

  for (int eg = e_lo; eg < e_hi; eg += EB) {       // groups of EB expressions (one pass over the packed masks per group)
    const int ne = min(EB, e_hi - eg);
    __syncthreads();
    for (int t = tid; t < EB * W32; t += blockDim.x) {
      const int j = t / W32, x = t - j * W32;
      ramp[t] = (j < ne && x < W) ? ramp_at(dirflag[eg + j], x, W) : 0.f;
    }
    __syncthreads();
    // conditioned heat of the band: A'' = (A - mn) * kk * ramp(x); one warp-task per (expression, word)
#pragma unroll
    for (int j = 0; j < EB; ++j) {
      float mn = 0.f, kk = 0.f;
      if (j < ne) heat_consts(ws.stats, eg + j, HW, mn, kk);
      float tot = 0.f;
      for (int wd = warp; wd < bwp; wd += kPoolWarps) {
        const int r = wd / WW, x = (wd - r * WW) * 32 + lane;
        float v = 0.f;
        if (j < ne && r < rows && x < W) v = (heat[(size_t)(eg + j) * HW + (size_t)(y0 + r) * W + x] - mn) * kk * ramp[j * W32 + x];
        cond[j * bpx + wd * 32 + lane] = v;
        const float s = warp_sum(v);
        if (lane == 0) wsum[j * bwp + wd] = s;
        tot += s;
      }
      if (lane == 0) tot_s[warp][j] = tot;
    }
    __syncthreads();
    if (tid < ne) {
      float s = 0.f;
      for (int w = 0; w < kPoolWarps; ++w) s += tot_s[w][tid];
      ws.part_tot[(size_t)band * E + eg + tid] = s;
    }

    // stream the masks of this image: warp per mask, the next mask's words are in flight while this one is reduced
    uint32_t nxt[kMaxLoads];
    auto fetch = [&](int k) {
      const uint32_t* mb = bits + ((size_t)(n_lo + k) * H + y0) * WW;
#pragma unroll
      for (int u = 0; u < kMaxLoads; ++u) nxt[u] = (u < nloads && u * 32 + lane < bw) ? __ldg(mb + u * 32 + lane) : 0u;
    };
    if (warp < n_cnt) fetch(warp);
    for (int k = warp; k < n_cnt; k += kPoolWarps) {
      uint32_t cur[kMaxLoads];
#pragma unroll
      for (int u = 0; u < kMaxLoads; ++u) cur[u] = nxt[u];
      if (k + kPoolWarps < n_cnt) fetch(k + kPoolWarps);
      float acc[EB];
#pragma unroll
      for (int j = 0; j < EB; ++j) acc[j] = 0.f;
      int cnt = 0;
      if (__ballot_sync(0xffffffffu, (cur[0] | cur[1] | cur[2] | cur[3]) != 0u) == 0u) {   // the mask does not touch this band
        if (lane == 0) {
#pragma unroll
          for (int j = 0; j < EB; ++j)
            if (j < ne) ws.part_sin[((size_t)band * E + eg + j) * max_n + k] = 0.f;
          if (eg == e_lo) ws.part_area[(size_t)band * M + n_lo + k] = 0;
        }
        continue;
      }
#pragma unroll
      for (int u = 0; u < kMaxLoads; ++u) {
        if (u < nloads) {
          const uint32_t mine = cur[u];
          cnt += __popc(mine);
          if (mine == 0xffffffffu) {                      // whole word inside the mask: its precomputed sum
#pragma unroll
            for (int j = 0; j < EB; ++j) acc[j] += wsum[j * bwp + u * 32 + lane];
          }
          uint32_t part = __ballot_sync(0xffffffffu, mine != 0u && mine != 0xffffffffu);
          while (part) {                                  // words cut by the outline: all lanes, one pixel each
            const int src = __ffs(part) - 1;
            part &= part - 1;
            const uint32_t word = __shfl_sync(0xffffffffu, mine, src);
            if ((word >> lane) & 1u) {
              const int o = (u * 32 + src) * 32 + lane;
#pragma unroll
              for (int j = 0; j < EB; ++j) acc[j] += cond[j * bpx + o];
            }
          }
        }
      }
#pragma unroll
      for (int j = 0; j < EB; ++j) acc[j] = warp_sum(acc[j]);
      cnt = warp_sum_i(cnt);
      if (lane == 0) {
#pragma unroll
        for (int j = 0; j < EB; ++j)
          if (j < ne) ws.part_sin[((size_t)band * E + eg + j) * max_n + k] = acc[j];
        if (eg == e_lo) ws.part_area[(size_t)band * M + n_lo + k] = cnt;
      }
    }
  }
}

__global__ void heat_finalize_kernel(const int32_t* __restrict__ expr_off, const int32_t* __restrict__ mask_off, const float* __restrict__ black,
                                     int B, int M, int E, int H, int W, int max_n, HeatWs ws, float* __restrict__ score_gem) {
  const int e = blockIdx.y;
  int b = 0;
  if (expr_off) { while (b + 1 < B && expr_off[b + 1] <= e) ++b; }
  int n_lo = 0, n_hi = M;
  if (mask_off) { n_lo = mask_off[b]; n_hi = mask_off[b + 1]; }
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= max_n) return;
  float out = 0.f;
  if (n < n_hi - n_lo) {
    float s_in = 0.f, s_tot = 0.f;
    int area = 0;
    for (int t = 0; t < ws.tiles; ++t) {
      s_in += ws.part_sin[((size_t)t * E + e) * max_n + n];
      s_tot += ws.part_tot[(size_t)t * E + e];
      area += ws.part_area[(size_t)t * M + n_lo + n];
    }
    const float bl = black[e];
    const float hw = (float)((size_t)H * W);
    out = (2.f - bl) * s_in / (float)area - bl * (s_tot - s_in) / (hw - (float)area);
  }
  score_gem[(size_t)e * max_n + n] = out;
}

template <int EB>
static int launch_pool(const float* heat, const int32_t* expr_off, const int32_t* dirflag, const uint32_t* bits, const int32_t* mask_off,
                       int B, int M, int E, int H, int W, int max_n, HeatWs ws, cudaStream_t st) {
  const int WW = (W + 31) >> 5;
  const size_t smem = ((size_t)EB * kPoolRows * WW * 32 + (size_t)EB * kPoolRows * WW + (size_t)EB * WW * 32) * 4;
  auto kern = heat_pool_kernel<EB>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("hgl_heat_pool: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return HGL_ECUDA; }
  dim3 grid(ws.tiles, B);
  kern<<<grid, kPoolWarps * 32, smem, st>>>(heat, expr_off, dirflag, bits, mask_off, M, E, H, W, max_n, ws);
  return launch_status("hgl_heat_pool(pool)");
}

}  // namespace hgl

// The expression-group width is picked from the average expressions per image (E/B); images with more run extra passes.
static int hgl_pool_eb(int B, int E) {
  const int avg = (E + B - 1) / B;
  return avg <= 1 ? 1 : (avg <= 2 ? 2 : (avg <= 3 ? 3 : (avg <= 4 ? 4 : (avg <= 6 ? 6 : 8))));
}

extern "C" int hgl_dir_mask(int dirflag, int H, int W, float* out, void* stream) {
  using namespace hgl;
  HGL_REQUIRE(out, "hgl_dir_mask: null pointer");
  HGL_REQUIRE(H >= 1 && W >= 1, "hgl_dir_mask: bad shape");
  const int blocks = (int)std::min<size_t>(((size_t)H * W + 255) / 256, (size_t)sm_count() * 8);
  dir_mask_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(dirflag, H, W, out);
  return launch_status("hgl_dir_mask");
}

extern "C" int64_t hgl_heat_pool_workspace_bytes(int B, int M, int E, int H, int W, int max_n) {
  using namespace hgl;
  if (B < 1 || M < 0 || E < 0 || H < 1 || W < 1 || max_n < 0) return -1;
  return (int64_t)carve(nullptr, M, E, H, W, max_n).bytes + 256;
}

extern "C" int hgl_heat_pool(const float* heat, const int32_t* expr_off, const int32_t* dirflag, const float* black,
                             const uint32_t* bits, const int32_t* mask_off, int B, int M, int E, int H, int W,
                             int max_n, float* score_gem, void* workspace, void* stream) {
  using namespace hgl;
  if (E == 0 || M == 0) return HGL_OK;
  HGL_REQUIRE(heat && dirflag && black && bits && score_gem && workspace, "hgl_heat_pool: null pointer");
  HGL_REQUIRE(B >= 1 && M >= 0 && E >= 0 && H >= 1 && W >= 1 && max_n >= 1, "hgl_heat_pool: bad shape");
  HGL_REQUIRE((mask_off && expr_off) || B == 1, "hgl_heat_pool: mask_off/expr_off required when B > 1");
  HGL_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "hgl_heat_pool: workspace must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  int eb = hgl_pool_eb(B, E);
  const int WW = (W + 31) >> 5;
  while (eb > 1 && (size_t)eb * (kPoolRows + 2) * WW * 32 * 4 > 96 * 1024) --eb;   // wide frames: fewer expressions per pass
  HGL_REQUIRE((size_t)eb * (kPoolRows + 2) * WW * 32 * 4 <= 200 * 1024 && kPoolRows * WW <= 128, "hgl_heat_pool: W=%d too wide", W);
  HeatWs ws = carve(workspace, M, E, H, W, max_n);
  heat_stats_kernel<<<dim3(kStatChunks, E), 256, 0, st>>>(heat, dirflag, H, W, ws.stats);
  int rc = launch_status("hgl_heat_pool(stats)");
  if (rc != HGL_OK) return rc;
  switch (eb) {
    case 1: rc = launch_pool<1>(heat, expr_off, dirflag, bits, mask_off, B, M, E, H, W, max_n, ws, st); break;
    case 2: rc = launch_pool<2>(heat, expr_off, dirflag, bits, mask_off, B, M, E, H, W, max_n, ws, st); break;
    case 3: rc = launch_pool<3>(heat, expr_off, dirflag, bits, mask_off, B, M, E, H, W, max_n, ws, st); break;
    case 4: rc = launch_pool<4>(heat, expr_off, dirflag, bits, mask_off, B, M, E, H, W, max_n, ws, st); break;
    case 5: case 6: rc = launch_pool<6>(heat, expr_off, dirflag, bits, mask_off, B, M, E, H, W, max_n, ws, st); break;
    default: rc = launch_pool<8>(heat, expr_off, dirflag, bits, mask_off, B, M, E, H, W, max_n, ws, st); break;
  }
  if (rc != HGL_OK) return rc;
  heat_finalize_kernel<<<dim3(ceil_div(max_n, 128), E), 128, 0, st>>>(expr_off, mask_off, black, B, M, E, H, W, max_n, ws, score_gem);
  return launch_status("hgl_heat_pool(finalize)");
}
