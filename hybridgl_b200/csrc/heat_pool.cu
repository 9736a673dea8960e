// (a10)+(a11) heat-map conditioning and per-mask pooling  --  Hybridgl_main.py:204-223, utils.py:135-161.
//
//   A'  = (A - min A) / (max A - min A) * ramp(dirflag)            :204-207
//   A'' = A' / mean(A')                                            :209
//   score_gem[n] = (2-black) * sum(A''*m_n)/|m_n|  -  black * sum(A''*(1-m_n)) / |1-m_n|     :218-223
// The reference loops over masks in Python (~8 full-frame kernels + one D2H sync per mask).  Here (SURVEY.md
// Appendix A-2): S_in[e,n] = masks[n,:] . A''_e, area[n] = |m_n|, S_tot[e] = sum A''_e, then a closed form.
//
// B200 design: the masks (M*H*W bytes) are the only large operand -> read each byte once per group of
// expressions, straight from HBM into registers (no shared-memory staging: there is no reuse of mask bytes).
//   pass 1  heat_stats : per expression min / max / sum(A*ramp)            (E*H*W*4 bytes, tiny)
//   pass 2  heat_pool  : a warp owns 32*PX contiguous pixels, keeps the conditioned heat values of up to EB
//                        expressions of its image in REGISTERS and streams every mask of that image past them
//                        (16-byte coalesced loads, predicated adds, one shuffle tree per (mask, warp));
//                        per-CTA partials are combined in a fixed order (deterministic, no float atomics)
//   pass 3  finalize   : sum the tile partials in order, closed form -> score_gem[E, max_n]
#include "hgl_common.cuh"

namespace hgl {

constexpr int kStatChunks = 32;
constexpr int kPoolWarps = 8;

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

struct HeatWs {            // workspace carve-up (all offsets in bytes, 16-aligned)
  float* stats;            // [E][kStatChunks][4]  min, max, sum(A*ramp), sum(ramp)
  float* part_sin;         // [tiles][E][max_n]
  float* part_tot;         // [tiles][E]
  int32_t* part_area;      // [tiles][M]
  int tiles;
  size_t bytes;
};

static HeatWs carve(void* ws, int M, int E, int H, int W, int max_n, int tile_px) {
  HeatWs h;
  const int tiles = (int)ceil_div64((int64_t)H * W, tile_px);
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
  const int e = blockIdx.y, ch = blockIdx.x;
  const size_t HW = (size_t)H * W;
  const size_t per = (HW + kStatChunks - 1) / kStatChunks;
  const size_t lo = (size_t)ch * per, hi = min(HW, lo + per);
  const float* A = heat + (size_t)e * HW;
  const int dir = dirflag[e];
  float mn = INFINITY, mx = -INFINITY, s1 = 0.f, s0 = 0.f;
  for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const float a = A[i];
    const float rp = ramp_at(dir, (int)(i % W), W);
    mn = fminf(mn, a); mx = fmaxf(mx, a);
    s1 += a * rp; s0 += rp;
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

template <int EB, int PX>
__global__ void __launch_bounds__(kPoolWarps * 32) heat_pool_kernel(const float* __restrict__ heat, const int32_t* __restrict__ expr_off,
                                                                    const int32_t* __restrict__ dirflag, const uint8_t* __restrict__ masks,
                                                                    const int32_t* __restrict__ mask_off, int M, int E, int H, int W, int max_n,
                                                                    int nchunk, HeatWs ws) {
  extern __shared__ float acc[];             // [kPoolWarps][nchunk][EB+1]  (last slot: area as int bits)
  const size_t HW = (size_t)H * W;
  const int tile = blockIdx.x, b = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int kWarpPx = 32 * PX;
  const size_t px0 = ((size_t)tile * kPoolWarps + warp) * kWarpPx + (size_t)lane * PX;   // first pixel of this lane
  int n_lo = 0, n_hi = M, e_lo = 0, e_hi = E;
  if (mask_off) { n_lo = mask_off[b]; n_hi = mask_off[b + 1]; }
  if (expr_off) { e_lo = expr_off[b]; e_hi = expr_off[b + 1]; }
  const int n_cnt = n_hi - n_lo;
  const bool vec_ok = (HW % PX == 0) && ((reinterpret_cast<uintptr_t>(masks) & (PX - 1)) == 0);

  for (int eg = e_lo; eg < e_hi; eg += EB) {       // groups of EB expressions (one pass over the masks per group)
    const int ne = min(EB, e_hi - eg);
    float cond[EB][PX];
    float tot[EB];
#pragma unroll
    for (int j = 0; j < EB; ++j) {
      tot[j] = 0.f;
      float mn = 0.f, kk = 0.f;
      int dir = 0;
      if (j < ne) { heat_consts(ws.stats, eg + j, HW, mn, kk); dir = dirflag[eg + j]; }
#pragma unroll
      for (int q = 0; q < PX; ++q) {
        const size_t pidx = px0 + q;
        float v = 0.f;
        if (j < ne && pidx < HW) {
          const float a = heat[(size_t)(eg + j) * HW + pidx];
          v = (a - mn) * kk * ramp_at(dir, (int)(pidx % W), W);
        }
        cond[j][q] = v;
        tot[j] += v;
      }
      tot[j] = warp_sum(tot[j]);
    }
    // deterministic S_tot partial: warp slots in shared memory, summed by thread 0
    __syncthreads();
    if (lane == 0)
      for (int j = 0; j < EB; ++j) acc[warp * EB + j] = tot[j];
    __syncthreads();
    if (threadIdx.x < ne) {
      float s = 0.f;
      for (int w = 0; w < kPoolWarps; ++w) s += acc[w * EB + threadIdx.x];
      ws.part_tot[(size_t)tile * E + eg + threadIdx.x] = s;
    }
    __syncthreads();

    for (int c0 = 0; c0 < n_cnt; c0 += nchunk) {   // chunks of masks whose accumulators fit in shared memory
      const int cn = min(nchunk, n_cnt - c0);
#pragma unroll 2
      for (int k = 0; k < cn; ++k) {
        const uint8_t* mp = masks + (size_t)(n_lo + c0 + k) * HW + px0;
        uint32_t w[PX / 4];
        if (vec_ok && px0 + PX <= HW) {
          if (PX == 16) {
            const uint4 v = ldg_stream(reinterpret_cast<const uint4*>(mp));
            w[0] = v.x; w[1] = v.y; w[PX / 4 - 2] = v.z; w[PX / 4 - 1] = v.w;
          } else {
            const uint2 v = *reinterpret_cast<const uint2*>(mp);
            w[0] = v.x; w[1] = v.y;
          }
        } else {
#pragma unroll
          for (int q = 0; q < PX / 4; ++q) w[q] = 0;
#pragma unroll
          for (int q = 0; q < PX; ++q)
            if (px0 + q < HW) w[q >> 2] |= (uint32_t)(mp[q] != 0) << ((q & 3) * 8);
        }
        float s[EB];
#pragma unroll
        for (int j = 0; j < EB; ++j) s[j] = 0.f;
        int cnt = 0;
#pragma unroll
        for (int q = 0; q < PX; ++q) {
          const bool on = ((w[q >> 2] >> ((q & 3) * 8)) & 0xffu) != 0;
          cnt += on;
#pragma unroll
          for (int j = 0; j < EB; ++j) s[j] += on ? cond[j][q] : 0.f;
        }
#pragma unroll
        for (int j = 0; j < EB; ++j) s[j] = warp_sum(s[j]);
        cnt = warp_sum_i(cnt);
        if (lane == 0) {
          float* a = acc + ((size_t)warp * nchunk + k) * (EB + 1);
#pragma unroll
          for (int j = 0; j < EB; ++j) a[j] = s[j];
          a[EB] = __int_as_float(cnt);
        }
      }
      __syncthreads();
      for (int t = threadIdx.x; t < cn * (EB + 1); t += blockDim.x) {
        const int k = t / (EB + 1), j = t - k * (EB + 1);
        if (j < EB) {
          if (j < ne) {
            float s = 0.f;
            for (int w2 = 0; w2 < kPoolWarps; ++w2) s += acc[((size_t)w2 * nchunk + k) * (EB + 1) + j];
            ws.part_sin[((size_t)tile * E + eg + j) * max_n + c0 + k] = s;
          }
        } else if (eg == e_lo) {
          int s = 0;
          for (int w2 = 0; w2 < kPoolWarps; ++w2) s += __float_as_int(acc[((size_t)w2 * nchunk + k) * (EB + 1) + EB]);
          ws.part_area[(size_t)tile * M + n_lo + c0 + k] = s;
        }
      }
      __syncthreads();
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

template <int EB, int PX>
static int launch_pool(const float* heat, const int32_t* expr_off, const int32_t* dirflag, const uint8_t* masks, const int32_t* mask_off,
                       int B, int M, int E, int H, int W, int max_n, HeatWs ws, cudaStream_t st) {
  int nchunk = max_n;
  const size_t per = (size_t)kPoolWarps * (EB + 1) * 4;
  if ((size_t)nchunk * per > 96 * 1024) nchunk = (int)(96 * 1024 / per);
  size_t smem = std::max<size_t>((size_t)nchunk * per, (size_t)kPoolWarps * EB * 4);
  auto kern = heat_pool_kernel<EB, PX>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("hgl_heat_pool: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return HGL_ECUDA; }
  dim3 grid(ws.tiles, B);
  kern<<<grid, kPoolWarps * 32, smem, st>>>(heat, expr_off, dirflag, masks, mask_off, M, E, H, W, max_n, nchunk, ws);
  return launch_status("hgl_heat_pool(pool)");
}

static int pool_px_for(int maxe) { return maxe <= 4 ? 16 : 8; }

}  // namespace hgl

// The expression-group width is picked from the average expressions per image (E/B); images with more run extra passes.
static int hgl_pool_eb(int B, int E) {
  const int avg = (E + B - 1) / B;
  return avg <= 1 ? 1 : (avg <= 2 ? 2 : (avg <= 3 ? 3 : (avg <= 4 ? 4 : 8)));
}

extern "C" int64_t hgl_heat_pool_workspace_bytes(int B, int M, int E, int H, int W, int max_n) {
  using namespace hgl;
  if (B < 1 || M < 0 || E < 0 || H < 1 || W < 1 || max_n < 0) return -1;
  const int px = pool_px_for(hgl_pool_eb(B, E));
  return (int64_t)carve(nullptr, M, E, H, W, max_n, kPoolWarps * 32 * px).bytes + 256;
}

extern "C" int hgl_heat_pool(const float* heat, const int32_t* expr_off, const int32_t* dirflag, const float* black,
                             const uint8_t* masks, const int32_t* mask_off, int B, int M, int E, int H, int W,
                             int max_n, float* score_gem, void* workspace, void* stream) {
  using namespace hgl;
  HGL_REQUIRE(heat && dirflag && black && masks && score_gem && workspace, "hgl_heat_pool: null pointer");
  HGL_REQUIRE(B >= 1 && M >= 0 && E >= 0 && H >= 1 && W >= 1 && max_n >= 1, "hgl_heat_pool: bad shape");
  HGL_REQUIRE((mask_off && expr_off) || B == 1, "hgl_heat_pool: mask_off/expr_off required when B > 1");
  HGL_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "hgl_heat_pool: workspace must be 16-byte aligned");
  if (E == 0 || M == 0) return HGL_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int eb = hgl_pool_eb(B, E);
  const int px = pool_px_for(eb);
  HeatWs ws = carve(workspace, M, E, H, W, max_n, kPoolWarps * 32 * px);
  heat_stats_kernel<<<dim3(kStatChunks, E), 256, 0, st>>>(heat, dirflag, H, W, ws.stats);
  int rc = launch_status("hgl_heat_pool(stats)");
  if (rc != HGL_OK) return rc;
  switch (eb) {
    case 1: rc = launch_pool<1, 16>(heat, expr_off, dirflag, masks, mask_off, B, M, E, H, W, max_n, ws, st); break;
    case 2: rc = launch_pool<2, 16>(heat, expr_off, dirflag, masks, mask_off, B, M, E, H, W, max_n, ws, st); break;
    case 3: rc = launch_pool<3, 16>(heat, expr_off, dirflag, masks, mask_off, B, M, E, H, W, max_n, ws, st); break;
    case 4: rc = launch_pool<4, 16>(heat, expr_off, dirflag, masks, mask_off, B, M, E, H, W, max_n, ws, st); break;
    default: rc = launch_pool<8, 8>(heat, expr_off, dirflag, masks, mask_off, B, M, E, H, W, max_n, ws, st); break;
  }
  if (rc != HGL_OK) return rc;
  heat_finalize_kernel<<<dim3(ceil_div(max_n, 128), E), 128, 0, st>>>(expr_off, mask_off, black, B, M, E, H, W, max_n, ws, score_gem);
  return launch_status("hgl_heat_pool(finalize)");
}
