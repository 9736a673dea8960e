// (a1) per-mask visual-prompt preprocessing: one batched gather/resample kernel.
//
// Replaces the per-mask Python loop of the reference (Hybridgl_main.py:92-125; utils.py:292-345):
//   global[n] = Normalize_IN( bilinear_S( where(m_n, img, bg) / 255 ) )
//   local[n]  = bilinear_S( where(m_n, Normalize_IN(img/255), clip_pixel_mean) )
// with the non-antialiased bilinear of T.Resize(..., antialias=None) (ATen upsample_bilinear2d, align_corners=False).
//
// B200 design (HBM-bound: reads M*H*W mask bytes once, writes 2*M*3*S*S outputs once):
//   * a CTA owns R output rows x S columns of ONE image and loops over that image's masks, so the image /
//     blur taps (shared by every mask) are gathered once per CTA and kept in registers as the two possible
//     answers of each output pixel: "all four taps inside" (FG) and "all four taps outside" (BG);
//   * the mask rows a tile needs are contiguous in memory -> one 1-D TMA bulk copy (cp.async.bulk, SASS UBLKCP)
//     per mask into a 4-deep shared-memory ring, completion on an mbarrier (no register staging);
//   * per mask a thread reads the 4 tap bytes of each of its 4 adjacent pixels, selects FG/BG, and only pixels
//     whose taps straddle the mask boundary re-evaluate the exact per-tap formula from tap bytes parked in smem;
//   * outputs leave as 8-byte (bf16) / 16-byte (f32) streaming stores, 4 adjacent pixels per thread.
// Arithmetic follows ATen's CPU kernel op for op (explicit __fmaf_rn/__fmul_rn/__fdiv_rn, no re-contraction),
// so f32 output is bit-identical to the reference on the S=224/336 paths (see oracle/hybridgl_oracle.py header).
#include "hgl_common.cuh"

namespace hgl {

constexpr int kPrepStages = 4;
constexpr int kPrepPx = 4;  // adjacent output pixels per thread

struct PrepParams {
  const uint8_t* image;
  const uint8_t* blur;
  const uint8_t* masks;
  const int32_t* mask_off;
  void* local_out;
  void* global_out;
  int B, M, H, W, S;
  int bg_mode;
  int R;            // output rows per CTA
  int tpr;          // threads per output row (S/4)
  int nsplit;       // CTAs sharing the masks of one (image, row tile)
  int stage_bytes;  // bytes of one mask stage in shared memory
};

struct Taps {
  int i0, d;       // first source index, i1 - i0 (0 or 1)
  float w0, w1;
};

// ATen area_pixel_compute_source_index + compute_source_index_and_lambda (float, align_corners=False)
__device__ __forceinline__ Taps make_taps(int dst, int in_size, int out_size) {
  Taps t;
  if (in_size == out_size) {
    t.i0 = dst; t.d = 0; t.w0 = 1.f; t.w1 = 0.f;
    return t;
  }
  const float scale = __fdiv_rn((float)in_size, (float)out_size);
  float src = __fmaf_rn(scale, (float)dst + 0.5f, -0.5f);
  src = fmaxf(src, 0.f);
  int i0 = (int)src;
  i0 = min(i0, in_size - 1);
  t.i0 = i0;
  t.d = (i0 < in_size - 1) ? 1 : 0;
  float w1 = __fsub_rn(src, (float)i0);
  w1 = fminf(fmaxf(w1, 0.f), 1.f);
  t.w1 = w1;
  t.w0 = __fsub_rn(1.f, w1);
  return t;
}

__device__ __forceinline__ float bilerp(float a, float b, float c, float d, float wx0, float wx1, float wy0, float wy1) {
  const float top = __fmaf_rn(a, wx0, __fmul_rn(b, wx1));
  const float bot = __fmaf_rn(c, wx0, __fmul_rn(d, wx1));
  return __fmaf_rn(top, wy0, __fmul_rn(bot, wy1));
}

__constant__ float c_in_mean[3] = {0.485f, 0.456f, 0.406f};
__constant__ float c_in_std[3] = {0.229f, 0.224f, 0.225f};
__constant__ float c_clip_mean[3] = {0.48145466f, 0.4578275f, 0.40821073f};

template <bool kBF16>
__device__ __forceinline__ void store4(void* base, size_t elem_off, const float v[4]) {
  if (kBF16) {
    uint2 u;
    u.x = pack_bf16x2(v[0], v[1]);
    u.y = pack_bf16x2(v[2], v[3]);
    stg_stream(reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(base) + elem_off), u);
  } else {
    uint4 u;
    u.x = __float_as_uint(v[0]); u.y = __float_as_uint(v[1]); u.z = __float_as_uint(v[2]); u.w = __float_as_uint(v[3]);
    stg_stream(reinterpret_cast<uint4*>(reinterpret_cast<float*>(base) + elem_off), u);
  }
}

template <bool kBF16>
__global__ void __launch_bounds__(384) prep_kernel(const PrepParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int H = p.H, W = p.W, S = p.S, R = p.R;
  const int tile = blockIdx.x, b = blockIdx.y, split = blockIdx.z;
  const int tid = threadIdx.x;
  const int r0 = tile * R;
  const int rows = min(R, S - r0);

  // ---- shared memory carve-up
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);                       // [kPrepStages]
  float* lut255 = reinterpret_cast<float*>(smem + 64);                      // [256]   v/255
  float* lutn = lut255 + 256;                                               // [3][256] (v/255-mean)/std
  uint32_t* tapw = reinterpret_cast<uint32_t*>(lutn + 768);                 // [6][R*S] tap bytes (img 3 words, bg 3 words)
  uint8_t* stage0 = reinterpret_cast<uint8_t*>(tapw + 6 * R * S);
  stage0 = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(stage0) + 127) & ~uintptr_t(127));

  int n_lo = 0, n_hi = p.M;
  if (p.mask_off) { n_lo = p.mask_off[b]; n_hi = p.mask_off[b + 1]; }
  {  // this CTA's share of the image's masks
    const int cnt = n_hi - n_lo, per = (cnt + p.nsplit - 1) / p.nsplit;
    n_lo = n_lo + split * per;
    n_hi = min(n_hi, n_lo + per);
  }
  const int n_cnt = max(n_hi - n_lo, 0);

  // source-row window of this tile
  const Taps ty_first = make_taps(r0, H, S);
  const Taps ty_last = make_taps(r0 + rows - 1, H, S);
  const int ylo = ty_first.i0, yhi = ty_last.i0 + ty_last.d;
  const uint32_t win_bytes = (uint32_t)(yhi - ylo + 1) * (uint32_t)W;
  const uint8_t* masks_end = p.masks + (size_t)p.M * H * W;

  auto issue = [&](int k) {  // one thread: TMA bulk copy of mask (n_lo+k)'s row window into stage k%kPrepStages
    const int s = k % kPrepStages;
    const uint8_t* g = p.masks + ((size_t)(n_lo + k) * H + ylo) * W;
    const uintptr_t ga = reinterpret_cast<uintptr_t>(g);
    const uint32_t lead = (uint32_t)(ga & 15);
    const uint8_t* gal = g - lead;
    uint32_t bytes = (lead + win_bytes + 15u) & ~15u;
    uint8_t* dst = stage0 + (size_t)s * p.stage_bytes;
    // never read past the 16-byte granule that holds the last mask byte
    const uintptr_t end_al = (reinterpret_cast<uintptr_t>(masks_end) + 15) & ~uintptr_t(15);
    if (reinterpret_cast<uintptr_t>(gal) + bytes > end_al) bytes = (uint32_t)(end_al - reinterpret_cast<uintptr_t>(gal));
    mbar_expect_tx(&full[s], bytes);
    bulk_g2s(dst, gal, bytes, &full[s]);
  };

  if (tid == 0) {
    for (int s = 0; s < kPrepStages; ++s) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  for (int v = tid; v < 256; v += blockDim.x) {
    const float f = __fdiv_rn((float)v, 255.f);
    lut255[v] = f;
#pragma unroll
    for (int c = 0; c < 3; ++c) lutn[c * 256 + v] = __fdiv_rn(__fsub_rn(f, c_in_mean[c]), c_in_std[c]);
  }
  __syncthreads();
  if (tid == 0) {
    const int pre = min(n_cnt, kPrepStages);
    for (int k = 0; k < pre; ++k) issue(k);
  }

  // ---- per-thread pixel set-up: taps, FG / BG answers
  const int r = tid / p.tpr, c4 = tid - r * p.tpr;
  const bool active = (r < rows);
  const int i = r0 + (active ? r : 0);
  const int j0 = c4 * kPrepPx;
  const Taps ty = make_taps(i, H, S);
  const int rowoff0 = (ty.i0 - ylo) * W, rowoff1 = rowoff0 + ty.d * W;

  int xo[kPrepPx];          // x0 | dx << 16
  float wx0[kPrepPx], wx1[kPrepPx];
  float fgl[3][kPrepPx], fgg[3][kPrepPx], bgl[3][kPrepPx], bgg[3][kPrepPx];
  const uint8_t* img = p.image + (size_t)b * H * W * 3;
  const uint8_t* bg = (p.bg_mode == HGL_BG_BLUR) ? p.blur + (size_t)b * H * W * 3 : nullptr;
  if (active) {
#pragma unroll
    for (int q = 0; q < kPrepPx; ++q) {
      const Taps tx = make_taps(j0 + q, W, S);
      xo[q] = tx.i0 | (tx.d << 16);
      wx0[q] = tx.w0; wx1[q] = tx.w1;
      const size_t o00 = ((size_t)ty.i0 * W + tx.i0) * 3, o01 = o00 + tx.d * 3;
      const size_t o10 = o00 + (size_t)ty.d * W * 3, o11 = o10 + tx.d * 3;
      uint32_t iw[3] = {0, 0, 0}, bw[3] = {0, 0, 0};  // 12 bytes each: [tap][channel]
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const uint32_t a = img[o00 + c], bb = img[o01 + c], cc = img[o10 + c], d = img[o11 + c];
        uint32_t ba = 0, bbq = 0, bc = 0, bd = 0;
        if (bg) { ba = bg[o00 + c]; bbq = bg[o01 + c]; bc = bg[o10 + c]; bd = bg[o11 + c]; }
        // byte k = tap*3 + c
        const uint32_t vals_i[4] = {a, bb, cc, d}, vals_b[4] = {ba, bbq, bc, bd};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int k = t * 3 + c;
          iw[k >> 2] |= vals_i[t] << ((k & 3) * 8);
          bw[k >> 2] |= vals_b[t] << ((k & 3) * 8);
        }
        const float mean = c_in_mean[c], stdv = c_in_std[c], pm = c_clip_mean[c];
        fgg[c][q] = __fdiv_rn(__fsub_rn(bilerp(lut255[a], lut255[bb], lut255[cc], lut255[d], tx.w0, tx.w1, ty.w0, ty.w1), mean), stdv);
        bgg[c][q] = __fdiv_rn(__fsub_rn(bilerp(lut255[ba], lut255[bbq], lut255[bc], lut255[bd], tx.w0, tx.w1, ty.w0, ty.w1), mean), stdv);
        fgl[c][q] = bilerp(lutn[c * 256 + a], lutn[c * 256 + bb], lutn[c * 256 + cc], lutn[c * 256 + d], tx.w0, tx.w1, ty.w0, ty.w1);
        bgl[c][q] = bilerp(pm, pm, pm, pm, tx.w0, tx.w1, ty.w0, ty.w1);
      }
      const int px = r * S + j0 + q;
#pragma unroll
      for (int w = 0; w < 3; ++w) {
        tapw[w * R * S + px] = iw[w];
        tapw[(3 + w) * R * S + px] = bw[w];
      }
    }
  }
  // each thread only re-reads its own tapw entries, no barrier needed

  const size_t plane = (size_t)S * S;
  for (int k = 0; k < n_cnt; ++k) {
    const int s = k % kPrepStages;
    const uint32_t parity = (uint32_t)(k / kPrepStages) & 1u;
    mbar_wait(&full[s], parity);
    if (active) {
      const int n = n_lo + k;
      const uint8_t* g = p.masks + ((size_t)n * H + ylo) * W;
      const uint32_t lead = (uint32_t)(reinterpret_cast<uintptr_t>(g) & 15);
      const uint8_t* ms = stage0 + (size_t)s * p.stage_bytes + lead;
      float ol[3][kPrepPx], og[3][kPrepPx];
#pragma unroll
      for (int q = 0; q < kPrepPx; ++q) {
        const int x0 = xo[q] & 0xffff, dx = xo[q] >> 16;
        const uint32_t m00 = ms[rowoff0 + x0] != 0, m01 = ms[rowoff0 + x0 + dx] != 0;
        const uint32_t m10 = ms[rowoff1 + x0] != 0, m11 = ms[rowoff1 + x0 + dx] != 0;
        const uint32_t code = m00 | (m01 << 1) | (m10 << 2) | (m11 << 3);
        if (code == 0u) {
#pragma unroll
          for (int c = 0; c < 3; ++c) { ol[c][q] = bgl[c][q]; og[c][q] = bgg[c][q]; }
        } else if (code == 15u) {
#pragma unroll
          for (int c = 0; c < 3; ++c) { ol[c][q] = fgl[c][q]; og[c][q] = fgg[c][q]; }
        } else {  // boundary pixel: exact per-tap evaluation
          const int px = r * S + j0 + q;
          uint32_t iw[3], bw[3];
#pragma unroll
          for (int w = 0; w < 3; ++w) { iw[w] = tapw[w * R * S + px]; bw[w] = tapw[(3 + w) * R * S + px]; }
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            float gv[4], lv[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const int kk = t * 3 + c;
              const uint32_t vi = (iw[kk >> 2] >> ((kk & 3) * 8)) & 0xffu;
              const uint32_t vb = (bw[kk >> 2] >> ((kk & 3) * 8)) & 0xffu;
              const bool in = (code >> t) & 1u;
              gv[t] = lut255[in ? vi : vb];
              lv[t] = in ? lutn[c * 256 + vi] : c_clip_mean[c];
            }
            og[c][q] = __fdiv_rn(__fsub_rn(bilerp(gv[0], gv[1], gv[2], gv[3], wx0[q], wx1[q], ty.w0, ty.w1), c_in_mean[c]), c_in_std[c]);
            ol[c][q] = bilerp(lv[0], lv[1], lv[2], lv[3], wx0[q], wx1[q], ty.w0, ty.w1);
          }
        }
      }
      const size_t base = (size_t)n * 3 * plane + (size_t)i * S + j0;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        store4<kBF16>(p.local_out, base + c * plane, ol[c]);
        store4<kBF16>(p.global_out, base + c * plane, og[c]);
      }
    }
    __syncthreads();  // every thread is done with stage s
    if (tid == 0 && k + kPrepStages < n_cnt) issue(k + kPrepStages);
  }
}

static size_t prep_smem_bytes(int R, int S, int stage_bytes) {
  return 64 + 1024 * 4 + (size_t)6 * R * S * 4 + 128 + (size_t)kPrepStages * stage_bytes;
}

}  // namespace hgl

extern "C" int hgl_prep(const uint8_t* image, const uint8_t* blur, const uint8_t* masks, const int32_t* mask_off,
                        int B, int M, int H, int W, int S, int bg_mode, int out_dtype,
                        void* local_out, void* global_out, void* stream) {
  using namespace hgl;
  if (M == 0 && B >= 1) return HGL_OK;   // nothing to do (empty tensors have null data pointers)
  HGL_REQUIRE(image && masks && local_out && global_out, "hgl_prep: null pointer");
  HGL_REQUIRE(bg_mode == HGL_BG_BLUR || bg_mode == HGL_BG_BLACK, "hgl_prep: bg_mode %d", bg_mode);
  HGL_REQUIRE(bg_mode != HGL_BG_BLUR || blur, "hgl_prep: blur frame required for HGL_BG_BLUR");
  HGL_REQUIRE(out_dtype == HGL_F32 || out_dtype == HGL_BF16, "hgl_prep: out_dtype %d", out_dtype);
  HGL_REQUIRE(B >= 1 && M >= 0 && H >= 1 && W >= 1 && W < 65536, "hgl_prep: bad shape B=%d M=%d H=%d W=%d", B, M, H, W);
  HGL_REQUIRE(S >= 4 && S % 4 == 0 && S <= 1024, "hgl_prep: S=%d must be a multiple of 4 in [4,1024]", S);
  HGL_REQUIRE(mask_off || B == 1, "hgl_prep: mask_off required when B > 1");
  if (M == 0) return HGL_OK;

  PrepParams p;
  p.image = image; p.blur = blur; p.masks = masks; p.mask_off = mask_off;
  p.local_out = local_out; p.global_out = global_out;
  p.B = B; p.M = M; p.H = H; p.W = W; p.S = S; p.bg_mode = bg_mode;
  p.tpr = S / kPrepPx;
  int R = 4;
  while (R > 1 && p.tpr * R > 384) R >>= 1;
  p.R = R;
  const int tiles = ceil_div(S, R);
  // rows of source needed by R output rows: floor((R-1)*scale)+3 is a safe bound
  const double scale = (double)H / (double)S;
  int max_rows = (int)((R - 1) * scale) + 3;
  if (max_rows > H) max_rows = H;
  p.stage_bytes = (max_rows * W + 15 + 16 + 127) & ~127;
  // enough CTAs for ~2 waves when the batch is small
  const int want = 2 * sm_count() * 3;
  int nsplit = 1;
  const int avg_masks = ceil_div(M, B);
  while (tiles * B * nsplit < want && nsplit * 8 <= avg_masks) nsplit <<= 1;
  p.nsplit = nsplit;
  int threads = ceil_div(p.tpr * R, 32) * 32;
  const size_t smem = prep_smem_bytes(R, S, p.stage_bytes);
  HGL_REQUIRE(smem <= 227 * 1024, "hgl_prep: tile needs %zu bytes of shared memory (W=%d too wide)", smem, W);
  auto kern = (out_dtype == HGL_BF16) ? prep_kernel<true> : prep_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) { set_error("hgl_prep: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return HGL_ECUDA; }
  dim3 grid(tiles, B, nsplit);
  kern<<<grid, threads, smem, (cudaStream_t)stream>>>(p);
  return launch_status("hgl_prep");
}
