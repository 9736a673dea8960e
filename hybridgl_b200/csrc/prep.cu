// (a1) per-mask visual-prompt preprocessing, and the packed-mask format every downstream kernel consumes.
//
// Replaces the per-mask Python loop of the reference (Hybridgl_main.py:92-125; utils.py:292-345):
//   global[n] = Normalize_IN( bilinear_S( where(m_n, img, bg) / 255 ) )
//   local[n]  = bilinear_S( where(m_n, Normalize_IN(img/255), clip_pixel_mean) )
// with the non-antialiased bilinear of T.Resize(..., antialias=None) (ATen upsample_bilinear2d, align_corners=False).
//
// B200 design -- three kernels, all HBM-bound:
//   pack   : the byte masks (M*H*W B, the largest input of the whole path) are read exactly ONCE, 16 B per lane fully
//            coalesced, and squeezed to 1 bit/pixel ([M,H,ceil(W/32)] u32).  Every later kernel (prep, mask grid,
//            heat-map pooling, IoU) works on this 8x smaller tensor, which stays L2-resident for typical batches.
//   setup  : per IMAGE (not per mask) the two possible answers of every output pixel -- "all four taps inside the mask"
//            (FG) and "all four outside" (BG) -- for the 6 output planes, already in the output dtype, plus the 24 tap
//            bytes of the pixel for the rare boundary case.  12 planes of S*S per image: L2-resident.
//   main   : streams over the masks of an image band by band.  A lane owns 8 adjacent output pixels (4 for f32) and keeps
//            their FG / BG answers in registers; a producer warp stages the band's bit rows with 1-D bulk async copies;
//            one funnel shift + two warp votes decide a whole warp's 256 pixels in the common case and the 6 planes leave
//            as 16-byte streaming stores, 512 contiguous bytes per warp.  Only pixels whose taps straddle the mask outline
//            re-evaluate the exact per-tap formula (see prep_main_kernel).  The kernel is write-dominated: 2*M*3*S*S
//            output elements against M*H*W/8 bytes of mask bits.
// Arithmetic follows ATen's CPU kernel op for op (explicit __fmaf_rn/__fmul_rn/__fdiv_rn, no re-contraction),
// so f32 output is bit-identical to the reference on the S=224/336 paths (see oracle/hybridgl_oracle.py header).
#include <stdlib.h>

#include "hgl_common.cuh"
#include "prep_math.cuh"

namespace hgl {

constexpr int kPrepPx = 4;        // adjacent output pixels per thread
constexpr int kPrepThreads = 256;

// ---------------------------------------------------------------------------------------------------------------------
// pack: byte masks -> bit masks
// ---------------------------------------------------------------------------------------------------------------------
// 4 mask bytes -> 4 bits (byte i non-zero -> bit i): high bit of every non-zero byte lane, gathered by one multiply
__device__ __forceinline__ uint32_t nibble_of(uint32_t v) {
  const uint32_t t = (((v & 0x7f7f7f7fu) + 0x7f7f7f7fu) | v) & 0x80808080u;
  return (t * 0x00204081u) >> 28;
}
// the same for bytes that are known to be 0 or 1 (torch.bool storage): the four bits sit at 0 / 8 / 16 / 24 and one multiply by
// 2^28 + 2^21 + 2^14 + 2^7 lines them up in the top nibble (every partial product lands on its own bit: no carries)
__device__ __forceinline__ uint32_t nibble_of_bool(uint32_t v) { return (v * 0x10204080u) >> 28; }
template <bool kBool>
__device__ __forceinline__ uint32_t half_of(const uint4 v) {   // 16 mask bytes -> 16 bits
  if (kBool) return nibble_of_bool(v.x) | (nibble_of_bool(v.y) << 4) | (nibble_of_bool(v.z) << 8) | (nibble_of_bool(v.w) << 12);
  return nibble_of(v.x) | (nibble_of(v.y) << 4) | (nibble_of(v.z) << 8) | (nibble_of(v.w) << 12);
}

// W % 32 == 0 and 16-byte aligned base: the masks are one flat byte stream, word w = pixels [32w, 32w+32).
// lane loads 16 B (fully coalesced 512 B per warp), neighbours pair up through one shuffle, even lanes store.
template <bool kBool>
__global__ void __launch_bounds__(256) pack_masks_flat_kernel(const uint4* __restrict__ src, size_t n16, uint32_t* __restrict__ bits) {
  constexpr int kU = 8;   // independent 16-byte loads in flight per thread
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  // The unrolled loop runs only while the WHOLE warp is in range (the test uses the warp's last lane, so it is warp-uniform and
  // the full-mask shuffle below is well defined); whatever is left goes through the tail loop.
  const size_t lane = threadIdx.x & 31;
  for (; (i - lane + 31) + (kU - 1) * stride < n16; i += kU * stride) {
    uint4 v[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) v[u] = ldg_stream(src + i + u * stride);
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const uint32_t h = half_of<kBool>(v[u]);
      const uint32_t o = __shfl_xor_sync(0xffffffffu, h, 1);
      if (!(threadIdx.x & 1)) bits[(i + u * stride) >> 1] = h | (o << 16);
    }
  }
  // tail: n16 is even (W % 32 == 0) and the stride is even, so lane pairs (2k, 2k+1) enter and leave together; the pair
  // exchanges through a ballot-derived mask of the lanes that are really here
  for (; (i - lane) < n16; i += stride) {
    const bool in = i < n16;
    const unsigned act = __ballot_sync(0xffffffffu, in);
    if (in) {
      const uint32_t h = half_of<kBool>(ldg_stream(src + i));
      const uint32_t o = __shfl_xor_sync(act, h, 1);
      if (!(threadIdx.x & 1)) bits[i >> 1] = h | (o << 16);
    }
  }
}

// general widths: one thread per (mask row, 32-pixel word)
template <bool kBool>
__global__ void __launch_bounds__(256) pack_masks_rows_kernel(const uint8_t* __restrict__ masks, size_t rows_total, int W, uint32_t* __restrict__ bits) {
  const int WW = (W + 31) >> 5;
  const size_t total = rows_total * WW;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const size_t row = t / WW;
    const int wd = (int)(t - row * WW);
    const uint8_t* sp = masks + row * W + 32 * wd;
    const int nv = min(32, W - 32 * wd);
    uint32_t word = 0;
    if (nv == 32 && (reinterpret_cast<uintptr_t>(sp) & 3) == 0) {
      const uint32_t* s4 = reinterpret_cast<const uint32_t*>(sp);
#pragma unroll
      for (int q = 0; q < 8; ++q) word |= (kBool ? nibble_of_bool(ldg_stream32(s4 + q)) : nibble_of(ldg_stream32(s4 + q))) << (4 * q);
    } else {
      for (int q = 0; q < nv; ++q) word |= (uint32_t)(sp[q] != 0) << q;
    }
    bits[t] = word;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// setup: per-image FG / BG answer planes + tap bytes
// ---------------------------------------------------------------------------------------------------------------------
// planes: [B][12][S*S] in the output dtype; plane k = 3*kind + channel, kind 0 = FG local, 1 = FG global, 2 = BG local, 3 = BG global
// taps:   [B][S*S][8] u32 (one 32-byte sector per output pixel); words 0-2 image bytes, 3-5 background bytes (6-7 unused), byte
//         index inside the 12 = tap*3 + channel.  One sector = one prefetch when an outline pixel is filed, two 16-byte loads when
//         it is evaluated (six planar words cost six sectors per pixel: 154 MB instead of 26 MB of sector traffic per pass).
template <bool kBF16>
__global__ void __launch_bounds__(256) prep_setup_kernel(const uint8_t* __restrict__ image, const uint8_t* __restrict__ blur, int H, int W, int S,
                                                         void* __restrict__ planes, uint32_t* __restrict__ taps, float* __restrict__ lut) {
  // per-byte maps: [v] = v/255 (T.ToTensor), [256 + 256*c + v] = Normalize_c(v/255): the correctly-rounded divisions are done
  // 4 x 256 times per CTA instead of ~40 times per pixel
  __shared__ float slut[1024];
  {
    const uint32_t v = threadIdx.x;
    slut[v] = to_unit(v);
    slut[256 + v] = to_norm(v, 0); slut[512 + v] = to_norm(v, 1); slut[768 + v] = to_norm(v, 2);
  }
  __syncthreads();
  const int b = blockIdx.y;
  if (blockIdx.x == 0 && b == 0)         // the same tables for the fix-up pass of prep_main_kernel
    for (int k = threadIdx.x; k < 1024; k += 256) lut[k] = slut[k];
  const int px = blockIdx.x * blockDim.x + threadIdx.x;
  const int SS = S * S;
  if (px >= SS) return;
  const int i = px / S, j = px - i * S;
  const Taps ty = make_taps(i, H, S), tx = make_taps(j, W, S);
  const uint8_t* img = image + (size_t)b * H * W * 3;
  const uint8_t* bg = blur ? blur + (size_t)b * H * W * 3 : nullptr;
  const size_t o00 = ((size_t)ty.i0 * W + tx.i0) * 3, o01 = o00 + tx.d * 3;
  const size_t o10 = o00 + (size_t)ty.d * W * 3, o11 = o10 + tx.d * 3;
  uint32_t iw[3] = {0, 0, 0}, bw[3] = {0, 0, 0};
  float v[12];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const uint32_t vi[4] = {img[o00 + c], img[o01 + c], img[o10 + c], img[o11 + c]};
    uint32_t vb[4] = {0, 0, 0, 0};
    if (bg) { vb[0] = bg[o00 + c]; vb[1] = bg[o01 + c]; vb[2] = bg[o10 + c]; vb[3] = bg[o11 + c]; }
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int k = t * 3 + c;
      iw[k >> 2] |= vi[t] << ((k & 3) * 8);
      bw[k >> 2] |= vb[t] << ((k & 3) * 8);
    }
    const float mean = c_in_mean[c], stdv = c_in_std[c], pm = c_clip_mean[c];
    const float* nl = slut + 256 + 256 * c;
    v[0 + c] = bilerp(nl[vi[0]], nl[vi[1]], nl[vi[2]], nl[vi[3]], tx.w0, tx.w1, ty.w0, ty.w1);
    v[3 + c] = __fdiv_rn(__fsub_rn(bilerp(slut[vi[0]], slut[vi[1]], slut[vi[2]], slut[vi[3]], tx.w0, tx.w1, ty.w0, ty.w1), mean), stdv);
    v[6 + c] = bilerp(pm, pm, pm, pm, tx.w0, tx.w1, ty.w0, ty.w1);
    v[9 + c] = __fdiv_rn(__fsub_rn(bilerp(slut[vb[0]], slut[vb[1]], slut[vb[2]], slut[vb[3]], tx.w0, tx.w1, ty.w0, ty.w1), mean), stdv);
  }
#pragma unroll
  for (int k = 0; k < 12; ++k) {
    const size_t o = ((size_t)b * 12 + k) * SS + px;
    if (kBF16) reinterpret_cast<__nv_bfloat16*>(planes)[o] = __float2bfloat16_rn(v[k]);
    else reinterpret_cast<float*>(planes)[o] = v[k];
  }
  uint4* tp = reinterpret_cast<uint4*>(taps + ((size_t)b * SS + px) * 8);
  tp[0] = make_uint4(iw[0], iw[1], iw[2], bw[0]);
  tp[1] = make_uint4(bw[1], bw[2], 0u, 0u);
}

// ---------------------------------------------------------------------------------------------------------------------
// main: stream the masks
// ---------------------------------------------------------------------------------------------------------------------
struct PrepParams {
  const uint32_t* bits;       // [M,H,WW]
  const int32_t* mask_off;
  const void* planes;         // [B,12,S*S]
  const uint32_t* taps;       // [B,S*S,8]
  const float* lut;           // [1024] byte -> unit / normalised value
  void* local_out;
  void* global_out;
  int B, M, H, W, S, WW;
  int stage_words;            // ring pitch of the shared-memory bit-row stages (words)
  int sub;                    // masks per stage
  int gw, gh, cw, nbx, strip; // thread -> pixel map (PrepGeom)
  int flush_every;            // masks between two flushes of a warp's outline list; > sub: only when the list is full and at the CTA's end
  int debug;                  // profiling only (HGL_PREP_DEBUG): 1 = skip the exact outline pixels, 2 = skip the stores
  int narrow;                 // 1 if the 8 taps of 4 adjacent pixels always fit one 32-bit window
  int gz;                     // mask-span splits per (image, band tile)
};

// PX adjacent output pixels of one plane, packed in the output dtype: NW 32-bit words (8- or 16-byte stores)
template <bool kBF16, int PX>
struct Pack {
  static constexpr int NW = kBF16 ? PX / 2 : PX;
  static_assert(NW == 2 || NW == 4, "a pack is one 8- or 16-byte store");
  uint32_t w[NW];
  __device__ __forceinline__ void load(const void* planes, size_t elem_off) {
    const uint8_t* ptr = reinterpret_cast<const uint8_t*>(planes) + elem_off * (kBF16 ? 2 : 4);
    if (NW == 2) { const uint2 v = *reinterpret_cast<const uint2*>(ptr); w[0] = v.x; w[1] = v.y; }
    else { const uint4 v = *reinterpret_cast<const uint4*>(ptr); w[0] = v.x; w[1] = v.y; w[NW - 2] = v.z; w[NW - 1] = v.w; }
  }
  // replace pixel q by value v
  __device__ __forceinline__ void put(int q, float v) {
    if (kBF16) {
      const uint32_t h = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v));
      uint32_t& x = w[q >> 1];
      x = (q & 1) ? ((x & 0x0000ffffu) | (h << 16)) : ((x & 0xffff0000u) | h);
    } else {
      w[q % NW] = __float_as_uint(v);
    }
  }
  __device__ __forceinline__ void store(uint8_t* ptr) const {   // ptr: address of the first of the PX pixels
    if (NW == 2) stg_stream(reinterpret_cast<uint2*>(ptr), make_uint2(w[0], w[1]));
    else stg_stream(reinterpret_cast<uint4*>(ptr), make_uint4(w[0], w[1], w[NW - 2], w[NW - 1]));
  }
};

// boundary pixel: exact per-tap evaluation (Hybridgl_main.py:106-121 restricted to the 4 taps of one output pixel).
// Only pixels whose taps straddle the mask outline come here, through the dense per-warp resolve step of prep_main_kernel.
// lut: tables (built by prep_setup_kernel, copied into shared memory by every CTA) of the two per-byte maps, [0..255] = v/255 (T.ToTensor), [256 + 256*c + v] = Normalize_c(v/255)
// -- the same correctly-rounded divisions as to_unit / to_norm, evaluated once per image batch instead of 27 times per pixel.
__device__ __forceinline__ void prep_boundary_pixel(const uint32_t* __restrict__ taps, size_t tap0, uint32_t code,
                                                    float wx0, float wx1, float wy0, float wy1, const float* __restrict__ lut,
                                                    float* __restrict__ out6) {
  // tap0 = (image * S*S + pixel) * 8: the pixel's 32-byte record (brought into L1 by the prefetch issued when it was filed)
  const uint4 t0 = __ldg(reinterpret_cast<const uint4*>(taps + tap0)), t1 = __ldg(reinterpret_cast<const uint4*>(taps + tap0) + 1);
  const uint32_t iw[3] = {t0.x, t0.y, t0.z}, bw[3] = {t0.w, t1.x, t1.y};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float gv[4], lv[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int kk = t * 3 + c;
      const uint32_t vi = (iw[kk >> 2] >> ((kk & 3) * 8)) & 0xffu;
      const uint32_t vb = (bw[kk >> 2] >> ((kk & 3) * 8)) & 0xffu;
      const bool in = (code >> t) & 1u;
      gv[t] = lut[in ? vi : vb];
      lv[t] = in ? lut[256 + 256 * c + vi] : c_clip_mean[c];
    }
    out6[3 + c] = __fdiv_rn(__fsub_rn(bilerp(gv[0], gv[1], gv[2], gv[3], wx0, wx1, wy0, wy1), c_in_mean[c]), c_in_std[c]);
    out6[c] = bilerp(lv[0], lv[1], lv[2], lv[3], wx0, wx1, wy0, wy1);
  }
}

#ifndef HGL_PREP_STAGES
#define HGL_PREP_STAGES 3
#endif
constexpr int kPrepStages = HGL_PREP_STAGES;      // shared-memory ring of bit-row stages (<= 8: the barrier block below)
constexpr int kPrepBarBytes = 128;  // full[kPrepStages] | empty[kPrepStages] mbarriers
constexpr int kPrepLutBytes = 4096; // the two per-byte maps (1024 floats) of the exact outline evaluation, copied in per CTA:
                                    // data-dependent lookups hit 32 banks at once instead of up to 32 L1 sectors one after the other
static_assert(2 * kPrepStages * 8 <= kPrepBarBytes, "barrier block");
constexpr int kPrepSubDefault = 4;  // masks per stage (PrepParams::sub; HGL_PREP_SUB overrides for tuning)
constexpr int kPrepWq = 512;        // per-warp list of outline pixels waiting for their exact value (>= 32 lanes x 8 pixels)
constexpr int kPrepMaxSub = 16;     // masks per stage, at most
constexpr int kPrepWarpBytes = kPrepWq * 4;

// Thread -> pixel map.  A lane owns PX adjacent output pixels (one 16-byte store per plane); the 32 lanes of a warp form a
// 2-D patch of gw groups x gh rows (gw * gh = 32) and the cw warps of a CTA sit side by side, so a CTA covers a band of gh
// output rows.  Compared with a warp = one long row strip, a mask outline crosses far fewer patches than strips, so the
// (longer) mixed-group path runs for ~10 % of the warp iterations instead of ~30 %.
struct PrepGeom {
  int gw, gh, cw, nbx;     // groups per patch row, rows per patch, warps (patches) per CTA, CTAs per band
  int strip;               // > 0: strip mode, = groups per output row (see hgl_prep_main)
};
__device__ __forceinline__ void prep_pixel_of(const PrepGeom& gm, int bxi, int byi, int t, int PX, int& i, int& j0) {
  if (gm.strip) {          // the band's groups in row-major order: a warp = 32 consecutive groups = one contiguous run of the plane
    const int r = t / gm.strip;
    i = byi * gm.gh + r;
    j0 = (t - r * gm.strip) * PX;
    return;
  }
  const int warp = t >> 5, lane = t & 31;
  const int gr = lane / gm.gw, gc = lane - gr * gm.gw;
  i = byi * gm.gh + gr;
  j0 = ((bxi * gm.cw + warp) * gm.gw + gc) * PX;
}

// One CTA = a band tile of one image x a span of masks (blockIdx.z-th share of the image's masks), processed as a pipeline
// of stages of kPrepSub masks with NO CTA-wide barrier in the steady state:
//   producer  (kTMA) one extra warp: its lane 0 keeps the ring of bit-row stages full: it waits for a slot to be released
//             (`empty` mbarrier, one arrival per consumer warp) and refills it -- one 1-D bulk async copy (cp.async.bulk, SASS
//             UBLKCP) per mask onto the slot's `full` mbarrier.  !kTMA (masks not 16-byte aligned: H * ceil(W/32) % 4 != 0):
//             cooperative loads between two __syncthreads per stage.
//   consumers per mask: window from shared memory, then two warp votes.  A warp whose 32 groups lie entirely inside or
//             entirely outside the mask (the common case) is six 16-byte streaming stores straight from the FG or BG answer
//             registers -- no select, no per-pixel work, no global load; the answers sit in registers for the CTA's whole
//             span.  Otherwise FG/BG are merged per pixel with bit masks; pixels whose own four taps straddle the outline get
//             the BG value as a placeholder and are filed in the WARP's list (one shuffle scan).
//   flush     at the end of a stage (or when the list is full) the warp evaluates its listed pixels exactly, one lane per
//             pixel side by side (dense; ATen's op order), and patches the freshly written lines (L2).  Warps never wait for
//             each other: they drift apart by up to kPrepStages stages, so a warp that crosses many outlines does not hold
//             the others up.
template <bool kBF16, int PX, bool kTMA, bool kNarrow>
__global__ void __launch_bounds__(kPrepThreads, 2) prep_main_kernel(const PrepParams p) {
  extern __shared__ __align__(128) uint8_t sm_prep[];
  uint64_t* full = reinterpret_cast<uint64_t*>(sm_prep);                       // [kPrepStages]  (kPrepBarBytes reserved for both)
  uint64_t* empty = full + kPrepStages;                                        // [kPrepStages]
  float* lut = reinterpret_cast<float*>(sm_prep + kPrepBarBytes);              // [1024]

  const int H = p.H, W = p.W, S = p.S, WW = p.WW;
  const int SS = S * S;
  const PrepGeom gm = {p.gw, p.gh, p.cw, p.nbx, p.strip};
  const int ncons = 32 * gm.cw;                                                // consumer threads (the producer warp comes after them)
  uint32_t* stage_base = reinterpret_cast<uint32_t*>(sm_prep + kPrepBarBytes + kPrepLutBytes + (size_t)gm.cw * kPrepWarpBytes);   // [kPrepStages][kPrepSub][rows * WW] + 16 B
  // blockIdx.x = band tile * gz + z: the gz CTAs that share a band tile's answer planes (and the CTAs of one image) are dispatched
  // back to back, so the planes are fetched from HBM once per image and served from L2 to the rest (with z as a grid dimension the
  // z-slices of an image ran a wave apart and every one of them re-fetched the planes: 2.3x read amplification in ncu)
  const int zi = blockIdx.x % p.gz, tile = blockIdx.x / p.gz;
  const int bxi = tile % gm.nbx, byi = tile / gm.nbx;
  const int b = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int n_lo = 0, n_hi = p.M;
  if (p.mask_off) { n_lo = p.mask_off[b]; n_hi = p.mask_off[b + 1]; }
  {
    const int per = (n_hi - n_lo + p.gz - 1) / p.gz;                           // this CTA's share of the image's masks
    n_lo += zi * per;
    n_hi = min(n_hi, n_lo + per);
  }
  if (n_lo >= n_hi) return;                                          // uniform for the whole CTA
  const int cnt = n_hi - n_lo;
  const int kPrepSub = p.sub;
  const int nst = (cnt + kPrepSub - 1) / kPrepSub;

  const float sc_y = tap_scale(H, S), sc_x = tap_scale(W, S);
  const int row_first = byi * gm.gh, row_last = min(S, row_first + gm.gh) - 1;   // output rows of the band
  const int ylo = make_taps(row_first, H, S, sc_y).i0;
  const Taps tl = make_taps(row_last, H, S, sc_y);
  const int nrows = tl.i0 + tl.d - ylo + 1;
  const size_t mask_words = (size_t)H * WW;
  // kTMA: every mask starts 16-byte aligned (H * WW % 4 == 0), the band's first row need not: a bulk copy starts `lead` words
  // early and is rounded up to whole 16-byte units (it stays inside the mask)
  const int lead = kTMA ? ((ylo * WW) & 3) : 0;
  const int per_mask = kTMA ? ((lead + nrows * WW + 3) & ~3) : nrows * WW;   // words of one mask's slot inside a stage
  const int stage_words = p.stage_words;                             // ring pitch (host: kPrepSub * max slot words)
  const uint32_t* src0 = p.bits + (size_t)n_lo * mask_words + (size_t)ylo * WW - lead;

  for (int t = tid; t < 1024; t += blockDim.x) lut[t] = __ldg(p.lut + t);
  if (kTMA) {
    if (tid == 0) {
      for (int s = 0; s < kPrepStages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, (uint32_t)gm.cw); }
      mbar_fence_init();
    }
    __syncthreads();
    if (warp == gm.cw) {                                             // ---- producer warp
      if (lane == 0 && !(p.debug & 4)) {
        for (int c = 0; c < nst; ++c) {
          const int slot = c % kPrepStages;
          if (c >= kPrepStages) mbar_wait(empty + slot, (uint32_t)((c / kPrepStages - 1) & 1));
          const int c0 = c * kPrepSub, cn = min(kPrepSub, cnt - c0);
          uint32_t* dst = stage_base + (size_t)slot * stage_words;
          mbar_expect_tx(full + slot, (uint32_t)(cn * per_mask * 4));
          for (int k = 0; k < cn; ++k) bulk_g2s(dst + k * per_mask, src0 + (size_t)(c0 + k) * mask_words, (uint32_t)(per_mask * 4), full + slot);
        }
      }
      return;
    }
  }

  // ---- consumers
  uint32_t* wq = reinterpret_cast<uint32_t*>(sm_prep + kPrepBarBytes + kPrepLutBytes + (size_t)warp * kPrepWarpBytes);   // [kPrepWq] mask << 16 | owner pixel << 4 | tap code
  int i, j0;
  prep_pixel_of(gm, bxi, byi, tid, PX, i, j0);
  const bool live = i <= row_last && j0 < S;                         // partial last band / idle lanes of the last warp (strip mode)
  if (!live) { i = row_last; j0 = 0; }
  const Taps ty = make_taps(i, H, S, sc_y);
  const int bx = make_taps(j0, W, S, sc_x).i0;
  uint32_t tapmask = 0;
  uint32_t tm[PX];                   // per pixel: the two tap columns as bits of the 32-bit window starting at bx
#pragma unroll
  for (int q = 0; q < PX; ++q) {
    tm[q] = 0;
    if (kNarrow) {
      const Taps tx = make_taps(j0 + q, W, S, sc_x);
      tm[q] = (1u << (tx.i0 - bx)) | (1u << (tx.i0 + tx.d - bx));
      tapmask |= tm[q];
    }
  }
  const int wi = bx >> 5, sh = bx & 31;                              // the word after the window's first is always readable (see host)
  Pack<kBF16, PX> ans[12];
  const size_t px0 = (size_t)i * S + j0;
#pragma unroll
  for (int k = 0; k < 12; ++k) ans[k].load(p.planes, ((size_t)b * 12 + k) * SS + px0);

  constexpr size_t kElem = kBF16 ? 2 : 4;
  const size_t plane_bytes = (size_t)SS * kElem;
  uint8_t* lp0 = reinterpret_cast<uint8_t*>(p.local_out) + ((size_t)n_lo * 3 * SS + px0) * kElem;    // planes of mask n_lo
  uint8_t* gp0 = reinterpret_cast<uint8_t*>(p.global_out) + ((size_t)n_lo * 3 * SS + px0) * kElem;
  const int row_off0 = lead + (ty.i0 - ylo) * WW, row_off1 = row_off0 + ty.d * WW;
  const uint32_t tap_base = (uint32_t)b * (uint32_t)SS * 8u;          // word offset of the image's tap records (< 2^32: host check)
  const uint32_t tap_px0 = tap_base + (uint32_t)px0 * 8u;

  // exact values of the listed outline pixels (Hybridgl_main.py:106-121 restricted to the 4 taps of one output pixel), one lane
  // per pixel; overwrites the placeholders this warp stored earlier (ordered by the __syncwarp)
  int wcount = 0;                                                    // entries in wq (warp-uniform)
  auto flush = [&]() {
    __syncwarp();
    for (int t = lane; t < wcount; t += 32) {
      const uint32_t ent = wq[t];
      const int lpx = (int)((ent >> 4) & 0xfffu), ol_ = lpx / PX, oq = lpx - ol_ * PX;      // owner lane, pixel of its group
      int pi, pj0;
      prep_pixel_of(gm, bxi, byi, (warp << 5) | ol_, PX, pi, pj0);
      const int pj = pj0 + oq, gpx = pi * S + pj;
      const Taps tyy = make_taps(pi, H, S, sc_y), txx = make_taps(pj, W, S, sc_x);
      float o6[6];
      prep_boundary_pixel(p.taps, tap_base + (uint32_t)gpx * 8u, ent & 15u, txx.w0, txx.w1, tyy.w0, tyy.w1, lut, o6);
      const size_t o = ((size_t)(n_lo + (int)(ent >> 16)) * 3) * SS + gpx;
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        if (kBF16) {
          reinterpret_cast<__nv_bfloat16*>(p.local_out)[o + (size_t)ch * SS] = __float2bfloat16_rn(o6[ch]);
          reinterpret_cast<__nv_bfloat16*>(p.global_out)[o + (size_t)ch * SS] = __float2bfloat16_rn(o6[3 + ch]);
        } else {
          reinterpret_cast<float*>(p.local_out)[o + (size_t)ch * SS] = o6[ch];
          reinterpret_cast<float*>(p.global_out)[o + (size_t)ch * SS] = o6[3 + ch];
        }
      }
    }
    __syncwarp();
    wcount = 0;
  };

  uint8_t* lp = lp0;
  uint8_t* gp = gp0;
  for (int c = 0; c < nst; ++c) {
    const int c0 = c * kPrepSub, cn = min(kPrepSub, cnt - c0);
    const int slot = c % kPrepStages;
    uint32_t* stage = stage_base + (size_t)slot * stage_words;
    if (kTMA) {
      if (!(p.debug & 4)) mbar_wait(full + slot, (uint32_t)((c / kPrepStages) & 1));
    } else {
      __syncthreads();                                               // everybody is done with the previous stage
      const uint32_t* src = src0 + (size_t)c0 * mask_words;
      for (int t = tid; t < cn * per_mask; t += ncons) {
        const int k = t / per_mask, o = t - k * per_mask;
        stage[t] = __ldg(src + (size_t)k * mask_words + o);
      }
      __syncthreads();
    }

    // Every lane runs the loop (idle lanes only skip their stores): the warp-level steps use full-warp votes and shuffles.
    const uint32_t* s0 = stage + row_off0;                           // tap rows inside the stage of mask 0
    const uint32_t* s1 = stage + row_off1;
    for (int k = 0; k < cn; ++k, s0 += per_mask, s1 += per_mask, lp += 3 * plane_bytes, gp += 3 * plane_bytes) {
      uint32_t a0 = 0, a1 = 0, in_bits = 0, bnd_bits = 0, codes = 0;   // bit q: pixel q fully inside / ON the outline; !kNarrow: 4-bit tap codes
      if (kNarrow) {
        if (!(p.debug & 8)) {
          a0 = __funnelshift_r(s0[wi], s0[wi + 1], sh);
          a1 = __funnelshift_r(s1[wi], s1[wi + 1], sh);
        }
        const uint32_t X = a0 & a1, Y = a0 | a1;                     // bit set: column inside on both / on either tap row
        const bool all_in = __all_sync(0xffffffffu, (X & tapmask) == tapmask);
        const bool all_out = __all_sync(0xffffffffu, (Y & tapmask) == 0u);
        if (all_in || all_out) {                                     // all 32 groups on one side of the outline (warp-uniform)
          if (live && !(p.debug & 2)) {
            if (all_in) {
#pragma unroll
              for (int ch = 0; ch < 3; ++ch) { ans[0 + ch].store(lp + (size_t)ch * plane_bytes); ans[3 + ch].store(gp + (size_t)ch * plane_bytes); }
            } else {
#pragma unroll
              for (int ch = 0; ch < 3; ++ch) { ans[6 + ch].store(lp + (size_t)ch * plane_bytes); ans[9 + ch].store(gp + (size_t)ch * plane_bytes); }
            }
          }
          continue;
        }
#pragma unroll
        for (int q = 0; q < PX; ++q) {
          const bool in_q = (X & tm[q]) == tm[q];
          in_bits |= (in_q ? 1u : 0u) << q;
          bnd_bits |= ((!in_q && (Y & tm[q]) != 0u) ? 1u : 0u) << q;
        }
      } else {
#pragma unroll
        for (int q = 0; q < PX; ++q) {
          const Taps tx = make_taps(j0 + q, W, S, sc_x);
          const int xa = tx.i0, xb = tx.i0 + tx.d;
          const uint32_t code = ((s0[xa >> 5] >> (xa & 31)) & 1u) | (((s0[xb >> 5] >> (xb & 31)) & 1u) << 1) |
                                (((s1[xa >> 5] >> (xa & 31)) & 1u) << 2) | (((s1[xb >> 5] >> (xb & 31)) & 1u) << 3);
          in_bits |= (code == 15u ? 1u : 0u) << q;
          bnd_bits |= ((code != 15u && code != 0u) ? 1u : 0u) << q;
          codes |= code << (4 * q);
        }
      }
      // groups cut by the outline: per-pixel FG/BG merge with bit masks; outline pixels keep the BG value as a placeholder
      Pack<kBF16, PX> ol[3], og[3];
#pragma unroll
      for (int w = 0; w < Pack<kBF16, PX>::NW; ++w) {
        uint32_t m;                        // all-ones in the lanes of pixels that are fully inside
        if (kBF16) m = (((in_bits >> (2 * w)) & 1u) ? 0x0000ffffu : 0u) | (((in_bits >> (2 * w + 1)) & 1u) ? 0xffff0000u : 0u);
        else m = ((in_bits >> w) & 1u) ? 0xffffffffu : 0u;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
          ol[ch].w[w] = (ans[0 + ch].w[w] & m) | (ans[6 + ch].w[w] & ~m);
          og[ch].w[w] = (ans[3 + ch].w[w] & m) | (ans[9 + ch].w[w] & ~m);
        }
      }
      if (live) {
        if (!(p.debug & 2)) {
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) {
            ol[ch].store(lp + (size_t)ch * plane_bytes);
            og[ch].store(gp + (size_t)ch * plane_bytes);
          }
        } else if (ol[0].w[0] == 0x12345u && og[2].w[1] == 0x54321u) {
          ol[0].store(lp);
        }
      }
      if (!live || (p.debug & 1)) bnd_bits = 0;
      // file the outline pixels.  Tap code: bit t = tap t inside; taps ordered (row0,x0) (row0,x1) (row1,x0) (row1,x1).
      if (__any_sync(0xffffffffu, bnd_bits != 0u)) {
        const uint32_t mine = __popc(bnd_bits);
        uint32_t incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t nb = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += nb;
        }
        const int total = (int)__shfl_sync(0xffffffffu, incl, 31);   // <= 32 * PX <= kPrepWq
        if (wcount + total > kPrepWq) flush();                        // earlier masks only: their placeholders are stored
        int slot_w = wcount + (int)(incl - mine);
#pragma unroll
        for (int q = 0; q < PX; ++q) {                                 // unrolled: tm[] stays in registers
          if (!((bnd_bits >> q) & 1u)) continue;
          uint32_t code;
          if (kNarrow) {
            const int pa = __ffs(tm[q]) - 1, pb = 31 - __clz(tm[q]);   // window bits of the pixel's two tap columns
            code = ((a0 >> pa) & 1u) | (((a0 >> pb) & 1u) << 1) | (((a1 >> pa) & 1u) << 2) | (((a1 >> pb) & 1u) << 3);
          } else {
            code = (codes >> (4 * q)) & 15u;
          }
          wq[slot_w++] = ((uint32_t)(c0 + k) << 16) | ((uint32_t)(lane * PX + q) << 4) | code;
          prefetch_l1(p.taps + tap_px0 + q * 8);              // the record is in L1 by the time the list is flushed
        }
        wcount += total;
      }
      if (wcount && ((k + 1) % p.flush_every) == 0) flush();          // (only for flush_every <= sub)
    }
    if (wcount && (p.flush_every <= kPrepSub || c == nst - 1)) flush();
    if (kTMA) {                                                       // this warp is done with the slot
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + slot);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// crop variant: every proposal is resampled from ITS OWN box (x, y, w, h) instead of the full frame
// ---------------------------------------------------------------------------------------------------------------------
// The reference casts SAM's box per proposal and never uses it (Hybridgl_main.py:101): it always resizes the full frame.  The
// north star names a bounding-box crop, SURVEY 8(b) proposes `crop_xywh` on the prep entry point; this is that option:
//   global[n] = Normalize_IN( bilinear_S( where(m_n, img, bg)[y:y+h, x:x+w] / 255 ) ),  local[n] likewise on the mean-filled view.
// Taps differ per proposal, so the per-image answer planes of prep_main do not apply: one thread evaluates one output pixel
// exactly (the boundary-pixel formula of prep_main, ATen's op order), 4 taps x (1 mask bit + 3 + 3 frame bytes).
template <bool kBF16>
__global__ void __launch_bounds__(256) prep_crop_kernel(const uint8_t* __restrict__ image, const uint8_t* __restrict__ blur,
                                                        const uint32_t* __restrict__ bits, const int32_t* __restrict__ mask_off,
                                                        const int32_t* __restrict__ crop, int B, int H, int W, int S,
                                                        void* __restrict__ local_out, void* __restrict__ global_out) {
  __shared__ float lut[1024];                          // [v] = v/255, [256 + 256*c + v] = Normalize_c(v/255)
  for (int t = threadIdx.x; t < 1024; t += blockDim.x) lut[t] = t < 256 ? to_unit((uint32_t)t) : to_norm((uint32_t)(t & 255), (t >> 8) - 1);
  __syncthreads();
  const int m = blockIdx.y, WW = (W + 31) >> 5, SS = S * S;
  int b = 0;
  if (mask_off) { while (b + 1 < B && mask_off[b + 1] <= m) ++b; }
  const int cx = crop[4 * m], cy = crop[4 * m + 1], cw = crop[4 * m + 2], ch = crop[4 * m + 3];
  const float sc_y = tap_scale(ch, S), sc_x = tap_scale(cw, S);
  const uint8_t* img = image + (size_t)b * H * W * 3;
  const uint8_t* bg = blur ? blur + (size_t)b * H * W * 3 : nullptr;
  const uint32_t* mb = bits + (size_t)m * H * WW;
  for (int px = blockIdx.x * blockDim.x + threadIdx.x; px < SS; px += gridDim.x * blockDim.x) {
    const int i = px / S, j = px - i * S;
    const Taps ty = make_taps(i, ch, S, sc_y), tx = make_taps(j, cw, S, sc_x);
    const int ys[2] = {cy + ty.i0, cy + ty.i0 + ty.d}, xs[2] = {cx + tx.i0, cx + tx.i0 + tx.d};
    float gv[3][4], lv[3][4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int yy = ys[t >> 1], xx = xs[t & 1];
      const bool in = (__ldg(mb + (size_t)yy * WW + (xx >> 5)) >> (xx & 31)) & 1u;
      const size_t o = ((size_t)yy * W + xx) * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const uint32_t vi = __ldg(img + o + c);
        const uint32_t vb = bg ? (uint32_t)__ldg(bg + o + c) : 0u;
        gv[c][t] = lut[in ? vi : vb];
        lv[c][t] = in ? lut[256 + 256 * c + vi] : c_clip_mean[c];
      }
    }
    const size_t o = (size_t)m * 3 * SS + px;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float g = __fdiv_rn(__fsub_rn(bilerp(gv[c][0], gv[c][1], gv[c][2], gv[c][3], tx.w0, tx.w1, ty.w0, ty.w1), c_in_mean[c]), c_in_std[c]);
      const float l = bilerp(lv[c][0], lv[c][1], lv[c][2], lv[c][3], tx.w0, tx.w1, ty.w0, ty.w1);
      if (kBF16) {
        reinterpret_cast<__nv_bfloat16*>(local_out)[o + (size_t)c * SS] = __float2bfloat16_rn(l);
        reinterpret_cast<__nv_bfloat16*>(global_out)[o + (size_t)c * SS] = __float2bfloat16_rn(g);
      } else {
        reinterpret_cast<float*>(local_out)[o + (size_t)c * SS] = l;
        reinterpret_cast<float*>(global_out)[o + (size_t)c * SS] = g;
      }
    }
  }
}

struct PrepWs {
  void* planes;
  uint32_t* taps;
  float* lut;
  size_t bytes;
};
static PrepWs prep_carve(void* ws, int B, int S, int out_dtype) {
  PrepWs w;
  const size_t SS = (size_t)S * S, elem = out_dtype == HGL_BF16 ? 2 : 4;
  size_t off = 0;
  auto take = [&](size_t n) { size_t o = off; off += (n + 255) & ~size_t(255); return o; };
  uint8_t* base = reinterpret_cast<uint8_t*>(ws);
  w.planes = base + take((size_t)B * 12 * SS * elem);
  w.taps = reinterpret_cast<uint32_t*>(base + take((size_t)B * 8 * SS * 4));
  w.lut = reinterpret_cast<float*>(base + take(1024 * 4));
  w.bytes = off;
  return w;
}

}  // namespace hgl

static int pack_masks_impl(const uint8_t* masks, int M, int H, int W, uint32_t* bits, bool is_bool, void* stream) {
  using namespace hgl;
  if (M == 0) return HGL_OK;
  HGL_REQUIRE(masks && bits, "hgl_pack_masks: null pointer");
  HGL_REQUIRE(M > 0 && H >= 1 && W >= 1, "hgl_pack_masks: bad shape");
  const size_t rows = (size_t)M * H;
  cudaStream_t st = (cudaStream_t)stream;
  if ((W & 31) == 0 && (reinterpret_cast<uintptr_t>(masks) & 15) == 0) {
    const size_t n16 = rows * W / 16;
    // persistent grid-stride CTAs.  HBM is saturated by the bytes in flight (8 x 16 B per thread), not by the thread count, so
    // the kernel only takes a quarter of every SM's thread slots (measured fastest alone, too): kernels of a concurrent stream (blur, prep setup, heat-map
    // tables in the batched pipeline) then run beside it instead of queueing behind a grid that fills the machine.
    int per_sm = 2;
#ifdef HGL_TUNING
    if (const char* pv = getenv("HGL_PACK_CTAS_PER_SM")) per_sm = std::max(1, std::min(8, atoi(pv)));
#endif
    const int blocks = std::max(1, (int)std::min<size_t>((n16 + 256 * 8 - 1) / (256 * 8), (size_t)sm_count() * per_sm));
    if (is_bool) pack_masks_flat_kernel<true><<<blocks, 256, 0, st>>>(reinterpret_cast<const uint4*>(masks), n16, bits);
    else pack_masks_flat_kernel<false><<<blocks, 256, 0, st>>>(reinterpret_cast<const uint4*>(masks), n16, bits);
  } else {
    const size_t total = rows * ((W + 31) >> 5);
    const int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)sm_count() * 16);
    if (is_bool) pack_masks_rows_kernel<true><<<blocks, 256, 0, st>>>(masks, rows, W, bits);
    else pack_masks_rows_kernel<false><<<blocks, 256, 0, st>>>(masks, rows, W, bits);
  }
  return launch_status("hgl_pack_masks");
}

extern "C" int hgl_pack_masks(const uint8_t* masks, int M, int H, int W, uint32_t* bits, void* stream) {
  return pack_masks_impl(masks, M, H, W, bits, false, stream);
}

extern "C" int hgl_pack_masks_bool(const uint8_t* masks, int M, int H, int W, uint32_t* bits, void* stream) {
  return pack_masks_impl(masks, M, H, W, bits, true, stream);
}

extern "C" int64_t hgl_prep_workspace_bytes(int B, int S, int out_dtype) {
  if (B < 1 || S < 4) return -1;
  return (int64_t)hgl::prep_carve(nullptr, B, S, out_dtype).bytes;
}

// per-image half: needs the frames only, not the masks -- a caller may enqueue it while the masks are still being packed
extern "C" int hgl_prep_setup(const uint8_t* image, const uint8_t* blur, int B, int H, int W, int S, int bg_mode, int out_dtype,
                              void* workspace, void* stream) {
  using namespace hgl;
  HGL_REQUIRE(image && workspace, "hgl_prep: null pointer");
  HGL_REQUIRE(bg_mode == HGL_BG_BLUR || bg_mode == HGL_BG_BLACK || bg_mode == HGL_BG_NONE, "hgl_prep: bg_mode %d", bg_mode);
  HGL_REQUIRE(bg_mode != HGL_BG_BLUR || blur, "hgl_prep: blur frame required for HGL_BG_BLUR");
  HGL_REQUIRE(out_dtype == HGL_F32 || out_dtype == HGL_BF16, "hgl_prep: out_dtype %d", out_dtype);
  HGL_REQUIRE(B >= 1 && H >= 1 && W >= 1, "hgl_prep: bad shape B=%d H=%d W=%d", B, H, W);
  HGL_REQUIRE(S >= 4 && S % 4 == 0 && S <= 1024, "hgl_prep: S=%d must be a multiple of 4 in [4,1024]", S);
  HGL_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "hgl_prep: workspace must be 256-byte aligned");
  HGL_REQUIRE(B <= 65535, "hgl_prep: batch too large for one launch (B=%d)", B);
  cudaStream_t st = (cudaStream_t)stream;
  const int SS = S * S;
  PrepWs ws = prep_carve(workspace, B, S, out_dtype);
  const uint8_t* bgp = bg_mode == HGL_BG_BLUR ? blur : bg_mode == HGL_BG_NONE ? image : nullptr;   // NONE: the frame is its own background
  if (out_dtype == HGL_BF16)
    prep_setup_kernel<true><<<dim3(ceil_div(SS, 256), B), 256, 0, st>>>(image, bgp, H, W, S, ws.planes, ws.taps, ws.lut);
  else
    prep_setup_kernel<false><<<dim3(ceil_div(SS, 256), B), 256, 0, st>>>(image, bgp, H, W, S, ws.planes, ws.taps, ws.lut);
  return launch_status("hgl_prep(setup)");
}

extern "C" int hgl_prep(const uint8_t* image, const uint8_t* blur, const uint32_t* bits, const int32_t* mask_off,
                        int B, int M, int max_n, int H, int W, int S, int bg_mode, int out_dtype,
                        void* local_out, void* global_out, void* workspace, void* stream) {
  if (M == 0 && B >= 1) return HGL_OK;   // nothing to do (empty tensors have null data pointers)
  int rc = hgl_prep_setup(image, blur, B, H, W, S, bg_mode, out_dtype, workspace, stream);
  if (rc != HGL_OK) return rc;
  return hgl_prep_main(bits, mask_off, B, M, max_n, H, W, S, out_dtype, local_out, global_out, workspace, stream);
}

// per-mask half: streams the packed masks over the answer planes hgl_prep_setup left in `workspace`
extern "C" int hgl_prep_main(const uint32_t* bits, const int32_t* mask_off, int B, int M, int max_n, int H, int W, int S, int out_dtype,
                             void* local_out, void* global_out, void* workspace, void* stream) {
  using namespace hgl;
  if (M == 0 && B >= 1) return HGL_OK;
  HGL_REQUIRE(bits && local_out && global_out && workspace, "hgl_prep: null pointer");
  HGL_REQUIRE((uint64_t)B * (uint64_t)S * S * 8 < (1ull << 32), "hgl_prep: B=%d images of %dx%d outputs exceed the 32-bit tap-record offsets", B, S, S);
  HGL_REQUIRE(out_dtype == HGL_F32 || out_dtype == HGL_BF16, "hgl_prep: out_dtype %d", out_dtype);
  HGL_REQUIRE(B >= 1 && M >= 0 && H >= 1 && W >= 1 && max_n >= 1, "hgl_prep: bad shape B=%d M=%d H=%d W=%d max_n=%d", B, M, H, W, max_n);
  HGL_REQUIRE(S >= 4 && S % 4 == 0 && S <= 1024, "hgl_prep: S=%d must be a multiple of 4 in [4,1024]", S);
  HGL_REQUIRE(mask_off || B == 1, "hgl_prep: mask_off required when B > 1");
  HGL_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "hgl_prep: workspace must be 256-byte aligned");
  HGL_REQUIRE(((reinterpret_cast<uintptr_t>(local_out) | reinterpret_cast<uintptr_t>(global_out)) & 15) == 0,
              "hgl_prep: outputs must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  PrepWs ws = prep_carve(workspace, B, S, out_dtype);

  PrepParams p;
  p.bits = bits; p.mask_off = mask_off; p.planes = ws.planes; p.taps = ws.taps; p.lut = ws.lut;
  p.local_out = local_out; p.global_out = global_out;
  p.B = B; p.M = M; p.H = H; p.W = W; p.S = S; p.WW = (W + 31) >> 5;
  const double sx = (double)W / (double)S;
  // pixels per thread: 8 (16-byte bf16 stores) when the 16 taps still fit one 32-bit window, else 4
  const int px = (out_dtype == HGL_BF16 && S % 8 == 0 && (int)(7.0 * sx) + 3 <= 31) ? 8 : kPrepPx;
  p.narrow = ((int)((px - 1) * sx) + 3 <= 31) ? 1 : 0;
  // thread -> pixel map: patch = gw groups x gh rows per warp, cw patches side by side per CTA (see PrepGeom)
  const int G = S / px;                                                       // groups per output row
  int gw = 1;
  while (gw < 8 && G % (gw * 2) == 0) gw *= 2;
  const int gh = 32 / gw, ppr = G / gw;
  // 1-D bulk copies need 16-byte aligned masks (H * WW % 4 == 0 and an aligned base; a band's first row may sit anywhere,
  // the copy then starts up to 3 words early); otherwise cooperative loads
  const bool tma = (((size_t)H * p.WW) % 4 == 0) && ((reinterpret_cast<uintptr_t>(bits) & 15) == 0) && !tuning_flag("HGL_PREP_NO_TMA");
  int cw = 1;
  for (int d = 1; d <= (tma ? 7 : 8); ++d) if (ppr % d == 0) cw = d;        // kTMA: one more warp (the producer) joins the CTA
  p.gw = gw; p.gh = gh; p.cw = cw; p.nbx = ppr / cw; p.strip = 0;
  // strip mode (default whenever a row's groups fit one CTA): a CTA owns R whole output rows and its lanes walk the band's
  // groups in row-major order, so every warp-wide store is one contiguous, 128-byte aligned 512-byte run of a plane (full
  // lines; measured 6.2 TB/s of pure stores against 5.1 TB/s for 2-D patches of 64-byte row pieces).  The last warp may
  // have idle lanes when R * G is not a multiple of 32.
  const int max_cons = tma ? kPrepThreads - 32 : kPrepThreads;
  if (G <= max_cons && !tuning_flag("HGL_PREP_PATCH")) {
    const int R = max_cons / G;
    p.strip = G; p.gh = R; p.cw = ceil_div(R * G, 32); p.nbx = 1; p.gw = 1;
  }
  const int gh_ = p.gh; cw = p.cw;
  const int nby = ceil_div(S, gh_);
  const int threads = 32 * cw + (tma ? 32 : 0);
  // the 2*px tap columns of a lane must fit one 32-bit window for the fast path
  // grid: (band tiles, images, z) -- z splits an image's masks so that the launch fills whole waves of resident CTAs
  const int gx = p.nbx * nby;
  const int per_image = (B == 1) ? M : std::min(max_n, M);
  const double sy = (double)H / (double)S;
  const int stage_rows = std::min(H, (int)(gh_ * sy) + 3);                    // source rows behind a band
  const size_t per_mask_bytes = (size_t)((stage_rows * p.WW + 3 + 3) & ~3) * 4;   // slot of one mask: rows + lead words, whole 16-byte units
  // barriers | per-warp scratch | stages (+ the word after the last row)
  const size_t fixed_smem = kPrepBarBytes + kPrepLutBytes + (size_t)cw * kPrepWarpBytes + 16;
  const size_t stage_budget = fixed_smem + 24 * 1024 <= 112 * 1024 ? 112 * 1024 - fixed_smem      // two CTAs per SM
                                                                     : (fixed_smem < 200 * 1024 ? 224 * 1024 - fixed_smem : 0);
  int kPrepSub = std::max(1, std::min(kPrepMaxSub, tuning_int("HGL_PREP_SUB", kPrepSubDefault)));
  while (kPrepSub > 1 && (size_t)kPrepStages * kPrepSub * per_mask_bytes > std::min<size_t>(stage_budget, 96 * 1024)) kPrepSub /= 2;
  p.sub = kPrepSub;
  HGL_REQUIRE((size_t)kPrepStages * kPrepSub * per_mask_bytes <= stage_budget, "hgl_prep: frame %dx%d (S=%d) too large for the shared-memory stages",
              H, W, S);
  p.stage_words = (int)(kPrepSub * per_mask_bytes / 4);
  // (HGL_PREP_SMEM_PAD_KB, profiling build: extra dynamic shared memory, e.g. 80 -> one CTA per SM, to measure prep at half residency)
  const size_t smem = fixed_smem + (size_t)kPrepStages * kPrepSub * per_mask_bytes + (size_t)tuning_int("HGL_PREP_SMEM_PAD_KB", 0) * 1024;
  const int resident = std::max(1, std::min(std::min(2 * kPrepThreads / threads, (int)((220 * 1024) / (smem + 1024))), 65536 / (threads * 128)));
  const long slots = (long)sm_count() * resident;
  int gz = 1;
  {
    double best = -1.0;
    const int gz_max = std::max(1, std::min(64, per_image / (2 * kPrepSub)));     // at least two stages per CTA
    for (int z = 1; z <= gz_max; ++z) {
      const long ctas = (long)gx * B * z;
      const double eff = (double)ctas / (double)(((ctas + slots - 1) / slots) * slots) - 0.004 * z;   // wave fill, mild bias to few splits
      if (eff > best) { best = eff; gz = z; }
    }
  }
  gz = std::max(1, tuning_int("HGL_PREP_GZ", gz));
  p.debug = 0;
  // outline lists are flushed as rarely as possible: when a warp's list is full, and after the CTA's last mask (a CTA lives ~15 us,
  // so the lines it patches are still in L2).  Dense flushes cost less than frequent ones: every 1 / 2 / 4 masks 224 / 210 / 204 us
  // alone, once per CTA 203 us and 0.466 instead of 0.474 ms per pass (profiles/r2_variants.md)
  p.flush_every = std::max(1, tuning_int("HGL_PREP_FLUSH", 16 * kPrepMaxSub));
#ifdef HGL_TUNING
  p.debug = tuning_int("HGL_PREP_DEBUG", 0);       // profiling builds only (results are wrong when set): never in the shipped library
#endif
  p.gz = gz;
  dim3 grid(gx * gz, B, 1);
  HGL_REQUIRE(grid.y <= 65535, "hgl_prep: batch too large for one launch (B=%d, max_n=%d)", B, max_n);
  int smem_rc = HGL_OK;
  auto launch = [&](auto kern) {
    smem_rc = ensure_dyn_smem(reinterpret_cast<const void*>(kern), smem, "hgl_prep(main)");
    if (smem_rc == HGL_OK) kern<<<grid, threads, smem, st>>>(p);
  };
  HGL_REQUIRE(per_image <= 65535 * gz, "hgl_prep: more than 65535 masks of one image per CTA span");
  const bool nw = p.narrow != 0;
  if (out_dtype == HGL_BF16) {
    if (px == 8) { if (tma) launch(prep_main_kernel<true, 8, true, true>); else launch(prep_main_kernel<true, 8, false, true>); }
    else if (nw) { if (tma) launch(prep_main_kernel<true, kPrepPx, true, true>); else launch(prep_main_kernel<true, kPrepPx, false, true>); }
    else { if (tma) launch(prep_main_kernel<true, kPrepPx, true, false>); else launch(prep_main_kernel<true, kPrepPx, false, false>); }
  } else {
    if (nw) { if (tma) launch(prep_main_kernel<false, kPrepPx, true, true>); else launch(prep_main_kernel<false, kPrepPx, false, true>); }
    else { if (tma) launch(prep_main_kernel<false, kPrepPx, true, false>); else launch(prep_main_kernel<false, kPrepPx, false, false>); }
  }
  if (smem_rc != HGL_OK) return smem_rc;
  return launch_status("hgl_prep(main)");
}

// per-proposal crop boxes: see prep_crop_kernel.  crop_xywh int32 [M,4] (x, y, w, h), every box inside its frame (w, h >= 1);
// the boxes are read on the device, so their validity is the caller's contract (ops.prep_visual_prompts checks it on the host side
// when the boxes are at hand).  No workspace.
extern "C" int hgl_prep_crop(const uint8_t* image, const uint8_t* blur, const uint32_t* bits, const int32_t* mask_off, const int32_t* crop_xywh,
                             int B, int M, int H, int W, int S, int bg_mode, int out_dtype, void* local_out, void* global_out, void* stream) {
  using namespace hgl;
  if (M == 0 && B >= 1) return HGL_OK;
  HGL_REQUIRE(image && bits && crop_xywh && local_out && global_out, "hgl_prep_crop: null pointer");
  HGL_REQUIRE(bg_mode == HGL_BG_BLUR || bg_mode == HGL_BG_BLACK || bg_mode == HGL_BG_NONE, "hgl_prep_crop: bg_mode %d", bg_mode);
  HGL_REQUIRE(bg_mode != HGL_BG_BLUR || blur, "hgl_prep_crop: blur frame required for HGL_BG_BLUR");
  HGL_REQUIRE(out_dtype == HGL_F32 || out_dtype == HGL_BF16, "hgl_prep_crop: out_dtype %d", out_dtype);
  HGL_REQUIRE(B >= 1 && M >= 0 && H >= 1 && W >= 1 && S >= 1 && S <= 4096, "hgl_prep_crop: bad shape");
  HGL_REQUIRE(mask_off || B == 1, "hgl_prep_crop: mask_off required when B > 1");
  HGL_REQUIRE(M <= 65535, "hgl_prep_crop: more than 65535 proposals per launch");
  const uint8_t* bgp = bg_mode == HGL_BG_BLUR ? blur : bg_mode == HGL_BG_NONE ? image : nullptr;   // NONE: the frame is its own background
  dim3 grid(std::min(ceil_div(S * S, 256), 64), M);
  cudaStream_t st = (cudaStream_t)stream;
  if (out_dtype == HGL_BF16) prep_crop_kernel<true><<<grid, 256, 0, st>>>(image, bgp, bits, mask_off, crop_xywh, B, H, W, S, local_out, global_out);
  else prep_crop_kernel<false><<<grid, 256, 0, st>>>(image, bgp, bits, mask_off, crop_xywh, B, H, W, S, local_out, global_out);
  return launch_status("hgl_prep_crop");
}
