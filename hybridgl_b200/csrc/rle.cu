// SAM run-length masks -> the packed-mask format, on the device.
//
// The reference receives its proposals from SamAutomaticMaskGenerator.generate (Hybridgl_main.py:84-87).  Inside SAM every
// proposal lives as an *uncompressed RLE* (third_party/segment-anything/segment_anything/utils/amg.py:107-135
// mask_to_rle_pytorch: column-major runs, first run counts zeros) and is only expanded to one byte per pixel on the host by
// rle_to_mask (amg.py:138-149) when output_mode == "binary_mask".  hgl_rle_to_bits takes the RLE form directly
// (output_mode="uncompressed_rle") and writes the same bits hgl_pack_masks would write for rle_to_mask(rle): the H*W-byte
// mask never exists, neither on the host (H2D shrinks ~250x) nor in HBM (no 1 B/pixel read pass).
//
// One CTA per mask (per column strip for frames that do not fit in shared memory):
//   1. toggles : the exclusive prefix sum of the counts gives the flat column-major position where every run starts; each
//                start flips one bit of a column-major bitmap in shared memory (atomicXor; zero-length runs cancel out).
//   2. fill    : a prefix-XOR over that bitmap (in-word shift/xor ladder + a block scan of word parities) turns toggles
//                into pixels.
//   3. turn    : 32x32 bit tiles are transposed with the 5-stage shuffle butterfly (empty tiles are skipped), giving
//                row-major words in shared memory.
//   4. store   : rows leave as 16-byte coalesced stores.
// Steps 2-3 only touch the columns between the first and the last toggle (a proposal covers ~1/4 of the frame's columns on
// average); the tile columns outside that range are written as zeros without ever being staged.
// HBM traffic is the packed output (M*H*ceil(W/32)*4 B) plus the counts; everything else is shared-memory work.
#include "hgl_common.cuh"

namespace hgl {

#ifndef HGL_RLE_THREADS
#define HGL_RLE_THREADS 512     // 16 warps per mask: 52 us for the bench batch against 59 us with 256 threads (phases are separated by CTA barriers)
#endif
constexpr int kRleThreads = HGL_RLE_THREADS;

__device__ __forceinline__ uint32_t prefix_xor32(uint32_t v) {   // bit j = XOR of bits 0..j
  v ^= v << 1; v ^= v << 2; v ^= v << 4; v ^= v << 8; v ^= v << 16;
  return v;
}

// in-place transpose of a 32x32 bit tile held one row per lane: on return lane k holds column k (bit j = old row j's bit k)
__device__ __forceinline__ uint32_t transpose32(uint32_t v, int lane) {
  uint32_t m = 0x0000ffffu;
#pragma unroll
  for (int j = 16; j != 0; j >>= 1, m ^= (m << j)) {
    const uint32_t p = __shfl_xor_sync(0xffffffffu, v, j);
    // pair (lo lane a, hi lane b = a + j): swap a's bits [j..] block with b's bits [..j) block under mask m
    if ((lane & j) == 0) v = (v & m) | ((p & m) << j);
    else v = (v & ~m) | ((p >> j) & m);
  }
  return v;
}

// smem layout: col[XS * HPW] (column-major toggles -> pixels; HPW odd), row[H * RS] (row-major words of the strip; RS odd)
__global__ void __launch_bounds__(kRleThreads) rle_to_bits_kernel(const int32_t* __restrict__ counts, const int32_t* __restrict__ rle_off,
                                                                  int H, int W, int WW, int HPW, int XS, int RS,
                                                                  uint32_t* __restrict__ bits) {
  extern __shared__ __align__(16) uint32_t sm[];
  uint32_t* col = sm;
  uint32_t* row = sm + (size_t)XS * HPW;
  __shared__ int s_warp[kRleThreads / 32];
  __shared__ int s_xmin, s_xmax;
  __shared__ uint32_t s_carry, s_flips;
  const int m = blockIdx.x, strip = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int x0 = strip * XS, x1 = min(W, x0 + XS);          // columns of this strip
  const int64_t p_lo = (int64_t)x0 * H, p_hi = (int64_t)x1 * H;
  const int NW = XS * HPW;                                   // XS is a multiple of 32: NW % 4 == 0

  for (int i = tid; i < NW / 4; i += kRleThreads) reinterpret_cast<uint4*>(col)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (tid == 0) { s_carry = 0u; s_flips = 0u; s_xmin = XS; s_xmax = -1; }
  __syncthreads();

  // ---- 1. toggles ---------------------------------------------------------------------------------------------------
  const int r_lo = rle_off[m], r_hi = rle_off[m + 1];
  uint32_t my_carry = 0u, my_flips = 0u;
  int xmin = XS, xmax = -1;                                  // strip-local columns that received a toggle
  int64_t base = 0;                                          // position where the chunk's first run starts (same in every thread)
  for (int r0 = r_lo; r0 < r_hi; r0 += kRleThreads) {
    const int r = r0 + tid;
    const int c = r < r_hi ? counts[r] : 0;
    int inc = c;                                              // inclusive warp scan
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    int64_t start = base + (inc - c);                         // position where run r starts
    int tot = 0;
#pragma unroll
    for (int w = 0; w < kRleThreads / 32; ++w) {
      const int v = s_warp[w];
      if (w < warp) start += v;
      tot += v;
    }
    if (r < r_hi && r > r_lo) {                               // run 0 starts at parity 0: no toggle
      if (start < p_lo) my_carry ^= 1u;
      else if (start < p_hi) {
        const int q = (int)(start - p_lo);
        const int x = q / H, y = q - x * H;
        atomicXor(&col[x * HPW + (y >> 5)], 1u << (y & 31));
        xmin = min(xmin, x); xmax = max(xmax, x);
        my_flips ^= 1u;
      }
    }
    base += tot;
    __syncthreads();                                          // s_warp is rewritten by the next chunk
    if (base >= p_hi) break;                                  // uniform: every later run starts right of the strip
  }
  my_carry = __ballot_sync(0xffffffffu, my_carry & 1u);
  my_flips = __ballot_sync(0xffffffffu, my_flips & 1u);
  xmin = __reduce_min_sync(0xffffffffu, xmin); xmax = __reduce_max_sync(0xffffffffu, xmax);
  if (lane == 0) {
    if (__popc(my_carry) & 1) atomicXor(&s_carry, 1u);
    if (__popc(my_flips) & 1) atomicXor(&s_flips, 1u);
    if (xmax >= 0) { atomicMin(&s_xmin, xmin); atomicMax(&s_xmax, xmax); }
  }
  __syncthreads();
  // columns that can hold pixels: from the first toggle (or the strip start when a run of ones enters the strip) to the
  // last toggle (or the strip end when a run of ones leaves it); everything else is zero and never touched below
  const uint32_t carry = s_carry;
  int xa = carry ? 0 : s_xmin, xb = (carry ^ s_flips) ? (x1 - x0 - 1) : s_xmax;
  const int XSW = (x1 - x0 + 31) >> 5, HB = (H + 31) >> 5;
  int twa = 0, twb = -1;                                      // tile columns [twa, twb] of the strip that are not all-zero
  if (xa <= xb) {
    twa = xa >> 5; twb = xb >> 5;

    // ---- 2. fill: prefix-XOR over the flat bitmap of columns xa..xb (padding bits y >= H carry the state through) ------
    const int f_lo = xa * HPW, f_n = (xb - xa + 1) * HPW;
    const int seg = ((f_n + kRleThreads - 1) / kRleThreads) | 1;   // odd stride between lanes: conflict-free
    const int w_lo = f_lo + min(f_n, tid * seg), w_hi = f_lo + min(f_n, (tid + 1) * seg);
    uint32_t par = 0u;
    for (int i = w_lo; i < w_hi; ++i) par ^= col[i];
    par = __popc(par) & 1u;
    const uint32_t wb = __ballot_sync(0xffffffffu, par);
    if (lane == 0) s_warp[warp] = __popc(wb) & 1;
    __syncthreads();
    uint32_t state = carry ^ (__popc(wb & ((1u << lane) - 1u)) & 1u);
    for (int w = 0; w < warp; ++w) state ^= (uint32_t)s_warp[w];
    for (int i = w_lo; i < w_hi; ++i) {
      uint32_t v = prefix_xor32(col[i]);
      if (state) v = ~v;
      col[i] = v;
      state = v >> 31;
    }
    __syncthreads();

    // ---- 3. turn: 32x32 tiles, lane = column on the way in, lane = row on the way out -------------------------------------
    const int ntile = (twb - twa + 1) * HB;
    int xw = twa, yb = warp;
    while (yb >= HB) { yb -= HB; ++xw; }
    for (int t = warp; t < ntile; t += kRleThreads / 32) {
      const int x = xw * 32 + lane;
      uint32_t v = (x >= xa && x <= xb) ? col[x * HPW + yb] : 0u;   // columns outside [xa, xb] hold no pixels (or carried state)
      if (__any_sync(0xffffffffu, v != 0u)) v = transpose32(v, lane);
      const int y = yb * 32 + lane;
      if (y < H) row[y * RS + xw] = v;
      yb += kRleThreads / 32;
      while (yb >= HB) { yb -= HB; ++xw; }
    }
    __syncthreads();
  }

  // ---- 4. store (tile columns outside [twa, twb] are zeros and never staged) ----------------------------------------------
  uint32_t* dst = bits + (size_t)m * H * WW + (x0 >> 5);
  if (XSW == WW && (WW & 3) == 0 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
    const int Q = WW >> 2;
    int y = tid / Q, q = tid - y * Q;                         // kRleThreads / Q rows per sweep
    const int dy = kRleThreads / Q, dq = kRleThreads - dy * Q;
    for (; y < H;) {
      const int j = q * 4;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (j + 3 >= twa && j <= twb) {
        const uint32_t* s = row + y * RS + j;
        v.x = (j >= twa && j <= twb) ? s[0] : 0u;
        v.y = (j + 1 >= twa && j + 1 <= twb) ? s[1] : 0u;
        v.z = (j + 2 >= twa && j + 2 <= twb) ? s[2] : 0u;
        v.w = (j + 3 >= twa && j + 3 <= twb) ? s[3] : 0u;
      }
      stg_stream(reinterpret_cast<uint4*>(dst + (size_t)y * WW + j), v);
      y += dy; q += dq;
      if (q >= Q) { q -= Q; ++y; }
    }
  } else {
    for (int i = tid; i < H * XSW; i += kRleThreads) {
      const int y = i / XSW, j = i - y * XSW;
      dst[(size_t)y * WW + j] = (j >= twa && j <= twb) ? row[y * RS + j] : 0u;
    }
  }
}

static int rle_geometry(int H, int W, int* HPW, int* XS, int* RS, size_t* smem) {
  const int hpw = ((H + 31) / 32) | 1;
  const int WW = (W + 31) / 32;
  // widest strip (multiple of 32 columns) whose two bitmaps fit in 200 KB of shared memory
  int xsw = WW;
  for (;;) {
    const size_t need = ((size_t)xsw * 32 * hpw + (size_t)H * (xsw | 1)) * 4;
    if (need <= 200 * 1024 || xsw == 1) { *smem = need; break; }
    xsw = (xsw + 1) / 2;
  }
  *HPW = hpw; *XS = xsw * 32; *RS = xsw | 1;
  return *smem <= 200 * 1024 ? HGL_OK : HGL_EINVAL;
}

}  // namespace hgl

extern "C" int hgl_rle_to_bits(const int32_t* counts, const int32_t* rle_off, int M, int H, int W, uint32_t* bits, void* stream) {
  using namespace hgl;
  HGL_REQUIRE(M >= 0 && H >= 1 && W >= 1 && (int64_t)H * W < (int64_t)1 << 31, "hgl_rle_to_bits: bad shape");
  if (M == 0) return HGL_OK;
  HGL_REQUIRE(counts && rle_off && bits, "hgl_rle_to_bits: null pointer");
  int HPW, XS, RS;
  size_t smem;
  HGL_REQUIRE(rle_geometry(H, W, &HPW, &XS, &RS, &smem) == HGL_OK, "hgl_rle_to_bits: frame height %d too large", H);
  {
    const int rc0 = ensure_dyn_smem(reinterpret_cast<const void*>(rle_to_bits_kernel), smem, "hgl_rle_to_bits");
    if (rc0 != HGL_OK) return rc0;
  }
  const int WW = (W + 31) / 32;
  const int strips = ceil_div(W, XS);
  rle_to_bits_kernel<<<dim3(M, strips), kRleThreads, smem, (cudaStream_t)stream>>>(counts, rle_off, H, W, WW, HPW, XS, RS, bits);
  return launch_status("hgl_rle_to_bits");
}
