// cv2.GaussianBlur(img, (15,15), 0) on uint8 with OpenCV's fixed-point arithmetic  --  Hybridgl_main.py:99.
//
// sigma = 0.3*((15-1)*0.5-1)+0.8 = 2.6; OpenCV quantises the taps to Q8 {1,3,6,12,20,30,36,40,36,...} (sum 256),
// runs the row pass in Q8 (fits u16), the column pass in Q16 (u32) and rounds half-up once: (v + 2^15) >> 16.
// Border = BORDER_REFLECT_101.  Restated (and checked bit-exact against cv2 4.13) in oracle/hybridgl_oracle.py.
//
// B200 design: integer, exact, and instruction-bound rather than HBM-bound (frames are < 1 MB each), so the work per
// output is what matters.  A CTA owns a 64x32 tile; the halo tile is de-interleaved into three byte planes in shared
// memory, then
//   row pass   4 adjacent outputs share 5 aligned words; each output is 5 dp4a (u8 x u8 taps, 4 products per instruction)
//              instead of 15 multiply-adds.  A thread does two rows and writes the (row, row+1) u16 pairs as one 16-byte store.
//   col pass   a thread owns one column: 23 pair-words in registers, each output is 8 dp2a (u16 x u8, 2 products each).
//   store      outputs are re-interleaved through shared memory and leave as coalesced 32-bit words.
#include "hgl_common.cuh"

namespace hgl {

constexpr int kTX = 64, kTY = 32;   // output tile (pixels)
constexpr int kBR = 7;              // kernel radius
constexpr int kInW = kTX + 2 * kBR; // 78 input pixels per tile row
constexpr int kInH = kTY + 2 * kBR; // 46 input rows
constexpr int kPlaneStride = 80;    // bytes per plane row (word multiple)
constexpr int kPairs = kInH / 2;    // 23 vertical (row, row+1) pairs
constexpr int kCols = kTX * 3;      // 192 (plane, x) columns

__host__ __device__ constexpr uint32_t gq(int k) {   // Q8 taps of getGaussianKernel(15, 0), zero outside
  constexpr uint32_t q[15] = {1, 3, 6, 12, 20, 30, 36, 40, 36, 30, 20, 12, 6, 3, 1};
  return (k >= 0 && k < 15) ? q[k] : 0u;
}
// row pass: coefficient quad of word wi (bytes 4wi..4wi+3 of the 20-byte window) for output j of the group (tap = byte - j)
// (plane column k holds pixel x0 - 8 + k, so byte b of the window is tap b - 1 - j of output j)
__host__ __device__ constexpr uint32_t row_coef(int j, int wi) {
  return gq(4 * wi - 1 - j) | (gq(4 * wi - j) << 8) | (gq(4 * wi + 1 - j) << 16) | (gq(4 * wi + 2 - j) << 24);
}
// col pass: coefficient pair p for an output row of parity `odd` (taps 2p, 2p+1 for even rows; 2p-1, 2p for odd rows)
__host__ __device__ constexpr uint32_t col_coef(int odd, int p) {
  return odd ? (gq(2 * p - 1) | (gq(2 * p) << 8)) : (gq(2 * p) | (gq(2 * p + 1) << 8));
}

__device__ __forceinline__ int reflect101(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return min(max(i, 0), n - 1);      // only positions that feed out-of-frame outputs of a partial tile are clamped
}

__global__ void __launch_bounds__(256) blur15_kernel(const uint8_t* __restrict__ img, uint8_t* __restrict__ out, int H, int W) {
  __shared__ __align__(16) uint8_t plane[3][kInH][kPlaneStride];     // 11040 B
  __shared__ __align__(16) uint32_t rsum[3][kPairs][kTX];            // 17664 B   u16 row sums of rows (2p, 2p+1)
  __shared__ __align__(16) uint8_t obuf[kTY][kCols];                 //  6144 B   interleaved outputs
  const int b = blockIdx.z, y0 = blockIdx.y * kTY, x0 = blockIdx.x * kTX;
  const uint8_t* src = img + (size_t)b * H * W * 3;
  const int tid = threadIdx.x;

  // ---- load + de-interleave (reflect-101 border).  Plane column k holds pixel x0 - 8 + k (column 0 is only there so that
  //      groups of 4 pixels = 12 interleaved bytes = 3 aligned words line up with the plane's words).
  const bool interior = x0 >= 8 && x0 + kTX + 8 <= W && (W & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 3) == 0;   // block-uniform
  if (interior) {
    // 4 pixels per item: three 32-bit loads, six byte permutes, three 32-bit shared-memory stores
    constexpr int kGroups = (kTX + 16) / 4;                              // 20 groups cover pixels x0-8 .. x0+71
    for (int t = tid; t < kInH * kGroups; t += 256) {
      const int ty = t / kGroups, gq4 = t - ty * kGroups;
      const int y = reflect101(y0 + ty - kBR, H);
      const uint32_t* wp = reinterpret_cast<const uint32_t*>(src + ((size_t)y * W + (x0 - 8)) * 3) + 3 * gq4;
      const uint32_t w0 = __ldg(wp), w1 = __ldg(wp + 1), w2 = __ldg(wp + 2);     // R0G0B0R1 G1B1R2G2 B2R3G3B3
      const uint32_t r = __byte_perm(__byte_perm(w0, w1, 0x0630), w2, 0x5210);
      const uint32_t g = __byte_perm(__byte_perm(w0, w1, 0x0741), w2, 0x6210);
      const uint32_t bl = __byte_perm(__byte_perm(w0, w1, 0x0052), w2, 0x7410);
      reinterpret_cast<uint32_t*>(&plane[0][ty][0])[gq4] = r;
      reinterpret_cast<uint32_t*>(&plane[1][ty][0])[gq4] = g;
      reinterpret_cast<uint32_t*>(&plane[2][ty][0])[gq4] = bl;
    }
  } else {
    for (int t = tid; t < kInH * kInW; t += 256) {
      const int ty = t / kInW, tx = t - ty * kInW;
      const int y = reflect101(y0 + ty - kBR, H), x = reflect101(x0 + tx - kBR, W);
      const uint8_t* px = src + ((size_t)y * W + x) * 3;
      plane[0][ty][tx + 1] = px[0]; plane[1][ty][tx + 1] = px[1]; plane[2][ty][tx + 1] = px[2];
    }
  }
  __syncthreads();

  // ---- row pass: item = (plane, row pair, group of 4 columns)
  for (int it = tid; it < 3 * kPairs * (kTX / 4); it += 256) {
    const int grp = it & (kTX / 4 - 1);
    const int rest = it / (kTX / 4);
    const int rp = rest % kPairs, c = rest / kPairs;
    uint32_t res[2][4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const uint32_t* wrow = reinterpret_cast<const uint32_t*>(&plane[c][2 * rp + h][0]) + grp;
      uint32_t w[5];
#pragma unroll
      for (int i = 0; i < 5; ++i) w[i] = wrow[i];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t s = 0;
#pragma unroll
        for (int i = 0; i < 5; ++i) s = __dp4a(w[i], row_coef(j, i), s);
        res[h][j] = s;
      }
    }
    uint4 o;
    o.x = res[0][0] | (res[1][0] << 16); o.y = res[0][1] | (res[1][1] << 16);
    o.z = res[0][2] | (res[1][2] << 16); o.w = res[0][3] | (res[1][3] << 16);
    *reinterpret_cast<uint4*>(&rsum[c][rp][4 * grp]) = o;
  }
  __syncthreads();

  // ---- column pass: thread = (plane, x)
  if (tid < kCols) {
    const int c = tid / kTX, x = tid - c * kTX;
    uint32_t r[kPairs];
#pragma unroll
    for (int p = 0; p < kPairs; ++p) r[p] = rsum[c][p][x];
#pragma unroll
    for (int oy = 0; oy < kTY; ++oy) {
      const int odd = oy & 1, base = (oy - odd) / 2;       // first pair index
      uint32_t s = 0;
#pragma unroll
      for (int p = 0; p < 8; ++p)
        if (base + p < kPairs) s = __dp2a_lo(r[base + p], col_coef(odd, p), s);
      s = (s + (1u << 15)) >> 16;
      obuf[oy][x * 3 + c] = (uint8_t)min(s, 255u);
    }
  }
  __syncthreads();

  // ---- coalesced store
  const int rows = min(kTY, H - y0), cols = min(kTX, W - x0) * 3;
  uint8_t* dst = out + ((size_t)b * H + y0) * W * 3 + (size_t)x0 * 3;
  if ((W & 3) == 0 && cols == kCols && (reinterpret_cast<uintptr_t>(out) & 3) == 0) {
    for (int t = tid; t < rows * (kCols / 4); t += 256) {
      const int ry = t / (kCols / 4), wd = t - ry * (kCols / 4);
      reinterpret_cast<uint32_t*>(dst + (size_t)ry * W * 3)[wd] = reinterpret_cast<const uint32_t*>(&obuf[ry][0])[wd];
    }
  } else {
    for (int t = tid; t < rows * cols; t += 256) {
      const int ry = t / cols, cb = t - ry * cols;
      dst[(size_t)ry * W * 3 + cb] = obuf[ry][cb];
    }
  }
}

}  // namespace hgl

extern "C" int hgl_gaussian_blur15(const uint8_t* image, uint8_t* out, int B, int H, int W, void* stream) {
  using namespace hgl;
  HGL_REQUIRE(image && out, "hgl_gaussian_blur15: null pointer");
  HGL_REQUIRE(B >= 1 && H >= 8 && W >= 8, "hgl_gaussian_blur15: frame %dx%d too small for a 15-tap reflect-101 border", H, W);
  HGL_REQUIRE(B <= 65535 && ceil_div(H, kTY) <= 65535, "hgl_gaussian_blur15: batch too large for one launch");
  dim3 grid(ceil_div(W, kTX), ceil_div(H, kTY), B);
  blur15_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(image, out, H, W);
  return launch_status("hgl_gaussian_blur15");
}
