// cv2.GaussianBlur(img, (15,15), 0) on uint8 with OpenCV's fixed-point arithmetic  --  Hybridgl_main.py:99.
//
// sigma = 0.3*((15-1)*0.5-1)+0.8 = 2.6; OpenCV quantises the taps to Q8 {1,3,6,12,20,30,36,40,36,...} (sum 256),
// runs the row pass in Q8 (fits u16), the column pass in Q16 (u32) and rounds half-up once: (v + 2^15) >> 16.
// Border = BORDER_REFLECT_101.  Restated (and checked bit-exact against cv2 4.13) in oracle/hybridgl_oracle.py.
// Frames are < 1 MB, so this is launch-latency work: one CTA per 32x32 tile, halo staged in shared memory.
#include "hgl_common.cuh"

namespace hgl {

constexpr int kBT = 32;          // tile side
constexpr int kBR = 7;           // kernel radius
__constant__ uint32_t c_gq[15] = {1, 3, 6, 12, 20, 30, 36, 40, 36, 30, 20, 12, 6, 3, 1};

__device__ __forceinline__ int reflect101(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

__global__ void __launch_bounds__(256) blur15_kernel(const uint8_t* __restrict__ img, uint8_t* __restrict__ out, int H, int W) {
  __shared__ uint8_t tile[kBT + 2 * kBR][(kBT + 2 * kBR) * 3];
  __shared__ uint16_t rowp[kBT + 2 * kBR][kBT * 3];
  const int b = blockIdx.z, y0 = blockIdx.y * kBT, x0 = blockIdx.x * kBT;
  const uint8_t* src = img + (size_t)b * H * W * 3;
  constexpr int TH = kBT + 2 * kBR, TW = (kBT + 2 * kBR) * 3;
  for (int t = threadIdx.x; t < TH * TW; t += blockDim.x) {
    const int ty = t / TW, tc = t - ty * TW;
    const int tx = tc / 3, c = tc - tx * 3;
    const int y = reflect101(y0 + ty - kBR, H), x = reflect101(x0 + tx - kBR, W);
    tile[ty][tc] = src[((size_t)y * W + x) * 3 + c];
  }
  __syncthreads();
  for (int t = threadIdx.x; t < TH * kBT * 3; t += blockDim.x) {
    const int ty = t / (kBT * 3), tc = t - ty * (kBT * 3);
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < 15; ++k) s += c_gq[k] * tile[ty][tc + 3 * k];
    rowp[ty][tc] = (uint16_t)s;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < kBT * kBT * 3; t += blockDim.x) {
    const int ty = t / (kBT * 3), tc = t - ty * (kBT * 3);
    const int y = y0 + ty, x = x0 + tc / 3;
    if (y < H && x < W) {
      uint32_t s = 0;
#pragma unroll
      for (int k = 0; k < 15; ++k) s += c_gq[k] * rowp[ty + k][tc];
      s = (s + (1u << 15)) >> 16;
      out[((size_t)b * H * W + (size_t)y * W) * 3 + (size_t)x0 * 3 + tc] = (uint8_t)min(s, 255u);
    }
  }
}

}  // namespace hgl

extern "C" int hgl_gaussian_blur15(const uint8_t* image, uint8_t* out, int B, int H, int W, void* stream) {
  using namespace hgl;
  HGL_REQUIRE(image && out, "hgl_gaussian_blur15: null pointer");
  HGL_REQUIRE(B >= 1 && H >= 8 && W >= 8, "hgl_gaussian_blur15: frame %dx%d too small for a 15-tap reflect-101 border", H, W);
  dim3 grid(ceil_div(W, kBT), ceil_div(H, kBT), B);
  blur15_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(image, out, H, W);
  return launch_status("hgl_gaussian_blur15");
}
