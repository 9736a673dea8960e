// Per-mask geometry straight from the packed masks: SAM's proposal boxes and the centroid / extent of mask2chw.
//
//   boxes  third_party/segment-anything/segment_anything/utils/amg.py:303-346 batched_mask_to_box (inclusive XYXY edges, [0,0,0,0]
//          for an empty mask) followed by box_xyxy_to_xywh (amg.py:91-95): w = x1 - x0, h = y1 - y0 -- the XYWH boxes the scoring
//          tail consumes (Hybridgl_main.py:89-90).  With hgl_rle_to_bits this keeps SAM's post-processing on the device.
//   chw    utils.py:280-289 mask2chw: centre = (int(mean(rows)), int(mean(cols))), height = rows.max() - rows.min() + 1, width likewise
//          (the 'circle' visual prompt, utils.py:322-335).
// One CTA per mask, a thread per bit row: popcount, sum of the set bits' columns (five masked popcounts per word), first / last
// set column (ffs / clz); 64-bit integer sums, so the means are exact: floor(sum / count) == int(np.mean(...)) for every frame size.
#include "hgl_common.cuh"

namespace hgl {

__device__ __forceinline__ int bitpos_sum(uint32_t v) {           // sum of the positions of the set bits of v
  return __popc(v & 0xaaaaaaaau) + 2 * __popc(v & 0xccccccccu) + 4 * __popc(v & 0xf0f0f0f0u) + 8 * __popc(v & 0xff00ff00u) +
         16 * __popc(v & 0xffff0000u);
}

__global__ void __launch_bounds__(256) mask_geometry_kernel(const uint32_t* __restrict__ bits, int H, int W, int WW,
                                                            int64_t* __restrict__ boxes, int32_t* __restrict__ chw) {
  const int m = blockIdx.x, tid = threadIdx.x;
  const uint32_t* mb = bits + (size_t)m * H * WW;
  long long cnt = 0, sy = 0, sx = 0;
  int y0 = H, y1 = -1, x0 = W, x1 = -1;
  for (int y = tid; y < H; y += blockDim.x) {
    const uint32_t* row = mb + (size_t)y * WW;
    int c = 0;
    long long rx = 0;
    for (int w = 0; w < WW; ++w) {
      const uint32_t v = __ldg(row + w);
      if (!v) continue;
      const int pc = __popc(v);
      c += pc;
      rx += (long long)32 * w * pc + bitpos_sum(v);
      x0 = min(x0, 32 * w + __ffs(v) - 1);
      x1 = max(x1, 32 * w + 31 - __clz(v));
    }
    if (c) { cnt += c; sy += (long long)y * c; sx += rx; y0 = min(y0, y); y1 = max(y1, y); }
  }
  __shared__ long long s_sum[3][8];
  __shared__ int s_ext[4][8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o); sy += __shfl_xor_sync(0xffffffffu, sy, o); sx += __shfl_xor_sync(0xffffffffu, sx, o);
    y0 = min(y0, __shfl_xor_sync(0xffffffffu, y0, o)); y1 = max(y1, __shfl_xor_sync(0xffffffffu, y1, o));
    x0 = min(x0, __shfl_xor_sync(0xffffffffu, x0, o)); x1 = max(x1, __shfl_xor_sync(0xffffffffu, x1, o));
  }
  if ((tid & 31) == 0) {
    const int w = tid >> 5;
    s_sum[0][w] = cnt; s_sum[1][w] = sy; s_sum[2][w] = sx;
    s_ext[0][w] = y0; s_ext[1][w] = y1; s_ext[2][w] = x0; s_ext[3][w] = x1;
  }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < 8; ++w) {
      cnt += s_sum[0][w]; sy += s_sum[1][w]; sx += s_sum[2][w];
      y0 = min(y0, s_ext[0][w]); y1 = max(y1, s_ext[1][w]); x0 = min(x0, s_ext[2][w]); x1 = max(x1, s_ext[3][w]);
    }
    if (boxes) {
      int64_t* b = boxes + (size_t)m * 4;
      if (cnt == 0) { b[0] = b[1] = b[2] = b[3] = 0; }
      else { b[0] = x0; b[1] = y0; b[2] = x1 - x0; b[3] = y1 - y0; }
    }
    if (chw) {
      int32_t* c = chw + (size_t)m * 4;
      if (cnt == 0) { c[0] = c[1] = -1; c[2] = c[3] = 0; }     // the reference raises on an empty mask (mean of nothing)
      else { c[0] = (int32_t)(sy / cnt); c[1] = (int32_t)(sx / cnt); c[2] = y1 - y0 + 1; c[3] = x1 - x0 + 1; }
    }
  }
}

}  // namespace hgl

extern "C" int hgl_mask_geometry(const uint32_t* bits, int M, int H, int W, int64_t* boxes_xywh, int32_t* chw, void* stream) {
  using namespace hgl;
  if (M == 0) return HGL_OK;
  HGL_REQUIRE(bits && (boxes_xywh || chw), "hgl_mask_geometry: null pointer");
  HGL_REQUIRE(M > 0 && H >= 1 && W >= 1, "hgl_mask_geometry: bad shape");
  mask_geometry_kernel<<<M, 256, 0, (cudaStream_t)stream>>>(bits, H, W, (W + 31) >> 5, boxes_xywh, chw);
  return launch_status("hgl_mask_geometry");
}
