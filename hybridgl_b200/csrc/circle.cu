// (a1') the 'circle' visual prompt -- utils.py:322-335:
//     (cy, cx), h, w = mask2chw(mask);  cv2.ellipse(image, (cx, cy), (w // 2, h // 2), 0, 0, 360, (255, 0, 0), 1)
// on the device, pixel for pixel what OpenCV draws.
//
// What cv2.ellipse does at thickness 1 / LINE_8 (OpenCV 4.x imgproc/src/drawing.cpp: ellipse -> EllipseEx -> ellipse2Poly -> PolyLine ->
// ThickLine -> Line -> LineIterator; restated and pinned against cv2 in oracle/hybridgl_oracle.py::ellipse_outline):
//   * a polygon with a vertex every `delta` degrees (delta from the larger half axis: < 3 -> 90, < 10 -> 30, < 15 -> 18, else 5);
//     a vertex is c * 2^16 + (axis * 2^16) * (double)SinTable[deg] in double (product and sum rounded separately), rounded half to
//     even to 16.16 fixed point and then to the nearest pixel, (v + 2^15) >> 16;
//   * every edge is an integer 8-connected Bresenham line drawn LEFT TO RIGHT after cv::clipLine moved end points outside the
//     frame onto its border (double arithmetic, truncated toward zero).
// Consecutive duplicate vertices, which OpenCV drops, only remove zero-length edges: the pixel set is the same without that step.
//
// Two entry points:
//   hgl_ellipse_outline   draws the outline IN PLACE into u8 frames [M,H,W,3] (image m gets the ellipse of chw[m]): the drop-in of
//                         utils.apply_visual_prompts(..., 'circle') for one prompted image per proposal.
//   hgl_prep_circle       the batched form for the prep outputs: after hgl_prep(...) wrote global[n] = Normalize(Resize(composite_n)),
//                         the output pixels whose bilinear taps touch proposal n's outline are re-evaluated exactly (one CTA per
//                         proposal: outline rasterised into a shared-memory bitmap, then a scan of the output pixels under the
//                         ellipse's bounding box).  The prompted frames [M,H,W,3] never exist.
// The order of the reference's if-chain is kept: blur composite -> circle -> black composite (an outline pixel outside the mask
// is zeroed by 'black').
#include "hgl_common.cuh"
#include "prep_math.cuh"

namespace hgl {

__constant__ float c_sin_deg[451] = {
#include "ellipse_sin.inc"
};

constexpr int kEllMaxSeg = 72;        // 360 / 5
constexpr int kEllThreads = 256;

__device__ __forceinline__ int ellipse_delta(int ax, int ay) {
  const int m = max(ax, ay);
  return m < 3 ? 90 : m < 10 ? 30 : m < 15 ? 18 : 5;
}
// vertex k (angle min(k * delta, 360)) in pixels
__device__ __forceinline__ void ellipse_vertex(int cx, int cy, int ax, int ay, int delta, int k, int& px, int& py) {
  const int ang = min(k * delta, 360);
  const double vx = __dadd_rn((double)((long long)cx * 65536), __dmul_rn((double)((long long)ax * 65536), (double)c_sin_deg[450 - ang]));
  const double vy = __dadd_rn((double)((long long)cy * 65536), __dmul_rn((double)((long long)ay * 65536), (double)c_sin_deg[ang]));
  px = (int)((__double2ll_rn(vx) + 32768) >> 16);
  py = (int)((__double2ll_rn(vy) + 32768) >> 16);
}

// cv::clipLine on integer end points (64-bit, like Point2l); false = nothing of the segment is inside the frame
__device__ __forceinline__ bool clip_line(int W, int H, long long& x1, long long& y1, long long& x2, long long& y2) {
  const long long right = W - 1, bottom = H - 1;
  int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
  int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
  if ((c1 & c2) == 0 && (c1 | c2) != 0) {
    long long a;
    if (c1 & 12) {
      a = c1 < 8 ? 0 : bottom;
      x1 += (long long)__ddiv_rn(__dmul_rn((double)(a - y1), (double)(x2 - x1)), (double)(y2 - y1));
      y1 = a;
      c1 = (x1 < 0) + (x1 > right) * 2;
    }
    if (c2 & 12) {
      a = c2 < 8 ? 0 : bottom;
      x2 += (long long)__ddiv_rn(__dmul_rn((double)(a - y2), (double)(x2 - x1)), (double)(y2 - y1));
      y2 = a;
      c2 = (x2 < 0) + (x2 > right) * 2;
    }
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
      if (c1) {
        a = c1 == 1 ? 0 : right;
        y1 += (long long)__ddiv_rn(__dmul_rn((double)(a - x1), (double)(y2 - y1)), (double)(x2 - x1));
        x1 = a;
        c1 = 0;
      }
      if (c2) {
        a = c2 == 1 ? 0 : right;
        y2 += (long long)__ddiv_rn(__dmul_rn((double)(a - x2), (double)(y2 - y1)), (double)(x2 - x1));
        x2 = a;
        c2 = 0;
      }
    }
  }
  return (c1 | c2) == 0;
}

// cv2.line(img, p1, p2, color, 1, LINE_8): clip, then LineIterator(..., connectivity 8, leftToRight = true)
template <typename Put>
__device__ __forceinline__ void line8(int H, int W, int px1, int py1, int px2, int py2, Put put) {
  long long x1 = px1, y1 = py1, x2 = px2, y2 = py2;
  if (!clip_line(W, H, x1, y1, x2, y2)) return;
  int dx = (int)(x2 - x1), dy = (int)(y2 - y1);
  int x = (int)x1, y = (int)y1;
  if (dx < 0) { x = (int)x2; y = (int)y2; dx = -dx; dy = -dy; }
  const int sy = dy >= 0 ? 1 : -1;
  dy = abs(dy);
  if (dy > dx) {                                   // y is the major axis
    int err = dy - 2 * dx;
    for (int k = 0; k <= dy; ++k) {
      put(x, y);
      if (err < 0) { err += 2 * dy - 2 * dx; ++x; } else { err -= 2 * dx; }
      y += sy;
    }
  } else {
    int err = dx - 2 * dy;
    for (int k = 0; k <= dx; ++k) {
      put(x, y);
      if (err < 0) { err += 2 * dx - 2 * dy; y += sy; } else { err -= 2 * dy; }
      ++x;
    }
  }
}

// thread t < edges draws edge t of the ellipse of (cy, cx, height, width) = chw; returns false for an empty proposal
template <typename Put>
__device__ __forceinline__ void ellipse_edges(const int32_t* __restrict__ chw, int H, int W, int t, int nthreads, Put put) {
  const int cy = chw[0], cx = chw[1], ay = chw[2] / 2, ax = chw[3] / 2;
  if (chw[2] <= 0 || chw[3] <= 0) return;          // empty proposal (the reference raises in mask2chw): nothing is drawn
  const int delta = ellipse_delta(ax, ay), edges = 360 / delta;
  for (int e = t; e < edges; e += nthreads) {
    int xa, ya, xb, yb;
    ellipse_vertex(cx, cy, ax, ay, delta, e, xa, ya);
    ellipse_vertex(cx, cy, ax, ay, delta, e + 1, xb, yb);
    line8(H, W, xa, ya, xb, yb, put);
  }
}

__global__ void __launch_bounds__(96) ellipse_draw_kernel(uint8_t* __restrict__ images, const int32_t* __restrict__ chw, int H, int W,
                                                          int r, int g, int b) {
  const int m = blockIdx.x;
  uint8_t* img = images + (size_t)m * H * W * 3;
  ellipse_edges(chw + 4 * m, H, W, threadIdx.x, blockDim.x, [&](int x, int y) {
    uint8_t* px = img + ((size_t)y * W + x) * 3;
    px[0] = (uint8_t)r; px[1] = (uint8_t)g; px[2] = (uint8_t)b;
  });
}

struct CircleParams {
  const uint8_t* image; const uint8_t* blur; const uint32_t* bits; const int32_t* mask_off; const int32_t* chw;
  int B, M, H, W, WW, S, bg_mode;
  int color[3];
  void* global_out;
};

// One CTA per proposal.  Shared memory: the outline bitmap [H][WW] u32.
template <bool kBF16>
__global__ void __launch_bounds__(kEllThreads) prep_circle_kernel(const CircleParams p) {
  extern __shared__ __align__(16) uint32_t obits[];
  const int m = blockIdx.x, tid = threadIdx.x;
  const int H = p.H, W = p.W, WW = p.WW, S = p.S, SS = S * S;
  const int32_t* chw = p.chw + 4 * m;
  if (chw[2] <= 0 || chw[3] <= 0) return;          // uniform for the CTA
  for (int t = tid; t < H * WW; t += kEllThreads) obits[t] = 0u;
  __syncthreads();
  ellipse_edges(chw, H, W, tid, kEllThreads, [&](int x, int y) { atomicOr(obits + y * WW + (x >> 5), 1u << (x & 31)); });
  __syncthreads();
  int b = 0;
  if (p.mask_off) { while (b + 1 < p.B && p.mask_off[b + 1] <= m) ++b; }
  const uint8_t* img = p.image + (size_t)b * H * W * 3;
  const uint8_t* bg = p.bg_mode == HGL_BG_BLUR ? p.blur + (size_t)b * H * W * 3 : nullptr;
  const uint32_t* mb = p.bits + (size_t)m * H * WW;
  // output pixels whose taps can reach the ellipse's bounding box (a tap sits at most one source pixel after its first)
  const int cy = chw[0], cx = chw[1], ay = chw[2] / 2, ax = chw[3] / 2;
  const float sc_y = tap_scale(H, S), sc_x = tap_scale(W, S);
  const int ylo = max(0, cy - ay - 1), yhi = min(H - 1, cy + ay + 1), xlo = max(0, cx - ax - 1), xhi = min(W - 1, cx + ax + 1);
  int i_lo = 0, i_hi = S - 1, j_lo = 0, j_hi = S - 1;                       // conservative bounds by bisection on the monotone tap index
  while (i_lo < S - 1 && make_taps(i_lo, H, S, sc_y).i0 + 1 < ylo) ++i_lo;
  while (i_hi > 0 && make_taps(i_hi, H, S, sc_y).i0 > yhi) --i_hi;
  while (j_lo < S - 1 && make_taps(j_lo, W, S, sc_x).i0 + 1 < xlo) ++j_lo;
  while (j_hi > 0 && make_taps(j_hi, W, S, sc_x).i0 > xhi) --j_hi;
  if (i_hi < i_lo || j_hi < j_lo) return;
  const int nj = j_hi - j_lo + 1, total = (i_hi - i_lo + 1) * nj;
  for (int t = tid; t < total; t += kEllThreads) {
    const int i = i_lo + t / nj, j = j_lo + t % nj;
    const Taps ty = make_taps(i, H, S, sc_y), tx = make_taps(j, W, S, sc_x);
    const int ys[2] = {ty.i0, ty.i0 + ty.d}, xs[2] = {tx.i0, tx.i0 + tx.d};
    uint32_t on = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) on |= ((obits[ys[q >> 1] * WW + (xs[q & 1] >> 5)] >> (xs[q & 1] & 31)) & 1u) << q;
    if (!on) continue;
    float gv[3][4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int yy = ys[q >> 1], xx = xs[q & 1];
      const bool in = (__ldg(mb + (size_t)yy * WW + (xx >> 5)) >> (xx & 31)) & 1u;
      const bool ol = (on >> q) & 1u;
      const size_t o = ((size_t)yy * W + xx) * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        uint32_t v;
        if (in) v = ol ? (uint32_t)p.color[c] : (uint32_t)__ldg(img + o + c);
        else if (p.bg_mode == HGL_BG_BLACK) v = 0u;                          // 'black' runs after 'circle'
        else v = ol ? (uint32_t)p.color[c] : (uint32_t)__ldg((bg ? bg : img) + o + c);
        gv[c][q] = to_unit(v);
      }
    }
    const size_t o = (size_t)m * 3 * SS + (size_t)i * S + j;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float g = __fdiv_rn(__fsub_rn(bilerp(gv[c][0], gv[c][1], gv[c][2], gv[c][3], tx.w0, tx.w1, ty.w0, ty.w1), c_in_mean[c]), c_in_std[c]);
      if (kBF16) reinterpret_cast<__nv_bfloat16*>(p.global_out)[o + (size_t)c * SS] = __float2bfloat16_rn(g);
      else reinterpret_cast<float*>(p.global_out)[o + (size_t)c * SS] = g;
    }
  }
}

}  // namespace hgl

extern "C" int hgl_ellipse_outline(uint8_t* images, const int32_t* chw, int M, int H, int W, int r, int g, int b, void* stream) {
  using namespace hgl;
  if (M == 0) return HGL_OK;
  HGL_REQUIRE(images && chw, "hgl_ellipse_outline: null pointer");
  HGL_REQUIRE(M > 0 && H >= 1 && W >= 1 && H <= 32767 && W <= 32767, "hgl_ellipse_outline: bad shape M=%d H=%d W=%d", M, H, W);
  ellipse_draw_kernel<<<M, 96, 0, (cudaStream_t)stream>>>(images, chw, H, W, r & 255, g & 255, b & 255);
  return launch_status("hgl_ellipse_outline");
}

extern "C" int hgl_prep_circle(const uint8_t* image, const uint8_t* blur, const uint32_t* bits, const int32_t* mask_off, const int32_t* chw,
                               int B, int M, int H, int W, int S, int bg_mode, int out_dtype, int r, int g, int b, void* global_out,
                               void* stream) {
  using namespace hgl;
  if (M == 0) return HGL_OK;
  HGL_REQUIRE(image && bits && chw && global_out, "hgl_prep_circle: null pointer");
  HGL_REQUIRE(B >= 1 && M > 0 && H >= 1 && W >= 1 && S >= 4 && H <= 32767 && W <= 32767, "hgl_prep_circle: bad shape");
  HGL_REQUIRE(mask_off || B == 1, "hgl_prep_circle: mask_off required when B > 1");
  HGL_REQUIRE(bg_mode == HGL_BG_BLUR || bg_mode == HGL_BG_BLACK || bg_mode == HGL_BG_NONE, "hgl_prep_circle: bg_mode %d", bg_mode);
  HGL_REQUIRE(bg_mode != HGL_BG_BLUR || blur, "hgl_prep_circle: blur frame required for HGL_BG_BLUR");
  HGL_REQUIRE(out_dtype == HGL_F32 || out_dtype == HGL_BF16, "hgl_prep_circle: out_dtype %d", out_dtype);
  CircleParams p = {};
  p.image = image; p.blur = blur; p.bits = bits; p.mask_off = mask_off; p.chw = chw;
  p.B = B; p.M = M; p.H = H; p.W = W; p.WW = (W + 31) >> 5; p.S = S; p.bg_mode = bg_mode;
  p.color[0] = r & 255; p.color[1] = g & 255; p.color[2] = b & 255;
  p.global_out = global_out;
  const size_t smem = (size_t)H * p.WW * 4;
  HGL_REQUIRE(smem <= 200 * 1024, "hgl_prep_circle: a %dx%d frame needs %zu B of shared memory for the outline bitmap (limit 200 KB)", H, W, smem);
  cudaStream_t st = (cudaStream_t)stream;
  if (out_dtype == HGL_BF16) {
    const int rc = ensure_dyn_smem(reinterpret_cast<const void*>(prep_circle_kernel<true>), smem, "hgl_prep_circle");
    if (rc != HGL_OK) return rc;
    prep_circle_kernel<true><<<M, kEllThreads, smem, st>>>(p);
  } else {
    const int rc = ensure_dyn_smem(reinterpret_cast<const void*>(prep_circle_kernel<false>), smem, "hgl_prep_circle");
    if (rc != HGL_OK) return rc;
    prep_circle_kernel<false><<<M, kEllThreads, smem, st>>>(p);
  }
  return launch_status("hgl_prep_circle");
}
