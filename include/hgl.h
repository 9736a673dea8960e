/*
 * libhgl -- C ABI of the B200-native HybridGL mask-proposal scoring path.
 *
 * The reference (fhgyuanshen/HybridGL) has no FFI / plugin interface: the path is reached through plain
 * Python calls.  Each entry point below therefore cites the reference *code block*
 * it replaces (file:line relative to the reference checkout); hybridgl_b200/ops.py binds them with ctypes
 * and hybridgl_b200/{backbone,utils,pipeline}.py re-expose the reference's own call surface on top.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host; the caller owns and sizes all buffers
 *   - `stream` is a cudaStream_t passed as void*; kernels are enqueued asynchronously on it
 *   - no allocation, no synchronisation, no host<->device copy inside the library (re-entrant per stream)
 *   - return value: 0 on success, a negative HGL_E* code otherwise; hgl_last_error() gives the text
 *     (thread-local).  Nothing throws across the boundary.
 *   - dtype codes: HGL_F32 / HGL_BF16 for floating tensors; masks are 1 byte per pixel (torch.bool)
 *   - ragged batches: masks of image b are rows mask_off[b] .. mask_off[b+1]-1 of the [M,...] tensors,
 *     expressions of image b are rows expr_off[b] .. expr_off[b+1]-1 of the [E,...] tensors
 */
#ifndef HGL_H_
#define HGL_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define HGL_API __attribute__((visibility("default")))
#else
#define HGL_API
#endif

#define HGL_OK 0
#define HGL_EINVAL (-1)   /* bad argument (shape, dtype, null pointer, unsupported size) */
#define HGL_ECUDA (-2)    /* CUDA runtime error at launch */
#define HGL_EARCH (-3)    /* device is not sm_100 */

#define HGL_F32 0
#define HGL_BF16 1

/* token-stream layouts: the reference keeps streams as [L+1, M, D] (LND, model/backbone.py:139); the B200 forward
 * keeps them batch-first [M, L+1, D] (NLD) so that attention runs as one batched SDPA call */
#define HGL_LND 0
#define HGL_NLD 1

/* background of the global view outside the mask (utils.py:292-345 apply_visual_prompts option set) */
#define HGL_BG_BLUR 0     /* Gaussian-blurred frame  (Hybridgl_main.py:99-113; utils.py:306-320) */
#define HGL_BG_BLACK 1    /* zeros                   (utils.py:336-341) */
#define HGL_BG_NONE 2     /* the frame itself: nothing is replaced outside the mask (the 'circle' prompt on its own) */

/* relation words of utils.py:240-268 (extract_rela_word, utils.py:207-237) */
enum { HGL_REL_NONE = 0, HGL_REL_LEFT, HGL_REL_RIGHT, HGL_REL_UP, HGL_REL_DOWN, HGL_REL_BIG, HGL_REL_SMALL, HGL_REL_WITHIN };
/* direction words of utils.py:102-161 (extract_dir_phrase / gen_dir_mask) */
enum { HGL_DIR_NONE = 0, HGL_DIR_LEFT, HGL_DIR_RIGHT, HGL_DIR_MIDDLE, HGL_DIR_UP, HGL_DIR_DOWN };

HGL_API const char* hgl_last_error(void);
HGL_API int hgl_version(void);
/* 0 if the current device is sm_100 and the kernels are loadable, HGL_EARCH otherwise */
HGL_API int hgl_check_device(void);

/* ---- packed mask format ------------------------------------------------------------------------------
 * bits u32 [M, H, WW], WW = ceil(W/32); bit i of word w of a row is pixel x = 32*w + i (non-zero byte -> 1).
 * The byte masks of SAM (torch.bool [M,H,W], Hybridgl_main.py:86-87) are read exactly once, here; prep, the mask
 * grid and the heat-map pooling all consume the 8x smaller packed tensor. */
HGL_API int hgl_pack_masks(const uint8_t* masks, int M, int H, int W, uint32_t* bits, void* stream);
/* The same for masks whose bytes are known to be 0 or 1 -- the storage of torch.bool, i.e. exactly what Hybridgl_main.py:86-87
 * builds (torch.from_numpy of SAM's bool 'segmentation' arrays): a cheaper byte -> bit squeeze (40 % fewer instructions; the pass
 * shares the SMs with the frame-only kernels of the path).  Any other byte value gives undefined bits: use hgl_pack_masks for uint8. */
HGL_API int hgl_pack_masks_bool(const uint8_t* masks, int M, int H, int W, uint32_t* bits, void* stream);

/* ---- per-mask geometry from the packed masks ------------------------------------------------------------
 * boxes_xywh int64 [M,4]: SAM's proposal boxes -- batched_mask_to_box + box_xyxy_to_xywh
 *   (third_party/segment-anything/segment_anything/utils/amg.py:303-346, :91-95): x0, y0, x1 - x0, y1 - y0 of the inclusive
 *   edges, zeros for an empty mask; the boxes Hybridgl_main.py:89-90 hands to relation_boxes.  May be NULL.
 * chw int32 [M,4]: mask2chw utils.py:280-289: (center_y, center_x, height, width) = (int(mean(rows)), int(mean(cols)),
 *   rows.max()-rows.min()+1, cols.max()-cols.min()+1); (-1,-1,0,0) for an empty mask (the reference raises).  May be NULL. */
HGL_API int hgl_mask_geometry(const uint32_t* bits, int M, int H, int W, int64_t* boxes_xywh, int32_t* chw, void* stream);

/* ---- (a1) per-mask visual-prompt preprocessing ------------------------------------------------------
 * Replaces the Python loop Hybridgl_main.py:92-125 (dups demo.py:79-112) and utils.py:292-345:
 *   global[n] = Normalize_IN(bilinear_S( where(mask_n, image, background) / 255 ))
 *   local[n]  = bilinear_S( where(mask_n, Normalize_IN(image/255), clip_pixel_mean) )
 * image/blur u8 [B,H,W,3]; bits = packed masks [M,H,WW]; mask_off int32 [B+1] (NULL => B==1, all M masks belong
 * to image 0); max_n >= masks of any one image; local_out/global_out [M,3,S,S] of out_dtype (16-byte aligned).
 * blur may be NULL for HGL_BG_BLACK.  workspace: hgl_prep_workspace_bytes() bytes, 256-byte aligned. */
HGL_API int64_t hgl_prep_workspace_bytes(int B, int S, int out_dtype);
HGL_API int hgl_prep(const uint8_t* image, const uint8_t* blur, const uint32_t* bits, const int32_t* mask_off,
             int B, int M, int max_n, int H, int W, int S, int bg_mode, int out_dtype,
             void* local_out, void* global_out, void* workspace, void* stream);

/* hgl_prep with a per-proposal crop (the north star's "bounding-box crop"; SURVEY 8(b) `crop_xywh`): proposal n is resampled from
 * its own box crop_xywh[n] = (x, y, w, h) (int32 [M,4], inside the frame, w, h >= 1) instead of from the full frame:
 *   global[n] = Normalize_IN(bilinear_S(where(mask_n, image, background)[y:y+h, x:x+w] / 255)), local[n] likewise on the mean-filled
 * view.  The reference itself never crops (it casts pred_box at Hybridgl_main.py:101 and does not use it).  No workspace. */
HGL_API int hgl_prep_crop(const uint8_t* image, const uint8_t* blur, const uint32_t* bits, const int32_t* mask_off, const int32_t* crop_xywh,
                  int B, int M, int H, int W, int S, int bg_mode, int out_dtype, void* local_out, void* global_out, void* stream);

/* ---- the 'circle' visual prompt (utils.py:322-335) -------------------------------------------------------
 *   (cy, cx), h, w = mask2chw(mask);  cv2.ellipse(image, (cx, cy), (w // 2, h // 2), 0, 0, 360, color = (255, 0, 0), thickness = 1)
 * chw int32 [M,4] = (center_y, center_x, height, width) per proposal, as hgl_mask_geometry writes it.  The outline is OpenCV's,
 * pixel for pixel (polygon vertices in 16.16 fixed point, cv::clipLine, 8-connected left-to-right Bresenham edges; pinned against
 * cv2 4.13 by the oracle).  An empty proposal (height or width 0; the reference raises) draws nothing.
 *
 * hgl_ellipse_outline: draws IN PLACE into images u8 [M,H,W,3]; image m gets the ellipse of chw[m] in colour (r, g, b).
 * hgl_prep_circle: the batched form on the prep outputs.  Call after hgl_prep / hgl_prep_main with the same image, blur, bits,
 *   mask_off, S, bg_mode and out_dtype: the pixels of global_out[n] whose bilinear taps touch proposal n's outline are re-evaluated
 *   with the prompt applied in the reference's order (blur composite -> circle -> black composite), so that
 *   global_out[n] == Normalize(Resize(ToTensor(apply_visual_prompts(frame, mask_n, types)))) (Hybridgl_main.py:103-118 with
 *   utils.py:292-345 in place of the inline composite).  local_out does not see the prompt.  H * ceil(W/32) * 4 <= 200 KB. */
HGL_API int hgl_ellipse_outline(uint8_t* images, const int32_t* chw, int M, int H, int W, int r, int g, int b, void* stream);
HGL_API int hgl_prep_circle(const uint8_t* image, const uint8_t* blur, const uint32_t* bits, const int32_t* mask_off, const int32_t* chw,
                            int B, int M, int H, int W, int S, int bg_mode, int out_dtype, int r, int g, int b, void* global_out,
                            void* stream);

/* hgl_prep in its two halves, for callers that overlap stages: hgl_prep_setup needs only the frames (per image: the
 * "all taps inside" / "all taps outside" answer planes and the tap bytes, left in `workspace`) and may be enqueued while
 * the masks are still being packed; hgl_prep_main streams the packed masks over them.  hgl_prep == setup then main on
 * one stream; same arguments, same workspace. */
HGL_API int hgl_prep_setup(const uint8_t* image, const uint8_t* blur, int B, int H, int W, int S, int bg_mode, int out_dtype,
                   void* workspace, void* stream);
HGL_API int hgl_prep_main(const uint32_t* bits, const int32_t* mask_off, int B, int M, int max_n, int H, int W, int S, int out_dtype,
                  void* local_out, void* global_out, void* workspace, void* stream);

/* cv2.GaussianBlur(img,(15,15),0) on uint8, BORDER_REFLECT_101, OpenCV's Q8 fixed-point taps
 * (Hybridgl_main.py:99).  image/out u8 [B,H,W,3]. */
HGL_API int hgl_gaussian_blur15(const uint8_t* image, uint8_t* out, int B, int H, int W, void* stream);

/* ---- (a2) mask -> patch grid ------------------------------------------------------------------------
 * Replaces TF.resize(pred_masks.float(), (g,g)) model/backbone.py:160.  antialias=1 is torchvision>=0.17
 * behaviour (this container), antialias=0 the reference's pinned torchvision 0.15.2.
 * bits = packed masks [M,H,WW] -> grid f32 [M,g,g]; area int32 [M] (pixel count of every mask; may be NULL).
 * workspace: hgl_mask_grid_workspace_bytes(M, g) bytes, 16-byte aligned (only read when antialias != 0; may be NULL otherwise). */
HGL_API int64_t hgl_mask_grid_workspace_bytes(int M, int g);
HGL_API int hgl_mask_grid(const uint32_t* bits, int M, int H, int W, int g, int antialias,
                  float* grid, int32_t* area, void* workspace, void* stream);

/* ---- (a3) CLS-row attention mask --------------------------------------------------------------------
 * Replaces CLIPViTFM.make_attn_mask model/backbone.py:108-115.
 * grid f32 [M,L] -> out u8 [M*heads, L+1, L+1], 1 = blocked: only (q=0, k>=1, grid[m,k-1]==0). */
HGL_API int hgl_attn_mask(const float* grid, int M, int L, int heads, uint8_t* out, void* stream);
/* Compact equivalent used by the B200 forward: additive key bias for the CLS query only,
 * bias f32 [M, L+1] = 0 or -inf (col 0 always 0). */
HGL_API int hgl_attn_bias(const float* grid, int M, int L, float* bias, void* stream);

/* ---- (f1) CLS-row attention under the key bitmap ------------------------------------------------------
 * The only masked row of the reference's attention mask (model/backbone.py:108-115; third_party/modified_CLIP/clip/model.py:220-257)
 * is the CLS query: out[m,h,:] = softmax_j(q[m,0,h].k[m,j,h] / sqrt(hd) + bias[m,j]) @ v[m,:,h] from the packed projection
 * qkv [M, L1, 3, heads, hd] of `dtype`; bias f32 [M, L1] (hgl_attn_bias; NULL = unmasked); out [M, heads, hd] of `dtype`.
 * The other rows come from one unmasked SDPA call; this replaces the N*heads*(L+1)^2 mask tensor. */
HGL_API int hgl_cls_attention(const void* qkv, const float* bias, int M, int L1, int heads, int hd, int dtype, void* out, void* stream);

/* ---- (a5) CLS head ------------------------------------------------------------------------------------
 * Replaces ln_post(x[:,0,:]) @ visual.proj (model/backbone.py:254-260, 220-225, 296-306): LayerNorm in f32 (eps, biased variance)
 * and the projection in one launch.  x: row m starts at x + m * row_stride elements (pass the [M, L+1, Dv] stream and
 * row_stride = (L+1)*Dv to read the CLS tokens in place); gamma / beta [Dv], proj [Dv, De] of w_dtype; out f32 [M, De],
 * overwritten or (accumulate != 0) added to -- G2L&L2G sums the heads of its two hybrid streams. */
HGL_API int hgl_cls_head(const void* x, int64_t row_stride, const void* gamma, const void* beta, const void* proj, int M, int Dv, int De,
                 double eps, int x_dtype, int w_dtype, int accumulate, float* out, void* stream);

/* ---- (a4) token masking + stream mix ----------------------------------------------------------------
 * Replaces the permute/view/mul/cat chains model/backbone.py:235-249, 214-216, 275-291:
 *   out[l,m,:] = a * w(l,m) * src[l,m,:] + b * add[l,m,:],  w = 1 for l==0 (CLS) else grid[m,l-1]
 * grid NULL => w == 1 (plain a*src + b*add);  add NULL => b term dropped.
 * src/add/out [L+1, M, D] (layout HGL_LND) or [M, L+1, D] (HGL_NLD) of dtype; grid f32 [M,L]. out may alias src or add. */
HGL_API int hgl_token_mask_fuse(const void* src, const void* add, const float* grid, float a, float b,
                        int L1, int M, int D, int dtype, int layout, void* out, void* stream);
/* (f1) The same fuse and the LayerNorm that follows it in the masked block (ln_1, third_party/modified_CLIP/clip/model.py:244-257,
 * computed in fp32 like CLIP's LayerNorm, :188-195) in one pass over the tokens, NLD layout [M, L+1, D]:
 *   out_x  = a * tokenmask(src, grid) + b * add      rounded to `dtype` (may be NULL when the stream is not needed again)
 *   out_ln = LayerNorm(out_x) * gamma + beta          f32 arithmetic on the stored out_x; gamma, beta f32 [D]
 * add / grid may be NULL (a = 1: plain copy + LayerNorm).  D % 8 == 0, D <= 2048.  out_x / out_ln may be slices of larger tensors
 * (contiguous [M, L+1, D] blocks): the hybrid forward writes its streams straight into the block's concatenated batch. */
HGL_API int hgl_token_mask_fuse_ln(const void* src, const void* add, const float* grid, float a, float b, const float* gamma,
                                   const float* beta, float eps, int L1, int M, int D, int dtype, void* out_x, void* out_ln, void* stream);

/* ---- (a10)+(a11) heat-map conditioning and mask pooling ---------------------------------------------
 * Replaces Hybridgl_main.py:204-223 and gen_dir_mask utils.py:135-161:
 *   A' = minmax(A) * ramp(dirflag);  A'' = A' / mean(A');
 *   score_gem[e,n] = (2-black_e) * sum(A''*m_n)/area(m_n) - black_e * sum(A''*(1-m_n))/area(1-m_n)
 * heat f32 [E,H,W] (GEM map after T.Resize, Hybridgl_main.py:201); expr_off int32 [B+1] (NULL => B==1);
 * dirflag int32 [E]; black f32 [E]; bits = packed masks [M,H,WW]; mask_off int32 [B+1] (NULL => B==1).
 * score_gem f32 [E, max_n] row e holds the n masks of its image (max_n = row stride; the tail is zero-filled).
 * workspace: hgl_heat_pool_workspace_bytes(...) bytes (row prefix tables of the heat-maps), zeroing not required. */
/* gen_dir_mask utils.py:135-161 as a tensor: out f32 [H,W] (left: 1->0, right: 0->1, middle: 0->1->0, others: ones) */
HGL_API int hgl_dir_mask(int dirflag, int H, int W, float* out, void* stream);
HGL_API int64_t hgl_heat_pool_workspace_bytes(int B, int M, int E, int H, int W, int max_n);
HGL_API int hgl_heat_pool(const float* heat, const int32_t* expr_off, const int32_t* dirflag, const float* black,
                  const uint32_t* bits, const int32_t* mask_off, int B, int M, int E, int H, int W,
                  int max_n, float* score_gem, void* workspace, void* stream);
/* (a2) + (a10) + (a11) in ONE pass over the packed masks: hgl_mask_grid(antialias=1) and hgl_heat_pool fused (same
 * arguments, same results); this is what the batched pipeline launches.  E == 0 degenerates to hgl_mask_grid.
 * workspace: hgl_grid_heat_pool_workspace_bytes(...) bytes. */
HGL_API int64_t hgl_grid_heat_pool_workspace_bytes(int B, int M, int E, int H, int W, int g, int max_n);
HGL_API int hgl_grid_heat_pool(const uint32_t* bits, const int32_t* mask_off, int B, int M, int H, int W, int g,
                       float* grid, int32_t* area,
                       const float* heat, const int32_t* expr_off, const int32_t* dirflag, const float* black, int E,
                       int max_n, float* score_gem, void* workspace, void* stream);

/* (a10), first step: imgattn = T.Resize((H,W), antialias=True)(gem(...)[0]) (Hybridgl_main.py:201) -- ATen
 * _upsample_bilinear2d_aa (separable triangle filter, horizontal pass then vertical pass), any scale.
 * heat_raw f32 [E,hh,hw] (the GEM map as the model returns it) -> out f32 [E,H,W]. */
HGL_API int hgl_heat_resize_aa(const float* heat_raw, int E, int hh, int hw, int H, int W, float* out, void* stream);
/* hgl_grid_heat_pool on the RAW GEM maps: heat_raw f32 [E,hh,hw]; the resize of Hybridgl_main.py:201 is evaluated inside the
 * prefix pass (the frame-sized heat-map never exists in HBM) when both axes are up-sampled, else materialised in the
 * workspace first.  Same results as hgl_heat_resize_aa followed by hgl_grid_heat_pool.
 * workspace: hgl_grid_heat_pool_raw_workspace_bytes(...) bytes. */
HGL_API int64_t hgl_grid_heat_pool_raw_workspace_bytes(int B, int M, int E, int H, int W, int g, int max_n, int hh, int hw);
HGL_API int hgl_grid_heat_pool_raw(const uint32_t* bits, const int32_t* mask_off, int B, int M, int H, int W, int g,
                           float* grid, int32_t* area,
                           const float* heat_raw, int hh, int hw, const int32_t* expr_off, const int32_t* dirflag,
                           const float* black, int E, int max_n, float* score_gem, void* workspace, void* stream);

/* hgl_grid_heat_pool{,_raw} in two halves, so that a caller can build the heat-map tables (needs the heat-maps only;
 * Hybridgl_main.py:201-209) while the masks are still being packed, and run the mask pass (Hybridgl_main.py:211-223 +
 * model/backbone.py:160) afterwards:
 *   hgl_heat_tables(heat, hh, hw, ...)       heat f32 [E,H,W] with hh = hw = 0, or the raw map [E,hh,hw]
 *   hgl_grid_heat_pool_rows(..., hh, hw, ...) same workspace, same hh / hw; results identical to the one-call forms.
 * workspace: hgl_grid_heat_pool_raw_workspace_bytes(...) (raw maps) or hgl_grid_heat_pool_workspace_bytes(...) bytes. */
HGL_API int hgl_heat_tables(const float* heat, int hh, int hw, const int32_t* dirflag, int E, int H, int W, void* workspace, void* stream);
HGL_API int hgl_grid_heat_pool_rows(const uint32_t* bits, const int32_t* mask_off, int B, int M, int H, int W, int g,
                            float* grid, int32_t* area, int hh, int hw, const int32_t* expr_off, const float* black, int E,
                            int max_n, float* score_gem, void* workspace, void* stream);

/* ---- (f3) GEM heat-map pooling in token space ---------------------------------------------------------
 * The same score_gem as hgl_grid_heat_pool_raw (Hybridgl_main.py:200-223) WITHOUT the frame-sized heat-map: the resize of
 * Hybridgl_main.py:201 is linear, so sum_p m_n(p) A''(p) = kk * (G_n . h - mn * sum(G_n)) with G_n = the mask resampled onto the
 * raw map's token grid by the adjoint of the up-sampler (direction ramp folded in) and mn / kk per-expression scalars
 * (SURVEY.md Appendix A-2).  No [E,H,W] tables are written or gathered from.  Results agree with the pixel-space entry points
 * to float rounding (tests: 1e-3 relative).  heat_raw f32 [E,hh,hw] with hh <= H, hw <= W (an up-sampled map), hh, hw <= 128.
 * workspace: hgl_gem_token_workspace_bytes(...) bytes, 16-byte aligned, zeroing not required. */
HGL_API int64_t hgl_gem_token_workspace_bytes(int B, int M, int E, int H, int W, int hh, int hw, int max_n);
HGL_API int hgl_gem_token_pool(const uint32_t* bits, const int32_t* mask_off, int B, int M, int H, int W,
                       const float* heat_raw, int hh, int hw, const int32_t* expr_off, const int32_t* dirflag, const float* black, int E,
                       int max_n, float* score_gem, void* workspace, void* stream);

/* ---- (b3') token-space mask pooling + L2 normalisation (tensor cores) ---------------------------------
 * The masks x tokens x D contraction of the north star; token-space form of the pooling loop Hybridgl_main.py:218-223
 * (S_in = (M~ . F^) . t) with the normalisation of model/backbone.py:79 fused:
 *   pooled[n,:] = sum_l weights[n,l] * tokens[b(n)][l,:];   out[n,:] = normalize ? pooled / ||pooled||_2 : pooled
 * weights f32 [M,L] (e.g. the soft grid masks of hgl_mask_grid, L = g*g); tokens bf16 [B,L,D] (D % 8 == 0, 16-byte aligned;
 * staged by TMA); mask_off int32 [B+1] (NULL => B==1); out [M,D] of out_dtype.  bf16 x bf16 -> f32 on tcgen05; the f32
 * weights enter as a hi + lo pair of bf16 operands (two MMAs), i.e. with 16 mantissa bits, not 8.
 * At most min(512, 128 * (512 / Nw)) proposals per image (Nw = 64 for D <= 512, 128 for D <= 1024, else 256).
 * workspace: unused (hgl_mask_pool_workspace_bytes() == 0; may be NULL). */
HGL_API int64_t hgl_mask_pool_workspace_bytes(int M, int D, int out_dtype);
HGL_API int hgl_mask_pool(const float* weights, const void* tokens, const int32_t* mask_off, int B, int M, int max_n, int L, int D,
                  int normalize, int out_dtype, void* out, void* workspace, void* stream);

/* ---- (b3') + (a6)-(a9),(a12) in ONE kernel ------------------------------------------------------------
 * hgl_mask_pool followed by hgl_score_select without the pooled features ever leaving the SM: the cosine scores
 * (model/backbone.py:74-87 on the text ensemble / negatives of Hybridgl_main.py:153-166) are taken from the f32 TMEM
 * accumulators, the selection tail (Hybridgl_main.py:168-196, 225-227; utils.py:240-268) runs in the same launch.
 * Arguments as in hgl_mask_pool (weights, tokens, L, D) and hgl_score_select (everything else).
 * features_out: [M,D] of out_dtype, the normalised pooled rows, or NULL (nothing but scores and picks is written). */
HGL_API int hgl_pool_score_select(const float* weights, const void* tokens, const int32_t* mask_off, const int32_t* expr_off,
                          int B, int M, int E, int max_n, int L, int D,
                          const float* sent, const float* noun, const float* others, const int32_t* other_off,
                          const int64_t* boxes, const int32_t* relaflag, const float* score_gem,
                          double logit_scale_exp, double r, double alpha, void* features_out, int out_dtype,
                          float* score_clip, int64_t* idx_hybrid, int64_t* idx_final, int32_t* top_idx, float* blended,
                          void* stream);

/* ---- (a6)-(a9),(a12) scoring, spatial-relationship re-ranking, per-expression argmax ----------------
 * Replaces Hybridgl_main.py:153-196 and :225-227 plus CLIPViTFM.calculate_score model/backbone.py:74-87 and
 * relation_boxes utils.py:240-268, one launch for a whole batch of images:
 *   text = r*sent + (1-r)*noun;  s = scale * cos(feat, text);  sneg = scale * cos(feat, mean(others))
 *   idx_hybrid = argmax s;  p = softmax(s), q = softmax(sneg);  top = topk(p, min(3,n)), topneg = topk(q, min(6,n))
 *   T_i = sum_j rel(box[top_i], box[J_j], p[top_i], Q[J_j])   (J,Q) = (top,p) if no other nouns else (topneg,q)
 *   T = softmax(T);  T_i = (1-alpha)*T_i + alpha*score_gem[top_i];  idx_final = top[argmax T]
 * feat [M,De] feat_dtype; sent/noun f32 [E,De]; others f32 [K,De] with other_off int32 [E+1];
 * boxes int64 [M,4] XYWH; relaflag int32 [E]; score_gem f32 [E,max_n] or NULL (=> alpha term skipped);
 * outputs: score_clip f32 [E,max_n] (pre-softmax), idx_hybrid/idx_final int64 [E] (index local to the image),
 * top_idx int32 [E,3] (-1 padded), blended f32 [E,3].  Masks of an image beyond max_n are ignored.
 * workspace: unused (hgl_score_select_workspace_bytes() == 0; may be NULL): one launch, scores meet in distributed shared memory. */
HGL_API int64_t hgl_score_select_workspace_bytes(int B, int E, int max_n);
HGL_API int hgl_score_select(const void* feat, int feat_dtype, const float* sent, const float* noun, const float* others,
                     const int32_t* other_off, const int64_t* boxes, const int32_t* relaflag, const float* score_gem,
                     const int32_t* mask_off, const int32_t* expr_off, int B, int M, int E, int De, int max_n,
                     double logit_scale_exp, double r, double alpha,
                     float* score_clip, int64_t* idx_hybrid, int64_t* idx_final, int32_t* top_idx, float* blended,
                     void* workspace, void* stream);

/* ---- (a13) IoU accounting ---------------------------------------------------------------------------
 * Replaces Compute_IoU utils.py:365-384 (called at Hybridgl_main.py:171,230):
 *   for every expression e and each of the two picks: I = |pred & gt|, U = |pred | gt|
 * masks u8 [M,H,W]; target u8 [B,H,W]; idx_hybrid/idx_final int64 [E] (image-local);
 * iu int64 [E,4] = {I_hybrid,U_hybrid,I_final,U_final} (overwritten);
 * cum int64 [4] += sums (caller zeroes at sweep start; these are the accumulators NCCL all-reduces). */
HGL_API int hgl_iou(const uint8_t* masks, const uint8_t* target, const int64_t* idx_hybrid, const int64_t* idx_final,
            const int32_t* mask_off, const int32_t* expr_off, int B, int M, int E, int H, int W,
            int64_t* iu, int64_t* cum, void* stream);

/* The same with the prediction read from the packed masks (bits u32 [M,H,ceil(W/32)], hgl_pack_masks / hgl_rle_to_bits)
 * instead of byte masks; identical integers. */
HGL_API int hgl_iou_bits(const uint32_t* bits, const uint8_t* target, const int64_t* idx_hybrid, const int64_t* idx_final,
            const int32_t* mask_off, const int32_t* expr_off, int B, int M, int E, int H, int W,
            int64_t* iu, int64_t* cum, void* stream);

/* ---- SAM run-length proposals -> packed masks --------------------------------------------------------
 * Replaces rle_to_mask third_party/segment-anything/segment_anything/utils/amg.py:138-149 (+ the torch.stack / .to(device)
 * of Hybridgl_main.py:86-87) for proposals kept in SAM's uncompressed RLE form (amg.py:107-135 mask_to_rle_pytorch;
 * SamAutomaticMaskGenerator(output_mode="uncompressed_rle")): column-major runs, the first run counts zeros.
 * counts int32 [R] = the `counts` lists of all M proposals back to back; rle_off int32 [M+1] = their boundaries in `counts`
 * (every proposal's counts sum to H*W; positions past H*W are ignored); bits u32 [M,H,ceil(W/32)] (overwritten) = exactly
 * what hgl_pack_masks writes for the expanded masks. */
HGL_API int hgl_rle_to_bits(const int32_t* counts, const int32_t* rle_off, int M, int H, int W, uint32_t* bits, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HGL_H_ */
