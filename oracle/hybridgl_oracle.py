"""CPU oracle for the HybridGL mask-proposal scoring path.  TEST INFRASTRUCTURE ONLY.

This file restates, in plain numpy (float32 arithmetic, explicit loop order), what the reference
computes on the hot path.  It exists to CHECK the CUDA kernels; nothing in ``hybridgl_b200/`` imports
it, and only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may.  Each function cites the reference lines it follows (paths relative to
the reference checkout, fhgyuanshen/HybridGL @ f7eb19b).

Parity pin: the reference has no tests or golden vectors for this path, so the
oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, executed in the build container on seeded
synthetic inputs by ``tests/golden/gen_golden.py`` (which exec's the reference's own source lines in
place from /root/reference and stores inputs + outputs under ``tests/golden/*.npz``);
``tests/test_oracle_golden.py`` replays those fixtures through this file.  Two boundaries stay
UNPINNED because their producers are third-party code that is not in the reference tree: the GEM
heat-map (gem-torch 1.0.1) and the spaCy parse (en_core_web_lg 3.7.1) -- both are plain inputs here.

Arithmetic notes (all probed against torch 2.11 CPU / torchvision 0.26 / OpenCV 4.13):
  * non-antialiased bilinear == ATen upsample_bilinear2d, align_corners=False, with the source
    coordinate and the two lerps contracted to FMAs exactly as the AVX2 build does:
        src = fma(scale, dst+0.5, -0.5) clamped at 0;  row = fma(a, wx0, b*wx1);  out = fma(top, wy0, bot*wy1)
    (bit-exact on 100 % of elements in the probe).
  * antialiased bilinear == ATen _upsample_bilinear2d_aa: separable triangle filter, horizontal pass
    first, weights in float32 with the double-precision intermediates of the C++ expression, each
    tap row normalised by the running float32 sum (<= 1 ulp off, identical zero pattern).
"""
from __future__ import annotations

import numpy as np

f32 = np.float32
f64 = np.float64

IMAGENET_MEAN = np.array([0.485, 0.456, 0.406], dtype=f32)          # Hybridgl_main.py:117
IMAGENET_STD = np.array([0.229, 0.224, 0.225], dtype=f32)           # Hybridgl_main.py:117
CLIP_PIXEL_MEAN = np.array([0.48145466, 0.4578275, 0.40821073], dtype=f32)  # Hybridgl_main.py:93


def _fma(a, b, c):
    """float32 fused multiply-add emulated through float64 (the product of two float32 is exact there)."""
    return (np.asarray(a, f64) * np.asarray(b, f64) + np.asarray(c, f64)).astype(f32)


# --------------------------------------------------------------------------------------------------
# resampling primitives
# --------------------------------------------------------------------------------------------------
def bilinear_taps(in_size: int, out_size: int):
    """Index/weight table of ATen upsample_bilinear2d (align_corners=False) -- used by
    T.Resize(..., antialias=None) at Hybridgl_main.py:116,121."""
    if in_size == out_size:
        i0 = np.arange(out_size, dtype=np.int64)
        return i0, i0.copy(), np.ones(out_size, f32), np.zeros(out_size, f32)
    scale = f32(in_size) / f32(out_size)
    d = np.arange(out_size, dtype=f32) + f32(0.5)
    src = np.maximum(_fma(scale, d, f32(-0.5)), f32(0))
    i0 = np.minimum(src.astype(np.int64), in_size - 1)
    i1 = i0 + (i0 < in_size - 1)
    w1 = np.clip((src - i0.astype(f32)).astype(f32), f32(0), f32(1))
    w0 = (f32(1) - w1).astype(f32)
    return i0, i1, w0, w1


def resize_bilinear(x: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """x f32 [..., H, W] -> [..., out_h, out_w], no antialias."""
    x = np.asarray(x, f32)
    y0, y1, wy0, wy1 = bilinear_taps(x.shape[-2], out_h)
    x0, x1, wx0, wx1 = bilinear_taps(x.shape[-1], out_w)
    r0 = x[..., y0, :]
    r1 = x[..., y1, :]
    top = _fma(r0[..., x0], wx0, (r0[..., x1] * wx1).astype(f32))
    bot = _fma(r1[..., x0], wx0, (r1[..., x1] * wx1).astype(f32))
    wy0 = wy0[:, None]; wy1 = wy1[:, None]
    return _fma(top, wy0, (bot * wy1).astype(f32))


def aa_taps(in_size: int, out_size: int):
    """Per-output (xmin, weights[f32]) of ATen's antialiased bilinear filter
    (_compute_indices_min_size_weights_aa); used by TF.resize(pred_masks.float(), (g,g)) under
    torchvision >= 0.17 (model/backbone.py:160)."""
    scale = f32(in_size) / f32(out_size)
    if scale >= 1:
        support = f32(f32(1.0) * scale); invscale = f32(f32(1.0) / scale)
    else:
        support = f32(1.0); invscale = f32(1.0)
    taps = []
    for i in range(out_size):
        center = f32(f64(scale) * (i + 0.5))
        xmin = max(int(f64(f32(center - support)) + 0.5), 0)
        xsize = max(min(int(f64(f32(center + support)) + 0.5), in_size) - xmin, 0)
        w = np.zeros(xsize, f32)
        tot = f32(0)
        for j in range(xsize):
            t = abs(f32((f64(f32(f32(j + xmin) - center)) + 0.5) * f64(invscale)))
            w[j] = f32(1.0) - t if t < 1 else f32(0)
            tot = f32(tot + w[j])
        if tot != 0:
            w = (w / tot).astype(f32)
        taps.append((xmin, w))
    return taps


def resize_bilinear_aa(x: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """x f32 [N,H,W] -> [N,out_h,out_w], antialiased (horizontal pass, then vertical)."""
    x = np.asarray(x, f32)
    n, h, w = x.shape
    tmp = np.zeros((n, h, out_w), f32)
    for j, (xmin, wt) in enumerate(aa_taps(w, out_w)):
        acc = np.zeros((n, h), f32)
        for k in range(wt.size):
            acc = (acc + x[:, :, xmin + k] * wt[k]).astype(f32)
        tmp[:, :, j] = acc
    out = np.zeros((n, out_h, out_w), f32)
    for i, (ymin, wt) in enumerate(aa_taps(h, out_h)):
        acc = np.zeros((n, out_w), f32)
        for k in range(wt.size):
            acc = (acc + tmp[:, ymin + k, :] * wt[k]).astype(f32)
        out[:, i, :] = acc
    return out


# --------------------------------------------------------------------------------------------------
# (a1) per-mask visual-prompt preprocessing -- Hybridgl_main.py:92-125, utils.py:292-345
# --------------------------------------------------------------------------------------------------
def to_tensor_u8(img_u8_hwc: np.ndarray) -> np.ndarray:
    """T.ToTensor on a uint8 HWC array: CHW float32 / 255 (Hybridgl_main.py:115)."""
    return (np.ascontiguousarray(img_u8_hwc.transpose(2, 0, 1)).astype(f32) / f32(255)).astype(f32)


def imagenet_normalize(x_chw: np.ndarray) -> np.ndarray:
    """T.Normalize([0.485,0.456,0.406],[0.229,0.224,0.225]) (Hybridgl_main.py:117; the dataset applies the
    same transform to produce image['image'], data/dataset_refer_bert.py:154-155)."""
    return ((x_chw - IMAGENET_MEAN[:, None, None]).astype(f32) / IMAGENET_STD[:, None, None]).astype(f32)


def prep(image_u8: np.ndarray, blur_u8: np.ndarray, masks: np.ndarray, S: int, background: str = "blur", circle: bool = False):
    """Hybridgl_main.py:92-125.  Returns (local_imgs, global_imgs), both f32 [N,3,S,S].

    global_n = Normalize(Resize(ToTensor(where(m_n, img, background))))     lines 103-118
    local_n  = Resize(img_norm * m_n + (1 - m_n) * pixel_mean)              lines 120-122
    ``background`` selects what replaces the pixels outside the mask in the global view:
    'blur' (the drivers, and apply_visual_prompts 'blur' utils.py:306-320), 'black' (utils.py:336-341), 'none' (the frame itself).
    ``circle``: the global view also carries the 'circle' prompt (utils.py:322-335), drawn after the blur composite and before the
    black one, as apply_visual_prompts orders them.
    """
    masks = np.asarray(masks).astype(bool)
    n = masks.shape[0]
    img_norm = imagenet_normalize(to_tensor_u8(image_u8))
    local = np.zeros((n, 3, S, S), f32)
    glob = np.zeros((n, 3, S, S), f32)
    if background == "blur":
        bg = blur_u8
    elif background == "black":
        bg = np.zeros_like(image_u8)
    elif background == "none":
        bg = image_u8
    else:
        raise ValueError(background)
    for i in range(n):
        m = masks[i]
        if circle:
            kinds = {"blur": ("blur", "circle"), "black": ("circle", "black"), "none": ("circle",)}[background]
            comp = apply_visual_prompt(image_u8, m, kinds, blur_u8)
        else:
            comp = np.where(m[:, :, None], image_u8, bg)                 # u8 composite, lines 106-113
        glob[i] = imagenet_normalize(resize_bilinear(to_tensor_u8(comp), S, S))
        masked = np.where(m[None], img_norm, CLIP_PIXEL_MEAN[:, None, None]).astype(f32)  # line 120
        local[i] = resize_bilinear(masked, S, S)
    return local, glob


def prep_crop(image_u8: np.ndarray, blur_u8: np.ndarray, masks: np.ndarray, crop_xywh: np.ndarray, S: int, background: str = "blur"):
    """prep() with a per-proposal crop box (x, y, w, h): the same two composites, cut to the box, then resized to S x S -- what
    Hybridgl_main.py:92-125 would compute had it used the pred_box it casts at :101 (it does not: an option of the build, pinned
    to prep() by the full-frame box)."""
    masks = np.asarray(masks).astype(bool)
    n = masks.shape[0]
    img_norm = imagenet_normalize(to_tensor_u8(image_u8))
    bg = blur_u8 if background == "blur" else np.zeros_like(image_u8)
    local = np.zeros((n, 3, S, S), f32); glob = np.zeros((n, 3, S, S), f32)
    for i in range(n):
        x, y, w, h = (int(v) for v in crop_xywh[i])
        m = masks[i, y:y + h, x:x + w]
        comp = np.where(m[:, :, None], image_u8[y:y + h, x:x + w], bg[y:y + h, x:x + w])
        glob[i] = imagenet_normalize(resize_bilinear(to_tensor_u8(comp), S, S))
        masked = np.where(m[None], img_norm[:, y:y + h, x:x + w], CLIP_PIXEL_MEAN[:, None, None]).astype(f32)
        local[i] = resize_bilinear(masked, S, S)
    return local, glob


def gaussian_kernel_u8(ksize: int = 15) -> np.ndarray:
    """Fixed-point (Q8) 1-D Gaussian taps used by cv2.GaussianBlur on CV_8U for sigma=0 -> derived sigma
    (getGaussianKernel: sigma = 0.3*((ksize-1)*0.5 - 1) + 0.8; softfloat kernel, then quantised so the
    Q8 taps sum to exactly 256 as in getGaussianKernelFixedPoint_ED)."""
    sigma = 0.3 * ((ksize - 1) * 0.5 - 1) + 0.8
    x = np.arange(ksize, dtype=f64) - (ksize - 1) * 0.5
    k = np.exp(-(x * x) / (2 * sigma * sigma))
    k /= k.sum()
    # error-diffusion quantisation from the centre outwards keeps symmetry and the exact sum
    half = ksize // 2
    q = np.zeros(ksize, dtype=np.int64)
    err = 0.0
    for i in range(half):
        v = k[i] * 256.0 + err
        q[i] = int(np.floor(v + 0.5)); err = v - q[i]
        q[ksize - 1 - i] = q[i]
    q[half] = 256 - 2 * int(q[:half].sum())
    return q


def gaussian_blur_u8(image_u8: np.ndarray, ksize: int = 15) -> np.ndarray:
    """cv2.GaussianBlur(img,(15,15),0) on uint8, BORDER_REFLECT_101 (Hybridgl_main.py:99).
    Row pass in Q8 (u16), column pass in Q16 (u32), round-half-up once at the end."""
    q = gaussian_kernel_u8(ksize).astype(np.uint32)
    r = ksize // 2
    h, w, _ = image_u8.shape
    xi = np.abs(np.arange(-r, w + r)); xi = np.where(xi >= w, 2 * (w - 1) - xi, xi)
    yi = np.abs(np.arange(-r, h + r)); yi = np.where(yi >= h, 2 * (h - 1) - yi, yi)
    src = image_u8.astype(np.uint32)
    row = np.zeros((h, w, 3), np.uint32)
    for k in range(ksize):
        row += src[:, xi[k:k + w], :] * q[k]
    col = np.zeros((h, w, 3), np.uint32)
    for k in range(ksize):
        col += row[yi[k:k + h], :, :] * q[k]
    return np.minimum((col + (1 << 15)) >> 16, 255).astype(np.uint8)


# --------------------------------------------------------------------------------------------------
# (a2)-(a5) mask grid, attention mask, token masking / stream fusion -- model/backbone.py
# --------------------------------------------------------------------------------------------------
def mask_to_grid(masks: np.ndarray, g: int, antialias: bool = True) -> np.ndarray:
    """model/backbone.py:160  TF.resize(pred_masks.float(), (g, g)).  antialias=True is what torchvision
    0.26 (this container) does; the reference's pinned 0.15.2 does not antialias tensors (App. B-1)."""
    m = np.asarray(masks).astype(f32)
    return resize_bilinear_aa(m, g, g) if antialias else resize_bilinear(m, g, g)


def make_attn_mask(grid: np.ndarray, heads: int) -> np.ndarray:
    """model/backbone.py:108-115.  grid f32 [N,g,g] -> bool [N*heads, L+1, L+1], True = blocked.
    Only row 0 (the CLS query) is masked, at patch keys whose soft mask value is exactly 0."""
    n = grid.shape[0]
    L = grid.shape[1] * grid.shape[2]
    keep = np.ones((n * heads, L + 1, L + 1), dtype=bool)
    keep[:, 0, 1:] = np.repeat((grid.reshape(n, L) != 0), heads, axis=0)
    return ~keep


def token_mask(x: np.ndarray, grid: np.ndarray) -> np.ndarray:
    """model/backbone.py:235-247.  x f32 [L+1,N,D] (LND): patch tokens scaled by the soft grid mask, CLS kept."""
    n = grid.shape[0]
    out = np.array(x, dtype=f32, copy=True)
    out[1:] = (x[1:] * grid.reshape(n, -1).T[:, :, None]).astype(f32)
    return out


def fuse_streams(x_masked_src: np.ndarray, grid, a: float, x_add: np.ndarray, b: float) -> np.ndarray:
    """The pre-block mixes of blocks masking_block..last (model/backbone.py:216,249,290,291):
    out = a * tokenmask(x_masked_src, grid) + b * x_add   (grid=None -> no token masking)."""
    t = token_mask(x_masked_src, grid) if grid is not None else np.asarray(x_masked_src, f32)
    return ((t * f32(a)).astype(f32) + (np.asarray(x_add, f32) * f32(b)).astype(f32)).astype(f32)


# --------------------------------------------------------------------------------------------------
# (a6)-(a9),(a12) scoring, selection, spatial relationship -- model/backbone.py:74-87, Hybridgl_main.py:153-228
# --------------------------------------------------------------------------------------------------
def calculate_score(image_features: np.ndarray, text_features: np.ndarray, logit_scale_exp: float) -> np.ndarray:
    """model/backbone.py:74-87: logit_scale.exp() * normalize(img) @ normalize(txt).T -> [N,T]."""
    f = np.asarray(image_features, f32)
    t = np.asarray(text_features, f32)
    with np.errstate(divide="ignore", invalid="ignore"):
        fn = f / np.sqrt((f * f).sum(1, keepdims=True, dtype=f32))
        tn = t / np.sqrt((t * t).sum(1, keepdims=True, dtype=f32))
        return ((f32(logit_scale_exp) * fn) @ tn.T).astype(f32)


def softmax0(x: np.ndarray) -> np.ndarray:
    """torch.nn.Softmax(0) (Hybridgl_main.py:60)."""
    x = np.asarray(x, f32)
    with np.errstate(invalid="ignore"):
        e = np.exp(x - np.max(x, axis=0, keepdims=True))
        return (e / e.sum(0, keepdims=True, dtype=f32)).astype(f32)


def topk_indices(x: np.ndarray, k: int) -> np.ndarray:
    """torch.topk(x, k)[1] on 1-D input: descending, NaN sorts as the largest, first index wins ties."""
    x = np.asarray(x, f32).reshape(-1)
    key = np.where(np.isnan(x), np.inf, x)
    return np.argsort(-key, kind="stable")[:k].astype(np.int64)


def relation_boxes(boxi, boxj, scorei, scorej, relaword):
    """utils.py:240-268 (XYWH boxes; centres are x+w/2, y+h/2)."""
    bi = [f32(v) for v in boxi]; bj = [f32(v) for v in boxj]
    si = f32(scorei); sj = f32(scorej)
    if relaword == "left":
        return f32(si * sj * f32((bi[0] + bi[2] / 2) < (bj[0] + bj[2] / 2)))
    if relaword == "right":
        return f32(si * sj * f32((bi[0] + bi[2] / 2) > (bj[0] + bj[2] / 2)))
    if relaword == "up":
        return f32(si * sj * f32((bi[1] + bi[3] / 2) < (bj[1] + bj[3] / 2)))
    if relaword == "down":
        return f32(si * sj * f32((bi[1] + bi[3] / 2) > (bj[1] + bj[3] / 2)))
    if relaword == "big":
        return f32(si * sj * f32((bi[2] * bi[3]) > (bj[2] * bj[3])))
    if relaword == "small":
        return f32(si * sj * f32((bi[2] * bi[3]) < (bj[2] * bj[3])))
    if relaword == "within":
        x1 = max(bi[0], bj[0]); x2 = max(x1, min(bi[0] + bi[2], bj[0] + bj[2]))
        y1 = max(bi[1], bj[1]); y2 = max(y1, min(bi[1] + bi[3], bj[1] + bj[3]))
        with np.errstate(divide="ignore", invalid="ignore"):
            return f32(f32(f32(si * sj) * f32(x2 - x1)) * f32(y2 - y1) / f32(bi[2] * bi[3]))
    return si  # 'none' and anything unknown


def gen_dir_mask(dirflag: str, height: int, width: int) -> np.ndarray:
    """utils.py:135-161: horizontal position ramp; up/down/none are all-ones (the vertical ramps are commented out)."""
    def linspace(a, b, n):  # torch.linspace float32 (ATen RangeFactories): evaluated from both ends
        if n == 1:
            return np.array([a], f32)
        step = (f32(b) - f32(a)) / f32(n - 1)
        i = np.arange(n)
        lo = _fma(step, i.astype(f32), f32(a))                      # the AVX2 build contracts start + step*i
        hi = _fma(-step, (n - 1 - i).astype(f32), f32(b))
        return np.where(i < n // 2, lo, hi).astype(f32)
    if dirflag == "left":
        row = linspace(1, 0, width)
    elif dirflag == "right":
        row = linspace(0, 1, width)
    elif dirflag == "middle":
        row = np.concatenate([linspace(0, 1, width // 2), linspace(1, 0, width - width // 2)])
    else:
        row = np.ones(width, f32)
    return np.broadcast_to(row[None, :], (height, width)).astype(f32)


def condition_heatmap(heatmap: np.ndarray, dirflag: str) -> np.ndarray:
    """Hybridgl_main.py:204-209: min-max normalise, multiply by the position ramp, divide by the mean."""
    a = np.asarray(heatmap, f32)
    a = ((a - a.min()) / (a.max() - a.min())).astype(f32)
    a = (a * gen_dir_mask(dirflag, a.shape[0], a.shape[1])).astype(f32)
    return (a / a.mean(dtype=f32)).astype(f32)


def black_for(relaflag: str) -> float:
    """Hybridgl_main.py:211-216."""
    return 1.95 if relaflag == "big" else (1.5 if relaflag == "small" else 1.8)


def gem_pool(cond_heatmap: np.ndarray, masks: np.ndarray, black: float) -> np.ndarray:
    """Hybridgl_main.py:218-223 in the closed form of SURVEY.md Appendix A-2:
    score_gem[n] = (2-black)*sum(A*m)/area(m) - black*sum(A*(1-m))/area(1-m)."""
    a = np.asarray(cond_heatmap, f64)
    m = np.asarray(masks).astype(bool)
    hw = a.size
    s_in = (a[None] * m).reshape(m.shape[0], -1).sum(1)
    area = m.reshape(m.shape[0], -1).sum(1).astype(f64)
    s_tot = a.sum()
    with np.errstate(divide="ignore", invalid="ignore"):
        return ((2 - black) * s_in / area - black * (s_tot - s_in) / (hw - area)).astype(f32)


def aa_matrix(in_size: int, out_size: int) -> np.ndarray:
    """The antialiased bilinear filter of aa_taps() as a dense [out_size, in_size] float64 matrix."""
    U = np.zeros((out_size, in_size), f64)
    for i, (xmin, wt) in enumerate(aa_taps(in_size, out_size)):
        U[i, xmin:xmin + wt.size] = wt
    return U


def gem_pool_token_space(heat_raw: np.ndarray, masks: np.ndarray, dirflag: str, black: float) -> np.ndarray:
    """Hybridgl_main.py:200-223 evaluated in token space: with A = Uy h Ux^T the pooled sums are
    S_in[n] = kk * (G_n . h - mn * sum G_n),  G_n = Uy^T (m_n * ramp) Ux,  kk = H*W / (C . h - mn * sum C),  C = G of the full frame,
    mn = min A, S_tot = H*W.  The matrix form of resize_bilinear_aa -> condition_heatmap -> gem_pool; tests pin it to that chain."""
    h = np.asarray(heat_raw, f64)
    m = np.asarray(masks).astype(bool)
    n, H, W = m.shape
    Uy, Ux = aa_matrix(h.shape[0], H), aa_matrix(h.shape[1], W)
    A = Uy @ h @ Ux.T
    ramp = gen_dir_mask(dirflag, 1, W)[0].astype(f64)
    mn = A.min()
    G = np.einsum("yi,nyx,xj->nij", Uy, m * ramp[None, None, :], Ux)
    C = np.einsum("yi,x,xj->ij", Uy, ramp, Ux)
    kk = H * W / ((C * h).sum() - mn * C.sum())
    s_in = kk * ((G * h[None]).sum((1, 2)) - mn * G.sum((1, 2)))
    area = m.reshape(n, -1).sum(1).astype(f64)
    with np.errstate(divide="ignore", invalid="ignore"):
        return ((2 - black) * s_in / area - black * (H * W - s_in) / (H * W - area)).astype(f32)


def score_and_select(features, sentence_feat, noun_feat, other_feats, boxes, relaflag,
                     score_gem=None, logit_scale_exp: float = 100.0, r: float = 0.5, alpha: float = 0.6,
                     k1: int = 3, k2: int = 6):
    """Hybridgl_main.py:153-196 and 225-227 for ONE expression.

    features f32 [N,De]; sentence_feat/noun_feat f32 [De]; other_feats f32 [K,De] (K may be 0);
    boxes int64 [N,4] XYWH; score_gem f32 [N] or None (-> final == relation-only re-ranking).
    Returns a dict with score_clip [N] (pre-softmax), idx_hybrid, top_idx [k1], relation [k1]
    (post-softmax, pre-blend), blended [k1], idx_final.
    """
    feats = np.asarray(features, f32)
    n = feats.shape[0]
    text = (f32(r) * np.asarray(sentence_feat, f32) + f32(1 - r) * np.asarray(noun_feat, f32)).astype(f32)   # :153
    score_clip = calculate_score(feats, text[None], logit_scale_exp)[:, 0]                                  # :154
    other_feats = np.asarray(other_feats, f32).reshape(-1, feats.shape[1])
    n_other = other_feats.shape[0]
    other = np.zeros(feats.shape[1], f32)                                                                    # :157
    for k in range(n_other):
        other = (other + other_feats[k]).astype(f32)                                                         # :161
    if n_other:
        other = (other / f32(n_other)).astype(f32)                                                           # :164
    score_neg = calculate_score(feats, other[None], logit_scale_exp)[:, 0]                                   # :166 (NaN if no others)
    idx_hybrid = int(np.argmax(score_clip))                                                                  # :168
    p = softmax0(score_clip); pneg = softmax0(score_neg)                                                     # :173-174
    k1 = min(k1, n); k2 = min(k2, n)                                                                         # :178-181
    top = topk_indices(p, k1); topneg = topk_indices(pneg, k2)                                               # :182-183
    rel = np.zeros(k1, f32)
    for i in range(k1):                                                                                      # :185-193
        js, q = (top, p) if n_other == 0 else (topneg, pneg)
        for j in js:
            rel[i] = f32(rel[i] + relation_boxes(boxes[top[i]], boxes[j], p[top[i]], q[j], relaflag))
    rel = softmax0(rel)                                                                                      # :196
    blended = rel.copy()
    if score_gem is not None:
        for i in range(k1):                                                                                  # :225-226
            blended[i] = f32(f32(rel[i] * f32(1 - alpha)) + f32(f32(alpha) * f32(score_gem[top[i]])))
    idx_final = int(top[int(np.argmax(blended))])                                                            # :227
    return dict(score_clip=score_clip, score_neg=score_neg, idx_hybrid=idx_hybrid, top_idx=top,
                relation=rel, blended=blended, idx_final=idx_final)


def mask_pool_tokens(weights: np.ndarray, tokens: np.ndarray, normalize: bool = True) -> np.ndarray:
    """Token-space form of the pooling loop Hybridgl_main.py:218-223:
    pooled[n,:] = sum_l weights[n,l] * tokens[l,:]  (the masks x tokens x D contraction), optionally followed by the
    L2 normalisation of model/backbone.py:79.  weights f32 [N,L] (used exactly as given), tokens [L,D]; accumulated in
    float64.  No reference code computes this directly; tests/test_oracle_golden.py pins it to gem_pool through the
    identity  sum_p m[p] * (U h)[p] == (U^T m) . h  for the linear up-sampling U of the heat-map."""
    w = np.asarray(weights, f32).reshape(weights.shape[0], -1).astype(f64)
    pooled = (w @ np.asarray(tokens, f64)).astype(f32)
    if normalize:
        with np.errstate(divide="ignore", invalid="ignore"):
            pooled = (pooled / np.sqrt((pooled.astype(f64) ** 2).sum(1, keepdims=True))).astype(f32)
    return pooled


# --------------------------------------------------------------------------------------------------
# (a13) IoU accounting -- utils.py:365-384, Hybridgl_main.py:240-247
# --------------------------------------------------------------------------------------------------
# --------------------------------------------------------------------------------------------------
# SAM run-length proposals (third_party/segment-anything/segment_anything/utils/amg.py)
# --------------------------------------------------------------------------------------------------
def mask_to_rle(mask: np.ndarray) -> dict:
    """amg.py:107-135 mask_to_rle_pytorch for one mask: column-major (Fortran) runs, first run counts zeros
    (a leading 0 is emitted when the first pixel is set, amg.py:132)."""
    h, w = mask.shape
    flat = np.asarray(mask, bool).T.reshape(-1)                      # amg.py:114 permute(0,2,1).flatten(1)
    change = np.nonzero(flat[1:] ^ flat[:-1])[0] + 1                  # amg.py:117-118, :126
    edges = np.concatenate([[0], change, [h * w]])                   # amg.py:123-129
    counts = np.diff(edges).tolist()                                 # amg.py:130
    if flat[0]:
        counts = [0] + counts                                        # amg.py:131
    return {"size": [h, w], "counts": counts}


def rle_to_mask(rle: dict) -> np.ndarray:
    """amg.py:138-149: fill alternating runs into a flat column-major buffer, reshape (w,h), transpose."""
    h, w = rle["size"]
    counts = np.asarray(rle["counts"], np.int64)
    parity = (np.arange(counts.size) & 1).astype(bool)               # amg.py:143-148
    flat = np.repeat(parity, counts)[: h * w]
    if flat.size < h * w:                                            # np.empty tail of a short RLE: define it as zeros
        flat = np.concatenate([flat, np.zeros(h * w - flat.size, bool)])
    return flat.reshape(w, h).T                                      # amg.py:148-149


def pack_bits(masks: np.ndarray) -> np.ndarray:
    """The packed-mask format of libhgl: uint32 [M,H,ceil(W/32)], bit i of word w = pixel 32*w+i (no reference
    counterpart -- a storage format; defined here so that tests can state it independently of the kernels)."""
    m = np.asarray(masks).astype(bool)
    M, H, W = m.shape
    WW = (W + 31) // 32
    pad = np.zeros((M, H, WW * 32), bool)
    pad[:, :, :W] = m
    by = np.packbits(pad.reshape(M, H, WW, 32), axis=-1, bitorder="little")      # [M,H,WW,4] bytes, little endian
    return by.view("<u4").reshape(M, H, WW)


def compute_iou(pred: np.ndarray, target: np.ndarray):
    """utils.py:365-384: integer I, U and this_iou = I/U (0 when U == 0)."""
    p = np.asarray(pred).astype(bool); t = np.asarray(target).astype(bool)
    i = int(np.logical_and(p, t).sum()); u = int(np.logical_or(p, t).sum())
    return i, u, (0.0 if u == 0 else float(f32(i) / f32(u)))


def report(cum_i: int, cum_u: int, ious) -> tuple:
    """Hybridgl_main.py:240-245: oIoU = cum_I*100/cum_U ; mIoU = mean(per-expression IoU)*100."""
    o = cum_i * 100.0 / cum_u if cum_u else float("nan")
    m = float(np.mean(np.asarray(ious, f32), dtype=f32)) * 100.0 if len(ious) else float("nan")
    return o, m


# --------------------------------------------------------------------------------------------------
# (a1') apply_visual_prompts utils.py:292-345 and its helpers: mask2chw (utils.py:280-289), SAM's boxes, cv2.ellipse outline
# --------------------------------------------------------------------------------------------------
def mask2chw(mask: np.ndarray):
    """utils.py:280-289: ((center_y, center_x), height, width); the reference raises on an empty mask (mean of nothing)."""
    rows, cols = np.nonzero(np.asarray(mask).astype(bool))
    if rows.size == 0:
        raise ValueError("empty mask")
    return (int(np.mean(rows)), int(np.mean(cols))), int(rows.max() - rows.min() + 1), int(cols.max() - cols.min() + 1)


def mask_to_box_xywh(mask: np.ndarray) -> np.ndarray:
    """amg.py:303-346 batched_mask_to_box + :91-95 box_xyxy_to_xywh for one mask: inclusive edges, w = x1 - x0, h = y1 - y0."""
    ys, xs = np.nonzero(np.asarray(mask).astype(bool))
    if ys.size == 0:
        return np.zeros(4, np.int64)
    return np.array([xs.min(), ys.min(), xs.max() - xs.min(), ys.max() - ys.min()], np.int64)


# ---- cv2.ellipse outline (the 'circle' visual prompt, utils.py:322-335) ---------------------------------------------
# OpenCV is a binary dependency of the reference (no source in its tree): cv2.ellipse(img, center, axes, 0, 0, 360, color, 1) as
# shipped in this container (4.13) is restated from OpenCV's published algorithm (imgproc/src/drawing.cpp: ellipse -> EllipseEx ->
# ellipse2Poly -> PolyLine -> ThickLine -> Line -> LineIterator) and pinned against cv2 itself: tests/test_oracle_golden.py draws
# thousands of random ellipses and integer lines with cv2 when it is importable, and the reference's own apply_visual_prompts
# outputs are in tests/golden/geometry.npz.  What the algorithm does at thickness 1, LINE_8:
#   * polygon vertices every `delta` degrees (delta from the larger axis: < 3 -> 90, < 10 -> 30, < 15 -> 18, else 5), computed in
#     16.16 fixed point as double: c*2^16 + (axis*2^16) * (double)SinTable[deg] (SinTable = sin of whole degrees as 7-decimal float
#     literals), rounded half-to-even, then to the nearest pixel with (v + 2^15) >> 16;
#   * every edge is an integer 8-connected Bresenham line drawn LEFT TO RIGHT (LineIterator(..., leftToRight=true)) after
#     cv::clipLine moved the end points that lie outside the frame onto its border (double arithmetic, truncated).
_SIN_TABLE = None


def _sin_table():
    global _SIN_TABLE
    if _SIN_TABLE is None:
        import math
        _SIN_TABLE = np.array([f32(f"{math.sin(math.radians(d)):.7f}") for d in range(451)], f32)
    return _SIN_TABLE


def _clip_line(W, H, x1, y1, x2, y2):
    """cv::clipLine on integer end points; returns (visible, x1, y1, x2, y2)."""
    right, bottom = W - 1, H - 1

    def code(x, y):
        return (x < 0) + (x > right) * 2 + (y < 0) * 4 + (y > bottom) * 8
    c1, c2 = code(x1, y1), code(x2, y2)
    if (c1 & c2) == 0 and (c1 | c2) != 0:
        if c1 & 12:
            a = 0 if c1 < 8 else bottom
            x1 += int(float(a - y1) * (x2 - x1) / (y2 - y1)); y1 = a
            c1 = (x1 < 0) + (x1 > right) * 2
        if c2 & 12:
            a = 0 if c2 < 8 else bottom
            x2 += int(float(a - y2) * (x2 - x1) / (y2 - y1)); y2 = a
            c2 = (x2 < 0) + (x2 > right) * 2
        if (c1 & c2) == 0 and (c1 | c2) != 0:
            if c1:
                a = 0 if c1 == 1 else right
                y1 += int(float(a - x1) * (y2 - y1) / (x2 - x1)); x1 = a; c1 = 0
            if c2:
                a = 0 if c2 == 1 else right
                y2 += int(float(a - x2) * (y2 - y1) / (x2 - x1)); x2 = a; c2 = 0
    return (c1 | c2) == 0, x1, y1, x2, y2


def line8_pixels(H, W, p1, p2):
    """Pixels (y, x) of cv2.line(img, p1, p2, color, 1, LINE_8) on an H x W frame (points are (x, y), may lie outside)."""
    ok, x1, y1, x2, y2 = _clip_line(W, H, int(p1[0]), int(p1[1]), int(p2[0]), int(p2[1]))
    if not ok:
        return []
    dx, dy = x2 - x1, y2 - y1
    if dx < 0:                                     # left to right
        x1, y1, x2, y2 = x2, y2, x1, y1
        dx, dy = -dx, -dy
    sy = 1 if dy >= 0 else -1
    dy = abs(dy)
    x, y, out = x1, y1, []
    if dy > dx:                                    # y is the major axis
        err, plus, minus = dy - 2 * dx, 2 * dy, -2 * dx
        for _ in range(dy + 1):
            out.append((y, x))
            if err < 0:
                err += minus + plus; x += 1
            else:
                err += minus
            y += sy
    else:
        err, plus, minus = dx - 2 * dy, 2 * dx, -2 * dy
        for _ in range(dx + 1):
            out.append((y, x))
            if err < 0:
                err += minus + plus; y += sy
            else:
                err += minus
            x += 1
    return out


def ellipse_vertices(cx: int, cy: int, ax: int, ay: int):
    """Pixel vertices (x, y) of the polygon cv2.ellipse draws for centre (cx, cy), half axes (ax, ay), angle 0, full arc."""
    st = _sin_table()
    m = max(ax, ay)
    delta = 90 if m < 3 else 30 if m < 10 else 18 if m < 15 else 5
    pts = []
    for i in range(0, 360 + delta, delta):
        ang = min(i, 360)
        vx = float(cx * 65536) + float(ax * 65536) * float(st[450 - ang])      # two roundings, like the C++ expression
        vy = float(cy * 65536) + float(ay * 65536) * float(st[ang])
        pts.append(((int(np.rint(vx)) + 32768) >> 16, (int(np.rint(vy)) + 32768) >> 16))
    return pts


def ellipse_outline(H: int, W: int, cx: int, cy: int, ax: int, ay: int) -> np.ndarray:
    """bool [H,W]: the pixels cv2.ellipse(img, (cx, cy), (ax, ay), 0, 0, 360, color, 1) sets."""
    out = np.zeros((H, W), bool)
    v = ellipse_vertices(cx, cy, ax, ay)
    for p0, p1 in zip(v[:-1], v[1:]):
        for y, x in line8_pixels(H, W, p0, p1):
            out[y, x] = True
    return out


def circle_outline_of_mask(mask: np.ndarray) -> np.ndarray:
    """utils.py:322-335: the ellipse at mask2chw's centre with half axes (width // 2, height // 2)."""
    (cy, cx), h, w = mask2chw(mask)
    return ellipse_outline(mask.shape[0], mask.shape[1], cx, cy, w // 2, h // 2)


CIRCLE_COLOR = np.array([255, 0, 0], np.uint8)      # utils.py:298 default colour


def apply_visual_prompt(image_u8: np.ndarray, mask: np.ndarray, kind, blur_u8: np.ndarray = None) -> np.ndarray:
    """utils.py:292-345.  image u8 [H,W,3]; `kind` a prompt type or a tuple of them, applied in the reference's order
    'blur' (:306-320) -> 'circle' (:322-335) -> 'black' (:336-341); returns u8 [H,W,3]."""
    kinds = (kind,) if isinstance(kind, str) else tuple(kind)
    for k in kinds:
        if k not in ("blur", "circle", "black"):
            raise ValueError(k)
    m = np.asarray(mask).astype(bool)
    img = image_u8
    if "blur" in kinds:
        img = np.where(m[:, :, None], img, blur_u8 if blur_u8 is not None else gaussian_blur_u8(img))
    if "circle" in kinds:
        img = np.where(circle_outline_of_mask(m)[:, :, None], CIRCLE_COLOR, img)
    if "black" in kinds:
        img = np.where(m[:, :, None], img, np.zeros_like(img))
    return img.astype(np.uint8)
