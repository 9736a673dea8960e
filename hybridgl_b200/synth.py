"""Seeded synthetic inputs at the boundary of the mask-proposal scoring path.

The reference consumes SAM output (``masks bool[N,H,W]``, ``boxes int64[N,4]`` XYWH,
Hybridgl_main.py:85-90), a raw RGB frame (``image['sam_img']``, data/dataset_refer_bert.py:113),
CLIP text embeddings (Hybridgl_main.py:150-166), a GEM heat-map (Hybridgl_main.py:200-201) and a
ground-truth mask (data/dataset_refer_bert.py:116-121).  None of the producers (SAM / GEM / spaCy /
RefCOCO) is available offline, so this module draws tensors of the same shape, dtype and statistics
(SURVEY.md section 8(d) and Appendix D).  Pure numpy, no device code: the same generator feeds the
CUDA path, the oracle and the golden-vector script.
"""
from __future__ import annotations

import dataclasses
from typing import List, Optional

import numpy as np

DIRFLAGS = ("none", "left", "right", "middle", "up", "down")          # utils.py:102-133
RELAFLAGS = ("none", "left", "right", "up", "down", "big", "small", "within")  # utils.py:240-268


@dataclasses.dataclass
class Expression:
    """Per-expression host-side inputs (what spaCy + encode_text + GEM hand to the path)."""
    sentence_feat: np.ndarray            # f32 [De]   encode_text(sentence)            Hybridgl_main.py:150
    noun_feat: np.ndarray                # f32 [De]   encode_text(noun_phrase)         Hybridgl_main.py:151
    other_feats: np.ndarray              # f32 [K,De] encode_text('a photo of '+noun)  Hybridgl_main.py:159-161 (K may be 0)
    dirflag: str                         # extract_dir_phrase                          Hybridgl_main.py:143
    relaflag: str                        # extract_rela_word                           Hybridgl_main.py:177
    heatmap: np.ndarray                  # f32 [H,W]  GEM map after T.Resize           Hybridgl_main.py:200-201
    heat_raw: Optional[np.ndarray] = None   # f32 [h',w'] the map as the GEM model returns it (before T.Resize, :200); an
    #                                         alternative input: resize_bilinear_aa(heat_raw) replaces `heatmap` when the raw path is used


@dataclasses.dataclass
class Item:
    """One image with its proposals and expressions (one DataLoader item of the reference)."""
    image: np.ndarray                    # u8 [H,W,3] RGB
    masks: np.ndarray                    # bool [N,H,W]
    boxes: np.ndarray                    # int64 [N,4] XYWH
    target: np.ndarray                   # u8 [H,W] {0,1}
    expressions: List[Expression]
    features: Optional[np.ndarray] = None  # f32 [N,De] hybrid CLIP features (stand-in for CLIPViTFM.forward)

    @property
    def n_masks(self) -> int:
        return int(self.masks.shape[0])


def _smooth_noise(rng: np.random.Generator, h: int, w: int, c: int, cells: int = 12) -> np.ndarray:
    """Low-pass noise: bilinear upsample of a coarse random lattice (keeps blur != identity)."""
    gy, gx = cells + 2, cells + 2
    lat = rng.random((gy, gx, c), dtype=np.float32)
    ys = np.linspace(0, gy - 1.001, h, dtype=np.float32)
    xs = np.linspace(0, gx - 1.001, w, dtype=np.float32)
    y0 = ys.astype(np.int64); x0 = xs.astype(np.int64)
    fy = (ys - y0)[:, None, None]; fx = (xs - x0)[None, :, None]
    a = lat[y0][:, x0]; b = lat[y0][:, x0 + 1]; cc = lat[y0 + 1][:, x0]; d = lat[y0 + 1][:, x0 + 1]
    return (a * (1 - fy) * (1 - fx) + b * (1 - fy) * fx + cc * fy * (1 - fx) + d * fy * fx).astype(np.float32)


def make_image(rng: np.random.Generator, h: int, w: int) -> np.ndarray:
    base = _smooth_noise(rng, h, w, 3)
    tex = rng.random((h, w, 3), dtype=np.float32)
    img = 0.75 * base + 0.25 * tex
    return np.clip(img * 255.0 + 0.5, 0, 255).astype(np.uint8)


def make_masks(rng: np.random.Generator, n: int, h: int, w: int,
               min_area: int = 800, max_frac: float = 0.6) -> np.ndarray:
    """SAM-style blobs: rotated ellipses, area log-uniform in [min_area, max_frac*H*W]
    (SAM is run with min_mask_region_area=800, Hybridgl_main.py:74)."""
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    out = np.zeros((n, h, w), dtype=bool)
    min_area = max(4, min(min_area, h * w // 8))
    lo, hi = np.log(float(min_area)), np.log(max_frac * h * w)
    for i in range(n):
        for _ in range(16):
            area = float(np.exp(rng.uniform(lo, hi)))
            ratio = float(np.exp(rng.uniform(-0.9, 0.9)))
            ra = np.sqrt(area / np.pi * ratio); rb = np.sqrt(area / np.pi / ratio)
            cy = rng.uniform(0.1 * h, 0.9 * h); cx = rng.uniform(0.1 * w, 0.9 * w)
            th = rng.uniform(0, np.pi)
            dx, dy = xx - cx, yy - cy
            u = dx * np.cos(th) + dy * np.sin(th); v = -dx * np.sin(th) + dy * np.cos(th)
            m = (u / ra) ** 2 + (v / rb) ** 2 <= 1.0
            if m.sum() >= min(min_area, h * w // 8) and not m.all():
                out[i] = m
                break
        else:  # degenerate tiny frames in unit tests
            out[i, h // 4: h // 2 + 1, w // 4: w // 2 + 1] = True
    return out


def masks_to_boxes(masks: np.ndarray) -> np.ndarray:
    """XYWH boxes exactly as SAM derives them: inclusive XYXY edges, w = x1 - x0, h = y1 - y0
    (segment_anything/utils/amg.py batched_mask_to_box + box_xyxy_to_xywh)."""
    n = masks.shape[0]
    boxes = np.zeros((n, 4), dtype=np.int64)
    for i in range(n):
        ys, xs = np.nonzero(masks[i])
        if ys.size == 0:
            continue
        x0, x1, y0, y1 = xs.min(), xs.max(), ys.min(), ys.max()
        boxes[i] = (x0, y0, x1 - x0, y1 - y0)
    return boxes


def masks_to_rle(masks: np.ndarray):
    """SAM's uncompressed RLE of every mask (amg.py:107-135: column-major runs, first run counts zeros), laid out for the
    C ABI: (counts int32 [R], rle_off int32 [M+1]).  Host-side data generation only."""
    m = np.asarray(masks, bool)
    M, h, w = m.shape
    flat = m.transpose(0, 2, 1).reshape(M, h * w)
    counts, off = [], [0]
    for i in range(M):
        change = np.nonzero(flat[i, 1:] ^ flat[i, :-1])[0] + 1
        c = np.diff(np.concatenate([[0], change, [h * w]]))
        if flat[i, 0]:
            c = np.concatenate([[0], c])
        counts.append(c.astype(np.int32))
        off.append(off[-1] + c.size)
    return (np.concatenate(counts) if counts else np.zeros(0, np.int32)), np.asarray(off, np.int32)


def masks_to_rle_device(masks):
    """The same on a torch device tensor bool [M,H,W] (benchmark-sized batches): returns (counts int32 [R], rle_off int32 [M+1])."""
    import torch
    M, h, w = masks.shape
    counts, lens = [], []
    for s in range(0, M, 64):                                                           # chunked: keeps temporaries small
        flat = masks[s:s + 64].permute(0, 2, 1).reshape(-1, h * w)
        n = flat.shape[0]
        first = flat[:, :1]
        # a run starts at position 0 (the leading zeros run, length 0 when the first pixel is set) and wherever the value changes
        chg = torch.cat([torch.ones_like(first), flat[:, 1:] ^ flat[:, :-1]], dim=1)
        mi, pos = chg.nonzero(as_tuple=True)
        lead = first[:, 0]                                                               # masks that need the extra leading 0
        per = torch.bincount(mi, minlength=n) + lead.long()
        nxt = torch.cat([pos[1:], pos.new_zeros(1)])
        last = torch.cat([mi[1:] != mi[:-1], mi.new_ones(1, dtype=torch.bool)])
        run = torch.where(last, h * w - pos, nxt - pos)
        # splice the leading zero counts in: output slot of every run = its rank + number of leads of masks <= its own
        lead_before = torch.cumsum(lead.long(), 0)
        slot = torch.arange(pos.numel(), device=masks.device) + lead_before[mi]
        out = torch.zeros(int(per.sum()), dtype=torch.int32, device=masks.device)
        out[slot] = run.to(torch.int32)
        counts.append(out); lens.append(per)
    lens = torch.cat(lens)
    off = torch.zeros(M + 1, dtype=torch.int64, device=masks.device)
    off[1:] = torch.cumsum(lens, 0)
    return torch.cat(counts), off.to(torch.int32)


def make_target(rng: np.random.Generator, masks: np.ndarray) -> np.ndarray:
    """Ground truth = one proposal shifted by a few pixels so that IoU is neither 0 nor 1."""
    k = int(rng.integers(0, masks.shape[0]))
    dy, dx = int(rng.integers(-6, 7)), int(rng.integers(-6, 7))
    t = np.roll(np.roll(masks[k], dy, axis=0), dx, axis=1)
    return t.astype(np.uint8)


def bf16_round(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even f32 -> bf16 -> f32 (the build's 'bf16 inputs' contract)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32).reshape(x.shape)


def make_expression(rng: np.random.Generator, h: int, w: int, de: int,
                    n_other: Optional[int] = None, dirflag: Optional[str] = None,
                    relaflag: Optional[str] = None, hm_grid=(28, 37), bf16: bool = True,
                    anchor: Optional[np.ndarray] = None) -> Expression:
    if n_other is None:
        n_other = int(rng.integers(0, 4))
    if dirflag is None:
        dirflag = DIRFLAGS[int(rng.integers(0, len(DIRFLAGS)))]
    if relaflag is None:
        relaflag = RELAFLAGS[int(rng.integers(0, len(RELAFLAGS)))]
    sent = rng.standard_normal(de).astype(np.float32)
    noun = (0.6 * sent + 0.8 * rng.standard_normal(de)).astype(np.float32)
    if anchor is not None:  # make the expression "refer to" one proposal so that scores are peaked
        sent = sent + 1.5 * anchor; noun = noun + 1.5 * anchor
    other = rng.standard_normal((n_other, de)).astype(np.float32)
    coarse = rng.random((hm_grid[0], hm_grid[1]), dtype=np.float32)
    ys = np.clip((np.arange(h) + 0.5) * hm_grid[0] / h - 0.5, 0, hm_grid[0] - 1)
    xs = np.clip((np.arange(w) + 0.5) * hm_grid[1] / w - 0.5, 0, hm_grid[1] - 1)
    y0 = np.floor(ys).astype(np.int64); x0 = np.floor(xs).astype(np.int64)
    y1 = np.minimum(y0 + 1, hm_grid[0] - 1); x1 = np.minimum(x0 + 1, hm_grid[1] - 1)
    fy = (ys - y0).astype(np.float32)[:, None]; fx = (xs - x0).astype(np.float32)[None, :]
    hm = (coarse[y0][:, x0] * (1 - fy) * (1 - fx) + coarse[y0][:, x1] * (1 - fy) * fx
          + coarse[y1][:, x0] * fy * (1 - fx) + coarse[y1][:, x1] * fy * fx).astype(np.float32)
    if bf16:
        sent, noun, other = bf16_round(sent), bf16_round(noun), bf16_round(other)
    return Expression(sent, noun, other.reshape(n_other, de), dirflag, relaflag, hm, coarse)


def make_item(seed: int, h: int = 480, w: int = 640, n_masks: int = 64, n_expr: int = 1,
              de: int = 512, with_features: bool = True, bf16: bool = True,
              n_other: Optional[int] = None, dirflag: Optional[str] = None,
              relaflag: Optional[str] = None) -> Item:
    rng = np.random.default_rng(seed)
    image = make_image(rng, h, w)
    masks = make_masks(rng, n_masks, h, w)
    boxes = masks_to_boxes(masks)
    target = make_target(rng, masks)
    feats = None
    if with_features:
        feats = rng.standard_normal((n_masks, de)).astype(np.float32)
        if bf16:
            feats = bf16_round(feats)
    exprs = []
    for _ in range(n_expr):
        anchor = None
        if feats is not None:
            anchor = feats[int(rng.integers(0, n_masks))] * 0.25
        exprs.append(make_expression(rng, h, w, de, n_other=n_other, dirflag=dirflag,
                                     relaflag=relaflag, bf16=bf16, anchor=anchor))
    return Item(image, masks, boxes, target, exprs, feats)


# BASELINE.json configs -> concrete shapes
CONFIGS = {
    1: dict(h=480, w=640, n_masks=64, n_expr=1, S=224, g=14, Dv=768, De=512, heads=12, layers=12, fusion_mode="G2L"),
    2: dict(h=480, w=640, n_masks=100, n_expr=3, S=224, g=14, Dv=768, De=512, heads=12, layers=12, fusion_mode="G2L&L2G"),
    3: dict(h=480, w=640, n_masks=150, n_expr=2, S=224, g=14, Dv=768, De=512, heads=12, layers=12, fusion_mode="L2G", n_other=3),
    4: dict(h=480, w=640, n_masks=200, n_expr=3, S=336, g=24, Dv=1024, De=768, heads=16, layers=24, fusion_mode="G2L&L2G"),
    5: dict(h=600, w=800, n_masks=200, n_expr=5, S=224, g=14, Dv=768, De=512, heads=12, layers=12, fusion_mode="G2L"),
}


# --------------------------------------------------------------------------------------------------
# device-side generator for benchmark-sized batches (same shapes / statistics, torch RNG on the GPU)
# --------------------------------------------------------------------------------------------------
def make_batch_device(seed: int, n_images: int, h: int, w: int, n_masks: int, n_expr: int, de: int,
                      device="cuda", n_other: int = 2, pinned_host: bool = False, grid: int = 14, raw_heat: bool = False):
    """RefCOCO-shaped batch laid out for the batched C ABI (ragged offsets, here uniform).
    Returns a dict of tensors on `device` (or pinned host tensors when pinned_host=True)."""
    import torch
    gen = torch.Generator(device=device).manual_seed(seed)
    B, N, E = n_images, n_masks, n_expr
    M, ET = B * N, B * E

    def rand(*shape):
        return torch.rand(*shape, generator=gen, device=device)

    coarse = rand(B, 3, 14, 14)
    base = torch.nn.functional.interpolate(coarse, size=(h, w), mode="bilinear", align_corners=False)
    img = (0.75 * base + 0.25 * rand(B, 3, h, w)).mul(255).add(0.5).clamp(0, 255).to(torch.uint8)
    image = img.permute(0, 2, 3, 1).contiguous()                                       # [B,H,W,3] u8
    lo, hi = float(np.log(800.0)), float(np.log(0.6 * h * w))
    area = torch.exp(lo + (hi - lo) * rand(M))
    ratio = torch.exp(-0.9 + 1.8 * rand(M))
    ra = torch.sqrt(area / np.pi * ratio); rb = torch.sqrt(area / np.pi / ratio)
    cy = (0.1 + 0.8 * rand(M)) * h; cx = (0.1 + 0.8 * rand(M)) * w
    th = rand(M) * np.pi
    yy = torch.arange(h, device=device, dtype=torch.float32)[None, :, None]
    xx = torch.arange(w, device=device, dtype=torch.float32)[None, None, :]
    masks = torch.empty((M, h, w), dtype=torch.bool, device=device)
    for s in range(0, M, 64):                                                           # chunked: keeps temporaries small
        sl = slice(s, min(M, s + 64))
        dx = xx - cx[sl, None, None]; dy = yy - cy[sl, None, None]
        c, sn = torch.cos(th[sl])[:, None, None], torch.sin(th[sl])[:, None, None]
        u = dx * c + dy * sn; v = -dx * sn + dy * c
        masks[sl] = (u / ra[sl, None, None]) ** 2 + (v / rb[sl, None, None]) ** 2 <= 1.0
    masks[:, h // 2, w // 2] = True                                                      # never empty
    ys = masks.any(dim=2); xs = masks.any(dim=1)
    y0 = ys.float().argmax(1); y1 = h - 1 - ys.flip(1).float().argmax(1)
    x0 = xs.float().argmax(1); x1 = w - 1 - xs.flip(1).float().argmax(1)
    boxes = torch.stack([x0, y0, x1 - x0, y1 - y0], dim=1).to(torch.int64)              # SAM XYWH convention
    pick = torch.randint(0, N, (B,), generator=gen, device=device) + torch.arange(B, device=device) * N
    target = torch.roll(masks[pick], shifts=(3, -4), dims=(1, 2)).to(torch.uint8)       # [B,H,W]
    feats = torch.randn((M, de), generator=gen, device=device).to(torch.bfloat16)
    tokens = torch.randn((B, grid * grid, de), generator=gen, device=device).to(torch.bfloat16)     # dense patch tokens in embedding space
    anchor = feats[torch.randint(0, N, (ET,), generator=gen, device=device)
                   + torch.arange(B, device=device).repeat_interleave(E) * N].float() * 0.25
    sent = (torch.randn((ET, de), generator=gen, device=device) + 1.5 * anchor).to(torch.bfloat16).float()
    noun = (0.6 * sent + 0.8 * torch.randn((ET, de), generator=gen, device=device) + 1.5 * anchor).to(torch.bfloat16).float()
    others = torch.randn((ET * n_other, de), generator=gen, device=device).to(torch.bfloat16).float()
    other_off = (torch.arange(ET + 1, device=device) * n_other).to(torch.int32)
    # GEM map: 448-short-side input, patch 16 -> 28 x 37.  raw_heat: hand the path the map as the model returns it
    # (the T.Resize((H,W), antialias=True) of Hybridgl_main.py:201 then runs on the device); else a frame-sized map.
    heat = rand(ET, 1, 28, 37)
    heat = (heat[:, 0].contiguous() if raw_heat else
            torch.nn.functional.interpolate(heat, size=(h, w), mode="bilinear", align_corners=False)[:, 0].contiguous())
    dirflag = torch.randint(0, 6, (ET,), generator=gen, device=device).to(torch.int32)
    relaflag = torch.randint(0, 8, (ET,), generator=gen, device=device).to(torch.int32)
    black = torch.where(relaflag == 5, 1.95, torch.where(relaflag == 6, 1.5, 1.8)).to(torch.float32)
    out = dict(image=image, masks=masks, boxes=boxes, target=target, features=feats, tokens=tokens, sent=sent, noun=noun, others=others,
               other_off=other_off, heat=heat, dirflag=dirflag, relaflag=relaflag, black=black,
               mask_off=(torch.arange(B + 1, device=device) * N).to(torch.int32),
               expr_off=(torch.arange(B + 1, device=device) * E).to(torch.int32))
    if pinned_host:
        out = {k: v.cpu().pin_memory() for k, v in out.items()}
    return out


# --------------------------------------------------------------------------------------------------
# seeded CLIP weights (no checkpoints offline): the same numbers on the reference side (tests/golden/gen_golden.py)
# and on this side (tests), derived from (seed, parameter name) only -- so fixtures store outputs, not weights
# --------------------------------------------------------------------------------------------------
def seeded_clip_state_dict(shapes, seed: int):
    """shapes: {parameter name: shape} (CLIP state_dict names).  Returns {name: float32 ndarray}.
    LayerNorm weights ~ 1 + 0.05 N(0,1), biases / 1-D tensors ~ 0.02 N(0,1), embeddings ~ 0.02 N(0,1), matrices ~ N(0,1) / sqrt(fan_in)
    (activations stay O(1) through 24 blocks), logit_scale = log(1 / 0.07)."""
    import zlib
    out = {}
    for name in sorted(shapes):
        shape = tuple(int(v) for v in shapes[name])
        rng = np.random.default_rng([int(seed), zlib.crc32(name.encode())])
        if name.endswith("logit_scale"):
            out[name] = np.array(np.log(1 / 0.07), np.float32).reshape(shape)
            continue
        x = rng.standard_normal(shape, dtype=np.float32)
        if len(shape) <= 1:
            is_ln_weight = name.endswith(".weight") and (".ln_" in name or name.startswith(("ln_", "visual.ln_")) or "ln_final" in name)
            x = (1.0 + 0.05 * x) if is_ln_weight else 0.02 * x
        elif "embedding" in name:
            x = 0.02 * x
        elif name.endswith("proj") or name.endswith("text_projection"):      # [width, out]: applied as x @ proj
            x = x / np.sqrt(shape[0])
        else:                                                                  # nn.Linear / conv weights: [out, in, ...]
            x = x / np.sqrt(float(np.prod(shape[1:])))
        out[name] = x.astype(np.float32)
    return out


def seeded_images(seed: int, n: int, size: int):
    """Two seeded CLIP input stacks (local, global) f32 [n,3,size,size] ~ N(0,1): stand-ins for the prep outputs."""
    rng = np.random.default_rng(seed)
    return rng.standard_normal((n, 3, size, size), dtype=np.float32), rng.standard_normal((n, 3, size, size), dtype=np.float32)
