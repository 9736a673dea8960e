"""Host-side mirror of the three path functions of the reference's utils.py, same names / arguments / returns:

    relation_boxes(boxi, boxj, scorei, scorej, relaword)                utils.py:240-268
    gen_dir_mask(dirflag, height, width, device)                        utils.py:135-161
    Compute_IoU(pred, target, cum_I, cum_U, mean_IoU=[])                utils.py:365-384

They exist so that Hybridgl_main.py can switch imports without touching its call sites.  The batched path
(pipeline.ScoringPath) does not use them: it evaluates the same formulas inside hgl_score_select / hgl_heat_pool /
hgl_iou for all expressions at once.  gen_dir_mask and Compute_IoU run on the GPU through libhgl; relation_boxes is
a handful of scalar comparisons on host-visible values (it is host code in the reference too).
"""
from __future__ import annotations

import torch

from . import ops

_REL = ("none", "left", "right", "up", "down", "big", "small", "within")


def relation_boxes(boxi, boxj, scorei, scorej, relaword):
    """Pairwise spatial-relationship score of two XYWH boxes (centres are x+w/2, y+h/2).  'none' and unknown words
    return scorei unchanged; every other word gates scorei*scorej by a comparison, 'within' scales it by the
    overlap area over the area of box i."""
    if relaword not in _REL or relaword == "none":
        return scorei
    xi, yi, wi, hi = boxi[0], boxi[1], boxi[2], boxi[3]
    xj, yj, wj, hj = boxj[0], boxj[1], boxj[2], boxj[3]
    pair = scorei * scorej
    if relaword in ("left", "right"):
        ci, cj = xi + wi / 2, xj + wj / 2
        return pair * ((ci < cj) if relaword == "left" else (ci > cj))
    if relaword in ("up", "down"):
        ci, cj = yi + hi / 2, yj + hj / 2
        return pair * ((ci < cj) if relaword == "up" else (ci > cj))
    if relaword in ("big", "small"):
        ai, aj = wi * hi, wj * hj
        return pair * ((ai > aj) if relaword == "big" else (ai < aj))
    left = max(xi, xj)
    right = max(left, min(xi + wi, xj + wj))
    top = max(yi, yj)
    bottom = max(top, min(yi + hi, yj + hj))
    return pair * (right - left) * (bottom - top) / (wi * hi)


def gen_dir_mask(dirflag, height, width, device):
    """Horizontal position ramp [H,W] (left 1->0, right 0->1, middle 0->1->0; up/down/none are all-ones exactly as in the
    reference, whose vertical ramps are commented out).  `device` falsy -> current CUDA device."""
    return ops.dir_mask(dirflag, int(height), int(width), device if device else "cuda")


def Compute_IoU(pred, target, cum_I, cum_U, mean_IoU=None):
    """I = |pred & target|, U = |pred | target| as exact integers (hgl_iou), this_iou = I/U (0.0 when U == 0).
    Returns (this_iou, mean_IoU, cum_I, cum_U) like the reference; cum_I / cum_U may be ints or tensors."""
    if mean_IoU is None:
        mean_IoU = []
    if target.dim() == 3:
        target = target[0]
    p = pred.to(torch.bool) if pred.dtype not in (torch.bool, torch.uint8) else pred
    t = target.to(torch.bool) if target.dtype not in (torch.bool, torch.uint8) else target
    zero = torch.zeros(1, dtype=torch.int64, device=p.device)
    iu = ops.iou_accumulate(p.contiguous()[None], t.contiguous(), zero, zero, None)
    I, U = iu[0, 0], iu[0, 1]
    this_iou = 0.0 if int(U) == 0 else I * 1.0 / U          # the reference's `if U == 0` is the same host sync
    cum_I = cum_I + I
    cum_U = cum_U + U
    mean_IoU.append(this_iou)
    return this_iou, mean_IoU, cum_I, cum_U


def mask2chw(arr):
    """utils.py:280-289 for a CUDA mask [H,W] (bool / uint8): ((center_y, center_x), height, width), computed by hgl_mask_geometry.
    Raises ValueError on an empty mask (the reference's np.mean of nothing is NaN and int(NaN) raises there too)."""
    m = arr if arr.dtype in (torch.bool, torch.uint8) else (arr == 1)
    cy, cx, h, w = (int(v) for v in ops.mask_geometry(m.contiguous()[None], want_boxes=False, want_chw=True)[0].tolist())
    if h == 0:
        raise ValueError("mask2chw: empty mask")
    return (cy, cx), h, w


def apply_visual_prompts(image_array, mask, visual_prompt_type=("circle",), color=(255, 0, 0), thickness=1, blur_strength=(15, 15)):
    """utils.py:292-345 on CUDA tensors, the prompt types applied in the reference's order: 'blur' (sharp inside the mask,
    cv2.GaussianBlur(15,15) outside), 'circle' (cv2.ellipse outline at mask2chw's centre with half axes (width // 2, height // 2):
    hgl_mask_geometry + hgl_ellipse_outline, OpenCV's pixels) and 'black' (zeros outside).  image_array u8 [H,W,3], mask bool/u8
    [H,W]; returns u8 [H,W,3].  Same defaults as the reference (visual_prompt_type=('circle',), red, thickness 1)."""
    if tuple(blur_strength) != (15, 15):
        raise ValueError("only the (15, 15) kernel of the reference's call is built (hgl_gaussian_blur15)")
    if thickness != 1:
        raise ValueError("only thickness 1 (the reference's default and only use) is built")
    img = image_array
    keep = (mask != 0)[:, :, None]
    for kind in ("blur", "circle", "black"):                # the order of the reference's if-chain
        if kind not in visual_prompt_type:
            continue
        if kind == "circle":
            chw = ops.mask_geometry((mask != 0)[None].contiguous(), want_boxes=False, want_chw=True)
            img = ops.ellipse_outline(img.contiguous().clone()[None], chw, color)[0]
            continue
        bg = ops.gaussian_blur15(img.contiguous()) if kind == "blur" else torch.zeros_like(img)
        img = torch.where(keep, img, bg)
    return img
