"""Thin tensor-level wrappers over the C ABI (include/hgl.h).

Every function takes CUDA tensors, checks dtype / contiguity / device, passes raw pointers and the
current CUDA stream to libhgl.so, and returns freshly allocated (or caller-supplied) CUDA tensors.
PyTorch is used for device memory and streams only -- no arithmetic of the path happens in torch.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import DIR_CODES, HGL_BF16, HGL_BG_BLACK, HGL_BG_BLUR, HGL_BG_NONE, HGL_F32, REL_CODES, check  # noqa: F401


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _req(t: torch.Tensor, dtype, name: str, ndim: Optional[int] = None) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise TypeError(f"{name}: expected a CUDA tensor")
    if dtype is not None and t.dtype not in (dtype if isinstance(dtype, tuple) else (dtype,)):
        raise TypeError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if ndim is not None and t.dim() != ndim:
        raise ValueError(f"{name}: expected {ndim} dims, got {tuple(t.shape)}")
    if not t.is_contiguous():
        raise ValueError(f"{name}: must be contiguous")
    return t


def _mask_bytes(masks: torch.Tensor, name: str = "masks") -> torch.Tensor:
    _req(masks, (torch.bool, torch.uint8), name, 3)
    return masks.view(torch.uint8) if masks.dtype == torch.bool else masks


def _dt(dtype: torch.dtype) -> int:
    if dtype == torch.float32:
        return HGL_F32
    if dtype == torch.bfloat16:
        return HGL_BF16
    raise TypeError(f"unsupported dtype {dtype} (float32 or bfloat16)")


def _offsets(off: Optional[torch.Tensor], B: int, name: str) -> Optional[torch.Tensor]:
    if off is None:
        if B != 1:
            raise ValueError(f"{name} is required when the batch holds more than one image")
        return None
    _req(off, torch.int32, name, 1)
    if off.numel() != B + 1:
        raise ValueError(f"{name}: expected {B + 1} entries, got {off.numel()}")
    return off


def device_ok() -> None:
    check(_lib.load().hgl_check_device(), "hgl_check_device")


# ---- (a1) ---------------------------------------------------------------------------------------------
def gaussian_blur15(image: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """cv2.GaussianBlur(img,(15,15),0) (Hybridgl_main.py:99).  image u8 [B,H,W,3] or [H,W,3]."""
    squeeze = image.dim() == 3
    img = image[None] if squeeze else image
    _req(img, torch.uint8, "image", 4)
    B, H, W, C = img.shape
    if C != 3:
        raise ValueError("image must be HWC with 3 channels")
    if out is None:
        out = torch.empty_like(img)
    check(_lib.load().hgl_gaussian_blur15(img.data_ptr(), out.data_ptr(), B, H, W, _stream()), "hgl_gaussian_blur15")
    return out[0] if squeeze else out


def pack_masks(masks: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """bool/u8 [M,H,W] -> packed int32 [M,H,ceil(W/32)] (bit i of word w = pixel 32*w+i).  The one pass over the byte masks."""
    m = _mask_bytes(masks)
    M, H, W = m.shape
    if out is None:
        out = torch.empty((M, H, (W + 31) // 32), dtype=torch.int32, device=m.device)
    else:
        _req(out, torch.int32, "bits", 3)
        if tuple(out.shape) != (M, H, (W + 31) // 32):
            raise ValueError(f"bits must be [M,H,ceil(W/32)] int32, got {tuple(out.shape)}")
    lib = _lib.load()
    # torch.bool storage is one byte of 0 / 1 per element: the cheaper squeeze applies; uint8 masks may hold any non-zero value
    fn = lib.hgl_pack_masks_bool if masks.dtype == torch.bool else lib.hgl_pack_masks
    check(fn(m.data_ptr(), M, H, W, out.data_ptr(), _stream()), "hgl_pack_masks")
    return out


def rle_to_bits(counts: torch.Tensor, rle_off: torch.Tensor, height: int, width: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """SAM uncompressed RLE (amg.py:107-149; column-major runs, first run counts zeros) -> packed int32 [M,H,ceil(W/32)],
    the same tensor pack_masks() gives for np.stack([rle_to_mask(r) for r in rles]).
    counts int32 [R]: the `counts` lists of all proposals back to back; rle_off int32 [M+1]: their boundaries."""
    _req(counts, torch.int32, "counts", 1)
    _req(rle_off, torch.int32, "rle_off", 1)
    M = rle_off.numel() - 1
    if M < 0:
        raise ValueError("rle_off must hold M+1 entries")
    H, W = int(height), int(width)
    if out is None:
        out = torch.empty((M, H, (W + 31) // 32), dtype=torch.int32, device=counts.device)
    else:
        _req(out, torch.int32, "bits", 3)
        if tuple(out.shape) != (M, H, (W + 31) // 32):
            raise ValueError(f"bits must be [M,H,ceil(W/32)] int32, got {tuple(out.shape)}")
    check(_lib.load().hgl_rle_to_bits(counts.data_ptr(), rle_off.data_ptr(), M, H, W, out.data_ptr(), _stream()), "hgl_rle_to_bits")
    return out


def mask_geometry(masks: torch.Tensor, width: Optional[int] = None, want_boxes: bool = True, want_chw: bool = False):
    """SAM's XYWH proposal boxes (amg.py:303-346 + :91-95) and / or mask2chw (utils.py:280-289) of every mask, on the device.
    `masks`: bool/u8 [M,H,W] (packed internally) or packed int32 [M,H,ceil(W/32)] together with `width`.
    Returns boxes int64 [M,4] (x, y, w, h), chw int32 [M,4] (center_y, center_x, height, width), or the pair."""
    if masks.dtype == torch.int32:
        if width is None:
            raise ValueError("width is required with packed masks")
        bits = _req(masks, torch.int32, "bits", 3)
        W = int(width)
        if bits.shape[2] != (W + 31) // 32:
            raise ValueError("packed masks do not match width")
    else:
        m = _mask_bytes(masks)
        W = m.shape[2]
        bits = pack_masks(m)
    M, H = bits.shape[0], bits.shape[1]
    boxes = torch.empty((M, 4), dtype=torch.int64, device=bits.device) if want_boxes else None
    chw = torch.empty((M, 4), dtype=torch.int32, device=bits.device) if want_chw else None
    check(_lib.load().hgl_mask_geometry(bits.data_ptr(), M, H, W, _ptr(boxes), _ptr(chw), _stream()), "hgl_mask_geometry")
    return (boxes, chw) if (want_boxes and want_chw) else (boxes if want_boxes else chw)


def _bits(masks_or_bits: torch.Tensor, W: Optional[int] = None) -> torch.Tensor:
    """Accept either byte masks (packed on the fly) or an already packed int32 tensor."""
    if masks_or_bits.dtype == torch.int32:
        return _req(masks_or_bits, torch.int32, "bits", 3)
    return pack_masks(masks_or_bits)


def _prep_frames(image: torch.Tensor, blur: Optional[torch.Tensor], background: str):
    img = image[None] if image.dim() == 3 else image
    _req(img, torch.uint8, "image", 4)
    bg = {"blur": HGL_BG_BLUR, "black": HGL_BG_BLACK, "none": HGL_BG_NONE}[background]
    bl = None
    if bg == HGL_BG_BLUR:
        if blur is None:
            blur = gaussian_blur15(img)
        bl = blur[None] if blur.dim() == 3 else blur
        _req(bl, torch.uint8, "blur", 4)
        if bl.shape != img.shape:
            raise ValueError("blur must have the shape of image")
    return img, bl, bg


def _prep_workspace(B: int, size: int, dtype: torch.dtype, device, workspace: Optional[torch.Tensor]) -> torch.Tensor:
    need = _lib.load().hgl_prep_workspace_bytes(B, size, _dt(dtype))
    if need < 0:
        raise ValueError(f"prep: unsupported size {size}")
    if workspace is None or workspace.numel() * workspace.element_size() < need:
        workspace = torch.empty((max(need, 1),), dtype=torch.uint8, device=device)
    return workspace


def prep_setup(image: torch.Tensor, blur: Optional[torch.Tensor], size: int, background: str = "blur",
               dtype: torch.dtype = torch.float32, workspace: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Per-image half of the prep loop (hgl_prep_setup): needs the frames only, so a caller can enqueue it while the masks are
    still being packed on another stream.  Returns the workspace prep_main() consumes."""
    img, bl, bg = _prep_frames(image, blur, background)
    B, H, W, _ = img.shape
    workspace = _prep_workspace(B, size, dtype, img.device, workspace)
    check(_lib.load().hgl_prep_setup(img.data_ptr(), _ptr(bl), B, H, W, size, bg, _dt(dtype), workspace.data_ptr(), _stream()),
          "hgl_prep_setup")
    return workspace


def prep_main(masks: torch.Tensor, frame_shape, size: int, workspace: torch.Tensor, mask_off: Optional[torch.Tensor] = None,
              max_n: Optional[int] = None, dtype: torch.dtype = torch.float32,
              out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Per-mask half of the prep loop (hgl_prep_main) over the workspace prep_setup() filled.  frame_shape = (B, H, W)."""
    B, H, W = (int(v) for v in frame_shape)
    bits = _bits(masks)
    M = bits.shape[0]
    if tuple(bits.shape[1:]) != (H, (W + 31) // 32):
        raise ValueError(f"masks {tuple(masks.shape)} do not match the frame {H}x{W}")
    off = _offsets(mask_off, B, "mask_off")
    if max_n is None:
        if B != 1:
            raise ValueError("max_n is required for batched calls")
        max_n = max(M, 1)
    if out is None:
        local = torch.empty((M, 3, size, size), dtype=dtype, device=bits.device)
        glob = torch.empty_like(local)
    else:
        local, glob = out
    _req(workspace, torch.uint8, "workspace", 1)
    if workspace.numel() < _lib.load().hgl_prep_workspace_bytes(B, size, _dt(dtype)):
        raise ValueError("prep_main: workspace smaller than hgl_prep_workspace_bytes()")
    check(_lib.load().hgl_prep_main(bits.data_ptr(), _ptr(off), B, M, max_n, H, W, size, _dt(dtype), local.data_ptr(), glob.data_ptr(),
                                    workspace.data_ptr(), _stream()), "hgl_prep_main")
    return local, glob


def prep_visual_prompts(image: torch.Tensor, blur: Optional[torch.Tensor], masks: torch.Tensor, size: int,
                        mask_off: Optional[torch.Tensor] = None, max_n: Optional[int] = None, background: str = "blur",
                        dtype: torch.dtype = torch.float32,
                        out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
                        workspace: Optional[torch.Tensor] = None, crop_xywh: Optional[torch.Tensor] = None,
                        circle: bool = False, circle_color=(255, 0, 0)) -> Tuple[torch.Tensor, torch.Tensor]:
    """The prep loop Hybridgl_main.py:92-125 for a whole batch.  Returns (local_imgs, global_imgs) [M,3,S,S].
    `masks` is bool/u8 [M,H,W] (packed internally) or the packed int32 [M,H,ceil(W/32)] from pack_masks().
    background: what replaces the frame outside the mask in the global view -- "blur" | "black" | "none" (the frame itself).
    crop_xywh int32/int64 [M,4]: resample every proposal from its own box (x, y, w, h) instead of the full frame (hgl_prep_crop;
    the reference always uses the full frame).
    circle: the global view also carries the 'circle' prompt of utils.apply_visual_prompts (cv2.ellipse at mask2chw's centre,
    utils.py:322-335), applied in the reference's order blur -> circle -> black (hgl_mask_geometry + hgl_prep_circle)."""
    if circle and crop_xywh is not None:
        raise ValueError("circle and crop_xywh cannot be combined")
    img, bl, bg = _prep_frames(image, blur, background)
    B, H, W, _ = img.shape
    bits = _bits(masks)
    M = bits.shape[0]
    if tuple(bits.shape[1:]) != (H, (W + 31) // 32):
        raise ValueError(f"masks {tuple(masks.shape)} do not match the frame {H}x{W}")
    off = _offsets(mask_off, B, "mask_off")
    if max_n is None:
        if B != 1:
            raise ValueError("max_n is required for batched calls")
        max_n = max(M, 1)
    if out is None:
        local = torch.empty((M, 3, size, size), dtype=dtype, device=img.device)
        glob = torch.empty_like(local)
    else:
        local, glob = out
    if crop_xywh is not None:
        _req(crop_xywh, (torch.int32, torch.int64), "crop_xywh", 2)
        if tuple(crop_xywh.shape) != (M, 4):
            raise ValueError("crop_xywh must be [M,4]")
        c = crop_xywh.to(torch.int32).contiguous()
        check(_lib.load().hgl_prep_crop(img.data_ptr(), _ptr(bl), bits.data_ptr(), _ptr(off), c.data_ptr(), B, M, H, W, size, bg, _dt(dtype),
                                        local.data_ptr(), glob.data_ptr(), _stream()), "hgl_prep_crop")
        return local, glob
    workspace = _prep_workspace(B, size, dtype, img.device, workspace)
    check(_lib.load().hgl_prep(img.data_ptr(), _ptr(bl), bits.data_ptr(), _ptr(off), B, M, max_n, H, W, size, bg, _dt(dtype),
                               local.data_ptr(), glob.data_ptr(), workspace.data_ptr(), _stream()), "hgl_prep")
    if circle and M > 0:
        chw = mask_geometry(bits, width=W, want_boxes=False, want_chw=True)
        r, g, b = (int(v) for v in circle_color)
        check(_lib.load().hgl_prep_circle(img.data_ptr(), _ptr(bl), bits.data_ptr(), _ptr(off), chw.data_ptr(), B, M, H, W, size, bg,
                                          _dt(dtype), r, g, b, glob.data_ptr(), _stream()), "hgl_prep_circle")
    return local, glob


def ellipse_outline(images: torch.Tensor, chw: torch.Tensor, color=(255, 0, 0)) -> torch.Tensor:
    """cv2.ellipse(image_m, (cx, cy), (w // 2, h // 2), 0, 0, 360, color, 1) drawn IN PLACE into images u8 [M,H,W,3] for
    chw int32 [M,4] = (center_y, center_x, height, width) per image (mask_geometry(..., want_chw=True)); utils.py:322-335."""
    _req(images, torch.uint8, "images", 4)
    _req(chw, torch.int32, "chw", 2)
    M, H, W, C = images.shape
    if C != 3 or tuple(chw.shape) != (M, 4):
        raise ValueError("images must be [M,H,W,3] and chw [M,4]")
    r, g, b = (int(v) for v in color)
    check(_lib.load().hgl_ellipse_outline(images.data_ptr(), chw.data_ptr(), M, H, W, r, g, b, _stream()), "hgl_ellipse_outline")
    return images


# ---- (a2)-(a4) ----------------------------------------------------------------------------------------
def masks_to_grid(masks: torch.Tensor, g: int, antialias: bool = True, want_area: bool = False, width: Optional[int] = None, out=None):
    """TF.resize(pred_masks.float(), (g, g)) model/backbone.py:160 -> f32 [M,g,g] (and int32 areas [M]).
    `masks`: bool/u8 [M,H,W] (packed internally) or packed int32 [M,H,ceil(W/32)] together with `width`."""
    if masks.dtype == torch.int32:
        if width is None:
            raise ValueError("width is required with packed masks")
        bits = _req(masks, torch.int32, "bits", 3)
        M, H, W = bits.shape[0], bits.shape[1], int(width)
        if bits.shape[2] != (W + 31) // 32:
            raise ValueError("packed masks do not match width")
    else:
        m = _mask_bytes(masks)
        M, H, W = m.shape
        bits = pack_masks(m)
    if out is not None:                      # caller-supplied (grid [M,g,g] f32, area [M] i32 or None)
        grid, area = out
    else:
        grid = torch.empty((M, g, g), dtype=torch.float32, device=bits.device)
        area = torch.empty((M,), dtype=torch.int32, device=bits.device) if want_area else None
    ws = None
    if antialias:
        ws = torch.empty((max(_lib.load().hgl_mask_grid_workspace_bytes(M, g), 1),), dtype=torch.uint8, device=bits.device)
    check(_lib.load().hgl_mask_grid(bits.data_ptr(), M, H, W, g, int(bool(antialias)), grid.data_ptr(), _ptr(area), _ptr(ws),
                                    _stream()), "hgl_mask_grid")
    return (grid, area) if want_area else grid


def make_attn_mask(grid: torch.Tensor, heads: int) -> torch.Tensor:
    """CLIPViTFM.make_attn_mask model/backbone.py:108-115 -> bool [M*heads, L+1, L+1] (True = blocked)."""
    _req(grid, torch.float32, "grid")
    M = grid.shape[0]
    L = grid[0].numel()
    out = torch.empty((M * heads, L + 1, L + 1), dtype=torch.uint8, device=grid.device)
    check(_lib.load().hgl_attn_mask(grid.data_ptr(), M, L, heads, out.data_ptr(), _stream()), "hgl_attn_mask")
    return out.view(torch.bool)


def attn_key_bias(grid: torch.Tensor) -> torch.Tensor:
    """Compact form of the same mask: additive bias [M, L+1] for the CLS query row (0 or -inf)."""
    _req(grid, torch.float32, "grid")
    M = grid.shape[0]
    L = grid[0].numel()
    out = torch.empty((M, L + 1), dtype=torch.float32, device=grid.device)
    check(_lib.load().hgl_attn_bias(grid.data_ptr(), M, L, out.data_ptr(), _stream()), "hgl_attn_bias")
    return out


def token_mask_fuse(src: torch.Tensor, add: Optional[torch.Tensor], grid: Optional[torch.Tensor], a: float = 1.0,
                    b: float = 1.0, out: Optional[torch.Tensor] = None, layout: str = "LND") -> torch.Tensor:
    """out = a * tokenmask(src, grid) + b * add (model/backbone.py:235-249, 216, 290-291) on token streams laid out
    [L+1, M, D] (layout="LND", the reference's) or [M, L+1, D] (layout="NLD", the B200 forward's)."""
    _req(src, (torch.float32, torch.bfloat16), "src", 3)
    lay = {"LND": _lib.HGL_LND, "NLD": _lib.HGL_NLD}[layout]
    if lay == _lib.HGL_LND:
        L1, M, D = src.shape
    else:
        M, L1, D = src.shape
    if add is not None:
        _req(add, src.dtype, "add", 3)
        if add.shape != src.shape:
            raise ValueError("add must have the shape of src")
    if grid is not None:
        _req(grid, torch.float32, "grid")
        if grid.shape[0] != M or grid[0].numel() != L1 - 1:
            raise ValueError(f"grid {tuple(grid.shape)} does not match streams {tuple(src.shape)}")
    if out is None:
        out = torch.empty_like(src)
    check(_lib.load().hgl_token_mask_fuse(src.data_ptr(), _ptr(add), _ptr(grid), float(a), float(b), L1, M, D, _dt(src.dtype),
                                          lay, out.data_ptr(), _stream()), "hgl_token_mask_fuse")
    return out


def token_mask_fuse_ln(src: torch.Tensor, add: Optional[torch.Tensor], grid: Optional[torch.Tensor], a: float, b: float,
                       gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5, out_x: Optional[torch.Tensor] = None,
                       out_ln: Optional[torch.Tensor] = None, want_x: bool = True):
    """token_mask_fuse (NLD layout) + the LayerNorm that follows it, one pass (hgl_token_mask_fuse_ln).  src / add [M, L+1, D];
    gamma / beta f32 [D].  Returns (out_x, out_ln); out_x is None with want_x=False.  out_x / out_ln may be slices along dim 0 of
    larger tensors (the block's concatenated batch)."""
    _req(src, (torch.float32, torch.bfloat16), "src", 3)
    M, L1, D = src.shape
    if add is not None:
        _req(add, src.dtype, "add", 3)
        if add.shape != src.shape:
            raise ValueError("add must have the shape of src")
    if grid is not None:
        _req(grid, torch.float32, "grid")
        if grid.shape[0] != M or grid[0].numel() != L1 - 1:
            raise ValueError(f"grid {tuple(grid.shape)} does not match streams {tuple(src.shape)}")
    _req(gamma, torch.float32, "gamma", 1)
    _req(beta, torch.float32, "beta", 1)
    if gamma.numel() != D or beta.numel() != D:
        raise ValueError("gamma / beta must be [D]")
    if out_ln is None:
        out_ln = torch.empty_like(src)
    if want_x and out_x is None:
        out_x = torch.empty_like(src)
    for t, name in ((out_ln, "out_ln"), (out_x, "out_x")):
        if t is not None:
            _req(t, src.dtype, name, 3)
            if t.shape != src.shape:
                raise ValueError(f"{name} must have the shape of src")
    check(_lib.load().hgl_token_mask_fuse_ln(src.data_ptr(), _ptr(add), _ptr(grid), float(a), float(b), gamma.data_ptr(), beta.data_ptr(),
                                             float(eps), L1, M, D, _dt(src.dtype), _ptr(out_x if want_x else None), out_ln.data_ptr(),
                                             _stream()), "hgl_token_mask_fuse_ln")
    return (out_x if want_x else None), out_ln


def cls_attention(qkv: torch.Tensor, bias: Optional[torch.Tensor], heads: int) -> torch.Tensor:
    """Attention output of the CLS query alone under the key bias (hgl_cls_attention).  qkv [M, L1, 3*D] (the packed in_proj output,
    f32 / bf16), bias f32 [M, L1] or None.  Returns [M, heads, hd] of qkv's dtype."""
    _req(qkv, (torch.float32, torch.bfloat16), "qkv", 3)
    M, L1, D3 = qkv.shape
    hd = D3 // (3 * heads)
    if hd * 3 * heads != D3:
        raise ValueError("qkv width is not 3 * heads * head_dim")
    if bias is not None:
        _req(bias, torch.float32, "bias", 2)
        if tuple(bias.shape) != (M, L1):
            raise ValueError("bias must be [M, L1]")
    out = torch.empty((M, heads, hd), dtype=qkv.dtype, device=qkv.device)
    check(_lib.load().hgl_cls_attention(qkv.data_ptr(), _ptr(bias), M, L1, heads, hd, _dt(qkv.dtype), out.data_ptr(), _stream()), "hgl_cls_attention")
    return out


def cls_head(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, proj: torch.Tensor, eps: float = 1e-5,
             out: Optional[torch.Tensor] = None, accumulate: bool = False) -> torch.Tensor:
    """ln_post(x[:, 0, :]) @ proj in one launch (hgl_cls_head).  x [M, L1, Dv] (the CLS token of every row is read in place) or
    [M, Dv]; gamma / beta [Dv], proj [Dv, De] (all of one dtype, f32 / bf16).  Returns f32 [M, De]; accumulate adds to `out`."""
    _req(x, (torch.float32, torch.bfloat16), "x")
    if x.dim() == 3:
        M, L1, Dv = x.shape
        stride = L1 * Dv
    else:
        M, Dv = x.shape
        stride = Dv
    for t, nm in ((gamma, "gamma"), (beta, "beta"), (proj, "proj")):
        _req(t, proj.dtype, nm)
    if proj.dtype not in (torch.float32, torch.bfloat16) or proj.shape[0] != Dv or gamma.numel() != Dv or beta.numel() != Dv:
        raise ValueError("cls_head: inconsistent weights")
    De = proj.shape[1]
    if out is None:
        if accumulate:
            raise ValueError("accumulate needs `out`")
        out = torch.empty((M, De), dtype=torch.float32, device=x.device)
    _req(out, torch.float32, "out", 2)
    check(_lib.load().hgl_cls_head(x.data_ptr(), stride, gamma.data_ptr(), beta.data_ptr(), proj.data_ptr(), M, Dv, De, float(eps), _dt(x.dtype),
                                   _dt(proj.dtype), int(bool(accumulate)), out.data_ptr(), _stream()), "hgl_cls_head")
    return out


# ---- (a10)+(a11) --------------------------------------------------------------------------------------
def dir_mask(dirflag: str, height: int, width: int, device=None) -> torch.Tensor:
    """gen_dir_mask utils.py:135-161 -> f32 [H,W] on the GPU."""
    out = torch.empty((height, width), dtype=torch.float32, device=device if device is not None else "cuda")
    if not out.is_cuda:
        raise TypeError("dir_mask: expected a CUDA device (no CPU fallback)")
    check(_lib.load().hgl_dir_mask(DIR_CODES.get(dirflag, 0), height, width, out.data_ptr(), _stream()), "hgl_dir_mask")
    return out


def heat_resize_aa(heat_raw: torch.Tensor, height: int, width: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """T.Resize((H,W), antialias=True)(gem(...)[0]) Hybridgl_main.py:201: raw GEM maps f32 [E,h,w] -> f32 [E,H,W]."""
    _req(heat_raw, torch.float32, "heat_raw", 3)
    E, hh, hw = heat_raw.shape
    if out is None:
        out = torch.empty((E, height, width), dtype=torch.float32, device=heat_raw.device)
    _req(out, torch.float32, "out", 3)
    check(_lib.load().hgl_heat_resize_aa(heat_raw.data_ptr(), E, hh, hw, height, width, out.data_ptr(), _stream()), "hgl_heat_resize_aa")
    return out


def heat_pool(heat: torch.Tensor, dirflag: torch.Tensor, black: torch.Tensor, masks: torch.Tensor,
              mask_off: Optional[torch.Tensor] = None, expr_off: Optional[torch.Tensor] = None, max_n: Optional[int] = None,
              workspace: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Hybridgl_main.py:204-223: conditioned GEM heat-map pooled inside / outside every mask -> score_gem f32 [E,max_n]."""
    _req(heat, torch.float32, "heat", 3)
    E, H, W = heat.shape
    m = _bits(masks)                      # byte masks are packed on the fly; packed int32 passes through
    M = m.shape[0]
    if tuple(m.shape[1:]) != (H, (W + 31) // 32):
        raise ValueError("masks and heat-map frames differ")
    _req(dirflag, torch.int32, "dirflag", 1)
    _req(black, torch.float32, "black", 1)
    B = 1 if mask_off is None else mask_off.numel() - 1
    moff = _offsets(mask_off, B, "mask_off")
    eoff = _offsets(expr_off, B, "expr_off")
    if max_n is None:
        if B != 1:
            raise ValueError("max_n is required for batched calls")
        max_n = max(M, 1)
    lib = _lib.load()
    need = lib.hgl_heat_pool_workspace_bytes(B, M, E, H, W, max_n)
    if workspace is None or workspace.numel() * workspace.element_size() < need:
        workspace = torch.empty((need,), dtype=torch.uint8, device=heat.device)
    out = torch.empty((E, max_n), dtype=torch.float32, device=heat.device)
    check(lib.hgl_heat_pool(heat.data_ptr(), _ptr(eoff), dirflag.data_ptr(), black.data_ptr(), m.data_ptr(), _ptr(moff),
                            B, M, E, H, W, max_n, out.data_ptr(), workspace.data_ptr(), _stream()), "hgl_heat_pool")
    return out


def grid_heat_pool(bits: torch.Tensor, width: int, g: int, heat: torch.Tensor, dirflag: torch.Tensor, black: torch.Tensor,
                   mask_off: Optional[torch.Tensor] = None, expr_off: Optional[torch.Tensor] = None, max_n: Optional[int] = None,
                   workspace: Optional[torch.Tensor] = None):
    """masks_to_grid(antialias=True, want_area=True) and heat_pool in one pass over the packed masks.
    Returns (grid f32 [M,g,g], area int32 [M], score_gem f32 [E,max_n])."""
    _req(bits, torch.int32, "bits", 3)
    _req(heat, torch.float32, "heat", 3)
    M, H, W = bits.shape[0], bits.shape[1], int(width)
    E = heat.shape[0]
    if bits.shape[2] != (W + 31) // 32:
        raise ValueError("packed masks and width disagree")
    raw = tuple(heat.shape[1:]) != (H, W)     # the GEM map before T.Resize((H,W), antialias=True), Hybridgl_main.py:201
    _req(dirflag, torch.int32, "dirflag", 1)
    _req(black, torch.float32, "black", 1)
    B = 1 if mask_off is None else mask_off.numel() - 1
    moff = _offsets(mask_off, B, "mask_off")
    eoff = _offsets(expr_off, B, "expr_off")
    if max_n is None:
        if B != 1:
            raise ValueError("max_n is required for batched calls")
        max_n = max(M, 1)
    lib = _lib.load()
    hh, hw = int(heat.shape[1]), int(heat.shape[2])
    need = (lib.hgl_grid_heat_pool_raw_workspace_bytes(B, M, E, H, W, g, max_n, hh, hw) if raw
            else lib.hgl_grid_heat_pool_workspace_bytes(B, M, E, H, W, g, max_n))
    if workspace is None or workspace.numel() * workspace.element_size() < need:
        workspace = torch.empty((need,), dtype=torch.uint8, device=bits.device)
    grid = torch.empty((M, g, g), dtype=torch.float32, device=bits.device)
    area = torch.empty((M,), dtype=torch.int32, device=bits.device)
    out = torch.empty((E, max_n), dtype=torch.float32, device=bits.device)
    if raw:
        check(lib.hgl_grid_heat_pool_raw(bits.data_ptr(), _ptr(moff), B, M, H, W, g, grid.data_ptr(), area.data_ptr(),
                                         heat.data_ptr(), hh, hw, _ptr(eoff), dirflag.data_ptr(), black.data_ptr(), E, max_n,
                                         out.data_ptr(), workspace.data_ptr(), _stream()), "hgl_grid_heat_pool_raw")
        return grid, area, out
    check(lib.hgl_grid_heat_pool(bits.data_ptr(), _ptr(moff), B, M, H, W, g, grid.data_ptr(), area.data_ptr(),
                                 heat.data_ptr(), _ptr(eoff), dirflag.data_ptr(), black.data_ptr(), E, max_n,
                                 out.data_ptr(), workspace.data_ptr(), _stream()), "hgl_grid_heat_pool")
    return grid, area, out


def gem_token_pool(bits: torch.Tensor, width: int, heat_raw: torch.Tensor, dirflag: torch.Tensor, black: torch.Tensor,
                   mask_off: Optional[torch.Tensor] = None, expr_off: Optional[torch.Tensor] = None, max_n: Optional[int] = None,
                   workspace: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """score_gem f32 [E,max_n] of Hybridgl_main.py:200-223 computed in TOKEN space (hgl_gem_token_pool): the packed masks are
    resampled onto the raw GEM map's grid by the adjoint of the up-sampler and contracted with the raw maps; no frame-sized
    heat-map or table exists.  bits int32 [M,H,ceil(W/32)], heat_raw f32 [E,hh,hw] (the map as the GEM model returns it)."""
    _req(bits, torch.int32, "bits", 3)
    _req(heat_raw, torch.float32, "heat_raw", 3)
    _req(dirflag, torch.int32, "dirflag", 1)
    _req(black, torch.float32, "black", 1)
    M, H, W = bits.shape[0], bits.shape[1], int(width)
    E, hh, hw = heat_raw.shape
    if bits.shape[2] != (W + 31) // 32:
        raise ValueError("packed masks and width disagree")
    B = 1 if mask_off is None else mask_off.numel() - 1
    moff = _offsets(mask_off, B, "mask_off")
    eoff = _offsets(expr_off, B, "expr_off")
    if max_n is None:
        if B != 1:
            raise ValueError("max_n is required for batched calls")
        max_n = max(M, 1)
    lib = _lib.load()
    need = lib.hgl_gem_token_workspace_bytes(B, M, E, H, W, hh, hw, max_n)
    if need < 0:
        raise ValueError("gem_token_pool: unsupported shape")
    if workspace is None or workspace.numel() * workspace.element_size() < need:
        workspace = torch.empty((max(need, 1),), dtype=torch.uint8, device=bits.device)
    if out is None:
        out = torch.empty((E, max_n), dtype=torch.float32, device=bits.device)
    check(lib.hgl_gem_token_pool(bits.data_ptr(), _ptr(moff), B, M, H, W, heat_raw.data_ptr(), hh, hw, _ptr(eoff), dirflag.data_ptr(),
                                 black.data_ptr(), E, max_n, out.data_ptr(), workspace.data_ptr(), _stream()), "hgl_gem_token_pool")
    return out


def heat_tables(heat: torch.Tensor, dirflag: torch.Tensor, H: int, W: int, workspace: torch.Tensor) -> None:
    """First half of grid_heat_pool: the heat-map tables (Hybridgl_main.py:201-209).  Needs the heat-maps only, so a caller can
    enqueue it while the masks are still being packed.  heat f32 [E,H,W], or the raw GEM map [E,hh,hw]."""
    _req(heat, torch.float32, "heat", 3)
    _req(dirflag, torch.int32, "dirflag", 1)
    raw = tuple(heat.shape[1:]) != (int(H), int(W))
    hh, hw = (int(heat.shape[1]), int(heat.shape[2])) if raw else (0, 0)
    check(_lib.load().hgl_heat_tables(heat.data_ptr(), hh, hw, dirflag.data_ptr(), heat.shape[0], int(H), int(W), workspace.data_ptr(),
                                      _stream()), "hgl_heat_tables")


def grid_heat_pool_rows(bits: torch.Tensor, width: int, g: int, heat_shape, black: torch.Tensor, mask_off: Optional[torch.Tensor],
                        expr_off: Optional[torch.Tensor], max_n: int, workspace: torch.Tensor, out=None):
    """Second half of grid_heat_pool: the pass over the packed masks, after heat_tables(...) on the same workspace.
    heat_shape = heat.shape of the tensor given to heat_tables.  Returns (grid, area, score_gem) like grid_heat_pool."""
    _req(bits, torch.int32, "bits", 3)
    _req(black, torch.float32, "black", 1)
    M, H, W = bits.shape[0], bits.shape[1], int(width)
    E = int(heat_shape[0])
    raw = tuple(heat_shape[1:]) != (H, W)
    hh, hw = (int(heat_shape[1]), int(heat_shape[2])) if raw else (0, 0)
    B = 1 if mask_off is None else mask_off.numel() - 1
    moff = _offsets(mask_off, B, "mask_off")
    eoff = _offsets(expr_off, B, "expr_off")
    if out is not None:                      # caller-supplied (grid [M,g,g] f32, area [M] i32, score_gem [E,max_n] f32)
        grid, area, out = out
    else:
        grid = torch.empty((M, g, g), dtype=torch.float32, device=bits.device)
        area = torch.empty((M,), dtype=torch.int32, device=bits.device)
        out = torch.empty((E, max_n), dtype=torch.float32, device=bits.device)
    check(_lib.load().hgl_grid_heat_pool_rows(bits.data_ptr(), _ptr(moff), B, M, H, W, g, grid.data_ptr(), area.data_ptr(), hh, hw,
                                              _ptr(eoff), black.data_ptr(), E, max_n, out.data_ptr(), workspace.data_ptr(), _stream()),
          "hgl_grid_heat_pool_rows")
    return grid, area, out


# ---- (b3') --------------------------------------------------------------------------------------------
def mask_pool(weights: torch.Tensor, tokens: torch.Tensor, mask_off: Optional[torch.Tensor] = None, max_n: Optional[int] = None,
              normalize: bool = True, dtype: torch.dtype = torch.float32, workspace: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Token-space mask pooling on the tensor cores: out[n] = normalize(sum_l weights[n,l] * tokens[b(n)][l,:]).
    weights f32 [M,L] or [M,g,g]; tokens bf16 [B,L,D] (or [L,D] for one image).  Returns [M,D] of `dtype`."""
    _req(weights, torch.float32, "weights")
    w = weights.reshape(weights.shape[0], -1)
    tok = tokens[None] if tokens.dim() == 2 else tokens
    _req(tok, torch.bfloat16, "tokens", 3)
    M, L = w.shape
    B, Lt, D = tok.shape
    if Lt != L:
        raise ValueError(f"weights have {L} cells per mask, tokens {Lt} per image")
    moff = _offsets(mask_off, B, "mask_off")
    if max_n is None:
        if B != 1:
            raise ValueError("max_n is required for batched calls")
        max_n = max(M, 1)
    lib = _lib.load()
    out = torch.empty((M, D), dtype=dtype, device=w.device)
    check(lib.hgl_mask_pool(w.data_ptr(), tok.data_ptr(), _ptr(moff), B, M, max_n, L, D, int(bool(normalize)), _dt(dtype),
                            out.data_ptr(), _ptr(workspace), _stream()), "hgl_mask_pool")
    return out


def pool_score_select(weights: torch.Tensor, tokens: torch.Tensor, sent: torch.Tensor, noun: torch.Tensor, others: torch.Tensor,
                      other_off: torch.Tensor, boxes: torch.Tensor, relaflag: torch.Tensor, score_gem: Optional[torch.Tensor],
                      mask_off: Optional[torch.Tensor] = None, expr_off: Optional[torch.Tensor] = None, max_n: Optional[int] = None,
                      logit_scale_exp: float = 100.0, r: float = 0.5, alpha: float = 0.6,
                      want_features: bool = False, dtype: torch.dtype = torch.bfloat16, out=None):
    """mask_pool + score_select in ONE kernel (hgl_pool_score_select): the pooled, normalised rows are scored straight from
    the f32 TMEM accumulators and only leave the SM when want_features is set.  Returns the score_select dict
    (+ "features" [M,D] of `dtype` when want_features)."""
    _req(weights, torch.float32, "weights")
    w = weights.reshape(weights.shape[0], -1)
    tok = tokens[None] if tokens.dim() == 2 else tokens
    _req(tok, torch.bfloat16, "tokens", 3)
    M, L = w.shape
    B, Lt, D = tok.shape
    if Lt != L:
        raise ValueError(f"weights have {L} cells per mask, tokens {Lt} per image")
    _req(sent, torch.float32, "sent", 2)
    _req(noun, torch.float32, "noun", 2)
    E = sent.shape[0]
    _req(others, torch.float32, "others", 2)
    _req(other_off, torch.int32, "other_off", 1)
    _req(boxes, torch.int64, "boxes", 2)
    _req(relaflag, torch.int32, "relaflag", 1)
    if sent.shape != (E, D) or noun.shape != (E, D) or other_off.numel() != E + 1 or boxes.shape != (M, 4):
        raise ValueError("pool_score_select: inconsistent shapes")
    moff = _offsets(mask_off, B, "mask_off")
    eoff = _offsets(expr_off, B, "expr_off")
    if max_n is None:
        if B != 1:
            raise ValueError("max_n is required for batched calls")
        max_n = max(M, 1)
    if score_gem is not None:
        _req(score_gem, torch.float32, "score_gem", 2)
        if score_gem.shape != (E, max_n):
            raise ValueError("score_gem must be [E, max_n]")
    dev = w.device
    if out is not None:          # caller-supplied result dict (score_clip, idx_hybrid, idx_final, top_idx, blended[, features])
        score_clip, idx_h, idx_f, top, blended = (out[k] for k in ("score_clip", "idx_hybrid", "idx_final", "top_idx", "blended"))
        feats = out.get("features") if want_features else None
    else:
        feats = torch.empty((M, D), dtype=dtype, device=dev) if want_features else None
        score_clip = torch.empty((E, max_n), dtype=torch.float32, device=dev)
        idx_h = torch.empty((E,), dtype=torch.int64, device=dev)
        idx_f = torch.empty((E,), dtype=torch.int64, device=dev)
        top = torch.empty((E, 3), dtype=torch.int32, device=dev)
        blended = torch.empty((E, 3), dtype=torch.float32, device=dev)
    check(_lib.load().hgl_pool_score_select(w.data_ptr(), tok.data_ptr(), _ptr(moff), _ptr(eoff), B, M, E, max_n, L, D,
                                            sent.data_ptr(), noun.data_ptr(), others.data_ptr(), other_off.data_ptr(),
                                            boxes.data_ptr(), relaflag.data_ptr(), _ptr(score_gem),
                                            float(logit_scale_exp), float(r), float(alpha), _ptr(feats), _dt(dtype),
                                            score_clip.data_ptr(), idx_h.data_ptr(), idx_f.data_ptr(), top.data_ptr(),
                                            blended.data_ptr(), _stream()), "hgl_pool_score_select")
    res = dict(score_clip=score_clip, idx_hybrid=idx_h, idx_final=idx_f, top_idx=top, blended=blended)
    if want_features:
        res["features"] = feats
    return res


# ---- (a6)-(a9),(a12) ----------------------------------------------------------------------------------
def score_select(feat: torch.Tensor, sent: torch.Tensor, noun: torch.Tensor, others: torch.Tensor, other_off: torch.Tensor,
                 boxes: torch.Tensor, relaflag: torch.Tensor, score_gem: Optional[torch.Tensor],
                 mask_off: Optional[torch.Tensor] = None, expr_off: Optional[torch.Tensor] = None,
                 max_n: Optional[int] = None, logit_scale_exp: float = 100.0, r: float = 0.5, alpha: float = 0.6,
                 workspace: Optional[torch.Tensor] = None, out=None):
    """Hybridgl_main.py:153-196,225-227 for a batch.  Returns dict(score_clip[E,max_n], idx_hybrid[E], idx_final[E],
    top_idx[E,3], blended[E,3])."""
    _req(feat, (torch.float32, torch.bfloat16), "feat", 2)
    M, De = feat.shape
    _req(sent, torch.float32, "sent", 2)
    _req(noun, torch.float32, "noun", 2)
    E = sent.shape[0]
    _req(others, torch.float32, "others", 2)
    _req(other_off, torch.int32, "other_off", 1)
    _req(boxes, torch.int64, "boxes", 2)
    _req(relaflag, torch.int32, "relaflag", 1)
    if sent.shape != (E, De) or noun.shape != (E, De) or other_off.numel() != E + 1 or boxes.shape != (M, 4):
        raise ValueError("score_select: inconsistent shapes")
    B = 1 if mask_off is None else mask_off.numel() - 1
    moff = _offsets(mask_off, B, "mask_off")
    eoff = _offsets(expr_off, B, "expr_off")
    if max_n is None:
        if B != 1:
            raise ValueError("max_n is required for batched calls")
        max_n = max(M, 1)
    if score_gem is not None:
        _req(score_gem, torch.float32, "score_gem", 2)
        if score_gem.shape != (E, max_n):
            raise ValueError("score_gem must be [E, max_n]")
    dev = feat.device
    if out is not None:          # caller-supplied result dict
        score_clip, idx_h, idx_f, top, blended = (out[k] for k in ("score_clip", "idx_hybrid", "idx_final", "top_idx", "blended"))
    else:
        score_clip = torch.empty((E, max_n), dtype=torch.float32, device=dev)
        idx_h = torch.empty((E,), dtype=torch.int64, device=dev)
        idx_f = torch.empty((E,), dtype=torch.int64, device=dev)
        top = torch.empty((E, 3), dtype=torch.int32, device=dev)
        blended = torch.empty((E, 3), dtype=torch.float32, device=dev)
    check(_lib.load().hgl_score_select(feat.data_ptr(), _dt(feat.dtype), sent.data_ptr(), noun.data_ptr(), others.data_ptr(),
                                       other_off.data_ptr(), boxes.data_ptr(), relaflag.data_ptr(), _ptr(score_gem),
                                       _ptr(moff), _ptr(eoff), B, M, E, De, max_n, float(logit_scale_exp), float(r), float(alpha),
                                       score_clip.data_ptr(), idx_h.data_ptr(), idx_f.data_ptr(), top.data_ptr(),
                                       blended.data_ptr(), _ptr(workspace), _stream()), "hgl_score_select")
    return dict(score_clip=score_clip, idx_hybrid=idx_h, idx_final=idx_f, top_idx=top, blended=blended)


# ---- (a13) --------------------------------------------------------------------------------------------
def iou_accumulate(masks: torch.Tensor, target: torch.Tensor, idx_hybrid: torch.Tensor, idx_final: torch.Tensor,
                   cum: Optional[torch.Tensor], mask_off: Optional[torch.Tensor] = None,
                   expr_off: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Compute_IoU utils.py:365-384 for both picks of every expression.  Returns iu int64 [E,4] =
    (I_hybrid, U_hybrid, I_final, U_final) and adds the column sums into cum int64 [4].
    `masks`: bool/u8 [M,H,W], or the packed int32 [M,H,ceil(W/32)] (the frame size is then taken from `target`)."""
    t = target[None] if target.dim() == 2 else target
    t = _mask_bytes(t, "target")
    B, H, W = t.shape
    packed = masks.dtype == torch.int32
    if packed:
        m = _req(masks, torch.int32, "bits", 3)
        if tuple(m.shape[1:]) != (H, (W + 31) // 32):
            raise ValueError("packed masks and target frames differ")
    else:
        m = _mask_bytes(masks)
        if tuple(m.shape[1:]) != (H, W):
            raise ValueError("masks and target frames differ")
    M = m.shape[0]
    _req(idx_hybrid, torch.int64, "idx_hybrid", 1)
    _req(idx_final, torch.int64, "idx_final", 1)
    E = idx_hybrid.numel()
    moff = _offsets(mask_off, B, "mask_off")
    eoff = _offsets(expr_off, B, "expr_off")
    if cum is not None:
        _req(cum, torch.int64, "cum", 1)
    iu = out if out is not None else torch.empty((E, 4), dtype=torch.int64, device=m.device)
    fn = _lib.load().hgl_iou_bits if packed else _lib.load().hgl_iou
    check(fn(m.data_ptr(), t.data_ptr(), idx_hybrid.data_ptr(), idx_final.data_ptr(), _ptr(moff), _ptr(eoff),
             B, M, E, H, W, iu.data_ptr(), _ptr(cum), _stream()), "hgl_iou_bits" if packed else "hgl_iou")
    return iu
