"""ctypes binding of libhgl.so (include/hgl.h).  Fails loudly: there is no CPU or PyTorch fallback."""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_void_p

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HGL_LIB") or os.path.join(_PKG, "libhgl.so")     # HGL_LIB: profiling builds (build.py, HGL_BUILD_TUNING=1)

HGL_F32, HGL_BF16 = 0, 1
HGL_BG_BLUR, HGL_BG_BLACK, HGL_BG_NONE = 0, 1, 2
HGL_LND, HGL_NLD = 0, 1
REL_CODES = {"none": 0, "left": 1, "right": 2, "up": 3, "down": 4, "big": 5, "small": 6, "within": 7}
DIR_CODES = {"none": 0, "left": 1, "right": 2, "middle": 3, "up": 4, "down": 5}

# name -> (restype, argtypes); mirrors include/hgl.h one to one (tests/test_abi.py checks the header against this)
SIGNATURES = {
    "hgl_last_error": (c_char_p, []),
    "hgl_version": (c_int, []),
    "hgl_check_device": (c_int, []),
    "hgl_pack_masks": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "hgl_pack_masks_bool": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "hgl_mask_geometry": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "hgl_prep_workspace_bytes": (c_int64, [c_int, c_int, c_int]),
    "hgl_prep": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                         c_void_p, c_void_p, c_void_p, c_void_p]),
    "hgl_prep_crop": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                              c_void_p, c_void_p, c_void_p]),
    "hgl_ellipse_outline": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "hgl_prep_circle": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                c_int, c_int, c_int, c_void_p, c_void_p]),
    "hgl_prep_setup": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "hgl_prep_main": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                              c_void_p]),
    "hgl_gaussian_blur15": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "hgl_mask_grid_workspace_bytes": (c_int64, [c_int, c_int]),
    "hgl_mask_grid": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "hgl_grid_heat_pool_workspace_bytes": (c_int64, [c_int, c_int, c_int, c_int, c_int, c_int, c_int]),
    "hgl_attn_mask": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "hgl_attn_bias": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "hgl_cls_attention": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "hgl_cls_head": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_double, c_int, c_int, c_int, c_void_p,
                             c_void_p]),
    "hgl_token_mask_fuse_ln": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_float, c_void_p, c_void_p, c_float, c_int, c_int, c_int, c_int,
                                       c_void_p, c_void_p, c_void_p]),
    "hgl_token_mask_fuse": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_float, c_int, c_int, c_int, c_int, c_int,
                                    c_void_p, c_void_p]),
    "hgl_dir_mask": (c_int, [c_int, c_int, c_int, c_void_p, c_void_p]),
    "hgl_heat_pool_workspace_bytes": (c_int64, [c_int, c_int, c_int, c_int, c_int, c_int]),
    "hgl_heat_pool": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                              c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "hgl_grid_heat_pool": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "hgl_heat_resize_aa": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "hgl_grid_heat_pool_raw_workspace_bytes": (c_int64, [c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int]),
    "hgl_grid_heat_pool_raw": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                       c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                       c_void_p]),
    "hgl_gem_token_workspace_bytes": (c_int64, [c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int]),
    "hgl_gem_token_pool": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int,
                                   c_int, c_void_p, c_void_p, c_void_p]),
    "hgl_heat_tables": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "hgl_grid_heat_pool_rows": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int,
                                        c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "hgl_score_select": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_double, c_double, c_double,
                                 c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "hgl_mask_pool_workspace_bytes": (c_int64, [c_int, c_int, c_int]),
    "hgl_mask_pool": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                              c_void_p]),
    "hgl_pool_score_select": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_double, c_double, c_double, c_void_p, c_int,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "hgl_score_select_workspace_bytes": (c_int64, [c_int, c_int, c_int]),
    "hgl_iou": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                        c_void_p, c_void_p, c_void_p]),
    "hgl_iou_bits": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                             c_void_p, c_void_p, c_void_p]),
    "hgl_rle_to_bits": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
}

_lib = None


class HglError(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """dlopen libhgl.so and attach prototypes.  Raises if the library has not been built
    (`python -m hybridgl_b200.build`); nothing in this package computes without it."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HglError(f"{LIB_PATH} is missing: build it with `python -m hybridgl_b200.build` "
                       "(there is no CPU / PyTorch fallback for the scoring path)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)       # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().hgl_last_error()
        raise HglError(f"{what} failed (code {rc}): {msg.decode() if msg else '?'}")
