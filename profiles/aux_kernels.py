#!/usr/bin/env python
"""Stand-alone timing of the path's kernels that bench.py's step does not launch (the backbone-side rows of SURVEY 8(a):
a2 non-antialiased grid, a3 attention mask / key bias, a4 token masking + stream mix) and of the drop-in single-purpose entry
points, at the BASELINE configs[1] per-image shape x 4 images.  L2 flushed before every launch, CUDA events, median of 20.

    python profiles/aux_kernels.py > gpurun_out/aux.md
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hybridgl_b200 import ops, synth  # noqa: E402

peak = 6550.0
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


cfg = synth.CONFIGS[2]
B, N, E, H, W, g, S = 4, cfg["n_masks"], cfg["n_expr"], cfg["h"], cfg["w"], cfg["g"], cfg["S"]
Dv, heads, L = cfg["Dv"], cfg["heads"], cfg["g"] ** 2
M = B * N
batch = synth.make_batch_device(4321, B, H, W, N, E, cfg["De"], device="cuda", grid=g, raw_heat=True)
bits = ops.pack_masks(batch["masks"])
grid = ops.masks_to_grid(bits, g, antialias=True, width=W)
WW = (W + 31) // 32
rows = []


def add(name, ref, fn, nbytes):
    ms = timeit(fn)
    gbs = nbytes / ms / 1e6
    rows.append(f"| `{name}` | {ref} | {nbytes / 1e6:.1f} | {ms * 1e3:.1f} | {gbs:.0f} | {gbs / peak:.3f} |")


for dt, nm in ((torch.bfloat16, "bf16"), (torch.float32, "f32")):
    x2 = torch.randn((L + 1, M, Dv), device="cuda").to(dt)
    x = torch.randn((L + 1, M, Dv), device="cuda").to(dt)
    out = torch.empty_like(x)
    eb = x.element_size()
    add(f"hgl_token_mask_fuse LND {nm}", "model/backbone.py:236-249 (a4)", lambda: ops.token_mask_fuse(x2, x, grid, 2.0, 1.0, out=out, layout="LND"),
        3 * (L + 1) * M * Dv * eb + M * L * 4)
    xn = x.permute(1, 0, 2).contiguous(); x2n = x2.permute(1, 0, 2).contiguous(); outn = torch.empty_like(xn)
    add(f"hgl_token_mask_fuse NLD {nm}", "same, batch-first streams", lambda: ops.token_mask_fuse(x2n, xn, grid, 2.0, 1.0, out=outn, layout="NLD"),
        3 * (L + 1) * M * Dv * eb + M * L * 4)
    del x, x2, out, xn, x2n, outn
add("hgl_attn_mask", "model/backbone.py:108-115 (a3, full bool tensor)", lambda: ops.make_attn_mask(grid, heads), M * heads * (L + 1) ** 2 + M * L * 4)
add("hgl_attn_bias", "a3, CLS-row key bias", lambda: ops.attn_key_bias(grid), M * (L + 1) * 4 + M * L * 4)
add("hgl_mask_grid antialias=1", "model/backbone.py:160 (a2), torchvision >= 0.17", lambda: ops.masks_to_grid(bits, g, antialias=True, width=W), M * H * WW * 4 + M * L * 4)
add("hgl_mask_grid antialias=0", "a2, torchvision 0.15.2 (the reference's pin)", lambda: ops.masks_to_grid(bits, g, antialias=False, width=W), M * H * WW * 4 + M * L * 4)
full = torch.empty((B * E, H, W), dtype=torch.float32, device="cuda")
add("hgl_heat_resize_aa", "Hybridgl_main.py:201", lambda: ops.heat_resize_aa(batch["heat"], H, W, out=full), B * E * (28 * 37 + H * W) * 4)
cum = torch.zeros(4, dtype=torch.int64, device="cuda")
idx = torch.zeros((B * E,), dtype=torch.int64, device="cuda")
add("hgl_iou (byte masks)", "utils.py:365-384 (a13)", lambda: ops.iou_accumulate(batch["masks"], batch["target"], idx, idx, cum, batch["mask_off"], batch["expr_off"]),
    2 * 2 * B * E * H * W)
add("hgl_iou_bits (packed masks)", "a13", lambda: ops.iou_accumulate(bits, batch["target"], idx, idx, cum, batch["mask_off"], batch["expr_off"]),
    2 * B * E * (H * W + H * WW * 4))
print(f"# Kernels outside the bench step, timed alone (B={B} images x {N} masks, {H}x{W}, L={L}, Dv={Dv}, heads={heads}; L2 flushed, median of 20)\n")
print(f"HBM peak used for `frac`: {peak:.0f} GB/s (MEASURED_PEAKS.json).\n")
print("| entry point | reference | algorithmic MB | us | GB/s | frac of HBM peak |\n|---|---|---|---|---|---|")
print("\n".join(rows))
