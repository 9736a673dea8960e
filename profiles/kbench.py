#!/usr/bin/env python
"""Stand-alone timing of every stage of the path at the bench shape (not part of the product).

    python profiles/kbench.py [tag] [stage ...]          # writes gpurun_out/<tag>_kbench.log

Each stage is timed alone on the default stream, L2 flushed before every launch (512 MB memset), CUDA events,
best / median of 20.  Used for kernel iteration; bench.py is the judged measurement.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hybridgl_b200 import ops, synth  # noqa: E402
from hybridgl_b200.pipeline import ScoringPath  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "run"
only = set(sys.argv[2:])
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
out = open(os.path.join(ROOT, "gpurun_out", f"{tag}_kbench.log"), "w")


def log(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True)
    out.write(s + "\n")
    out.flush()


flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=20, warm=3, flush=True):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush:
            flush_buf.zero_()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[0], ts[len(ts) // 2]


cfg = synth.CONFIGS[int(os.environ.get("KB_CFG", "2"))]
B = int(os.environ.get("KB_IMAGES", "16"))
H, W, N, E, S, g, De = cfg["h"], cfg["w"], cfg["n_masks"], cfg["n_expr"], cfg["S"], cfg["g"], cfg["De"]
batch = synth.make_batch_device(1234, B, H, W, N, E, De, device="cuda", grid=g, raw_heat=True)
counts, roff = synth.masks_to_rle_device(batch["masks"])
path = ScoringPath(size=S, grid=g, feature_source="tokens", overlap=False, keep_features=True)
res = path.run(batch, N)
torch.cuda.synchronize()
bits, grid, feats, sg = res["bits"].clone(), res["grid"], res["features"], res["score_gem"]
moff, eoff = batch["mask_off"], batch["expr_off"]
img = batch["image"]
blur = ops.gaussian_blur15(img)
M = B * N
lib = ops._lib.load()
bits2 = torch.empty_like(bits)
loc = torch.empty((M, 3, S, S), dtype=torch.bfloat16, device="cuda"); glo = torch.empty_like(loc)
pws = torch.empty((max(lib.hgl_prep_workspace_bytes(B, S, ops.HGL_BF16), 1),), dtype=torch.uint8, device="cuda")
hws = torch.empty((lib.hgl_grid_heat_pool_raw_workspace_bytes(B, M, B * E, H, W, g, N, batch["heat"].shape[1], batch["heat"].shape[2]),), dtype=torch.uint8, device="cuda")
cum = torch.zeros(4, dtype=torch.int64, device="cuda")
WW = (W + 31) // 32
stages = {
    "blur": (lambda: ops.gaussian_blur15(img, out=blur), 2 * B * H * W * 3),
    "pack": (lambda: ops.pack_masks(batch["masks"], out=bits2), M * H * W + M * H * WW * 4),
    "rle": (lambda: ops.rle_to_bits(counts, roff, H, W, out=bits2), counts.numel() * 4 + M * H * WW * 4),
    "prep": (lambda: ops.prep_visual_prompts(img, blur, bits, S, mask_off=moff, max_n=N, dtype=torch.bfloat16, out=(loc, glo), workspace=pws),
             M * H * WW * 4 + 2 * B * H * W * 3 + 2 * M * 3 * S * S * 2),
    "grid_heat_pool": (lambda: ops.grid_heat_pool(bits, W, g, batch["heat"], batch["dirflag"], batch["black"], moff, eoff, N, workspace=hws),
                       M * H * WW * 4 + M * g * g * 4 + 2 * B * E * H * W * 4 + B * E * N * 4),
    "grid_all": (lambda: ops.masks_to_grid(bits, g, width=W), M * H * WW * 4 + M * g * g * 4),
    "grid_one": (lambda: ops.masks_to_grid(bits[:1], g, width=W), H * WW * 4 + g * g * 4),
    "tables": (lambda: ops.heat_tables(batch["heat"], batch["dirflag"], H, W, hws), 2 * B * E * H * W * 4),
    "rows": (lambda: ops.grid_heat_pool_rows(bits, W, g, tuple(batch["heat"].shape), batch["black"], moff, eoff, N, hws),
             M * H * WW * 4 + M * g * g * 4 + B * E * N * 4),
    "mask_pool": (lambda: ops.mask_pool(grid, batch["tokens"], moff, N, normalize=True, dtype=torch.bfloat16),
                  M * g * g * 4 + B * g * g * De * 2 + M * De * 2),
    "pool_score": (lambda: ops.pool_score_select(grid, batch["tokens"], batch["sent"], batch["noun"], batch["others"], batch["other_off"],
                                                 batch["boxes"], batch["relaflag"], sg, moff, eoff, N),
                   M * g * g * 4 + B * g * g * De * 2 + 3 * B * E * De * 4 + 32 * M + 12 * B * E * N),
    "score_select": (lambda: ops.score_select(feats, batch["sent"], batch["noun"], batch["others"], batch["other_off"], batch["boxes"],
                                              batch["relaflag"], sg, moff, eoff, N),
                     M * De * 2 + 3 * B * E * De * 4 + 32 * M + 12 * B * E * N),
    "iou": (lambda: ops.iou_accumulate(batch["masks"], batch["target"], res["idx_hybrid"], res["idx_final"], cum, moff, eoff), 4 * B * E * H * W),
    "iou_bits": (lambda: ops.iou_accumulate(bits, batch["target"], res["idx_hybrid"], res["idx_final"], cum, moff, eoff),
                 2 * B * E * (H * W + H * WW * 4)),
}
log(f"shape: B={B} N={N} E={E} {H}x{W} S={S} g={g} De={De}; runs/mask {counts.numel() / M:.1f}")
for name, (fn, nbytes) in stages.items():
    if only and name not in only:
        continue
    best, med = timeit(fn)
    best_w, med_w = timeit(fn, flush=False)
    log(f"{name:15s} cold-L2 best {best * 1e3:7.1f} us  median {med * 1e3:7.1f} us  ({nbytes / med / 1e6:6.0f} GB/s of {nbytes / 1e6:.1f} MB)"
        f"   warm best {best_w * 1e3:7.1f} us median {med_w * 1e3:7.1f} us")
