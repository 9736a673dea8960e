#!/usr/bin/env python
"""Per-kernel tuning sweeps on one B200 (not part of the product; writes gpurun_out/<tag>_kernels.log).

    python profiles/bench_kernels.py [tag]

* pure-write / copy / pure-read HBM bandwidth with torch kernels (ceiling for write-dominated kernels such as prep)
* hgl_prep: z-split sweep (HGL_PREP_GZ), bulk-copy staging on / off (HGL_PREP_NO_TMA), bf16 and f32 outputs
Each timing: L2 flushed by the 0.96 GB of output the kernel itself writes; best and median of 20 runs, CUDA events.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hybridgl_b200 import ops, synth  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "run"
out = open(os.path.join(ROOT, "gpurun_out", f"{tag}_kernels.log"), "w")


def log(*a):
    s = " ".join(str(x) for x in a)
    print(s)
    out.write(s + "\n")
    out.flush()


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[0], ts[len(ts) // 2]


def membw():
    x = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
    y = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
    xf = x.view(torch.float32)
    for name, fn, nbytes in (("fill u8 1 GiB (pure write)", lambda: x.fill_(1), x.numel()),
                             ("zero f32 1 GiB (pure write)", lambda: xf.zero_(), x.numel()),
                             ("copy 1 GiB (read+write)", lambda: y.copy_(x), 2 * x.numel()),
                             ("sum f32 1 GiB (pure read)", lambda: xf.sum(), x.numel())):
        best, med = timeit(fn, 10)
        log(f"membw {name}: best {nbytes / best / 1e6:.0f} GB/s, median {nbytes / med / 1e6:.0f} GB/s")
    del x, y


def prep_sweep():
    cfg = synth.CONFIGS[2]
    B = 16
    batch = synth.make_batch_device(1234, B, cfg["h"], cfg["w"], cfg["n_masks"], cfg["n_expr"], cfg["De"], device="cuda")
    img = batch["image"]
    blur = ops.gaussian_blur15(img)
    bits = ops.pack_masks(batch["masks"])
    M = bits.shape[0]
    for dtype, eb in ((torch.bfloat16, 2), (torch.float32, 4)):
        loc = torch.empty((M, 3, cfg["S"], cfg["S"]), dtype=dtype, device="cuda"); glo = torch.empty_like(loc)
        ws = torch.empty((1 << 26,), dtype=torch.uint8, device="cuda")
        alg = M * cfg["h"] * ((cfg["w"] + 31) // 32) * 4 + 2 * B * cfg["h"] * cfg["w"] * 3 + 2 * M * 3 * cfg["S"] ** 2 * eb
        for sub in (8, 16, 4):
            os.environ["HGL_PREP_SUB"] = str(sub)
            for gz in (0, 1, 2, 3):
                os.environ.pop("HGL_PREP_GZ", None)
                if gz:
                    os.environ["HGL_PREP_GZ"] = str(gz)
                fn = lambda: ops.prep_visual_prompts(img, blur, bits, cfg["S"], mask_off=batch["mask_off"], max_n=cfg["n_masks"],  # noqa: E731
                                                     dtype=dtype, out=(loc, glo), workspace=ws)
                best, med = timeit(fn)
                log(f"prep {str(dtype)[6:]} sub={sub} gz={gz or 'auto'}: best {best * 1e3:.1f} us, median {med * 1e3:.1f} us, {alg / med / 1e6:.0f} GB/s")
        os.environ.pop("HGL_PREP_SUB", None); os.environ.pop("HGL_PREP_GZ", None)
        for tma in (1, 0):
            for gz in (0,):
                os.environ.pop("HGL_PREP_GZ", None); os.environ.pop("HGL_PREP_NO_TMA", None)
                if gz:
                    os.environ["HGL_PREP_GZ"] = str(gz)
                if not tma:
                    os.environ["HGL_PREP_NO_TMA"] = "1"
                fn = lambda: ops.prep_visual_prompts(img, blur, bits, cfg["S"], mask_off=batch["mask_off"], max_n=cfg["n_masks"],  # noqa: E731
                                                     dtype=dtype, out=(loc, glo), workspace=ws)
                best, med = timeit(fn)
                log(f"prep {str(dtype)[6:]} tma={tma} gz={gz or 'auto'}: best {best * 1e3:.1f} us, median {med * 1e3:.1f} us, "
                    f"{alg / med / 1e6:.0f} GB/s algorithmic (median)")
        os.environ.pop("HGL_PREP_GZ", None); os.environ.pop("HGL_PREP_NO_TMA", None)
        for dbg in (1, 2, 3):
            os.environ["HGL_PREP_DEBUG"] = str(dbg)
            fn = lambda: ops.prep_visual_prompts(img, blur, bits, cfg["S"], mask_off=batch["mask_off"], max_n=cfg["n_masks"],  # noqa: E731
                                                 dtype=dtype, out=(loc, glo), workspace=ws)
            best, med = timeit(fn)
            log(f"prep {str(dtype)[6:]} DEBUG={dbg} (1: no fix-up, 2: no P1 stores, 3: neither): best {best * 1e3:.1f} us, median {med * 1e3:.1f} us")
        os.environ.pop("HGL_PREP_DEBUG", None)
        del loc, glo


def mask_pool_sweep():
    import json
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    tf_peak = float(peaks.get("bf16_tflops", 1590.0)); bw_peak = float(peaks.get("hbm_gbs", 6650.0))
    for name, B, n, L, D in (("cfg2 B=16", 16, 100, 196, 768), ("cfg2 B=256", 256, 100, 196, 768), ("cfg4 B=16", 16, 200, 576, 1024),
                             ("cfg4 B=128", 128, 200, 576, 1024), ("cfg5 B=256", 256, 200, 196, 768)):
        M = B * n
        w = torch.rand((M, L), device="cuda"); w[w < 0.5] = 0
        tok = torch.randn((B, L, D), device="cuda").to(torch.bfloat16)
        moff = (torch.arange(B + 1, device="cuda") * n).to(torch.int32)
        for dtype, ob in ((torch.float32, 4), (torch.bfloat16, 2)):
            ws = torch.empty((M * D * 4 + 512,), dtype=torch.uint8, device="cuda")
            fn = lambda: ops.mask_pool(w, tok, moff, n, normalize=True, dtype=dtype, workspace=ws)  # noqa: E731
            best, med = timeit(fn)
            flop = 2.0 * M * L * D
            byts = M * L * 4 + B * L * D * 2 + M * D * ob
            log(f"mask_pool {name} out={str(dtype)[6:]}: median {med * 1e3:.1f} us, {flop / med / 1e9:.1f} TFLOP/s ({flop / med / 1e9 / tf_peak * 100:.1f}% of measured bf16 peak), "
                f"{byts / med / 1e6:.0f} GB/s algorithmic ({byts / med / 1e6 / bw_peak * 100:.1f}% of measured HBM peak), AI {flop / byts:.0f} flop/B")
        del w, tok


if __name__ == "__main__":
    ops.device_ok()
    which = sys.argv[2:] or ["membw", "prep", "mask_pool"]
    if "membw" in which:
        membw()
    if "prep" in which:
        prep_sweep()
    if "mask_pool" in which:
        mask_pool_sweep()
