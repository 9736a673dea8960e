#!/usr/bin/env python
"""Sensitivity of the pass to its phase-A kernels (tuning only): the pass without the blur (black background), and with RLE input."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hybridgl_b200 import synth  # noqa: E402
from hybridgl_b200.pipeline import ScoringPath  # noqa: E402
cfg = synth.CONFIGS[2]
B = 16
H, W, N, E, S, g, De = cfg["h"], cfg["w"], cfg["n_masks"], cfg["n_expr"], cfg["S"], cfg["g"], cfg["De"]
batches = [synth.make_batch_device(1234 + i, B, H, W, N, E, De, device="cuda", grid=g, raw_heat=True) for i in range(2)]
paths = {}
for name, kw in {"default": {}, "no blur (black background)": dict(background="black")}.items():
    path = ScoringPath(size=S, grid=g, feature_source="tokens", overlap=True, **kw)
    paths[name] = [path.capture(b, N) for b in batches]
for rep in range(3):
    for name, graphs in paths.items():
        for _ in range(6):
            for gr in graphs: gr.replay()
        torch.cuda.synchronize()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record()
        n = 400
        for i in range(n): graphs[i & 1].replay()
        b.record(); torch.cuda.synchronize()
        print(f"rep {rep} {name:30s}: {a.elapsed_time(b) / n:.4f} ms/pass", flush=True)
