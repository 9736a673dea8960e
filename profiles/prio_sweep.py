#!/usr/bin/env python
"""Stream-priority arrangements of the pass (tuning only):  python profiles/prio_sweep.py
side = pack -> mask pass -> pool+score -> IoU, pre = blur -> prep setup, tab = heat-map tables; main = prep main (priority 0)."""
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
print("priority range", torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, "priority_range") else "?")
cands = {"tab -3 side -2 pre -1 (HEAD)": (-2, -1, -3), "tab -3 side -2 pre -3": (-2, -3, -3), "tab -2 side -2 pre -3": (-2, -3, -2), "tab -3 side -1 pre -3": (-1, -3, -3),
         "all -1": (-1, -1, -1), "tab -3 side -3 pre -3": (-3, -3, -3)}
paths = {}
for name, (ps, pp, pt) in cands.items():
    path = ScoringPath(size=S, grid=g, feature_source="tokens", overlap=True)
    path._side = torch.cuda.Stream(priority=ps); path._pre = torch.cuda.Stream(priority=pp); path._tab = torch.cuda.Stream(priority=pt)
    paths[name] = (path, [path.capture(b, N) for b in batches])
for rep in range(3):
    for name, (path, graphs) in paths.items():
        for _ in range(6):
            for gr in graphs: gr.replay()
        torch.cuda.synchronize()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record()
        n = 400
        for i in range(n): graphs[i & 1].replay()
        b.record(); torch.cuda.synchronize()
        print(f"rep {rep} {name:36s}: {a.elapsed_time(b) / n:.4f} ms/pass", flush=True)
