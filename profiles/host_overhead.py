"""Host-side cost of one ScoringPath.run() call (launch overhead) against the device time of a step (not part of the product).
If the two are close, the step is launch-bound and a CUDA graph / fewer host calls would help."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hybridgl_b200 import synth
from hybridgl_b200.pipeline import ScoringPath

cfg = synth.CONFIGS[2]
B = 16
batches = [synth.make_batch_device(1000 + i, B, cfg["h"], cfg["w"], cfg["n_masks"], cfg["n_expr"], cfg["De"], device="cuda", grid=cfg["g"], raw_heat=True)
           for i in range(2)]
for overlap in (True, False):
    path = ScoringPath(size=cfg["S"], grid=cfg["g"], feature_source="tokens", overlap=overlap)
    for ev in (False, True):
        for _ in range(10):
            path.run(batches[0], cfg["n_masks"])
        torch.cuda.synchronize()
        n = 200
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record()
        for s in range(n):
            path.events = [] if ev else None
            path.run(batches[s % 2], cfg["n_masks"])
        t_host = time.perf_counter() - t0
        e1.record(); torch.cuda.synchronize()
        path.events = None
        print(f"overlap={overlap} stage_events={ev}: host {t_host / n * 1e3:.3f} ms per run() (enqueue only), device {e0.elapsed_time(e1) / n:.3f} ms per step", flush=True)
