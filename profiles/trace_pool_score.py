#!/usr/bin/env python
"""Phase trace of hgl_pool_score_select at the bench shape (tuning build only; not part of the product).

    HGL_BUILD_TUNING=1 python -m hybridgl_b200.build
    HGL_LIB=hybridgl_b200/libhgl_tuning.so python profiles/trace_pool_score.py

Thread 0 of every CTA stamps clock64() at the phase boundaries (PS_TRACE in csrc/pool_score.cu); prints the median duration
of every phase in microseconds (1.9 GHz assumed) for rank-0 CTAs and for the other CTAs of the clusters.
"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hybridgl_b200 import ops, synth  # noqa: E402
from hybridgl_b200.pipeline import ScoringPath  # noqa: E402

cfg = synth.CONFIGS[2]
B = 16
H, W, N, E, S, g, De = cfg["h"], cfg["w"], cfg["n_masks"], cfg["n_expr"], cfg["S"], cfg["g"], cfg["De"]
batch = synth.make_batch_device(1234, B, H, W, N, E, De, device="cuda", grid=g, raw_heat=True)
path = ScoringPath(size=S, grid=g, feature_source="tokens", overlap=False)
res = path.run(batch, N)
torch.cuda.synchronize()
grid, sg = res["grid"], res["score_gem"]
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
lib = ops._lib.load()
lib.hgl_debug_pool_score_trace.restype = ctypes.c_int
lib.hgl_debug_pool_score_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
NT = 4 if De > 64 else 1
names = ["prologue", "issue+text", "it0", "it1", "it2", "it3", "loop end", "wait acc", "tmem pass", "cluster sync", "reduce", "tail", "exit"]
for cold in (True, False):
    rows = []
    for rep in range(7):
        if cold:
            flush.zero_()
        a = torch.cuda.Event(enable_timing=True); b_ = torch.cuda.Event(enable_timing=True)
        a.record()
        ops.pool_score_select(grid, batch["tokens"], batch["sent"], batch["noun"], batch["others"], batch["other_off"], batch["boxes"],
                              batch["relaflag"], sg, batch["mask_off"], batch["expr_off"], N)
        b_.record()
        torch.cuda.synchronize()
        buf = np.zeros(16 * B * NT, np.int64)
        assert lib.hgl_debug_pool_score_trace(buf.ctypes.data, B * NT) == 0
        rows.append(buf.reshape(B * NT, 16))
    t = np.stack(rows[2:])                                    # [rep, cta, 16]
    d = np.diff(t[:, :, :14], axis=2) / 1900.0                # us
    r0 = d[:, 0::NT, :]; rp = d[:, [i for i in range(B * NT) if i % NT], :]
    print(f"--- {'cold L2' if cold else 'warm L2'}: phase durations, median us (rank 0 | peers); total rank0 {np.median(r0.sum(2)):.1f} peers {np.median(rp.sum(2)):.1f}")
    for k, nm in enumerate(names):
        print(f"  {nm:14s} {np.median(r0[:, :, k]):7.2f} | {np.median(rp[:, :, k]):7.2f}")
