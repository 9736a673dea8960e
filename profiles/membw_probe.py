"""Calibration probe (not part of the product): pure-write / copy / pure-read HBM bandwidth with torch kernels."""
import torch
x = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
y = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
xf = x.view(torch.float32)
def t(fn, nbytes, name, it=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(it):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    print(f"{name}: {nbytes / best / 1e6:.0f} GB/s ({best:.3f} ms)")
t(lambda: x.fill_(1), x.numel(), "fill 1 GiB (pure write)")
t(lambda: xf.zero_(), x.numel(), "zero f32 1 GiB (pure write)")
t(lambda: y.copy_(x), 2 * x.numel(), "copy 1 GiB (read+write)")
t(lambda: xf.sum(), x.numel(), "sum f32 1 GiB (pure read)")
t(lambda: torch.cuda.memset if False else x.zero_(), x.numel(), "zero u8")
