"""ncu target: a few hgl_mask_pool launches (not part of the product)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hybridgl_b200 import ops
B, n, L, D = (int(x) for x in (sys.argv[1:5] if len(sys.argv) > 4 else (256, 100, 196, 768)))
M = B * n
w = torch.rand((M, L), device="cuda"); w[w < 0.5] = 0
tok = torch.randn((B, L, D), device="cuda").to(torch.bfloat16)
moff = (torch.arange(B + 1, device="cuda") * n).to(torch.int32)
for _ in range(3):
    ops.mask_pool(w, tok, moff, n, normalize=True, dtype=torch.bfloat16)
torch.cuda.synchronize()
