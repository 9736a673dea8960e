"""ncu target: a few hgl_prep launches at the bench shape (not part of the product)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hybridgl_b200 import ops, synth
cfg = synth.CONFIGS[2]
B = 16
batch = synth.make_batch_device(1234, B, cfg["h"], cfg["w"], cfg["n_masks"], cfg["n_expr"], cfg["De"], device="cuda")
img = batch["image"]; blur = ops.gaussian_blur15(img); bits = ops.pack_masks(batch["masks"])
M = bits.shape[0]
loc = torch.empty((M, 3, cfg["S"], cfg["S"]), dtype=torch.bfloat16, device="cuda"); glo = torch.empty_like(loc)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    ops.prep_visual_prompts(img, blur, bits, cfg["S"], mask_off=batch["mask_off"], max_n=cfg["n_masks"], dtype=torch.bfloat16, out=(loc, glo))
torch.cuda.synchronize()
