#!/usr/bin/env python
"""The evaluation loop of Hybridgl_main.py:79-247 with libhgl behind it, on synthetic SAM / GEM / text tensors.

    python examples/eval_synthetic.py [--images 4] [--masks 100] [--expr 3] [--fusion_mode "G2L&L2G"] [--bf16]

Per image, in the reference's order (the comments name the lines each call replaces):
    masks, boxes            <- SAM (synthetic here)                               Hybridgl_main.py:84-90
    bits                    = ops.pack_masks(masks)                               (byte masks are read once)
    local, global           = ops.prep_visual_prompts(...)                        :92-125   (blur on the GPU, cv2's fixed-point taps)
    features                = CLIPViTFM(local, global, masks, masking_block, fusion_mode)     :128, model/backbone.py:117-306
    grid/area/score_gem     = ops.grid_heat_pool(bits, raw GEM maps, dirflag, black)          :201-223
    idx_hybrid, idx_final   = ops.score_select(features, text embeddings, boxes, relaflag, score_gem)   :153-196, :225-227
    I / U counters          = ops.iou_accumulate(...)                                         :171, :230
and the four numbers of the result log at the end (:240-247).  CLIP runs as plain PyTorch with random-init weights of the
named architecture (no checkpoints offline), so the IoU values are those of random features: the point is the call surface
and the time split between the path (libhgl) and the backbone.
"""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hybridgl_b200 import ops, sweep, synth  # noqa: E402
from hybridgl_b200.backbone import CLIPViTFM  # noqa: E402


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=4)
    ap.add_argument("--masks", type=int, default=100)
    ap.add_argument("--expr", type=int, default=3)
    ap.add_argument("--fusion_mode", default="G2L&L2G", choices=["G2L", "L2G", "G2L&L2G"])
    ap.add_argument("--masking_block", type=int, default=9)
    ap.add_argument("--bf16", action="store_true", help="run the ViT blocks and the prep outputs in bf16")
    ap.add_argument("--unfused-ln", action="store_true", help="token masking, torch.cat and ln_1 as separate passes (the round-1 forward) "
                    "instead of hgl_token_mask_fuse_ln")
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--width", type=int, default=640)
    a = ap.parse_args(argv)
    ops.device_ok()
    dev = torch.device("cuda", torch.cuda.current_device())
    dt = torch.bfloat16 if a.bf16 else torch.float32
    model = CLIPViTFM("ViT-B/16", device=dev, dtype=dt)                               # Hybridgl_main.py:47
    model.fused_ln = not a.unfused_ln
    S, g, de = 224, 14, 512
    cum = torch.zeros(4, dtype=torch.int64, device=dev)                              # cum_I, cum_U, cum_I_final, cum_U_final (:52-55)
    rows = []
    t_path = t_vit = 0.0
    for i in range(-1, a.images):                 # image -1: untimed warm-up (cuBLAS / SDPA heuristics, workspace allocation)
        if i == 0:
            cum.zero_(); rows.clear(); t_path = t_vit = 0.0
        b = synth.make_batch_device(100 + i, 1, a.height, a.width, a.masks, a.expr, de, device=dev, grid=g, raw_heat=True)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        bits = ops.pack_masks(b["masks"])
        blur = ops.gaussian_blur15(b["image"])
        local, glob = ops.prep_visual_prompts(b["image"], blur, bits, S, dtype=dt)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        feats = model(local, glob, b["masks"], masking_block=a.masking_block, fusion_mode=a.fusion_mode)
        torch.cuda.synchronize(); t2 = time.perf_counter()
        _, _, score_gem = ops.grid_heat_pool(bits, a.width, g, b["heat"], b["dirflag"], b["black"])
        res = ops.score_select(feats.float(), b["sent"], b["noun"], b["others"], b["other_off"], b["boxes"], b["relaflag"], score_gem,
                               logit_scale_exp=float(model.model.logit_scale.detach().exp()))
        iu = ops.iou_accumulate(bits, b["target"], res["idx_hybrid"], res["idx_final"], cum)
        torch.cuda.synchronize(); t3 = time.perf_counter()
        rows.append(iu)
        t_path += (t1 - t0) + (t3 - t2); t_vit += t2 - t1
    rep = sweep.report(cum, torch.cat(rows))
    n_expr = a.images * a.expr
    print(f"{a.images} images x {a.masks} proposals x {a.expr} expressions, {a.fusion_mode}, {'bf16' if a.bf16 else 'f32'}: "
          f"oIoU {rep['oIoU']:.2f} mIoU {rep['mIoU']:.2f} oIoU_final {rep['oIoU_final']:.2f} mIoU_final {rep['mIoU_final']:.2f} | "
          f"path (libhgl) {t_path / a.images * 1e3:.2f} ms/image, ViT blocks (PyTorch) {t_vit / a.images * 1e3:.1f} ms/image, "
          f"{n_expr / (t_path + t_vit):.1f} expressions/s end to end with the backbone")
    return rep


if __name__ == "__main__":
    main()
