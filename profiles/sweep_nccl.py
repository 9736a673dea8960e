"""BASELINE.json configs[4] on real GPUs: PhraseCut-shaped sweep sharded by image over `WORLD_SIZE` ranks (NCCL), checked
against the SAME sweep run by one process (rank 0 alone, no collective).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        profiles/sweep_nccl.py [--images 24] [--per-step 4]

The reference is single-process (Hybridgl_main_PhraseCut.py:66-70 loops over the DataLoader); images are independent, so
rank r owns images r, r+world, ... and the only exchange is hybridgl_b200/sweep.py::reduce_counters at the end.  Integer
accumulators and the gathered per-expression (I, U) rows must be IDENTICAL at every world size; prints one JSON line.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from hybridgl_b200 import sweep, synth          # noqa: E402
from hybridgl_b200.pipeline import ScoringPath  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=24)
    ap.add_argument("--per-step", type=int, default=4)
    ap.add_argument("--workload", type=int, default=5)
    a = ap.parse_args()
    cfg = synth.CONFIGS[a.workload]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    E = cfg["n_expr"]

    def make_batch(ids):
        # one seeded single-image batch per image id (so every world size sees the same image i), concatenated
        parts = [synth.make_batch_device(5000 + i, 1, cfg["h"], cfg["w"], cfg["n_masks"], E, cfg["De"], device=dev, grid=cfg["g"],
                                         raw_heat=True) for i in ids]
        n = len(parts)
        cat = {k: torch.cat([p[k] for p in parts]) for k in parts[0] if k not in ("mask_off", "expr_off", "other_off")}
        cat["mask_off"] = (torch.arange(n + 1, device=dev) * cfg["n_masks"]).to(torch.int32)
        cat["expr_off"] = (torch.arange(n + 1, device=dev) * E).to(torch.int32)
        k_other = parts[0]["others"].shape[0] // E
        cat["other_off"] = (torch.arange(n * E + 1, device=dev) * k_other).to(torch.int32)
        eids = torch.cat([torch.arange(E, dtype=torch.int64) + i * E for i in ids])
        return cat, cfg["n_masks"], eids

    path = ScoringPath(size=cfg["S"], grid=cfg["g"], feature_source="tokens")
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    out = sweep.run_sweep(a.images, make_batch, path, images_per_step=a.per_step)
    t1.record()
    torch.cuda.synchronize()
    ok = True
    single = None
    if rank == 0:
        # the same sweep in one process: temporarily hide the process group from sweep._world
        ref_path = ScoringPath(size=cfg["S"], grid=cfg["g"], feature_source="tokens")
        real = sweep._world
        sweep._world = lambda group=None: (0, 1)
        try:
            single = sweep.run_sweep(a.images, make_batch, ref_path, images_per_step=a.per_step)
        finally:
            sweep._world = real
        ok = (out["cum"].tolist() == single["cum"].tolist() and torch.equal(out["iu"].cpu(), single["iu"].cpu())
              and out["expr_ids"].tolist() == list(range(a.images * E))
              and all(out[k] == single[k] for k in ("oIoU", "mIoU", "oIoU_final", "mIoU_final")))
        print(json.dumps({"sweep": f"configs[{a.workload - 1}] {cfg['h']}x{cfg['w']}, {cfg['n_masks']} masks, {E} expressions/image",
                          "images": a.images, "world": world, "backend": "nccl" if world > 1 else "none",
                          "cum": out["cum"].tolist(), "oIoU": out["oIoU"], "mIoU": out["mIoU"], "oIoU_final": out["oIoU_final"],
                          "mIoU_final": out["mIoU_final"], "n_expressions": out["n_expressions"],
                          "identical_to_single_process": ok, "ms_sharded_incl_synthesis": t0.elapsed_time(t1)}), flush=True)
    if world > 1:
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.broadcast(flag, 0)
        ok = bool(flag.item())
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
