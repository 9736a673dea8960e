"""Image-sharded evaluation sweep (BASELINE.json configs[4]): one process per GPU, no data-path collective.

The reference is single-process (Hybridgl_main.py:79 loops over the DataLoader) and keeps four integer accumulators and
two per-expression IoU lists (Hybridgl_main.py:52-55, 171, 230) which it reports as oIoU = cum_I*100/cum_U and
mIoU = mean(list)*100 (Hybridgl_main.py:240-247).  Images are independent, so rank r of `world` owns the images
i = r, r+world, r+2*world, ... and the only exchange is, once per sweep:

    all_reduce(SUM) of cum int64[4]                  (order-independent => bit-exact at any world size)
    all_gather of the per-expression iu int64[E_r,4] rows + their global expression ids

The reference's mIoU is an fp32 mean over a Python list, which depends on the visiting order; here it is computed from
the gathered INTEGER (I, U) pairs re-ordered by global expression id, so every world size reports the same number.
NCCL is the backend on GPUs; the same code runs under gloo on CPU tensors (tests/test_sweep_gloo.py).
"""
from __future__ import annotations

from typing import Callable, Dict, List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_indices(n_items: int, rank: int, world: int) -> List[int]:
    """Strided image assignment (balances a variable number of masks per image better than contiguous blocks)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    return list(range(rank, n_items, world))


def _world(group=None) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def reduce_counters(cum: torch.Tensor, iu_rows: torch.Tensor, expr_ids: torch.Tensor, group=None) -> Dict[str, object]:
    """Combine the per-rank results of a sweep.

    cum      int64 [4]      this rank's (cum_I, cum_U, cum_I_final, cum_U_final)
    iu_rows  int64 [E_r,4]  (I_hybrid, U_hybrid, I_final, U_final) of every expression this rank scored
    expr_ids int64 [E_r]    global, dataset-order id of those expressions
    Returns the four numbers of the reference's result log plus the merged integer tables (identical on every rank)."""
    rank, world = _world(group)
    total = cum.clone()
    rows, ids = iu_rows, expr_ids
    if world > 1:
        dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)           # the path's only reduction: 32 bytes
        n_local = torch.tensor([iu_rows.shape[0]], dtype=torch.int64, device=cum.device)
        counts = [torch.zeros_like(n_local) for _ in range(world)]
        dist.all_gather(counts, n_local, group=group)
        n_max = int(max(int(c) for c in counts))
        pad = torch.zeros((n_max, 5), dtype=torch.int64, device=cum.device)
        pad[: iu_rows.shape[0], :4] = iu_rows
        pad[: iu_rows.shape[0], 4] = expr_ids
        gathered = [torch.zeros_like(pad) for _ in range(world)]
        dist.all_gather(gathered, pad, group=group)
        allrows = torch.cat([g[: int(c)] for g, c in zip(gathered, counts)], dim=0)
        rows, ids = allrows[:, :4], allrows[:, 4]
    order = torch.argsort(ids, stable=True)
    rows = rows[order]
    ids = ids[order]
    return dict(cum=total, iu=rows, expr_ids=ids, **report(total, rows))


def report(cum: torch.Tensor, iu_rows: torch.Tensor) -> Dict[str, float]:
    """oIoU / mIoU for both picks from integer tables (Hybridgl_main.py:240-247; this_iou = 0 when U == 0, utils.py:373-376)."""
    c = [int(v) for v in cum.tolist()]
    r = iu_rows.to("cpu", torch.float64)

    def mean_iou(i_col, u_col):
        if r.shape[0] == 0:
            return float("nan")
        u = r[:, u_col]
        iou = torch.where(u > 0, r[:, i_col] / torch.clamp(u, min=1), torch.zeros_like(u))
        return float(iou.mean()) * 100.0

    return dict(oIoU=c[0] * 100.0 / c[1] if c[1] else float("nan"), mIoU=mean_iou(0, 1),
                oIoU_final=c[2] * 100.0 / c[3] if c[3] else float("nan"), mIoU_final=mean_iou(2, 3),
                n_expressions=int(r.shape[0]))


def run_sweep(n_images: int, make_batch: Callable[[Sequence[int]], Tuple[Dict[str, torch.Tensor], int, torch.Tensor]],
              path, images_per_step: int = 16, group=None) -> Dict[str, object]:
    """Drive `path` (a pipeline.ScoringPath) over this rank's shard of `n_images` images.

    make_batch(image_ids) -> (device batch dict for those images, max_n, global expression ids int64 [E_batch]).
    Returns reduce_counters(...) of the whole sweep (same on every rank)."""
    rank, world = _world(group)
    mine = shard_indices(n_images, rank, world)
    path.cum.zero_()
    rows, ids = [], []
    for s in range(0, len(mine), images_per_step):
        batch, max_n, expr_ids = make_batch(mine[s: s + images_per_step])
        res = path.run(batch, max_n)
        rows.append(res["iu"].clone())
        ids.append(expr_ids.to(res["iu"].device))
    dev = path.cum.device
    iu = torch.cat(rows) if rows else torch.zeros((0, 4), dtype=torch.int64, device=dev)
    eid = torch.cat(ids) if ids else torch.zeros((0,), dtype=torch.int64, device=dev)
    return reduce_counters(path.cum, iu, eid, group)


def run_sweep_batches(n_batches: int, run_batch: Callable[[int], Dict[str, torch.Tensor]], path, exprs_per_batch: int,
                      group=None) -> Dict[str, object]:
    """The sweep over a dataset that is already cut into batches of images (bench.py --sweep): batch j belongs to rank
    j % world; run_batch(j) runs the path on it (ScoringPath.run or a captured GraphStep.replay) and returns its result dict,
    whose `iu` rows are the batch's `exprs_per_batch` expressions in dataset order.  Same reduction as run_sweep."""
    rank, world = _world(group)
    mine = shard_indices(n_batches, rank, world)
    path.cum.zero_()
    dev = path.cum.device
    rows, ids = [], []
    local_ids = torch.arange(exprs_per_batch, dtype=torch.int64, device=dev)
    for j in mine:
        res = run_batch(j)
        rows.append(res["iu"].clone())          # a replayed graph overwrites its result tensors
        ids.append(local_ids + j * exprs_per_batch)
    iu = torch.cat(rows) if rows else torch.zeros((0, 4), dtype=torch.int64, device=dev)
    eid = torch.cat(ids) if ids else torch.zeros((0,), dtype=torch.int64, device=dev)
    return reduce_counters(path.cum, iu, eid, group)
