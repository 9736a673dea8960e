"""World-size-2 (and 3) gloo test of the sweep's host logic: image sharding + the IoU all-reduce / gather.
CPU only: the per-image integer (I, U) rows come from the oracle (the checker), the code under test is
hybridgl_b200/sweep.py::shard_indices / reduce_counters / report."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hybridgl_b200 import sweep, synth
from oracle import hybridgl_oracle as O

N_IMAGES = 7


def _rows_for_image(i):
    """Oracle I/U rows of image i (2 expressions) and their global expression ids."""
    it = synth.make_item(900 + i, 48, 64, 5 + (i % 3), 2, de=64)
    rows = []
    for ex in it.expressions:
        sg = O.gem_pool(O.condition_heatmap(ex.heatmap, ex.dirflag), it.masks, O.black_for(ex.relaflag))
        r = O.score_and_select(it.features, ex.sentence_feat, ex.noun_feat, ex.other_feats, it.boxes, ex.relaflag, score_gem=sg)
        i0, u0, _ = O.compute_iou(it.masks[r["idx_hybrid"]], it.target)
        i1, u1, _ = O.compute_iou(it.masks[r["idx_final"]], it.target)
        rows.append([i0, u0, i1, u1])
    return np.array(rows, np.int64), np.array([2 * i, 2 * i + 1], np.int64)


def _single_process():
    rows = np.concatenate([_rows_for_image(i)[0] for i in range(N_IMAGES)])
    return rows.sum(0), rows


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = sweep.shard_indices(N_IMAGES, rank, world)
        parts = [_rows_for_image(i) for i in mine]
        rows = np.concatenate([p[0] for p in parts]) if parts else np.zeros((0, 4), np.int64)
        ids = np.concatenate([p[1] for p in parts]) if parts else np.zeros((0,), np.int64)
        out = sweep.reduce_counters(torch.from_numpy(rows.sum(0) if len(rows) else np.zeros(4, np.int64)),
                                    torch.from_numpy(rows), torch.from_numpy(ids))
        q.put((rank, out["cum"].tolist(), out["iu"].tolist(), out["expr_ids"].tolist(),
               out["oIoU"], out["mIoU"], out["oIoU_final"], out["mIoU_final"]))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_sweep_matches_single_process(world):
    cum_ref, rows_ref = _single_process()
    ref = sweep.report(torch.from_numpy(cum_ref), torch.from_numpy(rows_ref))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, cum, iu, ids, o, m, of, mf in results:
        assert cum == cum_ref.tolist()                      # integer accumulators: bit-exact at any world size
        assert ids == list(range(2 * N_IMAGES))             # dataset order restored
        assert iu == rows_ref.tolist()
        assert (o, m, of, mf) == (ref["oIoU"], ref["mIoU"], ref["oIoU_final"], ref["mIoU_final"])   # identical, not just close


def test_shard_indices_cover_and_partition():
    for n in (0, 1, 7, 16):
        for world in (1, 2, 3, 8):
            parts = [sweep.shard_indices(n, r, world) for r in range(world)]
            assert sorted(i for p in parts for i in p) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        sweep.shard_indices(4, 2, 2)


def test_report_handles_empty_union():
    rep = sweep.report(torch.tensor([5, 10, 0, 0]), torch.tensor([[5, 10, 0, 0]]))
    assert rep["oIoU"] == 50.0 and rep["mIoU"] == 50.0 and np.isnan(rep["oIoU_final"]) and rep["mIoU_final"] == 0.0


class _StubPath:
    """Stands in for pipeline.ScoringPath on the CPU: the sweep driver only touches `.cum`."""
    def __init__(self):
        self.cum = torch.zeros(4, dtype=torch.int64)


def _batch_rows(j):
    """Deterministic integer IU rows of 'batch' j (3 expressions), as ScoringPath.run would return them."""
    rng = np.random.default_rng(40 + j)
    u = rng.integers(1, 1000, (3, 2))
    i = (u * rng.random((3, 2))).astype(np.int64)
    return np.stack([i[:, 0], u[:, 0], i[:, 1], u[:, 1]], 1).astype(np.int64)


def _worker_batches(rank, world, port, q, n_batches):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        path = _StubPath()

        def run_batch(j):
            rows = torch.from_numpy(_batch_rows(j))
            path.cum += rows.sum(0)                      # what hgl_iou does on the device
            return {"iu": rows}
        out = sweep.run_sweep_batches(n_batches, run_batch, path, 3)
        q.put((rank, out["cum"].tolist(), out["iu"].tolist(), out["expr_ids"].tolist(), out["mIoU"], out["mIoU_final"]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_batch_sharded_sweep_matches_single_process(world):
    """bench.py --sweep (strong scaling): batches owned round-robin by the ranks, one all-reduce + all-gather at the end."""
    n_batches = 5
    rows_ref = np.concatenate([_batch_rows(j) for j in range(n_batches)])
    ref = sweep.report(torch.from_numpy(rows_ref.sum(0)), torch.from_numpy(rows_ref))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_batches, args=(r, world, port, q, n_batches)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, cum, iu, ids, m, mf in results:
        assert cum == rows_ref.sum(0).tolist() and iu == rows_ref.tolist() and ids == list(range(3 * n_batches))
        assert (m, mf) == (ref["mIoU"], ref["mIoU_final"])
