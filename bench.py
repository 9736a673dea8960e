#!/usr/bin/env python
"""Benchmark of the HybridGL mask-proposal scoring path on B200 (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload 2] [--images 16]

A *step* is one pass of the hot path (blur -> prep -> mask grid -> heat-map pooling -> score/select -> IoU)
over one batch of `--images` synthetic RefCOCO-shaped images per GPU (BASELINE.json configs[1]: 480x640,
100 masks/image, 3 expressions/image, ViT-B/16 geometry S=224 g=14 De=512, bf16 prep outputs).  One JSON line
on stdout (rank 0).  `value` = expressions/s with inputs resident in HBM, `e2e` = the same through
ScoringPath.run_host with pinned HOST buffers (H2D + kernels + D2H inside the timed region).
`--impl reference` times the CPU oracle port of the reference path on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "expressions/sec"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", type=int, default=2, help="BASELINE.json config index (1-based), default 2")
    ap.add_argument("--images", type=int, default=None, help="images per GPU per pass (default 16; 8 with --sweep)")
    ap.add_argument("--e2e-steps", type=int, default=30)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-backbone-view", action="store_true", help="skip the e2e_with_backbone view (path + random-init PyTorch CLIP ViT, N=1 only)")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from the host instead of replaying the step from a CUDA graph "
                    "(ScoringPath.capture); device time per pass is equal within noise, the graph removes ~0.3 ms of host work per pass")
    ap.add_argument("--inner", type=int, default=0, help="passes per timed step; 0 = sized so that the timed region lasts >= 0.6 s")
    ap.add_argument("--no-overlap", action="store_true", help="launch every stage in order on one stream (no helper streams)")
    ap.add_argument("--serial-steps", type=int, default=50, help="extra untimed-for-value pass with overlap off: each kernel timed alone")
    ap.add_argument("--rle-steps", type=int, default=200, help="extra pass with the proposals given as SAM uncompressed RLE (0 = skip)")
    ap.add_argument("--sweep", action="store_true", help="strong-scaling evaluation sweep (BASELINE.json configs[4]; use with --workload 5)")
    ap.add_argument("--sweep-images", type=int, default=4096)
    ap.add_argument("--sweep-pool", type=int, default=4, help="distinct synthetic batches the sweep cycles through")
    ap.add_argument("--prefetch", action="store_true", help="launch the next batch's frame-only chains (blur, prep setup, heat-map tables) inside the "
                    "current pass (ScoringPath.run(prefetch=...)); measured no faster on B200: the step is the SUM of the kernels' stand-alone times")
    ap.add_argument("--gem-space", default="pixel", choices=["pixel", "token"], help="score_gem from heat-map tables in pixel space (the reference's "
                    "formulation) or on the raw GEM map's token grid (hgl_gem_token_pool, no frame-sized tables); the other one is timed next to it")
    ap.add_argument("--rows-first", type=int, default=0, help="1: prep main waits for the mask pass (grid + heat-map pooling) instead of running beside it")
    ap.add_argument("--chunks", type=int, default=1, help="image groups a batch is cut into inside ScoringPath.run (stage pipelining within a pass)")
    ap.add_argument("--prep-dtype", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--features", default="tokens", choices=["tokens", "supplied"],
                    help="tokens: pool dense patch tokens under the grid masks on the tensor cores (hgl_mask_pool) and score those; "
                         "supplied: score hybrid features given as an input")
    a = ap.parse_args()
    if a.images is None:
        a.images = 8 if a.sweep else 16
    return a


def workload_config(idx: int, images: int):
    from hybridgl_b200 import synth
    c = dict(synth.CONFIGS[idx])
    c["images_per_gpu_per_step"] = images
    return c


# ------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------- CPU arm
def _oracle_one_image(args):
    """The reference path for ONE image on one core, through the numpy oracle port (oracle/hybridgl_oracle.py)."""
    seed, cfg, n_masks = args
    import numpy as np
    from hybridgl_b200 import synth
    from oracle import hybridgl_oracle as O
    it = synth.make_item(seed, cfg["h"], cfg["w"], n_masks, cfg["n_expr"], de=cfg["De"], n_other=2)
    t0 = time.perf_counter()
    blur = O.gaussian_blur_u8(it.image)
    O.prep(it.image, blur, it.masks, cfg["S"])
    grid = O.mask_to_grid(it.masks, cfg["g"], antialias=True)
    tokens = synth.bf16_round(np.random.default_rng(seed).standard_normal((cfg["g"] ** 2, cfg["De"])).astype(np.float32))
    feats = O.mask_pool_tokens(grid.reshape(grid.shape[0], -1), tokens)            # dense tokens pooled under every soft grid mask
    cum = np.zeros(4, np.int64)
    for ex in it.expressions:
        heat = O.resize_bilinear_aa(ex.heat_raw[None], cfg["h"], cfg["w"])[0]           # T.Resize((H,W), antialias=True), :201
        sg = O.gem_pool(O.condition_heatmap(heat, ex.dirflag), it.masks, O.black_for(ex.relaflag))
        r = O.score_and_select(feats, ex.sentence_feat, ex.noun_feat, ex.other_feats, it.boxes, ex.relaflag, score_gem=sg)
        i0, u0, _ = O.compute_iou(it.masks[r["idx_hybrid"]], it.target)
        i1, u1, _ = O.compute_iou(it.masks[r["idx_final"]], it.target)
        cum += np.array([i0, u0, i1, u1])
    return time.perf_counter() - t0


def _ref_available() -> bool:
    try:
        from oracle import ref_runner
        return ref_runner.ref_root() is not None
    except Exception:
        return False


def _reference_one_image(args):
    """The same image through the REAL reference code (oracle/_ref, copied by oracle/make_ref.py): Hybridgl_main.py:92-125 and
    :153-230 exec'd in place + TF.resize of the masks (model/backbone.py:160), one thread."""
    seed, cfg, n_masks = args
    from oracle import ref_runner
    return ref_runner.run_image(seed, cfg, n_masks, threads=1)["seconds"]


def _cpu_arm():
    """(worker function, kind, description) of the CPU arm: the reference's own code when oracle/_ref travelled, else the numpy port."""
    if _ref_available():
        return _reference_one_image, "reference", ("the reference's own code (oracle/_ref: Hybridgl_main.py:92-125 prep loop, model/backbone.py:160 mask "
                                                  "resize, Hybridgl_main.py:153-230 scoring / guidance / IoU per expression; CLIP features, GEM maps and "
                                                  "spaCy flags are inputs), torch CPU + cv2, one thread per worker")
    return _oracle_one_image, "port", "numpy oracle port (oracle/hybridgl_oracle.py); oracle/_ref (the reference's own files) is not present"


def cpu_baseline_sample(cfg, budget_s: float = 20.0):
    """Rank 0, N=1: the reference path on ONE host core over a bounded sample (whole images of the workload)."""
    fn, kind, what = _cpu_arm()
    n_masks = cfg["n_masks"]
    t_first = fn((9000, cfg, n_masks))
    imgs, total = 1, t_first
    while total + t_first < budget_s and imgs < 10:
        total += fn((9000 + imgs, cfg, n_masks)); imgs += 1
    return {"value": imgs * cfg["n_expr"] / total, "unit": METRIC, "cores": 1, "kind": kind,
            "sample": f"{imgs} image(s) x {n_masks} masks x {cfg['n_expr']} expressions of the same workload, {total:.1f} s on 1 core; {what}"}


def run_reference(args, cfg):
    """`--impl reference`: the reference's own CPU code for the path (oracle/_ref, copied from the reference checkout by
    oracle/make_ref.py; the numpy port of it only when that copy is absent) on all host cores: one image per worker process per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    workers = max(1, min(cores, 32))
    fn, kind, what = _cpu_arm()
    n_masks = cfg["n_masks"]
    t_img = fn((8000, cfg, n_masks))
    # keep the whole run within ~4 minutes: shrink the per-image mask count if one image per step would not fit
    budget = 240.0 / max(1, args.steps + args.warmup)
    scale = 1.0
    if t_img * 1.5 > budget:
        scale = max(0.05, budget / (t_img * 1.5))
        n_masks = max(4, int(cfg["n_masks"] * scale))
    ctx = mp.get_context("fork")
    with ctx.Pool(workers) as pool:
        for w in range(args.warmup):
            pool.map(fn, [(7000 + w * workers + i, cfg, n_masks) for i in range(workers)])
        t0 = time.perf_counter()
        for s in range(args.steps):
            pool.map(fn, [(6000 + s * workers + i, cfg, n_masks) for i in range(workers)])
        dt = time.perf_counter() - t0
    # throughput in expressions/s of the FULL workload: a step with n_masks' < n_masks does n_masks'/n_masks of the work
    eff = n_masks / cfg["n_masks"]
    value = args.steps * workers * cfg["n_expr"] * eff / dt
    sample = (f"{workers} images per step (one per worker process), {n_masks}/{cfg['n_masks']} masks per image"
              f"{' (throughput scaled by that fraction; per-mask work dominates)' if eff < 1 else ''}; {what}")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(cfg), **{k: cfg[k] for k in ("h", "w", "n_masks", "n_expr", "S", "g", "De")}},
            "cpu_baseline": {"value": value, "unit": METRIC, "cores": workers, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_name(cfg):
    vit = "ViT-L/14@336" if cfg["S"] == 336 else "ViT-B/16"
    return (f"RefCOCO-shaped synthetic batch: {cfg['images_per_gpu_per_step']} images/GPU/step, {cfg['h']}x{cfg['w']}, "
            f"{cfg['n_masks']} masks/image, {cfg['n_expr']} expressions/image, {vit} geometry (S={cfg['S']}, g={cfg['g']}, "
            f"De={cfg['De']}); features: dense patch tokens pooled per mask and scored on the tensor cores (hgl_pool_score_select) "
            f"unless --features supplied")


# ------------------------------------------------------------------------------------------------- GPU arm
def bind_to_gpu_numa_node(gpu_index: int):
    """One process per GPU: pin the process (and therefore its pinned host buffers, by first touch) to the CPU cores NVML
    reports as local to the GPU, so that the e2e H2D copies of 8 ranks do not cross the socket interconnect."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} cores local to GPU {gpu_index}"
    except Exception as e:      # no NVML / not permitted: keep the default placement
        return f"unbound ({type(e).__name__})"
    return "unbound"


def algorithmic_bytes(cfg, B, prep_bytes):
    """Per-launch algorithmic HBM bytes of each kernel (stated in DESIGN.md)."""
    H, W, N, E, S, De = cfg["h"], cfg["w"], cfg["n_masks"], cfg["n_expr"], cfg["S"], cfg["De"]
    M, ET = B * N, B * E
    return {
        "blur": 2 * B * H * W * 3,
        "pack": M * H * W + M * H * ((W + 31) // 32) * 4,
        # per-image half: frames in, 12 answer planes + 24 tap bytes per output pixel out
        "prep_setup": 2 * B * H * W * 3 + B * S * S * (12 * prep_bytes + 24),
        # per-mask half: packed masks + answer planes in, 2 x [M,3,S,S] out
        "prep": M * H * ((W + 31) // 32) * 4 + B * S * S * 12 * prep_bytes + 2 * M * 3 * S * S * prep_bytes,
        # heat-map tables (Hybridgl_main.py:201-209): raw GEM maps (28 x 37, resized on the fly) in, row-prefix tables of the
        # frame-sized conditioned maps out, written once
        "heat_tables": ET * 28 * 37 * 4 + ET * H * W * 4,
        # one pass over the packed masks (mask grid + heat-map pooling): packed masks in, grid + pooled scores out
        "grid_heat_pool": M * H * ((W + 31) // 32) * 4 + M * cfg["g"] ** 2 * 4 + ET * N * 4,
        # tensor-core pooling + scoring + selection in one kernel: soft masks f32 + tokens bf16 + text + boxes in, scores / picks out
        # (the pooled rows never reach HBM); flops: 2*M*L*De
        "pool_score": M * cfg["g"] ** 2 * 4 + B * cfg["g"] ** 2 * De * 2 + 3 * ET * De * 4 + 32 * M + 12 * ET * N,
        "score_select": M * De * 2 + 3 * ET * De * 4 + 32 * M + 12 * ET * N,
        "iou": 2 * 2 * ET * H * W,
    }


def setup_dist():
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local)       # before any pinned allocation: first-touch puts the host buffers next to the GPU
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout when the communicator is created: keep stdout for the ONE JSON line by
        # pointing fd 1 at stderr until the first collective has run
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    return world, rank, local, dev, numa


def load_peaks():
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    tf = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1400.0)))
    tf_src = ("measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if "bf16_tflops_sustained" in peaks
              else "fallback 1400 TFLOP/s sustained (B200_PROFILING.md)")
    return hbm, src, tf, tf_src


PCIE_GEN5_X16_GBS = 63.0      # nominal per direction (32 GT/s x 16 lanes, 128b/130b); ~52-55 GB/s is what pinned cudaMemcpyAsync delivers


def time_e2e(path, hosts, max_n, steps, world, dev, barrier, expr_per_step, pipelined=True):
    """e2e through the public host-buffer API.  pipelined: ScoringPath.run_host_iter (H2D of batch k+1 under the kernels of batch k,
    one CUDA graph per buffer set); else one blocking run_host call per step."""
    import torch
    import torch.distributed as dist
    seq = [hosts[s % len(hosts)] for s in range(steps)]
    if pipelined:
        for _ in path.run_host_iter(seq[:4], max_n, graph=True):
            pass
    else:
        for w in range(2):
            path.run_host(hosts[w % len(hosts)], max_n)
    barrier()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    if pipelined:
        for out in path.run_host_iter(seq, max_n, graph=True):
            pass
    else:
        for hb in seq:
            out = path.run_host(hb, max_n)
    t1.record()
    barrier()
    ems = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ems, op=dist.ReduceOp.MAX)
    sec = float(ems.item()) / 1e3
    h2d = path.h2d_bytes(hosts[0])
    link = h2d * steps / sec / 1e9
    del out
    return {"value": expr_per_step * steps / sec, "unit": METRIC, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": path.d2h_bytes(),
            "steps": steps, "ms_per_step": sec / steps * 1e3, "link_gbs_per_gpu": round(link, 2),
            "link_frac_of_pcie_gen5_x16": round(link / PCIE_GEN5_X16_GBS, 3),
            "api": ("hybridgl_b200.pipeline.ScoringPath.run_host_iter (pinned host tensors in / out; copies of step k+1 overlap the kernels of step k)"
                    if pipelined else "hybridgl_b200.pipeline.ScoringPath.run_host (pinned host tensors in / out, one blocking call per step)")}


def backbone_view(cfg, dev, images=2):
    """BASELINE.md section 3 'two views': the same path WITH the hybrid CLIP ViT (plain PyTorch, random-init weights of the named
    architecture, bf16) producing the features that are scored -- one image per call like the reference loop.  Rank 0, N=1 only."""
    import torch
    from hybridgl_b200 import ops, synth
    from hybridgl_b200.backbone import CLIPViTFM
    name = "ViT-L/14@336px" if cfg["S"] == 336 else "ViT-B/16"
    model = CLIPViTFM(name, device=dev, dtype=torch.bfloat16)
    mb = model.last_layer - 1
    t_path = t_vit = 0.0
    ev = lambda: torch.cuda.Event(enable_timing=True)      # noqa: E731
    for i in range(-1, images):                            # image -1: untimed warm-up
        b = synth.make_batch_device(300 + i, 1, cfg["h"], cfg["w"], cfg["n_masks"], cfg["n_expr"], cfg["De"], device=dev, grid=cfg["g"], raw_heat=True)
        e0, e1, e2, e3 = ev(), ev(), ev(), ev()
        e0.record()
        bits = ops.pack_masks(b["masks"])
        blur = ops.gaussian_blur15(b["image"])
        local, glob = ops.prep_visual_prompts(b["image"], blur, bits, cfg["S"], dtype=torch.bfloat16)
        e1.record()
        feats = model(local, glob, b["masks"], masking_block=mb, fusion_mode=cfg["fusion_mode"])
        e2.record()
        _, _, score_gem = ops.grid_heat_pool(bits, cfg["w"], cfg["g"], b["heat"], b["dirflag"], b["black"])
        res = ops.score_select(feats.contiguous(), b["sent"], b["noun"], b["others"], b["other_off"], b["boxes"], b["relaflag"], score_gem)
        cum = torch.zeros(4, dtype=torch.int64, device=dev)
        ops.iou_accumulate(bits, b["target"], res["idx_hybrid"], res["idx_final"], cum)
        e3.record()
        torch.cuda.synchronize()
        if i >= 0:
            t_path += e0.elapsed_time(e1) + e2.elapsed_time(e3); t_vit += e1.elapsed_time(e2)
    del model
    torch.cuda.empty_cache()
    return {"value": images * cfg["n_expr"] / ((t_path + t_vit) / 1e3), "unit": METRIC, "images": images,
            "path_ms_per_image": round(t_path / images, 3), "vit_ms_per_image": round(t_vit / images, 3),
            "backbone": f"{name} hybrid forward ({cfg['fusion_mode']}, masking_block={mb}), plain PyTorch bf16, random-init weights",
            "note": "device-resident inputs, one image per call like the reference loop; end to end the reference's loop is backbone-bound "
                    "(SURVEY.md section 8(d)); `value` / `e2e` above are the path with features / tokens as inputs, as the north star's I/O contract states"}


def run_ours(args, cfg):
    import torch
    import torch.distributed as dist

    from hybridgl_b200 import synth
    from hybridgl_b200.pipeline import ScoringPath

    world, rank, local, dev, numa = setup_dist()
    B = args.images
    prep_dtype = torch.bfloat16 if args.prep_dtype == "bf16" else torch.float32
    path = ScoringPath(size=cfg["S"], grid=cfg["g"], prep_dtype=prep_dtype, feature_source=args.features, overlap=not args.no_overlap,
                       chunks=args.chunks, rows_first=bool(args.rows_first), gem_space=args.gem_space)
    # two distinct device batches, alternated, each far larger than the 126 MB L2 (masks alone: B*N*H*W bytes)
    batches = [synth.make_batch_device(1000 + 17 * rank + i, B, cfg["h"], cfg["w"], cfg["n_masks"], cfg["n_expr"], cfg["De"], device=dev,
                                         grid=cfg["g"], raw_heat=True)
               for i in range(2)]
    max_n = cfg["n_masks"]
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for w in range(max(3, args.warmup)):
        path.run(batches[w % 2], max_n)
    # The step as a CUDA graph (one per device batch): the four-stream stage graph of ScoringPath.run is captured once and
    # replayed with ONE launch per pass; `prep` is bracketed by two event-record nodes inside the graph.
    graphs = None
    pf = args.prefetch and path.overlap and args.chunks == 1       # batch k's pass launches batch k+1's frame-only chains

    def capture_pair(stages):
        if not pf:
            return [path.capture(b, max_n, time_stages=stages) for b in batches]
        gs = [path.capture(batches[i], max_n, time_stages=stages, prefetch=batches[1 - i], frames_ready=True) for i in range(2)]
        path.prime(batches[0], max_n)                                       # the first replay (batch 0) finds its frame chains ready
        return gs
    if not args.no_graph:
        graphs = capture_pair(("prep",))
        for w in range(2 * max(2, args.warmup // 2)):                       # an even count: the next replay is batch 0 again
            graphs[w % 2].replay()
    # inner repeats: a timed "step" is `inner` passes over alternating batches, sized so that the timed region lasts >= ~0.6 s
    # whatever --steps is (a 10 ms region measures launch jitter and rank skew, not the path)
    inner = args.inner
    if inner <= 0:
        c0 = torch.cuda.Event(enable_timing=True); c1 = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        c0.record()
        for w in range(8):
            (graphs[w % 2].replay() if graphs else path.run(batches[w % 2], max_n, prefetch=batches[1 - w % 2] if pf else None))
        c1.record()
        torch.cuda.synchronize()
        t_pass = c0.elapsed_time(c1) / 8.0
        inner_t = torch.tensor([max(1, min(1024, int(600.0 / (max(args.steps, 1) * t_pass)) + 1))], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(inner_t, op=dist.ReduceOp.MAX)
        inner = int(inner_t.item())
    passes = args.steps * inner
    path.cum.zero_()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    # timed region: only the dominant stage (prep) is bracketed with events -- two records per pass instead of twenty, which
    # cost ~3 % of the step when every stage is bracketed (profiles/host_overhead.py)
    top_events = []
    top_samples = []
    path.events_only = {"prep"}
    t_start = torch.cuda.Event(enable_timing=True); t_end = torch.cuda.Event(enable_timing=True)
    t_start.record()
    if graphs is not None:
        prev = None
        for s in range(passes):
            g = graphs[s % 2]
            g.replay()
            if prev is not None and s % 8 == 1:   # a previous pass's pairs (one per image group), read while this pass runs
                prev.events[-1][2].synchronize()  # (re-stamped only by ITS next replay)
                top_samples.append(sum(e0.elapsed_time(e1) for _, e0, e1 in prev.events))
            prev = g
    else:
        for s in range(passes):
            path.events = []
            path.run(batches[s % 2], max_n, prefetch=batches[1 - s % 2] if pf else None)
            top_events.append(path.events)
    path.events = None
    path.events_only = None
    t_end.record()
    # the path's only collective, timed on its own: all-reduce of the IoU accumulators (32 bytes)
    # (an untimed all-reduce first lines the ranks up: without it k0 -> k1 on a fast rank is the wait for the slowest rank's
    # loop -- 10 ms at 8 ranks -- which the max over ranks of the loop time below already contains)
    cum = path.cum.clone()
    k0 = torch.cuda.Event(enable_timing=True); k1 = torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.all_reduce(torch.zeros(4, dtype=torch.int64, device=dev))
    k0.record()
    if world > 1:
        dist.all_reduce(cum)
    k1.record()
    barrier()
    clocks = sampler.stop() if sampler else None
    ms = torch.tensor([t_start.elapsed_time(t_end), k0.elapsed_time(k1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total, collective_ms = float(ms[0].item()), float(ms[1].item())
    expr_per_pass = world * B * cfg["n_expr"]
    value = expr_per_pass * passes / ((ms_total + collective_ms) / 1e3)        # the sweep's one collective is charged to the job

    # per-stage durations from event pairs, each pair on the stream its stage runs on
    def stage_avg(all_events):
        dur = {}
        for evs in all_events:
            d1 = {}
            for name, e0, e1 in evs:
                if name != "t0":
                    d1[name] = d1.get(name, 0.0) + e0.elapsed_time(e1)      # one launch per image group: summed per pass
            for name, v in d1.items():
                dur.setdefault(name, []).append(v)
        return {k: sum(v) / len(v) for k, v in dur.items()}
    top_ms = stage_avg(top_events) if graphs is None else {"prep": sum(top_samples) / max(1, len(top_samples))}
    # every stage bracketed, same overlapped stage graph, right after the timed region (durations UNDER the overlap: the
    # helper-stream chains run concurrently with prep).  The brackets are event-record NODES of a captured graph, so a small
    # kernel's duration is its device time, not the host's launch latency (an eager pass cannot resolve a 10 us kernel).
    ALL = ("t0", "pack", "rle", "blur", "prep_setup", "heat_tables", "grid_heat_pool", "prep", "pool_score", "score_select", "iou")
    timelines = {}

    def staged_pass(reps, tag):
        """-> {stage: mean ms}; also records timelines[tag] = {stage: [start, end]} in ms from the start of the pass."""
        dur, span = {}, {}

        def take(evs):
            # a stage is launched once per image group: its duration is the sum over the groups, its span first start .. last end
            t0 = [e for e in evs if e[0] == "t0"][0][1]
            d1, s1 = {}, {}
            for name, e0, e1 in evs:
                if name == "t0":
                    continue
                d1[name] = d1.get(name, 0.0) + e0.elapsed_time(e1)
                a, b_ = t0.elapsed_time(e0), t0.elapsed_time(e1)
                s1[name] = (min(a, s1[name][0]), max(b_, s1[name][1])) if name in s1 else (a, b_)
            for name in d1:
                dur.setdefault(name, []).append(d1[name])
                span.setdefault(name, []).append(s1[name])
        if args.no_graph:
            for s in range(reps):
                path.events = []
                path.run(batches[s % 2], max_n)
                torch.cuda.synchronize()
                take(path.events)
            path.events = None
        else:
            gs = capture_pair(ALL) if path.overlap else [path.capture(b, max_n, time_stages=ALL) for b in batches]
            for s in range(reps):
                g = gs[s % 2]
                g.replay()
                torch.cuda.synchronize()
                take(g.events)
        timelines[tag] = {k: [round(sum(a for a, _ in v) / len(v), 4), round(sum(b_ for _, b_ in v) / len(v), 4)] for k, v in span.items()}
        return {k: sum(v) / len(v) for k, v in dur.items()}

    cum_keep = path.cum.clone()
    avg_ms = staged_pass(50, "overlapped")
    avg_ms.update(top_ms)                # the dominant stage: the timed region's own measurement
    barrier()
    # the same stages launched back to back on one stream (no overlap): each kernel timed ALONE, after the timed region
    ms_serial = None
    alone_ms = {}
    if args.serial_steps > 0:
        path.overlap = False
        for w in range(3):
            path.run(batches[w % 2], max_n)
        alone_ms = staged_pass(args.serial_steps, "serial")
        sg_ = [path.capture(b, max_n) for b in batches] if not args.no_graph else None
        barrier()
        s0 = torch.cuda.Event(enable_timing=True); s1 = torch.cuda.Event(enable_timing=True)
        s0.record()
        for s in range(args.serial_steps):
            (sg_[s % 2].replay() if sg_ else path.run(batches[s % 2], max_n))
        s1.record()
        path.overlap = not args.no_overlap
        barrier()
        ms_serial = s0.elapsed_time(s1) / args.serial_steps
        del sg_
    path.cum.copy_(cum_keep)
    peak_hbm, peak_src, peak_tf, peak_tf_src = load_peaks()
    alg = algorithmic_bytes(cfg, B, 2 if prep_dtype == torch.bfloat16 else 4)
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:
        pass
    kernels = {}
    for k, b in alg.items():
        if k in avg_ms and avg_ms[k] > 0:
            ach = b / (avg_ms[k] * 1e-3) / 1e9
            kernels[k] = {"ms": round(avg_ms[k], 4), "algorithmic_bytes": b, "achieved_gbs": round(ach, 1), "frac": round(ach / peak_hbm, 4),
                          "share_of_step": round(avg_ms[k] / sum(avg_ms.values()), 4)}
            if alone_ms.get(k, 0) > 0:
                a1 = b / (alone_ms[k] * 1e-3) / 1e9
                kernels[k].update({"ms_alone": round(alone_ms[k], 4), "achieved_gbs_alone": round(a1, 1), "frac_alone": round(a1 / peak_hbm, 4)})
    if "pool_score" in kernels:
        # the tensor-core stage: algorithmic flops 2*M*L*De against the measured bf16 peak AND against its own ceiling
        # min(tensor peak, arithmetic intensity x HBM peak) -- at AI ~ 60 flop/B the HBM-limited ceiling is ~1/4 of the tensor peak
        flop = 2.0 * B * cfg["n_masks"] * cfg["g"] ** 2 * cfg["De"]
        ai = flop / alg["pool_score"]
        ceil_tf = min(peak_tf, ai * peak_hbm / 1e3)
        kk = kernels["pool_score"]
        kk.update({"algorithmic_flop": flop, "arithmetic_intensity_flop_per_byte": round(ai, 1), "tensor_peak_tflops": peak_tf,
                   "tensor_peak_source": peak_tf_src, "ceiling_tflops_min_tensor_ai_hbm": round(ceil_tf, 1)})
        for tag in ("", "_alone"):
            if "ms" + tag in kk:
                tf = flop / (kk["ms" + tag] * 1e-3) / 1e12
                kk.update({"achieved_tflops" + tag: round(tf, 2), "frac_tensor" + tag: round(tf / peak_tf, 5),
                           "frac_of_ceiling" + tag: round(tf / ceil_tf, 4)})
    top = "prep" if "prep" in kernels else max(kernels, key=lambda k: kernels[k]["ms"])      # the bandwidth-bound bulk of the step
    roofline = {"kernel": f"hgl_{top}", "bound": "hbm", "achieved": kernels[top]["achieved_gbs"], "peak": peak_hbm, "unit": "GB/s",
                "frac": kernels[top]["frac"], "traffic": traffic.get(top), "peak_source": peak_src,
                "timing": ("CUDA events around the stage on its own stream inside the timed region (event-record nodes of the replayed graph; "
                           "the other stages are bracketed in a second pass right after it: kernels{})" +
                           ("; the helper-stream chains (pack, mask grid + heat-map pooling, pooling + scoring, IoU; blur, prep setup; heat-map "
                            "tables) run concurrently with prep, so `frac` is prep's share of HBM while sharing it; `frac_alone` is the same "
                            "kernel timed alone in the serial pass" if path.overlap else "")),
                "frac_alone": kernels[top].get("frac_alone"), "achieved_alone": kernels[top].get("achieved_gbs_alone")}

    # ---- e2e: the public API with pinned HOST buffers (byte masks: what the reference's call surface receives)
    e2e = e2e_serial = None
    hosts = None
    if args.e2e_steps > 0:
        hosts = [{k: v.cpu().pin_memory() for k, v in b.items()} for b in batches]
        e2e = time_e2e(path, hosts, max_n, args.e2e_steps, world, dev, barrier, expr_per_pass, pipelined=True)
        e2e_serial = time_e2e(path, hosts, max_n, max(4, args.e2e_steps // 3), world, dev, barrier, expr_per_pass, pipelined=False)
        e2e["bound"] = (f"host link: {e2e['h2d_bytes_per_step'] / 1e6:.0f} MB of H2D per step ({B * cfg['n_masks'] * cfg['h'] * cfg['w'] / 1e6:.0f} MB of it "
                        "byte masks) against ~0.5 ms of kernels")
        e2e["serial_run_host"] = {k: e2e_serial[k] for k in ("value", "ms_per_step", "link_gbs_per_gpu")}

    # ---- the same workload with the proposals handed over as SAM's uncompressed RLE (amg.py:107-135) instead of byte masks:
    # hgl_rle_to_bits replaces hgl_pack_masks, the H*W-byte masks exist neither on the host nor in HBM.  Reported NEXT TO the
    # byte-mask numbers above (which stay the headline: byte masks are what the reference's call surface receives).
    rle_info = None
    if args.rle_steps > 0:
        rb = []
        for b in batches:
            c_, o_ = synth.masks_to_rle_device(b["masks"])
            d = {k: v for k, v in b.items() if k != "masks"}
            d["rle_counts"], d["rle_off"] = c_, o_
            rb.append(d)
        cum_bytes = path.cum.clone()
        for w in range(3):
            path.run(rb[w % 2], max_n)
        rgraphs = None
        if graphs is not None:
            rgraphs = [path.capture(b, max_n, time_stages=("rle",)) for b in rb]
            for w in range(3):
                rgraphs[w % 2].replay()
        barrier()
        r0 = torch.cuda.Event(enable_timing=True); r1 = torch.cuda.Event(enable_timing=True)
        rle_events = []
        r0.record()
        for s in range(args.rle_steps):
            if rgraphs is not None:
                rgraphs[s % 2].replay()
            else:
                path.events = []
                path.run(rb[s % 2], max_n)
                rle_events.append(path.events)
        r1.record()
        path.events = None
        barrier()
        if rgraphs is not None:            # the decode stage of the last two replays (event-record nodes inside the graphs)
            rle_events = [g.events for g in rgraphs]
        rms = torch.tensor([r0.elapsed_time(r1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(rms, op=dist.ReduceOp.MAX)
        rle_ms = float(rms.item()) / args.rle_steps
        rle_stage = stage_avg(rle_events).get("rle", 0.0)
        rle_bytes = rb[0]["rle_counts"].numel() * 4 + B * cfg["n_masks"] * cfg["h"] * ((cfg["w"] + 31) // 32) * 4
        rle_info = {"value": expr_per_pass / (rle_ms / 1e3), "unit": METRIC, "ms_per_step": rle_ms, "steps": args.rle_steps,
                    "runs_per_mask": rb[0]["rle_counts"].numel() / max(1, B * cfg["n_masks"]),
                    "rle_to_bits": {"ms": round(rle_stage, 4), "algorithmic_bytes": rle_bytes,
                                    "achieved_gbs": round(rle_bytes / max(rle_stage, 1e-9) / 1e6, 1),
                                    "frac": round(rle_bytes / max(rle_stage, 1e-9) / 1e6 / peak_hbm, 4)},
                    "note": "proposals as SAM uncompressed RLE (SamAutomaticMaskGenerator(output_mode='uncompressed_rle')); "
                            "results bit-identical to the byte-mask run (tests/test_gpu_parity.py::test_bench_workload_parity)"}
        del rgraphs
        if args.e2e_steps > 0:
            rhost = [{k: v.cpu().pin_memory() for k, v in b.items()} for b in rb]
            rle_info["e2e"] = time_e2e(path, rhost, max_n, args.e2e_steps * 10, world, dev, barrier, expr_per_pass, pipelined=True)
            rs = time_e2e(path, rhost, max_n, args.e2e_steps * 3, world, dev, barrier, expr_per_pass, pipelined=False)
            rle_info["e2e"]["serial_run_host"] = {k: rs[k] for k in ("value", "ms_per_step", "link_gbs_per_gpu")}
            del rhost
        path.cum.copy_(cum_bytes)

    # ---- the other formulation of the GEM pooling, timed next to the one the headline uses
    other = "token" if args.gem_space == "pixel" else "pixel"
    alt_path = ScoringPath(size=cfg["S"], grid=cfg["g"], prep_dtype=prep_dtype, feature_source=args.features, overlap=not args.no_overlap,
                           gem_space=other)
    alt_info = None
    try:
        ag = [alt_path.capture(b, max_n, time_stages=("grid_heat_pool", "heat_tables")) for b in batches]
        for w in range(4):
            ag[w % 2].replay()
        barrier()
        a0 = torch.cuda.Event(enable_timing=True); a1 = torch.cuda.Event(enable_timing=True)
        a0.record()
        for s in range(100):
            ag[s % 2].replay()
        a1.record()
        barrier()
        st_ms = {}
        for g_ in ag:
            for name, e0, e1 in g_.events:
                st_ms[name] = st_ms.get(name, 0.0) + e0.elapsed_time(e1) / len(ag)
        alt_info = {"gem_space": other, "ms_per_pass": a0.elapsed_time(a1) / 100, "value": expr_per_pass / (a0.elapsed_time(a1) / 100 / 1e3) / world * world,
                    "stage_ms_in_step": {k: round(v, 4) for k, v in st_ms.items()},
                    "note": "same pass with score_gem computed by the other formulation; results agree to 1e-3 "
                            "(tests/test_gpu_parity.py::test_pipeline_token_space_gem_equals_pixel_space_picks)"}
        del ag
    except Exception as ex:
        alt_info = {"gem_space": other, "unavailable": f"{type(ex).__name__}: {ex}"}
    del alt_path

    with_backbone = None
    if rank == 0 and world == 1 and not args.no_backbone_view:
        try:
            with_backbone = backbone_view(cfg, dev)
        except Exception as e:      # a reported view, never the headline: do not lose the line over it
            with_backbone = {"unavailable": f"{type(e).__name__}: {e}"}

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_baseline_sample(cfg)
        c = [int(v) for v in cum.tolist()]
        ms_per_step = (ms_total + collective_ms) / args.steps
        line = {"metric": METRIC, "value": value, "unit": METRIC, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16" if prep_dtype == torch.bfloat16 else "f32", "data": "synthetic",
                "config": {"workload": workload_name(cfg), **{k: cfg[k] for k in ("h", "w", "n_masks", "n_expr", "S", "g", "De")}},
                "image_groups_per_pass": args.chunks, "frame_chain_prefetch": bool(pf), "inner_repeats": inner, "passes_timed": passes, "ms_per_pass": ms_total / passes, "timed_region_s": ms_total / 1e3,
                "collective_ms": collective_ms,
                "collective": "one all-reduce(SUM) of the int64[4] IoU accumulators after the last pass (NCCL), timed on its own and charged to the job",
                "l2": f"two alternating batches; the byte masks alone are {B * cfg['n_masks'] * cfg['h'] * cfg['w'] / 1e6:.0f} MB per batch (> 126 MB L2)",
                "clocks": clocks, "host_affinity": numa, "e2e": e2e, "e2e_with_backbone": with_backbone,
                "gpu_launches": path.launches_per_run() * passes,
                "launch": ("one CUDA-graph replay per pass (ScoringPath.capture)" if graphs is not None else "eager: one host launch per kernel"),
                "streams": ("4 (prep on the caller's stream; pack -> mask pass -> pooling+scoring -> IoU, blur -> prep setup, heat-map tables on "
                            "high-priority helper streams)" if path.overlap else "1"),
                "ms_per_pass_serial": ms_serial,
                "gem_space": args.gem_space, "gem_other_formulation": alt_info,
                "roofline": roofline, "kernels": kernels, "timeline_ms": timelines, "rle_input": rle_info, "cpu_baseline": cpu,
                "iou": {"cum_I": c[0], "cum_U": c[1], "cum_I_final": c[2], "cum_U_final": c[3],
                        "oIoU": c[0] * 100.0 / max(c[1], 1), "oIoU_final": c[2] * 100.0 / max(c[3], 1)}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_sweep_mode(args, cfg):
    """`--sweep`: BASELINE.json configs[4] -- a PhraseCut-shaped evaluation sweep of `--sweep-images` images sharded by image over the
    ranks (STRONG scaling: the total is fixed), hybridgl_b200/sweep.py::run_sweep_batches + reduce_counters (one all-reduce of the
    int64[4] accumulators and one all-gather of the per-expression IU rows at the end; Hybridgl_main.py:52-55, 240-247).
    The dataset is `pool` distinct synthetic batches of `--images` images visited round-robin (180 GB of HBM cannot hold 4096 frames
    of byte masks; the generator is not part of the path): batch j of the sweep = pool[j % pool], owned by rank j % world."""
    import torch
    import torch.distributed as dist

    from hybridgl_b200 import sweep, synth
    from hybridgl_b200.pipeline import ScoringPath
    world, rank, local, dev, numa = setup_dist()
    B = args.images
    n_batches = max(1, args.sweep_images // B)
    pool = [synth.make_batch_device(7000 + i, B, cfg["h"], cfg["w"], cfg["n_masks"], cfg["n_expr"], cfg["De"], device=dev, grid=cfg["g"], raw_heat=True)
            for i in range(args.sweep_pool)]
    max_n = cfg["n_masks"]
    path = ScoringPath(size=cfg["S"], grid=cfg["g"], prep_dtype=torch.bfloat16, feature_source=args.features, chunks=args.chunks)
    steps = [path.capture(b, max_n) for b in pool]          # one CUDA graph per pool batch
    E_b = B * cfg["n_expr"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_sweep():
        return sweep.run_sweep_batches(n_batches, lambda j: steps[j % len(steps)].replay(), path, E_b)

    for w in range(max(1, min(args.warmup, 2))):
        one_sweep()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for s in range(args.steps):
        out = one_sweep()
    t1.record()
    barrier()
    clocks = sampler.stop() if sampler else None
    ms = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    sec = float(ms.item()) / 1e3
    n_expr = n_batches * E_b
    if rank == 0:
        line = {"metric": METRIC, "value": n_expr * args.steps / sec, "unit": METRIC, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": sec / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16",
                "data": "synthetic",
                "config": {"workload": f"PhraseCut-shaped eval sweep: {n_batches * B} images {cfg['h']}x{cfg['w']}, {cfg['n_masks']} masks and "
                                       f"{cfg['n_expr']} expressions per image, sharded by image over the ranks ({B} images per launch); a step = one whole sweep",
                           **{k: cfg[k] for k in ("h", "w", "n_masks", "n_expr", "S", "g", "De")}},
                "sweep": {"images": n_batches * B, "expressions": n_expr, "pool_batches": len(pool), "cum": out["cum"].tolist(),
                          "oIoU": out["oIoU"], "mIoU": out["mIoU"], "oIoU_final": out["oIoU_final"], "mIoU_final": out["mIoU_final"],
                          "collective": "all_reduce(int64[4]) + all_gather of the per-expression IU rows, once per sweep (inside the timed region)"},
                "clocks": clocks, "host_affinity": numa, "gpu_launches": path.launches_per_run() * ((n_batches + world - 1) // world) * args.steps,
                "e2e": None, "roofline": None, "cpu_baseline": None}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    cfg = workload_config(args.workload, args.images)
    if args.impl == "reference":
        run_reference(args, cfg)
    elif args.sweep:
        run_sweep_mode(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
