"""The mask-proposal scoring path as one batched call (the public API of this package).

Mirrors, for a batch of images at once, what one iteration of the reference's evaluation loop does between
"SAM has produced masks + boxes" (Hybridgl_main.py:86-90) and "IoU counters updated" (Hybridgl_main.py:230):

    blur -> prep (local/global CLIP inputs) -> mask grid (+area) -> [hybrid CLIP features: plain PyTorch backbone,
    or supplied by the caller] -> GEM heat-map pooling -> cosine scoring + spatial guidance + argmax -> IoU

Every arithmetic step runs in libhgl.so (hand-written sm_100a kernels behind the C ABI); this file only owns
buffers, streams and the order of launches.  Host language stays Python like the reference.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import ops

# host->device inputs of one step and their dtypes (what a caller must provide per batch)
INPUT_KEYS = ("image", "masks", "boxes", "target", "features", "tokens", "sent", "noun", "others", "other_off", "heat",
              "dirflag", "relaflag", "black", "mask_off", "expr_off")
# device->host results of one step
OUTPUT_KEYS = ("score_clip", "score_gem", "idx_hybrid", "idx_final", "top_idx", "blended", "iu")


class ScoringPath:
    """Batched scoring path.  `size` = CLIP input side (224 | 336), `grid` = patch grid side (14 | 24)."""

    def __init__(self, size: int = 224, grid: int = 14, prep_dtype: torch.dtype = torch.bfloat16, antialias: bool = True,
                 background: str = "blur", logit_scale_exp: float = 100.0, r: float = 0.5, alpha: float = 0.6,
                 feature_source: str = "supplied", device: Optional[torch.device] = None, overlap: bool = True,
                 keep_features: bool = False, chunks: int = 1, rows_first: bool = False, gem_space: str = "pixel"):
        """feature_source: "supplied" -> batch["features"] [M,De] (the hybrid CLIP features of CLIPViTFM.forward) are scored;
        "tokens" -> batch["tokens"] [B,L,De] (dense patch tokens, third_party/modified_CLIP/clip/model.py:302-307) are pooled
        under every proposal's soft grid mask on the tensor cores and scored in the same kernel (hgl_pool_score_select);
        keep_features: also write the pooled, normalised rows [M,De] (bf16) to HBM and return them as res["features"].
        chunks: groups of images a batch is cut into inside run() (see there); 1 = every stage once per batch (the default:
        measured on B200 at the bench shape, 2 / 4 groups are SLOWER -- 0.55 / 0.66 ms against 0.50 ms per pass -- because the
        small latency-bound kernels a group's prep waits for crawl while another group's pack saturates HBM).
        gem_space: "pixel" = heat-map tables + one fused pass over the masks (hgl_heat_tables + hgl_grid_heat_pool_rows); "token" =
        score_gem computed on the raw GEM map's token grid (hgl_gem_token_pool; needs raw maps) next to hgl_mask_grid."""
        if feature_source not in ("supplied", "tokens"):
            raise ValueError(feature_source)
        self.feature_source = feature_source
        self.keep_features = keep_features
        self.chunks = max(1, int(chunks))
        if gem_space not in ("pixel", "token"):
            raise ValueError(gem_space)
        self.gem_space = gem_space           # "token": score_gem from the raw GEM maps in token space (hgl_gem_token_pool): no frame-sized tables
        self.rows_first = bool(rows_first)   # prep main waits for the mask pass (both then run at their stand-alone speed, one after the other)
        self._plans: Dict[tuple, list] = {}
        self._ahead = None                  # frame chains a previous call prefetched: dict(key, slot, ev_setup, ev_tables)
        self._frame_slots: Dict[tuple, int] = {}
        ops.device_ok()
        self.size, self.grid = size, grid
        self.prep_dtype, self.antialias, self.background = prep_dtype, antialias, background
        self.logit_scale_exp, self.r, self.alpha = logit_scale_exp, r, alpha
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.cum = torch.zeros(4, dtype=torch.int64, device=self.device)     # cum_I, cum_U, cum_I_final, cum_U_final
        self._buf: Dict[str, torch.Tensor] = {}
        self.events = None            # optional per-stage CUDA event pairs (bench.py): [(stage, start, end)]
        self.events_only = None       # optional set of stage names to bracket (None = every stage)
        self.overlap = overlap        # fork the post-pack chain onto a high-priority side stream (see run())
        self._side: Optional[torch.cuda.Stream] = None
        self._pre: Optional[torch.cuda.Stream] = None
        self._tab: Optional[torch.cuda.Stream] = None
        self._capturing = False
        self._checked = set()         # (mask_off ptr, version, max_n) triples whose per-image mask counts were validated
        self._copy: Optional[torch.cuda.Stream] = None       # H2D stream of run_host_iter
        self._slots: Dict[int, Dict[str, torch.Tensor]] = {}  # run_host / run_host_iter: device input + pinned output buffer sets

    # ------------------------------------------------------------------------------------------------
    def _get(self, name: str, shape, dtype) -> torch.Tensor:
        t = self._buf.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = torch.empty(tuple(shape), dtype=dtype, device=self.device)
            self._buf[name] = t
        return t

    def _check_max_n(self, mask_off: torch.Tensor, max_n: int, host_off: Optional[torch.Tensor] = None) -> None:
        """max_n is the row stride of every [E, max_n] result and the per-image bound of every kernel: an image with more masks
        would silently lose its tail (the kernels clamp), so reject it here.  Checked once per offsets tensor (one small
        device->host read the first time a device-resident offsets tensor is seen; free when the host copy is at hand)."""
        key = (mask_off.data_ptr(), mask_off._version, int(max_n))
        if key in self._checked or self._capturing:
            return
        src = host_off if host_off is not None else mask_off
        worst = int((src[1:] - src[:-1]).max()) if src.numel() > 1 else 0
        if worst > max_n:
            raise ValueError(f"max_n={max_n} but an image of this batch has {worst} masks")
        if len(self._checked) > 64:
            self._checked.clear()
        self._checked.add(key)

    class _Span:
        """Brackets one stage with CUDA events on the stream it is launched on (bench.py reads them)."""
        def __init__(self, path, name):
            self.path, self.name = path, name

        def _on(self):
            p = self.path
            return p.events is not None and (p.events_only is None or self.name in p.events_only)

        def __enter__(self):
            if self._on():
                # inside a graph capture the pair becomes two event-record NODES (external events), re-stamped by every replay
                self.e0 = torch.cuda.Event(enable_timing=True, external=self.path._capturing)
                self.e0.record()

        def __exit__(self, *exc):
            if self._on():
                e1 = torch.cuda.Event(enable_timing=True, external=self.path._capturing)
                e1.record()
                self.path.events.append((self.name, self.e0, e1))
            return False

    def _span(self, name: str):
        return ScoringPath._Span(self, name)

    # ---- chunking ------------------------------------------------------------------------------------------------------
    def _chunk_plan(self, batch, chunks: int, host_off=None):
        """Cut a batch into `chunks` groups of whole images.  Returns a list of dicts with the image / mask / expression ranges and
        re-based offset tensors.  Needs the offsets on the host: free when the caller has them (run_host), one small
        device->host read the first time a device-resident batch is seen otherwise (cached per offsets tensor)."""
        moff, eoff = batch["mask_off"], batch["expr_off"]
        B = moff.numel() - 1
        chunks = max(1, min(chunks, B))
        key = (moff.data_ptr(), moff._version, eoff.data_ptr(), eoff._version, chunks)
        plan = self._plans.get(key)
        if plan is not None:
            return plan
        if self._capturing:
            raise RuntimeError("ScoringPath: a batch must run eagerly once before it is captured (capture() does)")
        mo = (host_off[0] if host_off is not None else moff).tolist()
        eo = (host_off[1] if host_off is not None else eoff).tolist()
        bounds = sorted({(c * B) // chunks for c in range(chunks + 1)})
        plan = []
        for b0, b1 in zip(bounds[:-1], bounds[1:]):
            m0, m1, e0, e1 = mo[b0], mo[b1], eo[b0], eo[b1]
            whole = b0 == 0 and b1 == B
            plan.append(dict(b=(b0, b1), m=(m0, m1), e=(e0, e1),
                             mask_off=moff if whole else (moff[b0:b1 + 1] - m0).contiguous(),
                             expr_off=eoff if whole else (eoff[b0:b1 + 1] - e0).contiguous()))
        if len(self._plans) > 32:
            self._plans.clear()
        self._plans[key] = plan
        return plan

    @staticmethod
    def _chunk_view(batch, ch, rle: bool):
        """The tensors of one chunk: views (no copies) of the batch's tensors; `others` / `rle_counts` stay whole because
        other_off / rle_off index them absolutely."""
        (b0, b1), (m0, m1), (e0, e1) = ch["b"], ch["m"], ch["e"]
        v = dict(image=batch["image"][b0:b1], target=batch["target"][b0:b1], boxes=batch["boxes"][m0:m1],
                 sent=batch["sent"][e0:e1], noun=batch["noun"][e0:e1], others=batch["others"], other_off=batch["other_off"][e0:e1 + 1],
                 heat=batch["heat"][e0:e1], dirflag=batch["dirflag"][e0:e1], relaflag=batch["relaflag"][e0:e1], black=batch["black"][e0:e1],
                 mask_off=ch["mask_off"], expr_off=ch["expr_off"])
        if rle:
            v["rle_counts"], v["rle_off"] = batch["rle_counts"], batch["rle_off"][m0:m1 + 1]
        else:
            v["masks"] = batch["masks"][m0:m1]
        if "tokens" in batch:
            v["tokens"] = batch["tokens"][b0:b1]
        if "features" in batch:
            v["features"] = batch["features"][m0:m1]
        return v

    @staticmethod
    def _frame_key(batch):
        """Identity of the inputs of the frame-only chains (frames, heat-maps, direction flags)."""
        return tuple((batch[k].data_ptr(), batch[k]._version, tuple(batch[k].shape)) for k in ("image", "heat", "dirflag"))

    def _frame_chain(self, views, plan, slot: int, max_n: int, H: int, W: int, raw: bool, split: bool, launch: bool):
        """Buffers of one set (`slot`) of frame-chain outputs and, if `launch`, the chains themselves: blur -> prep setup on the
        `pre` stream, heat-map tables on the `tab` stream, one launch per image group."""
        lib = ops._lib.load()
        S, g = self.size, self.grid
        pre = self._pre if self.overlap else torch.cuda.current_stream()
        tab = self._tab if self.overlap else torch.cuda.current_stream()
        n_ch = len(plan)
        B_all = plan[-1]["b"][1]
        blur_all = self._get(f"blur@{slot}", (B_all, H, W, 3), torch.uint8) if self.background == "blur" else None
        pws, hws = [], []
        ev_setup, ev_tables = [None] * n_ch, [None] * n_ch
        for c, ch in enumerate(plan):
            (b0, b1), (m0, m1), (e0, e1) = ch["b"], ch["m"], ch["e"]
            pws.append(self._get(f"prep_ws{c}@{slot}", (max(lib.hgl_prep_workspace_bytes(b1 - b0, S, ops._dt(self.prep_dtype)), 1),), torch.uint8))
            hshape = views[c]["heat"].shape
            need = (lib.hgl_grid_heat_pool_raw_workspace_bytes(b1 - b0, m1 - m0, e1 - e0, H, W, g, max_n, hshape[1], hshape[2]) if raw
                    else lib.hgl_grid_heat_pool_workspace_bytes(b1 - b0, m1 - m0, e1 - e0, H, W, g, max_n))
            hws.append(self._get(f"heat_ws{c}@{slot}", (need,), torch.uint8))
        if launch:
            def mark():
                if not self.overlap:
                    return None
                ev = torch.cuda.Event()
                ev.record()
                return ev
            with torch.cuda.stream(pre):
                for c, (ch, v) in enumerate(zip(plan, views)):
                    b0, b1 = ch["b"]
                    with self._span("blur"):
                        blur = ops.gaussian_blur15(v["image"], out=blur_all[b0:b1]) if self.background == "blur" else None
                    with self._span("prep_setup"):
                        ops.prep_setup(v["image"], blur, S, background=self.background, dtype=self.prep_dtype, workspace=pws[c])
                    ev_setup[c] = mark()
            if split:
                with torch.cuda.stream(tab):
                    for c, v in enumerate(views):
                        if v["heat"].shape[0] == 0:
                            continue
                        with self._span("heat_tables"):
                            ops.heat_tables(v["heat"], v["dirflag"], H, W, hws[c])
                        ev_tables[c] = mark()
        return dict(pws=pws, hws=hws, ev_setup=ev_setup, ev_tables=ev_tables)

    def _make_streams(self):
        """Helper streams above the caller's (prep main) in priority, so that their CTAs are dispatched first whenever SM
        resources free up: heat-map tables highest (they gate the mask pass, need no masks and are short), then the chain
        pack -> mask pass -> pooling + scoring -> IoU, then blur -> prep setup.  Measured on B200 at the bench shape
        (profiles/prio_sweep.py, three repeats each): 0.449 ms per pass against 0.487 ms with all three at the same priority --
        the mask pass then runs right behind the pack, before prep main starts, instead of sharing the SMs with it for its whole
        duration.  The host's enqueue order of the three root chains makes no difference (0.450 - 0.453 ms for all four orders)."""
        lo, hi = torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, "priority_range") else (0, -1)
        lvl = lambda k: max(hi, -k)      # noqa: E731  (priorities are negative numbers: hi is the most urgent the device offers)
        self._tab = torch.cuda.Stream(device=self.device, priority=lvl(3))
        self._side = torch.cuda.Stream(device=self.device, priority=lvl(2))
        self._pre = torch.cuda.Stream(device=self.device, priority=lvl(1))

    def prime(self, batch: Dict[str, torch.Tensor], max_n: int) -> None:
        """Run only the frame-only chains of `batch` (what a previous call's `prefetch=batch` would have done): the first pass of a
        prefetching loop -- or the first replay of graphs captured with frames_ready=True -- then finds them ready."""
        img = batch["image"]
        B, H, W, _ = img.shape
        E = batch["sent"].shape[0]
        M = batch["rle_off"].numel() - 1 if "rle_counts" in batch else batch["masks"].shape[0]
        raw = tuple(batch["heat"].shape[1:]) != (H, W)
        if self.overlap and self._side is None:
            self._make_streams()
        main = torch.cuda.current_stream()
        if self.overlap:
            for s_ in (self._pre, self._tab):
                s_.wait_stream(main)
        key = self._frame_key(batch)
        slot = self._frame_slots.setdefault(key, 0)
        fr = self._frame_chain([dict(image=img, heat=batch["heat"], dirflag=batch["dirflag"])], [dict(b=(0, B), m=(0, M), e=(0, E))], slot, max_n, H, W,
                               raw, self.antialias and E > 0 and M > 0, launch=True)
        self._ahead = dict(key=key, slot=slot, ev_setup=fr["ev_setup"], ev_tables=fr["ev_tables"])
        if self.overlap:
            main.wait_stream(self._pre)
            main.wait_stream(self._tab)

    def run(self, batch: Dict[str, torch.Tensor], max_n: int, features: Optional[torch.Tensor] = None, host_off=None,
            prefetch: Optional[Dict[str, torch.Tensor]] = None, frames_ready: bool = False) -> Dict[str, torch.Tensor]:
        """One pass over a device-resident batch.  Returns device tensors (see OUTPUT_KEYS) plus the prep
        outputs `local_imgs`, `global_imgs` [M,3,S,S], `grid` [M,g,g] and `area` [M].

        Stage graph (self.overlap=True), four streams forked from / joined to the caller's stream; the batch is cut into
        self.chunks groups of images (A, B, ...) and every stage is launched once per group, stage by stage:
            side  pack A, pack B | mask grid + heat-map pooling A, B | pooling + scoring A, B | IoU A, B
            pre   blur A, prep setup A, blur B, prep setup B
            tab   heat-map tables A, B
            main  (caller's stream)  prep main A (after pack A, setup A), prep main B (after pack B, setup B)
        Only pack and prep main are bandwidth-bound; everything else is small and latency-bound and runs up to 3x slower while HBM
        is saturated (profiles/r2_timeline.md).  With overlap=False every stage is launched in order on the caller's stream.

        prefetch: the batch the NEXT call will process.  Its frame-only chains (blur -> prep setup, heat-map tables: they need the
        frames and the heat-maps, not the masks) are launched now, behind this pass's pack, into a second set of buffers, and run in
        the shadow of this pass's prep writes; the next call finds them done and starts its prep the moment its pack ends -- the
        HBM-idle gap between pack and prep (blur -> setup crawling beside the pack) disappears from the steady state.
        frames_ready: (graph capture only) the frame chains of `batch` were produced by the previous replay's prefetch; skip them."""
        img = batch["image"]
        B, H, W, _ = img.shape
        rle = "rle_counts" in batch          # proposals as SAM uncompressed RLE instead of byte masks
        M = batch["rle_off"].numel() - 1 if rle else batch["masks"].shape[0]
        E = batch["sent"].shape[0]
        moff = batch["mask_off"]
        lib = ops._lib.load()
        self._check_max_n(moff, max_n, None if host_off is None else host_off[0])
        plan = self._chunk_plan(batch, self.chunks if self.overlap else 1, host_off)
        views = [self._chunk_view(batch, ch, rle) for ch in plan]
        main = torch.cuda.current_stream()
        side = pre = tab = main
        if self.overlap:
            if self._side is None:
                self._make_streams()
            side, pre, tab = self._side, self._pre, self._tab
            for s_ in (side, pre, tab):
                s_.wait_stream(main)
        if self.events is not None and (self.events_only is None or "t0" in self.events_only):
            t0 = torch.cuda.Event(enable_timing=True, external=self._capturing)      # origin of a stage timeline (bench.py)
            t0.record(main)
            self.events.append(("t0", t0, t0))

        def mark():
            if not self.overlap:
                return None
            ev = torch.cuda.Event()
            ev.record()
            return ev

        S, g = self.size, self.grid
        WW = (W + 31) // 32
        bits = self._get("bits", (M, H, WW), torch.int32)
        local = self._get("local", (M, 3, S, S), self.prep_dtype)
        glob = self._get("global", (M, 3, S, S), self.prep_dtype)
        heat = batch["heat"]           # frame-sized [E,H,W], or the raw GEM map [E,h,w] (resized like Hybridgl_main.py:201 on the fly)
        raw = tuple(heat.shape[1:]) != (H, W)
        gem_token = self.gem_space == "token" and raw and self.antialias and E > 0 and M > 0
        split = self.antialias and E > 0 and M > 0 and not gem_token
        use_tokens = self.feature_source == "tokens" and features is None
        feats_in = features if features is not None else batch.get("features")
        n_ch = len(plan)

        # ---- chain S (side), first stage: the one pass that produces the packed masks (from byte masks, or from SAM's RLE)
        ev_pack = [None] * n_ch
        with torch.cuda.stream(side):
            for c, (ch, v) in enumerate(zip(plan, views)):
                m0, m1 = ch["m"]
                if rle:
                    with self._span("rle"):
                        ops.rle_to_bits(v["rle_counts"], v["rle_off"], H, W, out=bits[m0:m1])
                else:
                    with self._span("pack"):
                        ops.pack_masks(v["masks"], out=bits[m0:m1])
                ev_pack[c] = mark()

        # ---- chains F (frames only: blur -> per-image half of prep) and T (heat-maps only: the table half of the pooling pass).
        #      Either launched now, beside the mask pack, or found ready because the previous call prefetched them.
        fkey = self._frame_key(batch)
        if len(self._frame_slots) > 64:
            self._frame_slots.clear()
        ahead = self._ahead if (self._ahead is not None and self._ahead["key"] == fkey and n_ch == 1) else None
        if frames_ready and self._capturing and n_ch == 1:
            ahead = dict(key=fkey, slot=self._frame_slots.setdefault(fkey, 0), ev_setup=[None], ev_tables=[None])
        if ahead is not None:
            fslot = ahead["slot"]
            fr = self._frame_chain(views, plan, fslot, max_n, H, W, raw, split, launch=False)
            fr["ev_setup"], fr["ev_tables"] = ahead["ev_setup"], ahead["ev_tables"]
        else:
            fslot = self._frame_slots.setdefault(fkey, 0) if n_ch == 1 else 0
            fr = self._frame_chain(views, plan, fslot, max_n, H, W, raw, split, launch=True)
        pws, hws, ev_setup, ev_tables = fr["pws"], fr["hws"], fr["ev_setup"], fr["ev_tables"]
        self._ahead = None

        # ---- chain S continued: everything that only needs the packed masks
        with torch.cuda.stream(side):
            grid = torch.empty((M, g, g), dtype=torch.float32, device=self.device)
            area = torch.empty((M,), dtype=torch.int32, device=self.device)
            score_gem = torch.empty((E, max_n), dtype=torch.float32, device=self.device)
            res = dict(score_clip=torch.empty((E, max_n), dtype=torch.float32, device=self.device),
                       idx_hybrid=torch.empty((E,), dtype=torch.int64, device=self.device),
                       idx_final=torch.empty((E,), dtype=torch.int64, device=self.device),
                       top_idx=torch.empty((E, 3), dtype=torch.int32, device=self.device),
                       blended=torch.empty((E, 3), dtype=torch.float32, device=self.device))
            iu = torch.empty((E, 4), dtype=torch.int64, device=self.device)
            feats = feats_in
            if use_tokens and self.keep_features:
                feats = torch.empty((M, batch["tokens"].shape[2]), dtype=torch.bfloat16, device=self.device)
            for c, (ch, v) in enumerate(zip(plan, views)):
                (m0, m1), (e0, e1) = ch["m"], ch["e"]
                if ev_tables[c] is not None:
                    side.wait_event(ev_tables[c])
                cb, ch_moff, ch_eoff = bits[m0:m1], v["mask_off"], v["expr_off"]
                with self._span("grid_heat_pool"):
                    if m1 == m0:
                        pass
                    elif gem_token:            # token-space GEM pooling: mask grid (antialiased) + adjoint-resampled masks . raw maps
                        ops.masks_to_grid(cb, g, antialias=True, want_area=True, width=W, out=(grid[m0:m1], area[m0:m1]))
                        if e1 > e0:
                            ops.gem_token_pool(cb, W, v["heat"], v["dirflag"], v["black"], ch_moff, ch_eoff, max_n,
                                               workspace=self._get(f"gem_ws{c}", (max(lib.hgl_gem_token_workspace_bytes(
                                                   ch["b"][1] - ch["b"][0], m1 - m0, e1 - e0, H, W, heat.shape[1], heat.shape[2], max_n), 1),), torch.uint8),
                                               out=score_gem[e0:e1])
                    elif split and e1 > e0:    # mask grid + heat-map pooling share one pass over the packed masks
                        ops.grid_heat_pool_rows(cb, W, g, v["heat"].shape, v["black"], ch_moff, ch_eoff, max_n, hws[c],
                                                out=(grid[m0:m1], area[m0:m1], score_gem[e0:e1]))
                    elif self.antialias:
                        g_, a_, s_ = ops.grid_heat_pool(cb, W, g, v["heat"], v["dirflag"], v["black"], ch_moff, ch_eoff, max_n, workspace=hws[c])
                        grid[m0:m1].copy_(g_); area[m0:m1].copy_(a_); score_gem[e0:e1].copy_(s_)
                    else:
                        hv = v["heat"]
                        if raw:
                            hv = ops.heat_resize_aa(hv, H, W, out=self._get(f"heat_full{c}", (e1 - e0, H, W), torch.float32))
                        g_, a_ = ops.masks_to_grid(cb, g, antialias=False, want_area=True, width=W)
                        grid[m0:m1].copy_(g_); area[m0:m1].copy_(a_)
                        score_gem[e0:e1].copy_(ops.heat_pool(hv, v["dirflag"], v["black"], cb, ch_moff, ch_eoff, max_n, workspace=hws[c]))

            ev_rows = mark() if self.rows_first else None
        if ev_rows is not None:
            main.wait_event(ev_rows)

        if prefetch is not None and self.overlap and n_ch == 1:
            pkey = self._frame_key(prefetch)
            pslot = 1 - fslot
            self._frame_slots[pkey] = pslot
            pB = prefetch["image"].shape[0]
            pplan = [dict(b=(0, pB), m=(0, 0), e=(0, prefetch["sent"].shape[0]))]
            pview = [dict(image=prefetch["image"], heat=prefetch["heat"], dirflag=prefetch["dirflag"])]
            pM = prefetch["rle_off"].numel() - 1 if "rle_counts" in prefetch else prefetch["masks"].shape[0]
            pplan[0]["m"] = (0, pM)
            praw = tuple(prefetch["heat"].shape[1:]) != tuple(prefetch["image"].shape[1:3])
            psplit = self.antialias and pplan[0]["e"][1] > 0 and pM > 0
            for s_ in (pre, tab):                       # start in the shadow of THIS pass's prep: not beside its pack, nor its mask pass
                s_.wait_event(ev_rows if ev_rows is not None else ev_pack[-1])
            pf = self._frame_chain(pview, pplan, pslot, max_n, prefetch["image"].shape[1], prefetch["image"].shape[2], praw, psplit, launch=True)
            self._ahead = dict(key=pkey, slot=pslot, ev_setup=pf["ev_setup"], ev_tables=pf["ev_tables"])

        # ---- chain P (caller's stream): the per-mask half of prep, the bandwidth-bound bulk of the step
        for c, (ch, v) in enumerate(zip(plan, views)):
            (b0, b1), (m0, m1) = ch["b"], ch["m"]
            if ev_pack[c] is not None:
                main.wait_event(ev_pack[c])
            if ev_setup[c] is not None:
                main.wait_event(ev_setup[c])
            if m1 == m0:
                continue
            with self._span("prep"):
                ops.prep_main(bits[m0:m1], (b1 - b0, H, W), S, pws[c], mask_off=v["mask_off"], max_n=max_n, dtype=self.prep_dtype,
                              out=(local[m0:m1], glob[m0:m1]))

        # ---- chain S, last part: pooling + scoring + selection -> IoU (small kernels, in the shadow of the prep writes)
        with torch.cuda.stream(side):
            for c, (ch, v) in enumerate(zip(plan, views)):
                (m0, m1), (e0, e1) = ch["m"], ch["e"]
                if e1 == e0:
                    continue
                sub = {k: t[e0:e1] for k, t in res.items()}
                if use_tokens:
                    # pooling (tcgen05) + cosine scoring + selection tail in ONE launch; the pooled rows stay on the SM unless asked for
                    if self.keep_features:
                        sub["features"] = feats[m0:m1]
                    with self._span("pool_score"):
                        ops.pool_score_select(grid[m0:m1], v["tokens"], v["sent"], v["noun"], v["others"], v["other_off"], v["boxes"],
                                              v["relaflag"], score_gem[e0:e1], v["mask_off"], v["expr_off"], max_n, self.logit_scale_exp,
                                              self.r, self.alpha, want_features=self.keep_features, dtype=torch.bfloat16, out=sub)
                else:
                    with self._span("score_select"):
                        ops.score_select(feats_in[m0:m1], v["sent"], v["noun"], v["others"], v["other_off"], v["boxes"], v["relaflag"],
                                         score_gem[e0:e1], v["mask_off"], v["expr_off"], max_n, self.logit_scale_exp, self.r, self.alpha, out=sub)
                with self._span("iou"):
                    ops.iou_accumulate(bits[m0:m1] if rle else v["masks"], v["target"], sub["idx_hybrid"], sub["idx_final"], self.cum,
                                       v["mask_off"], v["expr_off"], out=iu[e0:e1])
        if self.overlap:
            main.wait_stream(side)
            main.wait_stream(tab)          # (already ordered before the mask pass; keeps the join explicit when split is off)
            main.wait_stream(pre)
        res.update(score_gem=score_gem, iu=iu, local_imgs=local, global_imgs=glob, grid=grid, area=area, bits=bits,
                   features=feats if (not use_tokens or self.keep_features) else None)
        return res

    def capture(self, batch: Dict[str, torch.Tensor], max_n: int, time_stages=None, prefetch=None, frames_ready: bool = False) -> "GraphStep":
        """One step as a CUDA graph: the whole stage graph of run() (four streams, ~12 kernels, memsets, fork / join events) is
        captured once for THESE device buffers and replayed with a single launch; the per-launch host work of run() (~0.3 ms
        of ctypes / torch calls per step, profiles/host_overhead.py) disappears from the step.  `batch` must stay alive and keep
        its addresses (run_host's device buffers do; a sweep captures one graph per ring buffer).  time_stages: stage names
        to bracket with event-record nodes (GraphStep.events: re-stamped by every replay).
        Returns a GraphStep; GraphStep.replay() enqueues the step on the current stream and returns the (static) result dict."""
        saved = (self.events, self.events_only)
        self.events, self.events_only = None, None
        cum0 = self.cum.clone()
        self.run(batch, max_n, prefetch=prefetch)    # eager once: sizes every workspace, creates the streams, sets kernel attributes
        torch.cuda.current_stream().synchronize()
        self.cum.copy_(cum0)                   # the warm-up step must not count
        if frames_ready:                       # the capture below skips this batch's own frame chains: make sure the buffers hold them
            self.prime(batch, max_n)
            torch.cuda.current_stream().synchronize()
        graph = torch.cuda.CUDAGraph()
        events = []
        if time_stages:
            self.events, self.events_only = events, set(time_stages)
        self._capturing = True
        try:
            with torch.cuda.graph(graph):
                res = self.run(batch, max_n, prefetch=prefetch, frames_ready=frames_ready)
        finally:
            self._capturing = False
            self.events, self.events_only = saved
            self._ahead = None          # (a prefetch recorded inside the capture has not run, and its events belong to the graph)
        return GraphStep(graph, res, events, batch)

    # launches of OUR kernels per run(): blur 1, pack 1, prep 2, grid_heat_pool 3 (prefix, consts, rows), scoring 1 (hgl_score_select, or
    # hgl_pool_score_select with feature_source="tokens"), iou 2
    LAUNCHES_PER_RUN = 10

    def launches_per_run(self) -> int:
        return self.LAUNCHES_PER_RUN * self.chunks        # every stage once per image group

    def input_keys(self, host_batch=None):
        skip = {"features" if self.feature_source == "tokens" else "tokens"}
        keys = INPUT_KEYS
        if host_batch is not None and "rle_counts" in host_batch:       # RLE proposals replace the byte masks
            skip.add("masks")
            keys = keys + ("rle_counts", "rle_off")
        return tuple(k for k in keys if k not in skip)

    # ---- host-buffer API ---------------------------------------------------------------------------------------------
    def _slot_buffers(self, slot: int, host_batch: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        """Device input buffers of one slot, (re)allocated when a shape changes (which also drops the slot's captured graph)."""
        st = self._slots.setdefault(slot, {"in": {}, "out": {}, "graph": None})
        for k in self.input_keys(host_batch):
            src = host_batch[k]
            dst = st["in"].get(k)
            if dst is None or dst.shape != src.shape or dst.dtype != src.dtype:
                st["in"][k] = torch.empty(src.shape, dtype=src.dtype, device=self.device)
                st["graph"] = None
        for k in [k for k in st["in"] if k not in self.input_keys(host_batch)]:
            del st["in"][k]
            st["graph"] = None
        return st

    def _h2d(self, st, host_batch) -> None:
        for k, dst in st["in"].items():
            dst.copy_(host_batch[k], non_blocking=True)

    def _d2h(self, st, res) -> Dict[str, torch.Tensor]:
        out = {}
        for k in OUTPUT_KEYS:
            src = res[k]
            dst = st["out"].get(k)
            if dst is None or dst.shape != src.shape or dst.dtype != src.dtype:
                dst = torch.empty(src.shape, dtype=src.dtype).pin_memory()
                st["out"][k] = dst
            dst.copy_(src, non_blocking=True)
            out[k] = dst
        return out

    def run_host(self, host_batch: Dict[str, torch.Tensor], max_n: int) -> Dict[str, torch.Tensor]:
        """End-to-end call with HOST buffers (pinned): H2D of every input, the kernels, D2H of the results, then a stream
        synchronise.  One batch at a time, nothing overlaps; run_host_iter() is the pipelined form of the same call."""
        st = self._slot_buffers(0, host_batch)
        self._check_max_n(st["in"]["mask_off"], max_n, host_batch["mask_off"])
        self._h2d(st, host_batch)
        out = self._d2h(st, self.run(st["in"], max_n, host_off=(host_batch["mask_off"], host_batch["expr_off"])))
        torch.cuda.current_stream().synchronize()
        return out

    def run_host_iter(self, host_batches, max_n: int, depth: int = 2, graph: bool = False):
        """Generator over an iterable of pinned host batches; yields the pinned host result dict of every batch, in order.

        Software pipeline of `depth` buffer sets: the H2D copies of batch k+1 run on a copy stream under the kernels of batch k,
        whose (small) D2H copies are enqueued right behind its kernels and awaited only while batch k+1 is already running.
        A yielded dict is overwritten `depth` batches later.  graph=True replays the step from one CUDA graph per buffer set
        (ScoringPath.capture) -- the per-step host work drops from ~25 launches to one; tensor shapes must then repeat."""
        main = torch.cuda.current_stream()
        if self._copy is None:
            self._copy = torch.cuda.Stream(device=self.device)
        copy = self._copy
        done: Dict[int, torch.cuda.Event] = {}

        def stage_in(k, hb):
            st = self._slot_buffers(k % depth, hb)
            self._check_max_n(st["in"]["mask_off"], max_n, hb["mask_off"])
            if (k % depth) in done:
                copy.wait_event(done[k % depth])      # the step that last read this buffer set has finished
            else:
                copy.wait_stream(main)
            with torch.cuda.stream(copy):
                self._h2d(st, hb)
                ev = torch.cuda.Event()
                ev.record()
            return st, ev, hb

        it = iter(host_batches)
        hb = next(it, None)
        if hb is None:
            return
        nxt = stage_in(0, hb)
        pending = []
        k = 0
        while nxt is not None:
            st, ev_in, hb_cur = nxt
            hb = next(it, None)
            nxt = stage_in(k + 1, hb) if hb is not None else None      # queue the next batch's copies before this batch's kernels
            main.wait_event(ev_in)
            offs = (hb_cur["mask_off"], hb_cur["expr_off"])
            if graph:
                # a captured step has the image-group boundaries (pointers, sizes) baked in: it serves batches with the SAME offsets
                same = st["graph"] is not None and all(torch.equal(a, b_) for a, b_ in zip(st["graph_off"], offs))
                if not same:
                    main.synchronize()                                  # capture() runs the step once eagerly on these inputs
                    st["graph"] = self.capture(st["in"], max_n)
                    st["graph_off"] = tuple(t.clone() for t in offs)
                res = st["graph"].replay()
            else:
                res = self.run(st["in"], max_n, host_off=offs)
            out = self._d2h(st, res)
            ev_out = torch.cuda.Event()
            ev_out.record()
            done[k % depth] = ev_out
            pending.append((ev_out, out))
            if len(pending) >= depth:
                ev, o = pending.pop(0)
                ev.synchronize()
                yield o
            k += 1
        for ev, o in pending:
            ev.synchronize()
            yield o

    def h2d_bytes(self, host_batch: Dict[str, torch.Tensor]) -> int:
        return sum(host_batch[k].numel() * host_batch[k].element_size() for k in self.input_keys(host_batch))

    def d2h_bytes(self) -> int:
        st = self._slots.get(0)
        return sum(t.numel() * t.element_size() for t in st["out"].values()) if st else 0


class GraphStep:
    """A captured step (ScoringPath.capture).  `res` are static device tensors, overwritten by every replay."""

    def __init__(self, graph, res, events, batch):
        self.graph, self.res, self.events, self._batch = graph, res, events, batch

    def replay(self) -> Dict[str, torch.Tensor]:
        self.graph.replay()
        return self.res


def report(cum: torch.Tensor, ious_hybrid, ious_final):
    """The four numbers the reference appends to result_log (Hybridgl_main.py:240-247)."""
    c = [int(v) for v in cum.tolist()]
    o = c[0] * 100.0 / c[1] if c[1] else float("nan")
    of = c[2] * 100.0 / c[3] if c[3] else float("nan")
    m = float(torch.tensor(ious_hybrid, dtype=torch.float32).mean()) * 100.0 if len(ious_hybrid) else float("nan")
    mf = float(torch.tensor(ious_final, dtype=torch.float32).mean()) * 100.0 if len(ious_final) else float("nan")
    return dict(oIoU=o, mIoU=m, oIoU_final=of, mIoU_final=mf)
