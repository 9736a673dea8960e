#!/usr/bin/env python
"""Generate golden vectors by running the REFERENCE ITSELF (in place, from /root/reference).

Run in the build container only:   python tests/golden/gen_golden.py
The reference has no function boundary around its prep / scoring blocks (they are inline in
``main()``, Hybridgl_main.py:92-125 and :153-230), so this script reads those source lines from the
reference checkout at run time, dedents them and ``exec``s them in a namespace that holds seeded
synthetic tensors plus stubs for the producers that are unavailable offline (spaCy, GEM, encode_text).
No reference source is copied into this repository; only the numeric inputs/outputs are stored
(``tests/golden/*.npz``).  ``CLIPViTFM.forward`` / ``calculate_score`` / ``make_attn_mask`` /
``relation_boxes`` / ``gen_dir_mask`` / ``Compute_IoU`` are called directly.

The generated fixtures travel to the GPU box; /root/reference does not.
"""
import os
import sys
import textwrap
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("HGL_REFERENCE_ROOT", "/root/reference")

import torch  # noqa: E402

from hybridgl_b200 import synth  # noqa: E402


def load_reference():
    """SURVEY.md Appendix C import recipe."""
    for n in ("ftfy", "spacy", "matplotlib", "matplotlib.pyplot", "matplotlib.gridspec"):
        sys.modules.setdefault(n, types.ModuleType(n))
    sys.modules["ftfy"].fix_text = lambda s: s
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].gridspec = sys.modules["matplotlib.gridspec"]
    sys.path[:0] = [os.path.join(REF, "third_party/modified_CLIP"), REF]
    import clip  # noqa: F401
    from clip.model import CLIP
    from model.backbone import CLIPViTFM
    import utils as ref_utils
    return CLIP, CLIPViTFM, ref_utils


def ref_lines(path, first, last, must_contain):
    """Source lines [first,last] (1-based, inclusive) of a reference file, dedented; guarded by anchors."""
    with open(os.path.join(REF, path)) as f:
        lines = f.readlines()
    block = lines[first - 1:last]
    text = "".join(block)
    for off, needle in must_contain:
        assert needle in lines[off - 1], (path, off, needle, lines[off - 1])
    return textwrap.dedent(text)


def make_model(CLIP, CLIPViTFM, seed, embed=32, res=64, layers=4, width=64, patch=16, last_layer=2, heads=1):
    torch.manual_seed(seed)
    m = CLIPViTFM.__new__(CLIPViTFM)
    torch.nn.Module.__init__(m)
    m.last_layer, m.num_heads = last_layer, heads
    m.model = CLIP(embed, res, layers, width, patch, 77, 64, 32, 1, 1).eval()
    # LayerNorm / bias defaults are 1/0; perturb so the test exercises them
    with torch.no_grad():
        for name, p in m.model.visual.named_parameters():
            if p.ndim == 1:
                p.add_(0.05 * torch.randn_like(p))
    return m


# ---------------------------------------------------------------------------------------------
def gen_prep(ref_utils):
    import cv2
    import torchvision.transforms as T
    code = ref_lines("Hybridgl_main.py", 92, 125,
                     [(93, "pixel_mean"), (99, "GaussianBlur"), (116, "T.Resize"), (125, "local_imgs")])
    out = {}
    cases = [("a", 0, 96, 128, 32, 6), ("b", 1, 75, 101, 48, 5), ("c", 2, 224, 224, 224, 2), ("d", 3, 480, 640, 224, 3)]
    for tag, seed, h, w, S, n in cases:
        it = synth.make_item(seed, h, w, n, 0, with_features=False)
        sam = torch.from_numpy(it.image)[None]
        img_norm = T.Normalize([0.485, 0.456, 0.406], [0.229, 0.224, 0.225])(T.ToTensor()(it.image))[None]
        ns = dict(torch=torch, cv2=cv2, np=np, T=T, Height=S, Width=S, device="cpu",
                  image={"sam_img": sam, "image": img_norm}, original_img=img_norm,
                  masks=torch.from_numpy(it.masks), boxes=torch.from_numpy(it.boxes))
        exec(compile(code, "Hybridgl_main.py:92-125", "exec"), ns)
        loc = ns["local_imgs"].numpy(); glo = ns["global_imgs"].numpy()
        out[f"{tag}_meta"] = np.array([seed, h, w, S, n], np.int64)
        out[f"{tag}_blur"] = ns["blurred"]
        if loc.nbytes <= 400_000:
            out[f"{tag}_image"] = it.image
            out[f"{tag}_masks"] = np.packbits(it.masks, axis=-1)
            out[f"{tag}_local"] = loc; out[f"{tag}_global"] = glo
        else:  # big case: inputs are re-derived from the seed; keep a strided sample + float64 sums
            out[f"{tag}_image_sum"] = np.array([it.image.astype(np.int64).sum(), it.masks.sum()], np.int64)
            out[f"{tag}_blur"] = ns["blurred"][::8, ::8].copy()
            out[f"{tag}_local"] = loc[:, :, ::7, ::5].copy(); out[f"{tag}_global"] = glo[:, :, ::7, ::5].copy()
            out[f"{tag}_sums"] = np.array([loc.astype(np.float64).sum(), glo.astype(np.float64).sum()])
        print("prep", tag, loc.shape, float(np.abs(loc).mean()), float(np.abs(glo).mean()))
    np.savez_compressed(os.path.join(HERE, "prep.npz"), **out)


def gen_grid(CLIPViTFM):
    import torchvision.transforms.functional as TF
    out = {}
    m = CLIPViTFM.__new__(CLIPViTFM); torch.nn.Module.__init__(m); m.num_heads = 3
    m.model = types.SimpleNamespace(visual=types.SimpleNamespace(conv1=types.SimpleNamespace(weight=torch.zeros(1))))
    cases = [("a", 10, 480, 640, 14, 6), ("b", 11, 480, 640, 24, 4), ("c", 12, 333, 500, 14, 5),
             ("d", 13, 97, 131, 7, 4), ("e", 14, 600, 800, 14, 3), ("f", 15, 20, 24, 14, 3)]
    for tag, seed, h, w, g, n in cases:
        rng = np.random.default_rng(seed)
        masks = synth.make_masks(rng, n, h, w)
        if tag == "a":
            masks[0] = False; masks[1] = True      # empty and full-frame proposals
        t = torch.from_numpy(masks).type(torch.float32)
        aa = TF.resize(t, (g, g)).numpy()                       # model/backbone.py:160 as executed under torchvision 0.26
        na = TF.resize(t, (g, g), antialias=False).numpy()      # what the pinned torchvision 0.15.2 computes
        out[f"{tag}_meta"] = np.array([seed, h, w, g, n], np.int64)
        out[f"{tag}_masks"] = np.packbits(masks, axis=-1)
        out[f"{tag}_aa"] = aa; out[f"{tag}_noaa"] = na
        if tag in ("a", "d"):
            am = m.make_attn_mask(torch.from_numpy(aa)).numpy()  # model/backbone.py:108-115
            out[f"{tag}_attn_row0"] = am[:, 0, :].copy()
            assert not am[:, 1:, :].any()
        print("grid", tag, aa.shape, (aa != 0).mean(), (na != 0).mean())
    np.savez_compressed(os.path.join(HERE, "grid.npz"), **out)


def gen_forward(CLIP, CLIPViTFM):
    out = {}
    model = make_model(CLIP, CLIPViTFM, seed=1234)
    sd = {k: v.numpy() for k, v in model.model.visual.state_dict().items()}
    for k, v in sd.items():
        out["w/" + k] = v
    out["logit_scale"] = model.model.logit_scale.detach().numpy()
    n, h, w, S = 5, 60, 90, 64
    rng = np.random.default_rng(77)
    masks = synth.make_masks(rng, n, h, w, min_area=200)
    loc = rng.standard_normal((n, 3, S, S)).astype(np.float32)
    glo = rng.standard_normal((n, 3, S, S)).astype(np.float32)
    out["masks"] = masks; out["local"] = loc; out["global"] = glo
    with torch.no_grad():
        for mode in ("G2L", "L2G", "G2L&L2G", "token_masking", "attn_masking", "crop"):
            y = model(torch.from_numpy(loc), torch.from_numpy(glo), torch.from_numpy(masks),
                      masking_block=1, fusion_mode=mode)
            out["out/" + mode] = y.numpy()
            print("forward", mode, tuple(y.shape), float(y.abs().mean()))
        txt = torch.from_numpy(rng.standard_normal((2, 32)).astype(np.float32))
        out["score_text"] = txt.numpy()
        out["score"] = model.calculate_score(torch.from_numpy(out["out/G2L"]), txt).numpy()
    np.savez_compressed(os.path.join(HERE, "forward.npz"), **out)


def gen_scoring(CLIPViTFM, ref_utils):
    import torchvision.transforms as T
    code = ref_lines("Hybridgl_main.py", 153, 230,
                     [(153, "text_ensemble"), (166, "score_clip_Neg"), (183, "maxNegidxs"), (196, "softmax0(topscores)"),
                      (200, "gem_model"), (209, "imgattn.mean()"), (227, "max_index_final"), (230, "Compute_IoU")])
    Model = CLIPViTFM.__new__(CLIPViTFM); torch.nn.Module.__init__(Model)
    out = {}
    cases = []
    k = 0
    for rela in synth.RELAFLAGS:
        for n_other in (0, 2):
            cases.append((100 + k, 48, 64, 12, rela, synth.DIRFLAGS[k % 4], n_other, 100.0)); k += 1
    for d in synth.DIRFLAGS:
        cases.append((100 + k, 40, 57, 9, "none", d, 1, 14.285714)); k += 1
    cases.append((100 + k, 48, 64, 2, "left", "left", 3, 100.0)); k += 1     # N < k1 (clamping, Hybridgl_main.py:178-181)
    cases.append((100 + k, 48, 64, 5, "within", "middle", 4, 100.0)); k += 1  # k1 < N < k2
    cases.append((100 + k, 480, 640, 64, "big", "right", 2, 100.0)); k += 1   # config-1 shape
    for ci, (seed, h, w, n, rela, dirf, n_other, ls) in enumerate(cases):
        it = synth.make_item(seed, h, w, n, 1, de=512, n_other=n_other, dirflag=dirf, relaflag=rela)
        ex = it.expressions[0]
        rng = np.random.default_rng(seed + 5000)
        raw = rng.random((1, max(h // 16, 2), max(w // 16, 2)), dtype=np.float32)
        rec = {}

        class Resize:  # records the output of the reference's T.Resize(..., antialias=True) call
            def __init__(self, size, antialias=None):
                self.op = T.Resize(size, antialias=antialias)

            def __call__(self, x):
                y = self.op(x); rec["resized"] = y[0].numpy().copy(); return y

        def calculate_score(img, txt, _orig=CLIPViTFM.calculate_score):
            y = _orig(Model, img, txt); rec.setdefault("scores", []).append(y.numpy().copy()); return y

        Model.model = types.SimpleNamespace(
            logit_scale=torch.tensor(float(np.log(ls))),
            encode_text=lambda tok: torch.from_numpy(ex.other_feats[int(tok)][None]))
        ModelNS = types.SimpleNamespace(calculate_score=calculate_score, model=Model.model)
        names = [f"noun{i}" for i in range(n_other)]
        ns = dict(torch=torch, np=np, T=types.SimpleNamespace(Resize=Resize), device="cpu",
                  r=0.5, alpha=0.6, k1=3, k2=6, softmax0=torch.nn.Softmax(0), Model=ModelNS,
                  sentence_features=torch.from_numpy(ex.sentence_feat[None]),
                  noun_phrase_features=torch.from_numpy(ex.noun_feat[None]),
                  visual_feature=torch.from_numpy(it.features), sentence_for_spacy="stub", nlp=None,
                  noun_phrase="stub", dirflag=dirf,
                  extract_nouns=lambda s, nlp: (list(names), list(names)),
                  extract_rela_word=lambda s, nlp: rela,
                  clip=types.SimpleNamespace(tokenize=lambda s: torch.tensor(int(s.rsplit("noun", 1)[1]))),
                  relation_boxes=ref_utils.relation_boxes, gen_dir_mask=ref_utils.gen_dir_mask,
                  Compute_IoU=ref_utils.Compute_IoU,
                  gem_model=lambda img, texts: torch.from_numpy(raw)[None],
                  image={"tensor_img": torch.zeros(1), "height": torch.tensor([h]), "width": torch.tensor([w])},
                  masks=torch.from_numpy(it.masks), boxes=torch.from_numpy(it.boxes),
                  target=torch.from_numpy(it.target)[None],
                  cum_I=0, cum_U=0, m_IoU=[], cum_I_final=0, cum_U_final=0, m_IoU_final=[])
        exec(compile(code, "Hybridgl_main.py:153-230", "exec"), ns)
        p = f"c{ci:02d}_"
        out[p + "meta"] = np.array([seed, h, w, n, n_other], np.int64)
        out[p + "flags"] = np.array([rela, dirf]); out[p + "logit_scale_exp"] = np.float32(ls)
        out[p + "features"] = it.features; out[p + "sentence"] = ex.sentence_feat; out[p + "noun"] = ex.noun_feat
        out[p + "others"] = ex.other_feats; out[p + "boxes"] = it.boxes
        out[p + "masks"] = np.packbits(it.masks, axis=-1); out[p + "target"] = np.packbits(it.target.astype(bool), axis=-1)
        out[p + "heat_raw"] = raw
        if h * w <= 8192:
            out[p + "heat_resized"] = rec["resized"]
        out[p + "score_clip"] = rec["scores"][0][:, 0]; out[p + "score_neg"] = rec["scores"][1][:, 0]
        out[p + "idx_hybrid"] = np.int64(ns["max_index_hybrid"]); out[p + "top_idx"] = ns["maxidxs"].numpy()
        out[p + "topneg_idx"] = ns["maxNegidxs"].numpy()
        out[p + "score_gem"] = ns["score_gem"].numpy()[:, 0]; out[p + "blended"] = ns["topscores"].numpy()
        out[p + "idx_final"] = np.int64(ns["max_index_final"])
        out[p + "IU"] = np.array([int(ns["cum_I"]), int(ns["cum_U"]), int(ns["cum_I_final"]), int(ns["cum_U_final"])], np.int64)
        out[p + "iou"] = np.array([float(ns["m_IoU"][0]), float(ns["m_IoU_final"][0])], np.float32)
        print("score", ci, rela, dirf, n_other, "hybrid", int(ns["max_index_hybrid"]), "final", int(ns["max_index_final"]),
              "IU", out[p + "IU"].tolist())
    out["n_cases"] = np.int64(len(cases))
    np.savez_compressed(os.path.join(HERE, "scoring.npz"), **out)


def gen_misc(ref_utils):
    out = {}
    for d in synth.DIRFLAGS:
        for (h, w) in ((5, 7), (6, 10), (48, 64), (3, 1)):
            out[f"dir_{d}_{h}x{w}"] = ref_utils.gen_dir_mask(d, h, w, None).numpy()
    rng = np.random.default_rng(9)
    rows = []
    for _ in range(64):
        bi = rng.integers(1, 60, 4); bj = rng.integers(1, 60, 4)
        si, sj = np.float32(rng.random()), np.float32(rng.random())
        for rela in synth.RELAFLAGS + ("bogus",):
            v = ref_utils.relation_boxes(torch.from_numpy(bi), torch.from_numpy(bj), torch.tensor(si), torch.tensor(sj), rela)
            rows.append((bi, bj, si, sj, rela, float(v)))
    out["rel_bi"] = np.stack([r[0] for r in rows]); out["rel_bj"] = np.stack([r[1] for r in rows])
    out["rel_si"] = np.array([r[2] for r in rows], np.float32); out["rel_sj"] = np.array([r[3] for r in rows], np.float32)
    out["rel_word"] = np.array([r[4] for r in rows]); out["rel_out"] = np.array([r[5] for r in rows], np.float32)
    np.savez_compressed(os.path.join(HERE, "misc.npz"), **out)
    print("misc", len(rows))


def gen_rle():
    """SAM's own RLE codec (third_party/segment-anything/segment_anything/utils/amg.py:107-149), imported in place."""
    sys.path.insert(0, os.path.join(REF, "third_party/segment-anything"))
    from segment_anything.utils.amg import mask_to_rle_pytorch, rle_to_mask
    out = {}
    rng = np.random.default_rng(31)
    cases = []
    for ci, (h, w, n) in enumerate(((48, 64, 6), (37, 45, 5), (64, 33, 4), (5, 3, 3), (96, 160, 4))):
        m = synth.make_masks(rng, n, h, w, min_area=20)
        m[0] = False; m[0, 0, 0] = True                     # first pixel set -> leading zero count
        m[1] = True                                          # full frame: one run spanning every column
        if n > 2:
            m[2] = rng.random((h, w)) < 0.5                  # noise: thousands of 1-pixel runs
        if n > 3:
            m[3] = False                                     # empty mask: a single run of zeros
        cases.append(m)
        rles = mask_to_rle_pytorch(torch.from_numpy(m))
        counts = np.concatenate([np.asarray(r["counts"], np.int32) for r in rles])
        off = np.cumsum([0] + [len(r["counts"]) for r in rles]).astype(np.int32)
        back = np.stack([rle_to_mask(r) for r in rles])
        assert np.array_equal(back, m)
        out[f"c{ci}_masks"] = m; out[f"c{ci}_counts"] = counts; out[f"c{ci}_off"] = off; out[f"c{ci}_decoded"] = back
        print("rle", ci, (h, w, n), "runs", counts.size)
    out["n_cases"] = np.int64(len(cases))
    np.savez_compressed(os.path.join(HERE, "rle.npz"), **out)


def gen_geometry(ref_utils):
    """mask2chw (utils.py:280-289), apply_visual_prompts 'blur' / 'circle' / 'black' (utils.py:306-341) and SAM's boxes
    (amg.py:303-346 batched_mask_to_box + :91-95 box_xyxy_to_xywh), called in place."""
    sys.path.insert(0, os.path.join(REF, "third_party/segment-anything"))
    from segment_anything.utils.amg import batched_mask_to_box, box_xyxy_to_xywh
    out = {}
    rng = np.random.default_rng(77)
    for ci, (h, w, n) in enumerate(((48, 64, 6), (97, 131, 5), (480, 640, 8), (33, 32, 4), (600, 800, 3))):
        m = synth.make_masks(rng, n, h, w, min_area=12)
        m[0] = False; m[0, h // 3, w // 2] = True              # a single pixel
        m[1] = True                                             # the full frame
        if n > 3:
            m[3] = False; m[3, 0, :] = True; m[3, :, w - 1] = True   # an L along two frame edges
        boxes = torch.stack([box_xyxy_to_xywh(b) for b in batched_mask_to_box(torch.from_numpy(m))]).numpy().astype(np.int64)
        chw = np.array([[c[0][0], c[0][1], c[1], c[2]] for c in (ref_utils.mask2chw(mm.astype(np.uint8)) for mm in m)], np.int32)
        out[f"c{ci}_masks"] = np.packbits(m, axis=-1); out[f"c{ci}_hw"] = np.array([h, w, n], np.int64)
        out[f"c{ci}_boxes"] = boxes; out[f"c{ci}_chw"] = chw
        if ci < 2:
            img = synth.make_image(rng, h, w)
            out[f"c{ci}_image"] = img
            for kind in ("blur", "black"):
                out[f"c{ci}_{kind}"] = np.stack([ref_utils.apply_visual_prompts(img, mm.astype(np.uint8), visual_prompt_type=(kind,)) for mm in m])
            # 'circle' (utils.py:322-335: cv2.ellipse at mask2chw's centre), alone and combined the way the if-chain orders the types
            for tag, kinds in (("circle", ("circle",)), ("blur_circle", ("blur", "circle")), ("circle_black", ("circle", "black"))):
                out[f"c{ci}_{tag}"] = np.stack([ref_utils.apply_visual_prompts(img, mm.astype(np.uint8), visual_prompt_type=kinds) for mm in m])
        print("geometry", ci, (h, w, n), boxes[2].tolist(), chw[2].tolist())
    empty = batched_mask_to_box(torch.zeros((1, 8, 8), dtype=torch.bool))
    out["empty_box"] = box_xyxy_to_xywh(empty[0]).numpy().astype(np.int64)
    out["n_cases"] = np.int64(5)
    np.savez_compressed(os.path.join(HERE, "geometry.npz"), **out)


BACKBONE_CASES = {
    # name: (CLIP ctor args, last_layer, heads, masking_block, modes, n masks, seed)  -- model/backbone.py:16-21 for ViT-B/16; the
    # ViT-L/14@336 row is the SURVEY section 8(c) extension (last_layer=22, num_heads=16, masking_block=21 set by hand)
    "b16": ((512, 224, 12, 768, 16, 77, 49408, 512, 8, 12), 10, 12, 9, ("G2L", "L2G", "G2L&L2G"), 5, 11),
    "l14": ((768, 336, 24, 1024, 14, 77, 49408, 768, 12, 12), 22, 16, 21, ("G2L&L2G",), 2, 12),
}
TEXT_SENTENCES = ["man in black", "car behind bike", "the left-most zebra's head &amp; neck!!", "a photo of a cat", "woman w/ umbrella, 2nd from right",
                  "far right giraffe"]


def gen_backbone(CLIP, CLIPViTFM):
    """CLIPViTFM.forward (model/backbone.py:117-306) and CLIP.encode_text (clip/model.py:414-431) of the REAL geometries (12 heads,
    197 tokens, masking_block=9; 16 heads, 577 tokens, masking_block=21) with weights derived from a seed on both sides."""
    import clip
    out = {}
    for tag, (ctor, last_layer, heads, mb, modes, n, seed) in BACKBONE_CASES.items():
        m = CLIPViTFM.__new__(CLIPViTFM)
        torch.nn.Module.__init__(m)
        m.last_layer, m.num_heads = last_layer, heads
        m.model = CLIP(*ctor).eval()
        sd = m.model.state_dict()
        new = synth.seeded_clip_state_dict({k: v.shape for k, v in sd.items()}, seed)
        m.model.load_state_dict({k: torch.from_numpy(v) for k, v in new.items()})
        res = ctor[1]
        loc, glo = synth.seeded_images(seed, n, res)
        masks = synth.make_masks(np.random.default_rng(seed), n, 120, 160)
        out[f"{tag}_meta"] = np.array([seed, n, res, mb, 120, 160], np.int64)
        out[f"{tag}_masks"] = np.packbits(masks, axis=-1)
        for mode in modes:
            y = m(torch.from_numpy(loc), torch.from_numpy(glo), torch.from_numpy(masks), masking_block=mb, fusion_mode=mode)
            out[f"{tag}_out/{mode}"] = y.numpy()
            print("backbone", tag, mode, tuple(y.shape), float(y.abs().mean()))
        tok = clip.tokenize(TEXT_SENTENCES)
        txt = m.model.encode_text(tok)
        out[f"{tag}_tokens"] = tok.numpy()
        out[f"{tag}_text"] = txt.numpy()
        out[f"{tag}_score"] = m.calculate_score(torch.from_numpy(out[f"{tag}_out/{modes[-1]}"]), txt).numpy()
        print("text", tag, tuple(txt.shape), float(txt.abs().mean()))
    out["sentences"] = np.array(TEXT_SENTENCES)
    np.savez_compressed(os.path.join(HERE, "backbone.npz"), **out)


if __name__ == "__main__":
    CLIP, CLIPViTFM, ref_utils = load_reference()
    which = sys.argv[1:] or ["prep", "grid", "forward", "scoring", "misc", "rle", "backbone", "geometry"]
    if "rle" in which:
        gen_rle()
    with torch.no_grad():
        if "prep" in which:
            gen_prep(ref_utils)
        if "grid" in which:
            gen_grid(CLIPViTFM)
        if "forward" in which:
            gen_forward(CLIP, CLIPViTFM)
        if "scoring" in which:
            gen_scoring(CLIPViTFM, ref_utils)
        if "misc" in which:
            gen_misc(ref_utils)
        if "backbone" in which:
            gen_backbone(CLIP, CLIPViTFM)
        if "geometry" in which:
            gen_geometry(ref_utils)
