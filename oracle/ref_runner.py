"""Runs the REAL reference path (the files oracle/make_ref.py copied into oracle/_ref, or /root/reference in the build container)
on one synthetic image: the prep loop Hybridgl_main.py:92-125, the mask -> grid resize model/backbone.py:160 and, per expression,
the scoring / spatial-guidance / IoU block Hybridgl_main.py:153-230 -- the reference's own source lines, exec'd in place with
stubs for the producers that are unavailable offline (spaCy flags, GEM map, encode_text outputs are synthetic inputs).
TEST / BASELINE INFRASTRUCTURE ONLY (bench.py's CPU arms and tests/); the product never imports it.
"""
from __future__ import annotations

import os
import sys
import textwrap
import time
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def ref_root():
    """oracle/_ref when the recipe has run, else the reference checkout of the build container, else None."""
    for cand in (os.path.join(HERE, "_ref"), os.environ.get("HGL_REFERENCE_ROOT", "/root/reference")):
        if cand and os.path.exists(os.path.join(cand, "Hybridgl_main.py")):
            return cand
    return None


_STATE = {}


def _load():
    if _STATE:
        return _STATE
    root = ref_root()
    if root is None:
        raise RuntimeError("the reference is not available (run oracle/make_ref.py in the build container)")
    for n in ("ftfy", "spacy", "matplotlib", "matplotlib.pyplot", "matplotlib.gridspec"):
        sys.modules.setdefault(n, types.ModuleType(n))
    sys.modules["ftfy"].fix_text = lambda s: s
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].gridspec = sys.modules["matplotlib.gridspec"]
    sys.path[:0] = [os.path.join(root, "third_party/modified_CLIP"), root]
    import torch
    import utils as ref_utils                      # the reference's utils.py
    from model.backbone import CLIPViTFM           # the reference's wrapper (only calculate_score / TF.resize are used here)

    def lines(first, last, anchors):
        with open(os.path.join(root, "Hybridgl_main.py")) as f:
            src = f.readlines()
        for off, needle in anchors:
            assert needle in src[off - 1], (off, needle)
        return compile(textwrap.dedent("".join(src[first - 1:last])), f"Hybridgl_main.py:{first}-{last}", "exec")
    _STATE.update(root=root, torch=torch, utils=ref_utils, CLIPViTFM=CLIPViTFM,
                  prep=lines(92, 125, [(93, "pixel_mean"), (99, "GaussianBlur"), (125, "local_imgs")]),
                  score=lines(153, 230, [(153, "text_ensemble"), (200, "gem_model"), (230, "Compute_IoU")]))
    return _STATE


def run_image(seed: int, cfg: dict, n_masks: int, threads: int = 1) -> dict:
    """One image of the workload through the reference's own code.  Returns dict(seconds, idx_hybrid, idx_final, IU)."""
    st = _load()
    torch = st["torch"]
    import cv2
    import torchvision.transforms as T
    import torchvision.transforms.functional as TF
    from hybridgl_b200 import synth
    torch.set_num_threads(threads)
    cv2.setNumThreads(threads)
    h, w, S, g, De = cfg["h"], cfg["w"], cfg["S"], cfg["g"], cfg["De"]
    it = synth.make_item(seed, h, w, n_masks, cfg["n_expr"], de=De, n_other=2)
    sam = torch.from_numpy(it.image)[None]
    img_norm = T.Normalize([0.485, 0.456, 0.406], [0.229, 0.224, 0.225])(T.ToTensor()(it.image))[None]
    Model = st["CLIPViTFM"].__new__(st["CLIPViTFM"])
    torch.nn.Module.__init__(Model)
    t0 = time.perf_counter()
    with torch.no_grad():
        ns = dict(torch=torch, cv2=cv2, np=np, T=T, Height=S, Width=S, device="cpu", image={"sam_img": sam, "image": img_norm},
                  original_img=img_norm, masks=torch.from_numpy(it.masks), boxes=torch.from_numpy(it.boxes))
        exec(st["prep"], ns)                                                             # Hybridgl_main.py:92-125
        TF.resize(torch.from_numpy(it.masks).type(torch.float32), (g, g))                # model/backbone.py:160
        out = dict(idx_hybrid=[], idx_final=[], IU=[])
        for ex in it.expressions:
            Model.model = types.SimpleNamespace(logit_scale=torch.tensor(float(np.log(100.0))),
                                                encode_text=lambda tok, ex=ex: torch.from_numpy(ex.other_feats[int(tok)][None]))
            MS = types.SimpleNamespace(calculate_score=lambda a, b: st["CLIPViTFM"].calculate_score(Model, a, b), model=Model.model)
            names = [f"noun{i}" for i in range(ex.other_feats.shape[0])]
            sc = dict(torch=torch, np=np, T=T, device="cpu", r=0.5, alpha=0.6, k1=3, k2=6, softmax0=torch.nn.Softmax(0), Model=MS,
                      sentence_features=torch.from_numpy(ex.sentence_feat[None]), noun_phrase_features=torch.from_numpy(ex.noun_feat[None]),
                      visual_feature=torch.from_numpy(it.features), sentence_for_spacy="stub", nlp=None, noun_phrase="stub", dirflag=ex.dirflag,
                      extract_nouns=lambda s, nlp: (list(names), list(names)), extract_rela_word=lambda s, nlp, ex=ex: ex.relaflag,
                      clip=types.SimpleNamespace(tokenize=lambda s: torch.tensor(int(s.rsplit("noun", 1)[1]))),
                      relation_boxes=st["utils"].relation_boxes, gen_dir_mask=st["utils"].gen_dir_mask, Compute_IoU=st["utils"].Compute_IoU,
                      gem_model=lambda img, texts, ex=ex: torch.from_numpy(ex.heat_raw)[None, None],
                      image={"tensor_img": torch.zeros(1), "height": torch.tensor([h]), "width": torch.tensor([w])},
                      masks=torch.from_numpy(it.masks), boxes=torch.from_numpy(it.boxes), target=torch.from_numpy(it.target)[None],
                      cum_I=0, cum_U=0, m_IoU=[], cum_I_final=0, cum_U_final=0, m_IoU_final=[])
            exec(st["score"], sc)                                                        # Hybridgl_main.py:153-230
            out["idx_hybrid"].append(int(sc["max_index_hybrid"])); out["idx_final"].append(int(sc["max_index_final"]))
            out["IU"].append([int(sc["cum_I"]), int(sc["cum_U"]), int(sc["cum_I_final"]), int(sc["cum_U_final"])])
    out["seconds"] = time.perf_counter() - t0
    return out
