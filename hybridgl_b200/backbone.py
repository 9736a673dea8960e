"""Host-side mirror of the reference's hybrid CLIP wrapper (model/backbone.py) on top of libhgl.

Same call surface as the reference class:

    CLIPViTFM(model_name='ViT-B/16').forward(local_imgs, global_imgs, pred_masks, masking_block=None, fusion_mode='G2L')
    CLIPViTFM.calculate_score(image_features, text_features)          (model/backbone.py:74-87)
    CLIPViTFM.make_attn_mask(pred_masks, size=None)                   (model/backbone.py:108-115)

What is B200-native here (the path's rows a2-a5 of SURVEY.md section 8):
  * mask -> patch grid (model/backbone.py:160)              -> ops.masks_to_grid      (hgl_mask_grid, packed masks)
  * CLS-row attention mask (model/backbone.py:108-115)      -> ops.attn_key_bias      (hgl_attn_bias): a [M, L+1] additive key
    bias for the CLS query only.  The reference materialises N*heads*(L+1)^2 booleans (1.07 GB at ViT-L/14@336 with 200
    masks) and nn.MultiheadAttention converts them to a float mask of the same shape; here every query row runs through
    one unmasked SDPA call and only the CLS row is recomputed with the bias (identical result: only that row is masked).
  * token masking + stream mixes (model/backbone.py:214-216, 235-249, 275-291) -> ops.token_mask_fuse (hgl_token_mask_fuse)
    on batch-first [M, L+1, D] streams; the 2-4 streams of a fused block go through ONE resblock call (weights read once).
The transformer blocks themselves are plain PyTorch with random-init weights of the named architecture (north_star):
there is no checkpoint offline.  Parameter names follow CLIP's `visual.*` state_dict so real weights load unchanged.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .clip_text import tokenize  # noqa: F401  (re-export: `clip.tokenize` of the reference's loop, Hybridgl_main.py:146-160)

# name -> (embed_dim, image_resolution, layers, width, patch, heads); last_layer as in model/backbone.py:16-21
# (ViT-L/14@336 is the SURVEY section 8(c) extension: last_layer=22, heads=16)
# text tower (clip/model.py CLIP.__init__): context 77, vocabulary 49408, width / heads / layers per model
ARCH = {
    "ViT-B/32": dict(embed_dim=512, res=224, layers=12, width=768, patch=32, heads=12, last_layer=10, text_width=512, text_heads=8, text_layers=12),
    "ViT-B/16": dict(embed_dim=512, res=224, layers=12, width=768, patch=16, heads=12, last_layer=10, text_width=512, text_heads=8, text_layers=12),
    "ViT-L/14@336px": dict(embed_dim=768, res=336, layers=24, width=1024, patch=14, heads=16, last_layer=22, text_width=768, text_heads=12,
                           text_layers=12),
}


class LayerNormF32(nn.LayerNorm):
    """CLIP's LayerNorm: computed in fp32 whatever the stream dtype (third_party/modified_CLIP/clip/model.py:188-195)."""

    def forward(self, x):
        return F.layer_norm(x.float(), self.normalized_shape, self.weight.float(), self.bias.float(), self.eps).to(x.dtype)


class _Attn(nn.Module):
    """Parameter container with nn.MultiheadAttention's names (in_proj_weight, in_proj_bias, out_proj.*)."""

    def __init__(self, width: int):
        super().__init__()
        self.in_proj_weight = nn.Parameter(torch.empty(3 * width, width))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * width))
        self.out_proj = nn.Linear(width, width)
        nn.init.xavier_uniform_(self.in_proj_weight)


class ResBlock(nn.Module):
    """x + MHA(ln_1(x)); x + mlp(ln_2(x))  (third_party/modified_CLIP/clip/model.py:244-257), batch-first.

    `cls_bias` f32 [M, L+1] (0 / -inf) masks patch keys for the CLS query only -- the reference's attn_mask has no other
    blocked entry."""

    def __init__(self, width: int, heads: int):
        super().__init__()
        self.heads = heads
        self.attn = _Attn(width)
        self.ln_1 = LayerNormF32(width)
        self.mlp = nn.Sequential(OrderedDict([("c_fc", nn.Linear(width, 4 * width)), ("gelu", nn.Identity()),
                                              ("c_proj", nn.Linear(4 * width, width))]))
        self.ln_2 = LayerNormF32(width)

    def attention(self, h: torch.Tensor, cls_bias: Optional[torch.Tensor], causal: bool = False, rows=None) -> torch.Tensor:
        """cls_bias f32 [M, L1] (0 / -inf): key bias of the CLS query; rows: optional list of (lo, hi) row ranges that carry a
        non-trivial bias (the other rows' bias is all zero and their CLS output is SDPA's own)."""
        M, L1, D = h.shape
        hd = D // self.heads
        qkv3 = F.linear(h, self.attn.in_proj_weight, self.attn.in_proj_bias)     # [M, L1, 3*D], packed (q | k | v) x heads x hd
        qkv = qkv3.view(M, L1, 3, self.heads, hd)
        q, k, v = (qkv[:, :, i].transpose(1, 2) for i in range(3))             # [M, heads, L1, hd]
        o = F.scaled_dot_product_attention(q, k, v, is_causal=causal)           # every query row; causal = the text tower's mask
        if cls_bias is not None:
            # only the CLS query is masked in the reference (model/backbone.py:108-115): recompute that one row under the key
            # bias (hgl_cls_attention, one warp per (proposal, head)) and drop it into SDPA's output in place
            for lo, hi in (rows if rows is not None else [(0, M)]):
                if hi > lo:
                    o[lo:hi, :, 0, :] = ops.cls_attention(qkv3[lo:hi], cls_bias[lo:hi].contiguous(), self.heads)
        return self.attn.out_proj(o.transpose(1, 2).reshape(M, L1, D))

    def forward(self, x: torch.Tensor, cls_bias: Optional[torch.Tensor] = None, causal: bool = False, rows=None,
                h: Optional[torch.Tensor] = None) -> torch.Tensor:
        """h: ln_1(x) when the caller already has it (hgl_token_mask_fuse_ln writes it together with the fused streams)."""
        x = x + self.attention(self.ln_1(x) if h is None else h, cls_bias, causal, rows)
        h = self.mlp.c_fc(self.ln_2(x))
        h = h * torch.sigmoid(1.702 * h)                                        # QuickGELU
        return x + self.mlp.c_proj(h)


class _Transformer(nn.Module):
    def __init__(self, width, layers, heads):
        super().__init__()
        self.resblocks = nn.ModuleList([ResBlock(width, heads) for _ in range(layers)])


class VisionTransformer(nn.Module):
    """CLIP ViT with the state_dict layout of clip.model.VisionTransformer (conv1, class_embedding, positional_embedding,
    ln_pre, transformer.resblocks.*, ln_post, proj)."""

    def __init__(self, res, patch, width, layers, heads, out_dim):
        super().__init__()
        self.conv1 = nn.Conv2d(3, width, kernel_size=patch, stride=patch, bias=False)
        scale = width ** -0.5
        self.class_embedding = nn.Parameter(scale * torch.randn(width))
        self.positional_embedding = nn.Parameter(scale * torch.randn((res // patch) ** 2 + 1, width))
        self.ln_pre = LayerNormF32(width)
        self.transformer = _Transformer(width, layers, heads)
        self.ln_post = LayerNormF32(width)
        self.proj = nn.Parameter(scale * torch.randn(width, out_dim))

    def embed(self, imgs: torch.Tensor) -> torch.Tensor:
        """conv1 + class token + positional embedding + ln_pre -> [M, L+1, D]  (model/backbone.py:130-137)."""
        x = self.conv1(imgs.to(self.conv1.weight.dtype)).flatten(2).transpose(1, 2)
        cls = self.class_embedding.to(x.dtype).expand(x.shape[0], 1, -1)
        return self.ln_pre(torch.cat([cls, x], dim=1) + self.positional_embedding.to(x.dtype))

    def head(self, x: torch.Tensor, out: Optional[torch.Tensor] = None, accumulate: bool = False) -> torch.Tensor:
        """ln_post(CLS) @ proj  (model/backbone.py:254-258) in one launch (hgl_cls_head): f32 [M, De]."""
        if not x.is_cuda:
            return self.ln_post(x[:, 0, :]) @ self.proj
        return ops.cls_head(x.contiguous(), self.ln_post.weight, self.ln_post.bias, self.proj, self.ln_post.eps, out=out, accumulate=accumulate)


class _Clip(nn.Module):
    """What the evaluation loop touches of clip.model.CLIP: `visual`, `encode_text` (third_party/modified_CLIP/clip/model.py:414-431) and
    `logit_scale`, with CLIP's state_dict names (token_embedding, positional_embedding, transformer.resblocks.*, ln_final,
    text_projection) so that real weights load unchanged."""

    def __init__(self, a):
        super().__init__()
        self.visual = VisionTransformer(a["res"], a["patch"], a["width"], a["layers"], a["heads"], a["embed_dim"])
        tw, th, tl = a.get("text_width", 512), a.get("text_heads", 8), a.get("text_layers", 12)
        self.context_length, self.vocab_size = a.get("context_length", 77), a.get("vocab_size", 49408)
        self.token_embedding = nn.Embedding(self.vocab_size, tw)
        self.positional_embedding = nn.Parameter(torch.empty(self.context_length, tw))
        self.transformer = _Transformer(tw, tl, th)
        self.ln_final = LayerNormF32(tw)
        self.text_projection = nn.Parameter(torch.empty(tw, a["embed_dim"]))
        self.logit_scale = nn.Parameter(torch.ones([]) * math.log(1 / 0.07))     # clip/model.py CLIP.__init__
        # clip/model.py initialize_parameters (random-init stand-in for the checkpoint that is unavailable offline)
        nn.init.normal_(self.token_embedding.weight, std=0.02)
        nn.init.normal_(self.positional_embedding, std=0.01)
        nn.init.normal_(self.text_projection, std=tw ** -0.5)
        proj_std, attn_std, fc_std = (tw ** -0.5) * ((2 * tl) ** -0.5), tw ** -0.5, (2 * tw) ** -0.5
        for blk in self.transformer.resblocks:
            nn.init.normal_(blk.attn.in_proj_weight, std=attn_std)
            nn.init.normal_(blk.attn.out_proj.weight, std=proj_std)
            nn.init.normal_(blk.mlp.c_fc.weight, std=fc_std)
            nn.init.normal_(blk.mlp.c_proj.weight, std=proj_std)

    @property
    def dtype(self):
        return self.visual.conv1.weight.dtype

    @torch.no_grad()
    def encode_text(self, text: torch.Tensor, target_noun_index=None) -> torch.Tensor:
        """clip/model.py:414-431: token + positional embedding, causal transformer, ln_final, the <|endoftext|> position (the
        highest token id of every row; or target_noun_index + 1 when given) projected by text_projection -> [n, embed_dim]."""
        text = text.to(self.token_embedding.weight.device)
        x = self.token_embedding(text.long()).to(self.dtype) + self.positional_embedding.to(self.dtype)
        for blk in self.transformer.resblocks:
            x = blk(x, causal=True)
        x = self.ln_final(x).to(self.dtype)
        pos = (target_noun_index + 1) if target_noun_index else text.argmax(dim=-1)
        return x[torch.arange(x.shape[0], device=x.device), pos] @ self.text_projection


class CLIPViTFM(nn.Module):
    """Drop-in for model/backbone.py::CLIPViTFM (random-init; `load_clip_state_dict` accepts real CLIP weights)."""

    def __init__(self, model_name: str = "ViT-B/16", size: int = 224, arch: Optional[dict] = None, antialias: bool = True,
                 device="cuda", dtype: torch.dtype = torch.float32):
        super().__init__()
        a = dict(ARCH[model_name]) if arch is None else dict(arch)
        self.last_layer = a["last_layer"]
        self.num_heads = a["heads"]
        self.antialias = antialias          # TF.resize semantics of model/backbone.py:160
        self.model = _Clip(a).to(device=device, dtype=dtype).eval()

    @property
    def device(self):
        return self.model.visual.conv1.weight.device

    @property
    def dtype(self):
        return self.model.visual.conv1.weight.dtype

    def load_clip_state_dict(self, sd: dict) -> None:
        """Load a CLIP state_dict: `visual.*` (or a bare visual state_dict) plus, when present, the text tower and logit_scale."""
        sd = {k: torch.as_tensor(v) for k, v in sd.items()}
        vis = {k[len("visual."):]: v for k, v in sd.items() if k.startswith("visual.")}
        text_keys = ("token_embedding.", "positional_embedding", "transformer.", "ln_final.", "text_projection")
        if vis:
            self.model.visual.load_state_dict(vis)
            txt = {k: v for k, v in sd.items() if k.startswith(text_keys)}
            if txt:
                missing, unexpected = self.model.load_state_dict(txt, strict=False)
                bad = [k for k in missing if k.startswith(text_keys)] + list(unexpected)
                if bad:
                    raise KeyError(f"text tower state_dict mismatch: {bad[:5]}")
        else:
            self.model.visual.load_state_dict({k: v for k, v in sd.items() if k != "logit_scale"})
        if "logit_scale" in sd:
            with torch.no_grad():
                self.model.logit_scale.copy_(sd["logit_scale"])

    # ---- model/backbone.py:74-87 -----------------------------------------------------------------------------------
    def calculate_score(self, image_features: torch.Tensor, text_features: torch.Tensor, visual_norm_dim: int = 1) -> torch.Tensor:
        """logit_scale.exp() * normalize(img) @ normalize(txt).T -> [N, T]; runs in hgl_score_select (scores only)."""
        if visual_norm_dim != 1:
            raise ValueError("only visual_norm_dim=1 (the reference's call sites) is supported")
        T = text_features.shape[0]
        feat = image_features
        if feat.dtype not in (torch.float32, torch.bfloat16):     # the reference's CUDA model is fp16: score such features in f32
            feat = feat.float()
        feat = feat.contiguous()
        txt = text_features.float().contiguous()
        N = feat.shape[0]
        dev = feat.device
        zero_off = torch.zeros(T + 1, dtype=torch.int32, device=dev)
        # one 'expression' per text row: r=1 takes the sentence vector alone
        res = ops.score_select(feat, txt, txt, torch.zeros((0, feat.shape[1]), device=dev), zero_off,
                               torch.zeros((N, 4), dtype=torch.int64, device=dev), torch.zeros(T, dtype=torch.int32, device=dev),
                               None, logit_scale_exp=float(self.model.logit_scale.detach().exp()), r=1.0, alpha=0.0)
        return res["score_clip"][:, :N].t().contiguous()

    # ---- model/backbone.py:108-115 ---------------------------------------------------------------------------------
    def make_attn_mask(self, pred_masks: torch.Tensor, size: Optional[int] = None) -> torch.Tensor:
        """grid masks [N,g,g] (float) -> bool [N*heads, L+1, L+1], True = blocked.  Kept for drop-in use; forward() uses
        the compact CLS-row bias instead."""
        if size is not None:
            pred_masks = ops.masks_to_grid(pred_masks.to(torch.bool) if pred_masks.dtype != torch.bool else pred_masks, size,
                                           antialias=self.antialias)
        return ops.make_attn_mask(pred_masks.float().contiguous(), self.num_heads)

    # ---- model/backbone.py:117-309 ---------------------------------------------------------------------------------
    fused_ln = True      # False: hgl_token_mask_fuse + torch.cat + the block's own LayerNorm (the round-1 path; examples/eval_synthetic.py times both)

    def _fuse_ln(self, blk, streams):
        """streams: [(src, add, grid, a, b)] of equal shape [N, L1, D].  Returns (x, ln_1(x)) of the concatenated batch [len * N, L1, D],
        every stream fused and normalised by one launch straight into its slice."""
        if not self.fused_ln:
            parts = [src if (add is None and grid is None) else ops.token_mask_fuse(src.contiguous(), add, grid, a, b, layout="NLD")
                     for src, add, grid, a, b in streams]
            return (parts[0] if len(parts) == 1 else torch.cat(parts, dim=0)), None
        src0 = streams[0][0].contiguous()
        N = src0.shape[0]
        xcat = torch.empty((len(streams) * N,) + tuple(src0.shape[1:]), dtype=src0.dtype, device=src0.device)
        hcat = torch.empty_like(xcat)
        gam, bet = blk.ln_1.weight.float().contiguous(), blk.ln_1.bias.float().contiguous()
        for k, (src, add, grid, a, b) in enumerate(streams):
            ops.token_mask_fuse_ln(src.contiguous(), None if add is None else add.contiguous(), grid, a, b, gam, bet, blk.ln_1.eps,
                                   out_x=xcat[k * N:(k + 1) * N], out_ln=hcat[k * N:(k + 1) * N])
        return xcat, hcat

    @torch.no_grad()
    def forward(self, local_imgs, global_imgs, pred_masks, masking_block: Optional[int] = None, fusion_mode: str = "G2L"):
        if masking_block is None:
            masking_block = self.last_layer
        vit = self.model.visual
        blocks = vit.transformer.resblocks
        x = vit.embed(local_imgs)                                                 # local stream  [M, L+1, D]
        if fusion_mode == "crop":                                                 # model/backbone.py:126-128
            for blk in blocks:
                x = blk(x)
            return vit.head(x)
        x2 = vit.embed(global_imgs) if global_imgs is not None else None          # global stream
        M, L1, D = x.shape
        g = int(math.isqrt(L1 - 1))
        assert g * g == L1 - 1
        grid = ops.masks_to_grid(pred_masks, g, antialias=self.antialias)         # model/backbone.py:160
        N = grid.shape[0]
        final = self.last_layer + 1

        if fusion_mode == "token_masking":                                        # model/backbone.py:161-184
            for i, blk in enumerate(blocks):
                if i >= masking_block:
                    if x.shape[0] != N:
                        x = x.expand(N, -1, -1).contiguous()
                    xf, h = self._fuse_ln(blk, [(x, None, grid, 1.0, 0.0)])          # token masking + ln_1 in one pass
                    x = blk(xf, h=h)
                    if i == final:
                        return vit.head(x)
                else:
                    x = blk(x)
            return x.transpose(0, 1)

        bias = ops.attn_key_bias(grid)                                            # [N, L+1], 0 / -inf
        if fusion_mode == "attn_masking":                                         # model/backbone.py:186-203
            for i, blk in enumerate(blocks):
                if i >= masking_block:
                    if i == masking_block and x.shape[0] != N:
                        x = x.expand(N, -1, -1).contiguous()
                    x = blk(x, bias)
                    if i == self.last_layer:
                        return vit.head(x)
                else:
                    x = blk(x)
            return x.transpose(0, 1)

        if fusion_mode not in ("L2G", "G2L", "G2L&L2G"):
            return x.transpose(0, 1)                                              # model/backbone.py:309 (LND like the reference)

        zero = torch.zeros_like(bias)
        both = torch.cat([x, x2], dim=0)                                          # blocks < masking_block: both streams, one call
        xhl = xhg = None
        for i, blk in enumerate(blocks):
            if i < masking_block:
                both = blk(both)
                continue
            if i == masking_block:
                x, x2 = both[:N], both[N:]
                if fusion_mode == "G2L&L2G":
                    xhl, xhg = x, x2
            # Each stream of the block is written ONCE, already mixed, into its slice of the block's batch together with its ln_1
            # (hgl_token_mask_fuse_ln): no torch.cat copy, no separate LayerNorm pass.  (src, add, grid, a, b) = a*tokenmask(src) + b*add
            if fusion_mode == "G2L":                                              # model/backbone.py:227-260
                xcat, hcat = self._fuse_ln(blk, [(x2, x, grid, 2.0, 1.0),         # 2*tokenmask(x2) + x
                                                 (x2, None, None, 1.0, 0.0)])
                out = blk(xcat, torch.cat([zero, bias], dim=0), rows=[(N, 2 * N)], h=hcat)
                x, x2 = out[:N], out[N:]
                result = x
            elif fusion_mode == "L2G":                                            # model/backbone.py:206-225
                xcat, hcat = self._fuse_ln(blk, [(x, None, None, 1.0, 0.0),
                                                 (x2, x, None, 2.0, 1.0)])        # x_old + 2*x2
                out = blk(xcat, torch.cat([zero, bias], dim=0), rows=[(N, 2 * N)], h=hcat)
                x, x2 = out[:N], out[N:]
                result = x2
            else:                                                                 # model/backbone.py:262-306
                xcat, hcat = self._fuse_ln(blk, [(x, None, None, 1.0, 0.0), (x2, None, None, 1.0, 0.0),
                                                 (x2, xhl, grid, 2.0, 1.0),       # xhl + 2*tokenmask(x2)
                                                 (xhg, x, None, 2.0, 1.0)])       # x + 2*xhg
                out = blk(xcat, torch.cat([zero, bias, zero, bias], dim=0), rows=[(N, 2 * N), (3 * N, 4 * N)], h=hcat)
                x, x2, xhl, xhg = out[:N], out[N:2 * N], out[2 * N:3 * N], out[3 * N:]
                result = None
            if i == final:
                if fusion_mode == "G2L&L2G":
                    return vit.head(xhg, out=vit.head(xhl), accumulate=True)      # CLS(xhl) + CLS(xhg), model/backbone.py:296-306
                return vit.head(result)
        return x.transpose(0, 1)
