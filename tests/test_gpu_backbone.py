"""GPU parity of the host-side mirrors (hybridgl_b200/backbone.py, hybridgl_b200/utils.py) against outputs recorded from
the reference's own CLIPViTFM.forward / calculate_score / relation_boxes / gen_dir_mask / Compute_IoU
(tests/golden/forward.npz, misc.npz, scoring.npz -- produced by tests/golden/gen_golden.py from /root/reference)."""
import numpy as np
import pytest
import torch

from conftest import unpack_masks
from oracle import hybridgl_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"

TINY = dict(embed_dim=32, res=64, layers=4, width=64, patch=16, heads=1, last_layer=2)   # gen_golden.make_model


@pytest.fixture(scope="module")
def model(golden):
    from hybridgl_b200.backbone import CLIPViTFM
    g = golden("forward")
    torch.backends.cudnn.allow_tf32 = False          # fp32 parity: no TF32 in conv1 / matmuls
    torch.backends.cuda.matmul.allow_tf32 = False
    m = CLIPViTFM(arch=TINY, device=DEV)
    sd = {k[2:]: g[k] for k in g.files if k.startswith("w/")}
    m.load_clip_state_dict(sd)
    with torch.no_grad():
        m.model.logit_scale.copy_(torch.as_tensor(g["logit_scale"]))
    return m


@pytest.mark.parametrize("mode", ["G2L", "L2G", "G2L&L2G", "token_masking", "attn_masking", "crop"])
def test_forward_matches_reference(model, golden, mode):
    g = golden("forward")
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)  # noqa: E731
    y = model(cu(g["local"]), cu(g["global"]), cu(g["masks"]), masking_block=1, fusion_mode=mode)
    ref = g["out/" + mode]
    assert tuple(y.shape) == ref.shape
    # fp32 backbone on the GPU (TF32 off) vs the reference's CPU fp32: tolerance 2e-4 abs on O(1) features
    np.testing.assert_allclose(y.float().cpu().numpy(), ref, rtol=2e-3, atol=2e-4)


def test_calculate_score_matches_reference(model, golden):
    g = golden("forward")
    feat = torch.from_numpy(g["out/G2L"]).to(DEV)
    txt = torch.from_numpy(g["score_text"]).to(DEV)
    got = model.calculate_score(feat, txt).cpu().numpy()
    np.testing.assert_allclose(got, g["score"], rtol=1e-3, atol=1e-4)      # north-star tolerance: 1e-3 relative


def test_make_attn_mask_dropin(model, golden):
    g = golden("grid")
    grid = torch.from_numpy(g["d_aa"]).to(DEV)
    model.num_heads, keep = 3, model.num_heads
    try:
        am = model.make_attn_mask(grid).cpu().numpy()
    finally:
        model.num_heads = keep
    assert np.array_equal(am[:, 0, :], g["d_attn_row0"]) and not am[:, 1:, :].any()


def test_cls_bias_attention_equals_full_mask(model):
    """The compact CLS-row bias must reproduce nn.MultiheadAttention with the reference's full boolean mask."""
    torch.manual_seed(0)
    blk = model.model.visual.transformer.resblocks[1]
    M, L1, D = 6, 17, 64
    x = torch.randn(M, L1, D, device=DEV)
    grid = torch.rand(M, 4, 4, device=DEV); grid[grid < 0.5] = 0
    from hybridgl_b200 import ops
    bias = ops.attn_key_bias(grid)
    full = ops.make_attn_mask(grid, 1)                                    # [M*1, L1, L1] True = blocked
    mha = torch.nn.MultiheadAttention(D, 1).to(DEV)
    with torch.no_grad():
        mha.in_proj_weight.copy_(blk.attn.in_proj_weight); mha.in_proj_bias.copy_(blk.attn.in_proj_bias)
        mha.out_proj.weight.copy_(blk.attn.out_proj.weight); mha.out_proj.bias.copy_(blk.attn.out_proj.bias)
        h = blk.ln_1(x)
        ref = mha(h.transpose(0, 1), h.transpose(0, 1), h.transpose(0, 1), need_weights=False, attn_mask=full)[0].transpose(0, 1)
        got = blk.attention(h, bias)
    torch.testing.assert_close(got, ref, rtol=1e-4, atol=1e-5)


def test_utils_mirror(golden):
    from hybridgl_b200 import utils as U
    g = golden("misc")
    for key in g.files:
        if key.startswith("dir_"):
            _, d, hw = key.split("_")
            h, w = map(int, hw.split("x"))
            got = U.gen_dir_mask(d, h, w, DEV).cpu().numpy()
            assert np.array_equal(got, g[key]), key                         # bit-exact ramp (ATen linspace restated)
    for bi, bj, si, sj, word, ref in zip(g["rel_bi"], g["rel_bj"], g["rel_si"], g["rel_sj"], g["rel_word"], g["rel_out"]):
        v = U.relation_boxes(torch.from_numpy(bi), torch.from_numpy(bj), torch.tensor(si), torch.tensor(sj), str(word))
        np.testing.assert_allclose(float(v), ref, rtol=2e-7, atol=0)
    s = golden("scoring")
    for ci in (0, 7, 24):
        p = f"c{ci:02d}_"
        _, h, w, n, _ = s[p + "meta"].tolist()
        masks = unpack_masks(s[p + "masks"], w); target = unpack_masks(s[p + "target"], w)
        cum_I, cum_U, lst = 0, 0, []
        for pick, (ri, ru) in zip((int(s[p + "idx_hybrid"]), int(s[p + "idx_final"])), (s[p + "IU"][:2], s[p + "IU"][2:])):
            iou, lst, ci_, cu_ = U.Compute_IoU(torch.from_numpy(masks[pick]).to(DEV), torch.from_numpy(target.astype(np.uint8))[None].to(DEV), 0, 0, lst)
            assert [int(ci_), int(cu_)] == [int(ri), int(ru)]
            i0, u0, ref_iou = O.compute_iou(masks[pick], target)
            assert abs(float(iou) - ref_iou) < 1e-6
        assert len(lst) == 2


@pytest.mark.parametrize("mode", ["G2L", "L2G", "G2L&L2G"])
def test_example_eval_loop_runs_all_fusion_modes(mode):
    """examples/eval_synthetic.py: the reference's evaluation loop with libhgl behind it, end to end with the ViT blocks."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("eval_synthetic", os.path.join(os.path.dirname(os.path.dirname(__file__)), "examples", "eval_synthetic.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rep = mod.main(["--images", "2", "--masks", "12", "--expr", "2", "--fusion_mode", mode, "--height", "240", "--width", "320"])
    assert rep["n_expressions"] == 4 and 0.0 <= rep["oIoU"] <= 100.0


# ------------------------------------------------------------------------------------------------ the real geometries
@pytest.mark.parametrize("tag,name,modes", [("b16", "ViT-B/16", ("G2L", "L2G", "G2L&L2G")), ("l14", "ViT-L/14@336px", ("G2L&L2G",))])
def test_forward_real_geometry_matches_reference(golden, tag, name, modes):
    """CLIPViTFM.forward at ViT-B/16 (12 heads, 197 tokens, masking_block=9) and ViT-L/14@336 (16 heads, 577 tokens, masking_block=21;
    the SURVEY 8(c) extension) against outputs of the reference's own forward (gen_golden.py::gen_backbone); weights are re-derived
    from the seed on both sides.  Also the text tower (encode_text) and calculate_score on those features."""
    from hybridgl_b200 import synth
    from hybridgl_b200.backbone import CLIPViTFM
    g = golden("backbone")
    seed, n, res, mb, h, w = (int(v) for v in g[f"{tag}_meta"])
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    m = CLIPViTFM(name, device=DEV)
    shapes = {k: tuple(v.shape) for k, v in m.model.state_dict().items()}
    sd = synth.seeded_clip_state_dict(shapes, seed)
    m.load_clip_state_dict(sd)
    loc, glo = synth.seeded_images(seed, n, res)
    masks = unpack_masks(g[f"{tag}_masks"], w)
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)  # noqa: E731
    for mode in modes:
        y = m(cu(loc), cu(glo), cu(masks), masking_block=mb, fusion_mode=mode)
        ref = g[f"{tag}_out/{mode}"]
        assert tuple(y.shape) == ref.shape
        # fp32 on the GPU (TF32 off) vs the reference's CPU fp32 through 12 / 24 blocks: 1e-3 of the feature scale
        np.testing.assert_allclose(y.float().cpu().numpy(), ref, rtol=2e-3, atol=2e-3 * float(np.abs(ref).mean()))
    txt = m.model.encode_text(cu(g[f"{tag}_tokens"]))
    np.testing.assert_allclose(txt.cpu().numpy(), g[f"{tag}_text"], rtol=2e-3, atol=1e-4)
    score = m.calculate_score(cu(g[f"{tag}_out/{modes[-1]}"]), cu(g[f"{tag}_text"])).cpu().numpy()
    np.testing.assert_allclose(score, g[f"{tag}_score"], rtol=1e-3, atol=1e-3)
    half = m.calculate_score(cu(g[f"{tag}_out/{modes[-1]}"]).half(), cu(g[f"{tag}_text"]).half()).cpu().numpy()      # fp16 features are accepted
    np.testing.assert_allclose(half, g[f"{tag}_score"], rtol=2e-2, atol=2e-2)


@pytest.mark.parametrize("M,L1,heads,hd,dtype", [(5, 197, 12, 64, torch.float32), (3, 577, 16, 64, torch.bfloat16), (2, 17, 1, 64, torch.float32), (4, 50, 8, 32, torch.bfloat16)])
def test_cls_attention_kernel_vs_torch(M, L1, heads, hd, dtype):
    """hgl_cls_attention against softmax(q0 k^T / sqrt(hd) + bias) v computed by torch in f32."""
    from hybridgl_b200 import ops
    torch.manual_seed(M * L1)
    qkv = torch.randn(M, L1, 3 * heads * hd, device=DEV).to(dtype)
    bias = torch.zeros(M, L1, device=DEV)
    bias[:, 1:][torch.rand(M, L1 - 1, device=DEV) < 0.5] = float("-inf")
    got = ops.cls_attention(qkv, bias, heads).float()
    v5 = qkv.float().view(M, L1, 3, heads, hd)
    q0, k, v = v5[:, 0, 0], v5[:, :, 1], v5[:, :, 2]
    s = torch.einsum("mhd,mjhd->mhj", q0, k) / hd ** 0.5 + bias[:, None, :]
    ref = torch.einsum("mhj,mjhd->mhd", torch.softmax(s, -1), v)
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    torch.testing.assert_close(got, ref, rtol=tol, atol=tol)
    none = ops.cls_attention(qkv, None, heads).float()
    ref0 = torch.einsum("mhj,mjhd->mhd", torch.softmax(torch.einsum("mhd,mjhd->mhj", q0, k) / hd ** 0.5, -1), v)
    torch.testing.assert_close(none, ref0, rtol=tol, atol=tol)


@pytest.mark.parametrize("M,L1,Dv,De,dtype", [(100, 197, 768, 512, torch.float32), (7, 5, 64, 32, torch.float32), (33, 577, 1024, 768, torch.bfloat16)])
def test_cls_head_kernel_vs_torch(M, L1, Dv, De, dtype):
    """hgl_cls_head == F.layer_norm(x[:, 0, :]) @ proj (f32 reference), plain and accumulated."""
    from hybridgl_b200 import ops
    torch.manual_seed(Dv)
    x = torch.randn(M, L1, Dv, device=DEV).to(dtype)
    gamma = (1 + 0.1 * torch.randn(Dv, device=DEV)).to(dtype); beta = (0.1 * torch.randn(Dv, device=DEV)).to(dtype)
    proj = (torch.randn(Dv, De, device=DEV) * Dv ** -0.5).to(dtype)
    ref = torch.nn.functional.layer_norm(x[:, 0, :].float(), (Dv,), gamma.float(), beta.float(), 1e-5) @ proj.float()
    got = ops.cls_head(x, gamma, beta, proj, 1e-5)
    tol = 2e-5 if dtype == torch.float32 else 2e-2
    torch.testing.assert_close(got, ref, rtol=tol, atol=tol)
    ops.cls_head(x, gamma, beta, proj, 1e-5, out=got, accumulate=True)
    torch.testing.assert_close(got, 2 * ref, rtol=tol, atol=2 * tol)


@pytest.mark.parametrize("M,L1,D,dtype", [(5, 17, 64, torch.float32), (3, 197, 768, torch.bfloat16), (2, 577, 1024, torch.bfloat16),
                                          (4, 50, 768, torch.float32), (1, 2, 2048, torch.float32)])
def test_token_mask_fuse_ln_equals_fuse_then_layernorm(M, L1, D, dtype):
    """hgl_token_mask_fuse_ln (f1): the fused stream is bit-identical to hgl_token_mask_fuse, its LayerNorm equals CLIP's fp32
    LayerNorm (clip/model.py:188-195) of the STORED stream; also the plain copy + LayerNorm form and slice outputs."""
    from hybridgl_b200 import ops
    gen = torch.Generator(device=DEV).manual_seed(M * 1000 + D)
    src = torch.randn((M, L1, D), generator=gen, device=DEV).to(dtype)
    add = torch.randn((M, L1, D), generator=gen, device=DEV).to(dtype)
    grid = torch.rand((M, L1 - 1), generator=gen, device=DEV)
    grid[:, ::3] = 0.0
    gam = 1 + 0.1 * torch.randn(D, generator=gen, device=DEV); bet = 0.1 * torch.randn(D, generator=gen, device=DEV)
    tol = dict(rtol=2e-5, atol=2e-5) if dtype == torch.float32 else dict(rtol=1.6e-2, atol=1.6e-2)
    for a_, g_, a, b in ((add, grid, 2.0, 1.0), (add, None, 2.0, 1.0), (None, grid, 1.0, 0.0), (None, None, 1.0, 0.0)):
        ref_x = ops.token_mask_fuse(src, a_, g_, a, b, layout="NLD")
        ref_h = torch.nn.functional.layer_norm(ref_x.float(), (D,), gam, bet, 1e-5)
        big_x = torch.zeros((M + 2, L1, D), dtype=dtype, device=DEV); big_h = torch.zeros_like(big_x)
        x, h = ops.token_mask_fuse_ln(src, a_, g_, a, b, gam, bet, 1e-5, out_x=big_x[1:M + 1], out_ln=big_h[1:M + 1])
        assert torch.equal(x, ref_x)
        torch.testing.assert_close(h.float(), ref_h, **tol)
        if dtype == torch.bfloat16:                                  # off by at most one bf16 step from the rounded fp32 LayerNorm
            assert (h.float() - ref_h.to(dtype).float()).abs().max() <= 2.0 ** -6 * max(1.0, float(ref_h.abs().max()))
        assert not big_x[0].any() and not big_x[M + 1].any() and not big_h[0].any() and not big_h[M + 1].any()
        x2, h2 = ops.token_mask_fuse_ln(src, a_, g_, a, b, gam, bet, 1e-5, want_x=False)
        assert x2 is None and torch.equal(h2, h)


def test_fused_and_unfused_ln_forward_agree(model, golden):
    """The hybrid forward with hgl_token_mask_fuse_ln (default) against the same forward with hgl_token_mask_fuse + torch.cat +
    the block's own LayerNorm, every fusion mode of the toy model."""
    g = golden("forward")
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)  # noqa: E731
    loc, glo, masks = cu(g["local"]), cu(g["global"]), cu(g["masks"])
    for mode in ("G2L", "L2G", "G2L&L2G", "token_masking"):
        with torch.no_grad():
            model.fused_ln = True
            a = model(loc, glo, masks, masking_block=1, fusion_mode=mode)
            model.fused_ln = False
            try:
                b = model(loc, glo, masks, masking_block=1, fusion_mode=mode)
            finally:
                model.fused_ln = True
        torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-5)
