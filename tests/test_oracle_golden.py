"""Pin the numpy oracle against outputs of the reference itself (tests/golden/*.npz, produced by
tests/golden/gen_golden.py from /root/reference).  CPU only."""
import numpy as np
import pytest

from conftest import unpack_masks
from hybridgl_b200 import synth
from oracle import hybridgl_oracle as O


@pytest.mark.parametrize("tag", ["a", "b", "c", "d"])
def test_prep_matches_reference(golden, tag):
    g = golden("prep")
    seed, h, w, S, n = g[f"{tag}_meta"].tolist()
    it = synth.make_item(seed, h, w, n, 0, with_features=False)
    blur = O.gaussian_blur_u8(it.image)
    if f"{tag}_image" in g:
        assert np.array_equal(g[f"{tag}_image"], it.image)
        assert np.array_equal(unpack_masks(g[f"{tag}_masks"], w), it.masks)
        assert np.array_equal(g[f"{tag}_blur"], blur)          # cv2.GaussianBlur restated bit-exactly
        loc, glo = O.prep(it.image, blur, it.masks, S)
        if tag == "b":   # odd tiny frame: ATen takes a differently-contracted scalar path here (<= 3 ulp off)
            np.testing.assert_allclose(loc, g[f"{tag}_local"], rtol=0, atol=1e-6)
            np.testing.assert_allclose(glo, g[f"{tag}_global"], rtol=0, atol=1e-6)
        else:
            assert np.array_equal(loc, g[f"{tag}_local"])       # bit-exact (FMA order restated)
            assert np.array_equal(glo, g[f"{tag}_global"])
    else:
        assert g[f"{tag}_image_sum"].tolist() == [int(it.image.astype(np.int64).sum()), int(it.masks.sum())]
        assert np.array_equal(g[f"{tag}_blur"], blur[::8, ::8])
        loc, glo = O.prep(it.image, blur, it.masks, S)
        assert np.array_equal(loc[:, :, ::7, ::5], g[f"{tag}_local"])
        assert np.array_equal(glo[:, :, ::7, ::5], g[f"{tag}_global"])
        np.testing.assert_allclose([loc.astype(np.float64).sum(), glo.astype(np.float64).sum()], g[f"{tag}_sums"], rtol=1e-12)


@pytest.mark.parametrize("tag", ["a", "b", "c", "d", "e", "f"])
def test_grid_matches_reference(golden, tag):
    g = golden("grid")
    seed, h, w, gs, n = g[f"{tag}_meta"].tolist()
    masks = unpack_masks(g[f"{tag}_masks"], w)
    aa = O.mask_to_grid(masks, gs, antialias=True)
    na = O.mask_to_grid(masks, gs, antialias=False)
    # small outputs take a differently-contracted ATen loop (<= 2 ulp); taps and zero pattern are exact
    np.testing.assert_allclose(na, g[f"{tag}_noaa"], rtol=0, atol=2.4e-7)
    assert np.array_equal(na != 0, g[f"{tag}_noaa"] != 0)
    np.testing.assert_allclose(aa, g[f"{tag}_aa"], rtol=0, atol=1.2e-7)   # <= 1 ulp: ATen vectorises the tap sum
    assert np.array_equal(aa != 0, g[f"{tag}_aa"] != 0)                   # the zero pattern drives the attention mask
    if f"{tag}_attn_row0" in g:
        am = O.make_attn_mask(g[f"{tag}_aa"], heads=3)
        assert np.array_equal(am[:, 0, :], g[f"{tag}_attn_row0"])
        assert not am[:, 1:, :].any()


def test_dir_mask_and_relation(golden):
    g = golden("misc")
    for key in g.files:
        if key.startswith("dir_"):
            _, d, hw = key.split("_")
            h, w = map(int, hw.split("x"))
            assert np.array_equal(O.gen_dir_mask(d, h, w), g[key]), key
    for bi, bj, si, sj, word, ref in zip(g["rel_bi"], g["rel_bj"], g["rel_si"], g["rel_sj"], g["rel_word"], g["rel_out"]):
        got = O.relation_boxes(bi, bj, si, sj, str(word))
        np.testing.assert_allclose(got, ref, rtol=2e-7, atol=0)


def _case(g, ci):
    p = f"c{ci:02d}_"
    seed, h, w, n, n_other = g[p + "meta"].tolist()
    d = {k[len(p):]: g[k] for k in g.files if k.startswith(p)}
    d.update(h=h, w=w, n=n, n_other=n_other, rela=str(d["flags"][0]), dirf=str(d["flags"][1]))
    d["masks"] = unpack_masks(d["masks"], w); d["target"] = unpack_masks(d["target"], w)
    return d


def test_scoring_matches_reference(golden):
    g = golden("scoring")
    for ci in range(int(g["n_cases"])):
        c = _case(g, ci)
        it = synth.make_item(int(c["meta"][0]), c["h"], c["w"], c["n"], 1, de=512, n_other=c["n_other"],
                             dirflag=c["dirf"], relaflag=c["rela"])
        heat = c["heat_resized"] if "heat_resized" in c else O.resize_bilinear_aa(c["heat_raw"], c["h"], c["w"])[0]
        cond = O.condition_heatmap(heat, c["dirf"])
        sg = O.gem_pool(cond, c["masks"], O.black_for(c["rela"]))
        np.testing.assert_allclose(sg, c["score_gem"], rtol=2e-4, atol=2e-5)
        r = O.score_and_select(c["features"], c["sentence"], c["noun"], c["others"], c["boxes"], c["rela"],
                               score_gem=sg, logit_scale_exp=float(c["logit_scale_exp"]))
        np.testing.assert_allclose(r["score_clip"], c["score_clip"], rtol=1e-4, atol=1e-4)
        if c["n_other"]:
            np.testing.assert_allclose(r["score_neg"], c["score_neg"], rtol=1e-4, atol=1e-4)
        else:
            assert np.isnan(c["score_neg"]).all() and np.isnan(r["score_neg"]).all()   # Appendix B-5
        assert r["idx_hybrid"] == int(c["idx_hybrid"]), ci
        assert np.array_equal(r["top_idx"], c["top_idx"]), ci
        np.testing.assert_allclose(r["blended"], c["blended"], rtol=2e-4, atol=2e-5)
        assert r["idx_final"] == int(c["idx_final"]), ci
        i0, u0, iou0 = O.compute_iou(c["masks"][r["idx_hybrid"]], c["target"])
        i1, u1, iou1 = O.compute_iou(c["masks"][r["idx_final"]], c["target"])
        assert [i0, u0, i1, u1] == c["IU"].tolist()
        np.testing.assert_allclose([iou0, iou1], c["iou"], rtol=1e-6)


def test_heatmap_resize_aa_matches_reference(golden):
    g = golden("scoring")
    seen = 0
    for ci in range(int(g["n_cases"])):
        c = _case(g, ci)
        if "heat_resized" in c:
            got = O.resize_bilinear_aa(c["heat_raw"], c["h"], c["w"])[0]
            np.testing.assert_allclose(got, c["heat_resized"], rtol=0, atol=2e-7)
            seen += 1
    assert seen >= 10


def test_token_space_pooling_equals_pixel_space_pooling():
    """Pins O.mask_pool_tokens (the tensor-core kernel's oracle) to the pinned pixel-space O.gem_pool:
    if the heat-map is A = U h (U = bilinear up-sampling of a token-grid map h = F t), then for every mask
    sum_p m[p] A[p] == (U^T m) . h == ((U^T m) F) . t  -- SURVEY.md Appendix A-2."""
    rng = np.random.default_rng(5)
    gh, gw, h, w, d, n = 6, 8, 24, 32, 16, 5
    F = synth.bf16_round(rng.standard_normal((gh * gw, d)).astype(np.float32))
    t = rng.standard_normal(d).astype(np.float32)
    # U as an explicit [h*w, gh*gw] matrix: resize_bilinear applied to the basis maps
    basis = np.eye(gh * gw, dtype=np.float32).reshape(gh * gw, gh, gw)
    U = O.resize_bilinear(basis, h, w).reshape(gh * gw, h * w).T.astype(np.float64)
    masks = synth.make_masks(rng, n, h, w, min_area=20)
    heat = (U @ (F.astype(np.float64) @ t.astype(np.float64))).reshape(h, w)
    s_in_pixel = (heat[None] * masks).reshape(n, -1).sum(1)
    weights = (masks.reshape(n, -1).astype(np.float64) @ U).astype(np.float32)          # U^T m, one row per mask
    # exact identity with unrounded weights ...
    np.testing.assert_allclose((weights.astype(np.float64) @ F.astype(np.float64)) @ t, s_in_pixel, rtol=1e-5)
    # ... and the oracle (f32 weights as given, f32 pooled rows) within the north star's 1e-3
    pooled = O.mask_pool_tokens(weights, F, normalize=False)
    np.testing.assert_allclose(pooled.astype(np.float64) @ t, s_in_pixel, rtol=1e-3, atol=1e-3 * np.abs(s_in_pixel).max())
    # closed form of Hybridgl_main.py:220 from the pooled sums == gem_pool on the pixel map
    black = 1.8
    area = masks.reshape(n, -1).sum(1)
    ref = O.gem_pool(heat.astype(np.float32), masks, black)
    s_in = (weights.astype(np.float64) @ F.astype(np.float64)) @ t
    got = (2 - black) * s_in / area - black * (heat.sum() - s_in) / (h * w - area)
    np.testing.assert_allclose(got, ref, rtol=1e-4, atol=1e-5)
    # normalised rows have unit length
    pn = O.mask_pool_tokens(weights, F, normalize=True)
    np.testing.assert_allclose(np.linalg.norm(pn.astype(np.float64), axis=1), 1.0, rtol=1e-6)


def test_rle_codec_matches_sam(golden):
    """oracle mask_to_rle / rle_to_mask against SAM's own codec (amg.py:107-149) recorded by gen_golden.py::gen_rle."""
    g = golden("rle")
    for ci in range(int(g["n_cases"])):
        masks, counts, off = g[f"c{ci}_masks"], g[f"c{ci}_counts"], g[f"c{ci}_off"]
        c2, off2 = synth.masks_to_rle(masks)
        assert np.array_equal(c2, counts) and np.array_equal(off2, off)
        for i in range(masks.shape[0]):
            r = O.mask_to_rle(masks[i])
            assert r["counts"] == counts[off[i]:off[i + 1]].tolist() and r["size"] == list(masks.shape[1:])
            assert np.array_equal(O.rle_to_mask(r), g[f"c{ci}_decoded"][i])


def test_pack_bits_layout():
    rng = np.random.default_rng(3)
    m = rng.random((3, 5, 70)) < 0.4
    b = O.pack_bits(m)
    assert b.shape == (3, 5, 3) and b.dtype == np.uint32
    for (n, y, x) in [(0, 0, 0), (1, 2, 31), (2, 4, 32), (2, 3, 69)]:
        assert bool((b[n, y, x // 32] >> (x % 32)) & 1) == bool(m[n, y, x])
    assert (b[:, :, 2] >> 6).max() == 0            # bits past W are zero


def test_geometry_and_prompts_match_reference(golden):
    """oracle mask2chw / SAM boxes / apply_visual_prompts('blur' | 'black') against the reference's own functions
    (gen_golden.py::gen_geometry: utils.mask2chw, utils.apply_visual_prompts, amg.batched_mask_to_box + box_xyxy_to_xywh)."""
    from conftest import unpack_masks
    g = golden("geometry")
    for ci in range(int(g["n_cases"])):
        h, w, n = g[f"c{ci}_hw"].tolist()
        masks = unpack_masks(g[f"c{ci}_masks"], w)
        for i in range(n):
            (cy, cx), hh, ww = O.mask2chw(masks[i])
            assert [cy, cx, hh, ww] == g[f"c{ci}_chw"][i].tolist()
            assert O.mask_to_box_xywh(masks[i]).tolist() == g[f"c{ci}_boxes"][i].tolist()
            assert synth.masks_to_boxes(masks[i:i + 1])[0].tolist() == g[f"c{ci}_boxes"][i].tolist()     # the synthetic generator's boxes
        if f"c{ci}_image" in g.files:
            img = g[f"c{ci}_image"]
            blur = O.gaussian_blur_u8(img)
            for i in range(n):
                assert np.array_equal(O.apply_visual_prompt(img, masks[i], "blur", blur), g[f"c{ci}_blur"][i])
                assert np.array_equal(O.apply_visual_prompt(img, masks[i], "black"), g[f"c{ci}_black"][i])
                # the 'circle' prompt: cv2.ellipse at mask2chw's centre, alone and in the if-chain's order with the other two
                assert np.array_equal(O.apply_visual_prompt(img, masks[i], "circle"), g[f"c{ci}_circle"][i])
                assert np.array_equal(O.apply_visual_prompt(img, masks[i], ("blur", "circle"), blur), g[f"c{ci}_blur_circle"][i])
                assert np.array_equal(O.apply_visual_prompt(img, masks[i], ("circle", "black")), g[f"c{ci}_circle_black"][i])
    assert O.mask_to_box_xywh(np.zeros((8, 8), bool)).tolist() == g["empty_box"].tolist() == [0, 0, 0, 0]


def test_ellipse_outline_equals_cv2():
    """The restated rasteriser of the 'circle' prompt against OpenCV itself (a binary dependency of the reference: no source in its
    tree): random ellipses -- tiny, frame-sized, centred near the border so that edges are clipped -- and random integer lines."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    for t in range(400):
        H, W = int(rng.integers(20, 200)), int(rng.integers(20, 260))
        p = rng.integers(-30, [W + 30, H + 30, W + 30, H + 30])
        a = np.zeros((H, W), np.uint8)
        cv2.line(a, (int(p[0]), int(p[1])), (int(p[2]), int(p[3])), 255, 1, cv2.LINE_8, 0)
        b = np.zeros((H, W), bool)
        for y, x in O.line8_pixels(H, W, (p[0], p[1]), (p[2], p[3])):
            b[y, x] = True
        assert np.array_equal(a > 0, b), (H, W, p.tolist())
    for t in range(600):
        H, W = ((480, 640), (97, 131), (600, 800), (33, 32))[t % 4] if t % 3 == 0 else (int(rng.integers(20, 200)), int(rng.integers(20, 260)))
        cx, cy = int(rng.integers(0, W)), int(rng.integers(0, H))
        ax, ay = (int(rng.integers(0, 16)), int(rng.integers(0, 16))) if t % 5 == 0 else (int(rng.integers(0, W // 2 + 1)), int(rng.integers(0, H // 2 + 1)))
        ref = cv2.ellipse(np.zeros((H, W, 3), np.uint8), (cx, cy), (ax, ay), 0, 0, 360, (255, 0, 0), 1)[:, :, 0] > 0
        assert np.array_equal(ref, O.ellipse_outline(H, W, cx, cy, ax, ay)), (H, W, cx, cy, ax, ay)


def test_token_space_gem_pooling_equals_pixel_space_chain():
    """O.gem_pool_token_space (the oracle of hgl_gem_token_pool: masks resampled onto the raw GEM map's grid, affine conditioning as
    per-expression scalars) against the pinned pixel-space chain resize_bilinear_aa -> condition_heatmap -> gem_pool, 1e-3."""
    it = synth.make_item(3, 96, 128, 6, 6, de=32)
    it.masks[0] = False; it.masks[0, 40:44, 60:70] = True
    for ex, d, bl in zip(it.expressions, synth.DIRFLAGS, (1.8, 1.95, 1.5, 1.8, 1.8, 1.5)):
        full = O.resize_bilinear_aa(ex.heat_raw[None], 96, 128)[0]
        ref = O.gem_pool(O.condition_heatmap(full, d), it.masks, bl)
        np.testing.assert_allclose(O.gem_pool_token_space(ex.heat_raw, it.masks, d, bl), ref, rtol=1e-3, atol=1e-4)
