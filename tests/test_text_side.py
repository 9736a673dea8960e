"""The text side of the drop-in boundary: `tokenize` == clip.tokenize and `_Clip.encode_text` == CLIP.encode_text
(third_party/modified_CLIP/clip/clip.py:197-237, clip/model.py:414-431), against outputs recorded from the reference itself
(tests/golden/backbone.npz, gen_golden.py::gen_backbone; weights re-derived from the seed on both sides).  Pure PyTorch: runs on the CPU."""
import os

import numpy as np
import pytest
import torch

from hybridgl_b200 import synth

VOCAB = os.path.join(os.environ.get("HGL_REFERENCE_ROOT", "/root/reference"), "third_party", "modified_CLIP", "clip", "bpe_simple_vocab_16e6.txt.gz")
needs_vocab = pytest.mark.skipif(not (os.path.exists(VOCAB) or os.environ.get("HGL_CLIP_BPE")),
                                 reason="CLIP's BPE merge table is not installed (INTEGRATION.md, text side)")


@needs_vocab
def test_tokenize_matches_reference_tokens(golden):
    from hybridgl_b200.backbone import tokenize
    g = golden("backbone")
    tok = tokenize([str(s) for s in g["sentences"]])
    assert tok.dtype == torch.int32 and tuple(tok.shape) == (len(g["sentences"]), 77)
    assert np.array_equal(tok.numpy(), g["b16_tokens"])
    assert np.array_equal(tokenize("man in black").numpy(), g["b16_tokens"][:1])          # a bare string is one row
    with pytest.raises(RuntimeError, match="too long"):
        tokenize("word " * 100)
    t = tokenize("word " * 100, truncate=True)
    assert int(t[0, -1]) == 49407 and int(t[0, 0]) == 49406                                # <|endoftext|> closes a truncated row


@pytest.mark.parametrize("tag,name", [("b16", "ViT-B/16")])
def test_encode_text_matches_reference(golden, tag, name):
    from hybridgl_b200.backbone import ARCH, _Clip
    g = golden("backbone")
    seed = int(g[f"{tag}_meta"][0])
    m = _Clip(dict(ARCH[name])).eval()
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items() if not k.startswith("visual.")}
    sd = synth.seeded_clip_state_dict(shapes, seed)
    missing, unexpected = m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
    assert not unexpected and all(k.startswith("visual.") for k in missing)
    txt = m.encode_text(torch.from_numpy(g[f"{tag}_tokens"]))
    np.testing.assert_allclose(txt.numpy(), g[f"{tag}_text"], rtol=1e-4, atol=1e-5)
    # target_noun_index picks position index + 1 instead of the <|endoftext|> position (clip/model.py:424-427)
    t2 = m.encode_text(torch.from_numpy(g[f"{tag}_tokens"][:2]), target_noun_index=2)
    assert tuple(t2.shape) == (2, 512) and not torch.allclose(t2, txt[:2])
