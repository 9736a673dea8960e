"""CLIP's text-side entry points the reference's evaluation loop calls: `clip.tokenize` (third_party/modified_CLIP/clip/clip.py:197-237,
byte-level BPE of clip/simple_tokenizer.py) -- `Model.model.encode_text` lives in backbone._Clip.

    from hybridgl_b200.clip_text import tokenize          # drop-in for clip.tokenize at Hybridgl_main.py:146,147,160

Byte-pair encoding, restated from the published algorithm (Sennrich et al.; the GPT-2 / CLIP byte-level variant):
text is cleaned and lower-cased, split by CLIP's token pattern, every piece is mapped byte by byte onto printable unicode
code points, the last symbol gets the end-of-word marker `</w>`, and adjacent symbol pairs are merged in the order of the
merge table until no listed pair is left.  Vocabulary: 256 byte symbols, the same 256 with `</w>`, one entry per merge,
`<|startoftext|>`, `<|endoftext|>` (49 408 entries for CLIP's table).

The merge table (`bpe_simple_vocab_16e6.txt.gz`, 1.3 MB) is CLIP's data file and is NOT part of this repository.  It is looked up in:
  1. $HGL_CLIP_BPE
  2. hybridgl_b200/bpe_simple_vocab_16e6.txt.gz                                  (install step: copy it next to this file)
  3. $HGL_REFERENCE_ROOT/third_party/modified_CLIP/clip/bpe_simple_vocab_16e6.txt.gz   (default /root/reference: the HybridGL checkout)
"""
from __future__ import annotations

import gzip
import html
import os
from functools import lru_cache
from typing import List, Union

import regex
import torch

_N_MERGES = 49152 - 256 - 2          # merges CLIP keeps from the table (the file holds more)
_PATTERN = r"""<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+"""


def bpe_path() -> str:
    here = os.path.dirname(os.path.abspath(__file__))
    cands = [os.environ.get("HGL_CLIP_BPE"), os.path.join(here, "bpe_simple_vocab_16e6.txt.gz"),
             os.path.join(os.environ.get("HGL_REFERENCE_ROOT", "/root/reference"), "third_party", "modified_CLIP", "clip",
                          "bpe_simple_vocab_16e6.txt.gz")]
    for c in cands:
        if c and os.path.exists(c):
            return c
    raise FileNotFoundError("CLIP's BPE merge table bpe_simple_vocab_16e6.txt.gz was not found; set HGL_CLIP_BPE or copy the file next to "
                            "hybridgl_b200/clip_text.py (INTEGRATION.md, 'text side')")


@lru_cache()
def _byte_symbols():
    """256 bytes -> printable code points: the printable latin-1 ranges map to themselves, the rest to 256, 257, ..."""
    keep = list(range(ord("!"), ord("~") + 1)) + list(range(0xA1, 0xAC + 1)) + list(range(0xAE, 0xFF + 1))
    table, nxt = {}, 0
    for b in keep:
        table[b] = chr(b)
    for b in range(256):
        if b not in table:
            table[b] = chr(256 + nxt)
            nxt += 1
    return table, keep + [b for b in range(256) if b not in keep]


class BPETokenizer:
    def __init__(self, path: str = None):
        path = path or bpe_path()
        with gzip.open(path, "rt", encoding="utf-8") as f:
            rows = f.read().split("\n")
        merges = [tuple(r.split()) for r in rows[1:1 + _N_MERGES]]
        table, order = _byte_symbols()
        base = [table[b] for b in order]
        vocab = base + [s + "</w>" for s in base] + ["".join(m) for m in merges] + ["<|startoftext|>", "<|endoftext|>"]
        self.byte_sym = table
        self.encoder = {s: i for i, s in enumerate(vocab)}
        self.decoder = {i: s for s, i in self.encoder.items()}
        self.rank = {m: i for i, m in enumerate(merges)}
        self.pat = regex.compile(_PATTERN, regex.IGNORECASE)
        self._memo = {"<|startoftext|>": ["<|startoftext|>"], "<|endoftext|>": ["<|endoftext|>"]}

    def _merge(self, piece: str) -> List[str]:
        """Symbols of one pre-token after all applicable merges."""
        hit = self._memo.get(piece)
        if hit is not None:
            return hit
        syms = list(piece[:-1]) + [piece[-1] + "</w>"]
        while len(syms) > 1:
            best, best_rank = None, None
            for a, b in zip(syms, syms[1:]):
                r = self.rank.get((a, b))
                if r is not None and (best_rank is None or r < best_rank):
                    best, best_rank = (a, b), r
            if best is None:
                break
            out, i = [], 0
            while i < len(syms):
                if i + 1 < len(syms) and (syms[i], syms[i + 1]) == best:
                    out.append(syms[i] + syms[i + 1]); i += 2
                else:
                    out.append(syms[i]); i += 1
            syms = out
        self._memo[piece] = syms
        return syms

    @staticmethod
    def clean(text: str) -> str:
        try:                                   # the reference runs ftfy.fix_text first; optional here (absent offline)
            import ftfy
            if hasattr(ftfy, "fix_text"):
                text = ftfy.fix_text(text)
        except ImportError:
            pass
        text = html.unescape(html.unescape(text)).strip()
        return regex.sub(r"\s+", " ", text).strip().lower()

    def encode(self, text: str) -> List[int]:
        ids = []
        for piece in self.pat.findall(self.clean(text)):
            mapped = "".join(self.byte_sym[b] for b in piece.encode("utf-8"))
            ids.extend(self.encoder[s] for s in self._merge(mapped))
        return ids

    def decode(self, ids) -> str:
        inv = {c: b for b, c in self.byte_sym.items()}
        text = "".join(self.decoder[int(i)] for i in ids)
        return bytearray(inv[c] for c in text.replace("</w>", " ") if c in inv).decode("utf-8", errors="replace")


@lru_cache()
def _tokenizer() -> BPETokenizer:
    return BPETokenizer()


def tokenize(texts: Union[str, List[str]], context_length: int = 77, truncate: bool = False) -> torch.Tensor:
    """clip.tokenize (clip/clip.py:197-237): int32 [len(texts), context_length], <|startoftext|> ... <|endoftext|>, zero padded.
    Raises RuntimeError for over-long input unless truncate=True (then the last kept token becomes <|endoftext|>)."""
    if isinstance(texts, str):
        texts = [texts]
    tk = _tokenizer()
    sot, eot = tk.encoder["<|startoftext|>"], tk.encoder["<|endoftext|>"]
    out = torch.zeros(len(texts), context_length, dtype=torch.int)
    for i, t in enumerate(texts):
        ids = [sot] + tk.encode(t) + [eot]
        if len(ids) > context_length:
            if not truncate:
                raise RuntimeError(f"Input {t} is too long for context length {context_length}")
            ids = ids[:context_length]
            ids[-1] = eot
        out[i, :len(ids)] = torch.tensor(ids)
    return out
