"""CLIP byte-level BPE tokenizer (host-side input prep; reference: CLIP/clip/simple_tokenizer.py:66-141 and
`tokenize`, clip_explainability.py:237-273). Own implementation over the merge table shipped as a data asset
(`assets/clip_bpe_merges.txt.xz`); ftfy is not available offline, so text cleaning is html-unescape + strip
(identity for the plain ASCII labels this path is used with)."""
from __future__ import annotations

import html
import lzma
from functools import lru_cache
from pathlib import Path
from typing import List, Sequence, Union

import regex as re
import torch

_ASSET = Path(__file__).resolve().parent.parent / "assets" / "clip_bpe_merges.txt.xz"
SOT, EOT = "<|startoftext|>", "<|endoftext|>"


@lru_cache()
def _byte_alphabet():
    """Reversible byte -> printable unicode map used by GPT-2/CLIP BPE."""
    keep = list(range(ord("!"), ord("~") + 1)) + list(range(ord("¡"), ord("¬") + 1)) + list(range(ord("®"), ord("ÿ") + 1))
    chars = keep[:]
    extra = 0
    for b in range(256):
        if b not in keep:
            keep.append(b)
            chars.append(256 + extra)
            extra += 1
    return {b: chr(c) for b, c in zip(keep, chars)}


class ClipTokenizer:
    def __init__(self, merges_path: Path = _ASSET):
        lines = lzma.open(merges_path).read().decode("utf-8").split("\n")
        merges = [tuple(l.split()) for l in lines[1:] if l]  # line 0 is a provenance comment
        assert len(merges) == 49152 - 256 - 2, len(merges)
        alphabet = list(_byte_alphabet().values())
        vocab = alphabet + [c + "</w>" for c in alphabet] + ["".join(m) for m in merges] + [SOT, EOT]
        self.encoder = {tok: i for i, tok in enumerate(vocab)}
        self.rank = {m: i for i, m in enumerate(merges)}
        self.byte_map = _byte_alphabet()
        self.splitter = re.compile(
            r"""<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+""",
            re.IGNORECASE,
        )
        self._cache = {SOT: [SOT], EOT: [EOT]}

    def _merge_word(self, token: str) -> List[str]:
        if token in self._cache:
            return self._cache[token]
        parts = list(token[:-1]) + [token[-1] + "</w>"]
        while len(parts) > 1:
            best, best_rank = None, None
            for i in range(len(parts) - 1):
                rk = self.rank.get((parts[i], parts[i + 1]))
                if rk is not None and (best_rank is None or rk < best_rank):
                    best, best_rank = (parts[i], parts[i + 1]), rk
            if best is None:
                break
            merged, i = [], 0
            while i < len(parts):
                if i + 1 < len(parts) and (parts[i], parts[i + 1]) == best:
                    merged.append(parts[i] + parts[i + 1])
                    i += 2
                else:
                    merged.append(parts[i])
                    i += 1
            parts = merged
        self._cache[token] = parts
        return parts

    def encode(self, text: str) -> List[int]:
        text = html.unescape(html.unescape(text)).strip()
        text = re.sub(r"\s+", " ", text).strip().lower()
        ids: List[int] = []
        for word in self.splitter.findall(text):
            mapped = "".join(self.byte_map[b] for b in word.encode("utf-8"))
            ids.extend(self.encoder[p] for p in self._merge_word(mapped))
        return ids


@lru_cache()
def _default_tokenizer() -> ClipTokenizer:
    return ClipTokenizer()


def tokenize(texts: Union[str, Sequence[str]], context_length: int = 77, truncate: bool = False) -> torch.LongTensor:
    """Same contract as the reference `tokenize`: [n, context_length] int64, SOT + ids + EOT, zero padded."""
    if isinstance(texts, str):
        texts = [texts]
    tk = _default_tokenizer()
    sot, eot = tk.encoder[SOT], tk.encoder[EOT]
    out = torch.zeros(len(texts), context_length, dtype=torch.long)
    for i, t in enumerate(texts):
        ids = [sot] + tk.encode(t) + [eot]
        if len(ids) > context_length:
            if not truncate:
                raise RuntimeError(f"Input {t} is too long for context length {context_length}")
            ids = ids[:context_length]
            ids[-1] = eot
        out[i, : len(ids)] = torch.tensor(ids)
    return out
