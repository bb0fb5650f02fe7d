"""Byte-level BPE tokenizer for CLIP prompts (host-side string work; stands in for the reference's
clip/simple_tokenizer.py:62-132 and is checked against it in tests/test_host_logic.py).

The 49 408-entry vocabulary is OpenAI CLIP's merge table ``bpe_simple_vocab_16e6.txt.gz``. It is a data file,
not shipped in this repo; it is looked up at (1) $PROTOCLIP_BPE_VOCAB, (2) next to this file, (3) the reference
checkout's clip/ directory, (4) ~/.cache/clip/. Without it ``tokenize`` raises; the encoders themselves only
need token ids.
"""
from __future__ import annotations

import gzip
import html
import os
from functools import lru_cache
from typing import Dict, List, Tuple

import regex

VOCAB_FILE = "bpe_simple_vocab_16e6.txt.gz"
N_MERGES = 49152 - 256 - 2  # merge rules actually used (the file has a header line and trailing extras)


def find_vocab() -> str:
    here = os.path.dirname(os.path.abspath(__file__))
    candidates = [os.environ.get("PROTOCLIP_BPE_VOCAB", ""), os.path.join(here, VOCAB_FILE),
                  os.path.join(os.environ.get("PROTOCLIP_REFERENCE_ROOT", "/root/reference"), "clip", VOCAB_FILE),
                  os.path.expanduser(os.path.join("~/.cache/clip", VOCAB_FILE))]
    for c in candidates:
        if c and os.path.isfile(c):
            return c
    raise RuntimeError(f"CLIP BPE vocabulary {VOCAB_FILE} not found; set PROTOCLIP_BPE_VOCAB to its path "
                       f"(searched: {[c for c in candidates if c]})")


@lru_cache()
def byte_alphabet() -> Dict[int, str]:
    """Reversible byte -> printable unicode character table (printable latin-1 bytes map to themselves, the rest
    are shifted above U+0100), so BPE never sees whitespace / control bytes."""
    keep = list(range(ord("!"), ord("~") + 1)) + list(range(0xA1, 0xAD)) + list(range(0xAE, 0x100))
    table = {b: chr(b) for b in keep}  # insertion order = vocabulary order: printable bytes first
    extra = 0
    for b in range(256):
        if b not in table:
            table[b] = chr(256 + extra)
            extra += 1
    return table


def _clean(text: str) -> str:
    try:  # ftfy is optional; it is the identity on the ASCII class names / templates used here
        import ftfy
        text = ftfy.fix_text(text)
    except ImportError:
        pass
    text = html.unescape(html.unescape(text))
    return regex.sub(r"\s+", " ", text.strip()).strip()


class BPETokenizer:
    WORD_RE = regex.compile(
        r"<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+", regex.IGNORECASE)

    def __init__(self, vocab_path: str = ""):
        path = vocab_path or find_vocab()
        with gzip.open(path, "rt", encoding="utf-8") as f:
            lines = f.read().split("\n")
        merges: List[Tuple[str, str]] = [tuple(l.split()) for l in lines[1:1 + N_MERGES]]  # type: ignore[misc]
        alphabet = list(byte_alphabet().values())
        symbols = alphabet + [a + "</w>" for a in alphabet] + ["".join(m) for m in merges]
        symbols += ["<|startoftext|>", "<|endoftext|>"]
        self.encoder = {s: i for i, s in enumerate(symbols)}
        self.decoder = {i: s for s, i in self.encoder.items()}
        self.rank = {m: i for i, m in enumerate(merges)}
        self._cache = {"<|startoftext|>": ("<|startoftext|>",), "<|endoftext|>": ("<|endoftext|>",)}

    def _merge_word(self, token: str) -> Tuple[str, ...]:
        if token in self._cache:
            return self._cache[token]
        parts: List[str] = list(token[:-1]) + [token[-1] + "</w>"]
        while len(parts) > 1:
            best, best_rank = -1, None
            for i in range(len(parts) - 1):
                r = self.rank.get((parts[i], parts[i + 1]))
                if r is not None and (best_rank is None or r < best_rank):
                    best, best_rank = i, r
            if best_rank is None:
                break
            a, b = parts[best], parts[best + 1]
            merged, i = [], 0
            while i < len(parts):  # merge every occurrence of the best pair, left to right
                if i < len(parts) - 1 and parts[i] == a and parts[i + 1] == b:
                    merged.append(a + b)
                    i += 2
                else:
                    merged.append(parts[i])
                    i += 1
            parts = merged
        out = tuple(parts)
        self._cache[token] = out
        return out

    def encode(self, text: str) -> List[int]:
        ids: List[int] = []
        table = byte_alphabet()
        for word in self.WORD_RE.findall(_clean(text).lower()):
            mapped = "".join(table[b] for b in word.encode("utf-8"))
            ids.extend(self.encoder[s] for s in self._merge_word(mapped))
        return ids

    def decode(self, ids) -> str:
        inv = {c: b for b, c in byte_alphabet().items()}
        text = "".join(self.decoder[int(i)] for i in ids)
        return bytearray(inv.get(c, 32) for c in text.replace("</w>", " ")).decode("utf-8", errors="replace")
