"""Deterministic stand-in for ``LlamaTokenizerWrapper`` (reference `modeling_minicpmv.py:404-438`).

No SentencePiece model ships with the reference and there is no network, so tests/benchmarks use this
character-level tokenizer. It exposes exactly the attributes the hot path reads
(`modeling_minicpmv.py:175-186,248-252,595-609`): special-token strings, ``add_bos_token``, ``bos_id``,
``im_start_id``/``im_end_id``, ``encode(str) -> List[int]``. Ids only index the embedding table.
"""
from __future__ import annotations

from typing import List


class StubTokenizer:
    im_start = "<image>"
    im_end = "</image>"
    unk_token = "<unk>"
    slice_start = "<slice>"
    slice_end = "</slice>"

    # id layout: 0 pad/unk, 1 bos, 2 eos, 3..6 image/slice markers, >= 8 characters
    _SPECIAL = {"<unk>": 0, "<image>": 3, "</image>": 4, "<slice>": 5, "</slice>": 6}
    _FIRST_CHAR_ID = 8

    def __init__(self, vocab_size: int = 122753, add_bos_token: bool = True):
        if vocab_size <= self._FIRST_CHAR_ID + 1:
            raise ValueError("vocab too small")
        self.vocab_size = vocab_size
        self.add_bos_token = add_bos_token
        self._ordered = sorted(self._SPECIAL, key=len, reverse=True)

    @property
    def bos_id(self) -> int:
        return 1

    @property
    def eos_id(self) -> int:
        return 2

    @property
    def unk_id(self) -> int:
        return 0

    @property
    def im_start_id(self) -> int:
        return self._SPECIAL[self.im_start]

    @property
    def im_end_id(self) -> int:
        return self._SPECIAL[self.im_end]

    def encode(self, text: str) -> List[int]:
        ids: List[int] = [self.bos_id] if self.add_bos_token else []
        i, n = 0, len(text)
        span = self.vocab_size - self._FIRST_CHAR_ID
        while i < n:
            for tok in self._ordered:
                if text.startswith(tok, i):
                    ids.append(self._SPECIAL[tok])
                    i += len(tok)
                    break
            else:
                ids.append(self._FIRST_CHAR_ID + (ord(text[i]) * 2654435761 % span))
                i += 1
        return ids
