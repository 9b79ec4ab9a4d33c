"""Tokenizer surface the model needs (`/root/reference/llama/tokenizer.py:14-42`).

The reference's prompt builders (`encode_vqa/vaq/qav`, `:44-302`) are CPU data preparation and out
of the accelerated path's scope; the model itself only reads
`n_words`, `eos_id`, `a_token_id`, `q_token_id` (`llama/model.py:201-204`). With a real
`tokenizer.model` this class wraps SentencePiece exactly like the reference; `SyntheticTokenizer`
stands in when no tokenizer file exists (tests, benchmarks)."""
from __future__ import annotations

import os
from typing import List


class Tokenizer:
    def __init__(self, model_path: str, args=None):
        self.args = args
        if not os.path.isfile(model_path):
            raise FileNotFoundError(f"{model_path}: no SentencePiece model; pass tokenizer=SyntheticTokenizer(n_words) for synthetic runs")
        from sentencepiece import SentencePieceProcessor
        self.sp_model = SentencePieceProcessor(model_file=model_path)
        self.n_words: int = self.sp_model.vocab_size()
        self.bos_id: int = self.sp_model.bos_id()
        self.eos_id: int = self.sp_model.eos_id()
        self.pad_id: int = self.sp_model.pad_id()
        # hard-coded ids of 'Video', 'Question', 'Answer' pieces and newline (`tokenizer.py:28-31`)
        self.v_token_id, self.q_token_id, self.a_token_id, self.nl_id = 15167, 16492, 22550, 13

    def encode(self, s: str, bos: bool, eos: bool) -> List[int]:
        t = self.sp_model.encode(s)
        return ([self.bos_id] if bos else []) + t + ([self.eos_id] if eos else [])

    def decode(self, t: List[int]) -> str:
        return self.sp_model.decode(t)


class SyntheticTokenizer:
    """Id-only stand-in (no text): vocabulary size plus the special ids the model stores."""

    def __init__(self, n_words: int = 32000):
        self.n_words, self.bos_id, self.eos_id, self.pad_id = n_words, 1, 2, -1
        self.v_token_id, self.q_token_id, self.a_token_id, self.nl_id = 15167, 16492, 22550, 13

    def decode(self, t):
        return ""
