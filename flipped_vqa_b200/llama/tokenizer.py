"""Tokenizer surface the model needs (`/root/reference/llama/tokenizer.py:14-42`) and the prompt builders of the
three objectives (`encode_vqa` / `encode_vaq` / `encode_qav`, `tokenizer.py:44-211`).

The model itself only reads `n_words`, `eos_id`, `a_token_id`, `q_token_id` (`llama/model.py:201-204`). With a real
`tokenizer.model` this class wraps SentencePiece exactly like the reference; `SyntheticTokenizer` stands in when no
tokenizer file exists (tests, benchmarks). The prompt builders live in `PromptBuilder`, which only needs an object
with `.encode(str) -> List[int]` as `sp_model`, so they are testable without a SentencePiece model file. The subtitle
(`--sub`) builders of TVQA / VLEP (`encode_dvqa` / `encode_dvaq` / `encode_dqav`, `tokenizer.py:218-302`) are mirrored too."""
from __future__ import annotations

import os
from typing import List


class PromptBuilder:
    """Token sequences of the three objectives. Layouts (F = max_feats video placeholders, id -2):

      vqa : [bos] enc(instr 'Video:') | F x -2 | nl | enc(question [choices] 'Answer: The answer is ' <option>) eos
      vaq : [bos] enc(instr 'Video:') | F x -2 | nl | enc([choices] answer-part <option> '\n' question) eos
      qav : [bos] enc(instr question [choices] answer-part <option> '\nVideo:') | F x -2 | eos

    `split == 'train'` builds ONE sequence (the correct option), otherwise one per option (`tokenizer.py:69-77`).
    Multiple-choice runs (`args.is_generation_task` false) append `answer_mapping[k]` ('(A)' ...) and include the choices
    text; generation-task runs append the option TEXT and omit the choices (`tokenizer.py:79-101,132-164,192-211`).
    prefix_index = first position the loss reads: the 'Answer' piece + 5 (vqa), the 'Question' piece + 2 (vaq), the
    'Video' piece + 2 (qav), located in the sequence of the correct option (vaq in generation mode: of option 0)."""

    _INSTR = {"vqa": "Instruction: Predict the answer based on the video and question.\n",
              "vaq": "Instruction: Predict the question based on the video and answer.\n",
              "qav": "Instruction: Predict the video based on the question and answer.\n"}

    def _gen(self) -> bool:
        return bool(getattr(self.args, "is_generation_task", False))

    def _option_strings(self, split, answer_mapping, answer, options):
        pool = list(options) if self._gen() else list(answer_mapping.values())
        if split == "train":
            return [options[answer] if self._gen() else answer_mapping[answer]]
        return pool

    def _video_head(self, task: str):
        head = [self.bos_id] + self.sp_model.encode(self._INSTR[task] + "Video:")
        return head, len(head)

    def encode_vqa(self, text=None, max_feats: int = 10, split: str = "train", answer_mapping=None, answer=None, options=None):
        head, video_start = self._video_head("vqa")
        stem = text["q_text"] + ("" if self._gen() else text["o_text"]) + text["a_text"]
        seqs = [head + [-2] * max_feats + [self.nl_id] + self.sp_model.encode(stem + o) + [self.eos_id]
                for o in self._option_strings(split, answer_mapping, answer, options)]
        ref = seqs[0] if split == "train" else seqs[answer]
        return seqs, ref.index(self.a_token_id) + 5, video_start

    def encode_vaq(self, text=None, max_feats: int = 10, split: str = "train", answer_mapping=None, answer=None, options=None):
        head, video_start = self._video_head("vaq")
        q = text["q_text"].strip()
        stem = ("\n" + text["a_text"]) if self._gen() else (text["o_text"] + text["a_text"])
        seqs = [head + [-2] * max_feats + [self.nl_id] + self.sp_model.encode(stem + o + "\n" + q) + [self.eos_id]
                for o in self._option_strings(split, answer_mapping, answer, options)]
        ref = seqs[0] if (split == "train" or self._gen()) else seqs[answer]
        return seqs, ref.index(self.q_token_id) + 2, video_start

    def encode_qav(self, text=None, max_feats: int = 10, split: str = "train", answer_mapping=None, answer=None, options=None):
        stem = self._INSTR["qav"] + text["q_text"] + ("" if self._gen() else text["o_text"]) + text["a_text"]
        seqs = [[self.bos_id] + self.sp_model.encode(stem + o + "\n" + "Video:") + [-2] * max_feats + [self.eos_id]
                for o in self._option_strings(split, answer_mapping, answer, options)]
        ref = seqs[0] if split == "train" else seqs[answer]
        return seqs, ref.index(self.v_token_id) + 2


    # ---- dialogue (`--sub`) variants, `tokenizer.py:218-302`: the subtitle text d_text sits between the video block and the question
    # (vqa / vaq) or between the instruction and the question (qav). They always append `answer_mapping[k]`, return the token spans
    # of the dialogue (prefix_i = its first position, prefix_main = first position after it) so that the dataset can cut the
    # dialogue when the sequence overflows max_seq_len (`dataloader/tvqa.py:75-108`), and locate prefix_index in sequence 0.
    _INSTR_D = {"vqa": "Instruction: Predict the answer based on the dialogue, video and question.\n",
                "vaq": "Instruction: Predict the question based on the dialogue, video and answer.\n",
                "qav": "Instruction: Predict the video based on the dialogue, question and answer.\n"}

    def _mapped(self, split, answer_mapping, answer):
        return [answer_mapping[answer]] if split == "train" else list(answer_mapping.values())

    def encode_dvqa(self, text=None, max_feats: int = 10, split: str = "train", answer_mapping=None, answer=None):
        head = [self.bos_id] + self.sp_model.encode(self._INSTR_D["vqa"] + "Video:")
        video_start = len(head)
        prefix_i = video_start + max_feats + 1
        dialogue = self.sp_model.encode(text["d_text"])
        stem = text["q_text"] + text["o_text"] + text["a_text"]
        seqs = [head + [-2] * max_feats + [self.nl_id] + dialogue + self.sp_model.encode(stem + v) + [self.eos_id]
                for v in self._mapped(split, answer_mapping, answer)]
        return seqs, len(seqs[0]) - 4, video_start, prefix_i, prefix_i + len(dialogue)

    def encode_dvaq(self, text=None, max_feats: int = 10, split: str = "train", answer_mapping=None, answer=None):
        head = [self.bos_id] + self.sp_model.encode(self._INSTR_D["vaq"] + "Video:")
        video_start = len(head)
        prefix_i = video_start + max_feats + 1
        dialogue = self.sp_model.encode(text["d_text"])
        stem, q = text["o_text"] + text["a_text"], text["q_text"].strip()
        seqs = [head + [-2] * max_feats + [self.nl_id] + dialogue + self.sp_model.encode(stem + v + "\n" + q) + [self.eos_id]
                for v in self._mapped(split, answer_mapping, answer)]
        return seqs, seqs[0].index(self.q_token_id) + 2, video_start, prefix_i, prefix_i + len(dialogue)

    def encode_dqav(self, text=None, max_feats: int = 10, max_seq_len: int = 128, split: str = "train", answer_mapping=None, answer=None):
        head = [self.bos_id] + self.sp_model.encode(self._INSTR_D["qav"])
        dialogue = self.sp_model.encode(text["d_text"])
        stem = text["q_text"] + text["o_text"] + text["a_text"]
        seqs = [head + dialogue + self.sp_model.encode(stem + v + "\n" + "Video:") + [-2] * max_feats + [self.eos_id]
                for v in self._mapped(split, answer_mapping, answer)]
        return seqs, len(seqs[0]) - max_feats - 1, len(head), len(head) + len(dialogue)


class Tokenizer(PromptBuilder):
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

    def __init__(self, n_words: int = 32000, a_token_id: int = 22550):
        self.n_words, self.bos_id, self.eos_id, self.pad_id = n_words, 1, 2, -1
        self.v_token_id, self.q_token_id, self.a_token_id, self.nl_id = 15167, 16492, a_token_id, 13

    def decode(self, t):
        return ""
