"""Synthetic weights and batches for the LLaMA-VQA training step.

No datasets, tokenizer model or LLaMA checkpoints are available offline, so every test and the
benchmark run on synthetic inputs that imitate the reference's batch contract
(`/root/reference/dataloader/__init__.py:28-90`, `dataloader/base_dataset.py:30-174`) and the token
layout produced by `llama/tokenizer.py:44-211`.

Two generators live here:

* ``hash_normal`` – a pure integer-hash pseudo-normal generator (splitmix64 + Irwin-Hall). It is
  bit-reproducible on every platform/torch version, which lets the golden fixtures under
  ``tests/golden/`` store only *results* (losses, gradients) and regenerate the weights.
* ``synthetic_batch`` – the NExT-QA / DramaQA / TVQA-shaped batch dict.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch

_MASK64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _MASK64
    z = x
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _MASK64
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _MASK64
    return z ^ (z >> np.uint64(31))


def hash_normal(shape, seed: int, std: float = 1.0, mean: float = 0.0) -> torch.Tensor:
    """Approximately normal fp32 tensor from integer hashing only (exactly reproducible).

    Sum of four 16-bit uniforms (Irwin-Hall, n=4) centred and scaled to unit variance; the result
    is then rounded to bf16-representable values so that fp16, bf16 and fp32 consumers all see
    *identical* numbers (values below 2^-13 in magnitude are flushed to zero to stay inside fp16's
    normal range)."""
    n = int(np.prod(shape)) if len(shape) else 1
    with np.errstate(over="ignore"):
        idx = np.arange(n, dtype=np.uint64) + (np.uint64(seed) << np.uint64(32))
        h = _splitmix64(idx)
    s = np.zeros(n, dtype=np.int64)
    for k in range(4):
        s += ((h >> np.uint64(16 * k)) & np.uint64(0xFFFF)).astype(np.int64)
    # each uniform: mean 32767.5, var (65536^2-1)/12 ; sum of 4
    z = (s.astype(np.float64) - 4 * 32767.5) / math.sqrt(4 * (65536.0 ** 2 - 1) / 12.0)
    t = torch.from_numpy((z * std + mean).astype(np.float32)).reshape(shape)
    t = t.to(torch.bfloat16).to(torch.float32)
    t[t.abs() < 2.0 ** -13] = 0.0
    return t


AUDIO_DIM = 1024        # ImageBind audio features (`dataloader/base_dataset.py:13`)


def audio_mode(args) -> Optional[str]:
    """The five input-fusion branches of `llama/model.py:209-227,306-322` as one tag:
    None (video only) | 'audio_only' | 'concat' | 'sum' | 'attention'."""
    if not getattr(args, "audio", False):
        return None
    if getattr(args, "audio_only", False):
        return "audio_only"
    m = getattr(args, "audio_merge", "none")
    if m == "concat":
        return "concat"
    if m != "" and m in "sum":                     # the reference tests `audio_merge in 'sum'` (`model.py:215,313`)
        return "sum"
    if m == "attention":
        return "attention"
    return None


def synthetic_audio_state(params, mode: str, seed: int = 0, video_dim: int = 768) -> Dict[str, torch.Tensor]:
    """Extra / replaced parameters of an audio-fusion variant, reference names (`llama/model.py:209-227,145-150`)."""
    d = params.dim
    sd: Dict[str, torch.Tensor] = {}
    k = seed * 1000 + 900
    if mode in ("audio_only", "sum"):
        sd["audio_proj.weight"] = hash_normal((d, AUDIO_DIM), k + 1, 1.0 / math.sqrt(AUDIO_DIM))
    if mode == "attention":
        sd["audio_proj.weight"] = hash_normal((video_dim, AUDIO_DIM), k + 1, 1.0 / math.sqrt(AUDIO_DIM))
        for i, nm in enumerate(("query", "key", "value")):
            sd[f"video_audio_cross_attn.{nm}.weight"] = hash_normal((video_dim, video_dim), k + 2 + 2 * i, 1.0 / math.sqrt(video_dim))
            sd[f"video_audio_cross_attn.{nm}.bias"] = hash_normal((video_dim,), k + 3 + 2 * i, 0.05)
    if mode == "concat":
        sd["visual_proj.weight"] = hash_normal((d, video_dim + AUDIO_DIM), k + 8, 1.0 / math.sqrt(video_dim + AUDIO_DIM))
    return sd


def synthetic_audio(bsz: int, frames: int, seed: int = 0) -> torch.Tensor:
    g = torch.Generator(device="cpu")
    g.manual_seed(77_000 + seed)
    return torch.randn(bsz, frames, AUDIO_DIM, generator=g)


def synthetic_state_dict(params, seed: int = 0, max_feats: int = 10, bias: float = 3.5,
                         video_dim: int = 768) -> Dict[str, torch.Tensor]:
    """Random-init state dict with the reference's parameter names (`llama/model.py:190-248`,
    SURVEY.md §8(b)). Frozen ≥2-D weights ~ N(0, 0.02²); norm weights 1+N(0, 0.1²); adapter and
    temporal embeddings N(0,1); gate1 ~ N(0, 0.5²) (NOT zero: the reference's zero init makes every
    adapter gradient identically zero, SURVEY.md §7 'Hard parts'); gate2 = -bias + N(0, 0.1²)."""
    d, L, H, V = params.dim, params.n_layers, params.n_heads, params.vocab_size
    hid = ffn_hidden_dim(d, params.multiple_of)
    sd: Dict[str, torch.Tensor] = {}
    k = [seed * 1000 + 1]

    def nxt():
        k[0] += 1
        return k[0]

    sd["tok_embeddings.weight"] = hash_normal((V, d), nxt(), 0.02)
    sd["output.weight"] = hash_normal((V, d), nxt(), 0.02)
    sd["norm.weight"] = hash_normal((d,), nxt(), 0.1, 1.0)
    for i in range(L):
        p = f"layers.{i}."
        for nm in ("wq", "wk", "wv", "wo"):
            sd[p + f"attention.{nm}.weight"] = hash_normal((d, d), nxt(), 0.02)
        sd[p + "feed_forward.w1.weight"] = hash_normal((hid, d), nxt(), 0.02)
        sd[p + "feed_forward.w2.weight"] = hash_normal((d, hid), nxt(), 0.02)
        sd[p + "feed_forward.w3.weight"] = hash_normal((hid, d), nxt(), 0.02)
        sd[p + "attention_norm.weight"] = hash_normal((d,), nxt(), 0.1, 1.0)
        sd[p + "ffn_norm.weight"] = hash_normal((d,), nxt(), 0.1, 1.0)
        sd[p + "attention.gate1"] = hash_normal((1, H, 1, 1), nxt(), 0.5)
        sd[p + "attention.gate2"] = hash_normal((1, H, 1, 1), nxt(), 0.1, -bias)
    sd["adapter_query.weight"] = hash_normal((params.adapter_len * params.adapter_layer, d), nxt(), 1.0)
    sd["visual_proj.weight"] = hash_normal((d, video_dim), nxt(), 1.0 / math.sqrt(video_dim))
    sd["temporal_emb.weight"] = hash_normal((max_feats, d), nxt(), 1.0)
    return sd


def ffn_hidden_dim(dim: int, multiple_of: int) -> int:
    """SwiGLU hidden size rule of `llama/model.py:134-135` applied to hidden_dim=4*dim (`:179`)."""
    hidden = int(2 * (4 * dim) / 3)
    return multiple_of * ((hidden + multiple_of - 1) // multiple_of)


def synthetic_batch(bsz: int, seqlen: int, vocab: int, max_feats: int = 10, seed: int = 0,
                    video_start: int = 18, n_options: int = 1, video_dim: int = 768,
                    full_length: bool = False, vaq_label_span=(10, 25),
                    generator_device: str = "cpu") -> Dict:
    """Batch dict with the contract of `dataloader/__init__.py:28-90` (CPU tensors).

    VQA/VAQ: ``[BOS, instr…] ‖ F video slots ‖ nl ‖ text ‖ EOS ‖ pad(0)`` with the video block at
    ``video_start``; VQA labels = last 4 real tokens, VAQ labels = last U[span] real tokens, 0 =
    ignore (`base_dataset.py:65-77`). QAV: video slots at ``[len-F-1, len-1)``, label 0..F-1 there
    and -1 elsewhere, ``video_index = arange(prefix, prefix+F)`` (`base_dataset.py:80-91,120`).
    ``n_options > 1`` produces the validation layout: per sample ``n_options`` sequences identical
    except for the answer span (`llama/tokenizer.py:69-90`), plus ``answer`` in [0, n_options).
    """
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    F = max_feats
    assert seqlen >= video_start + F + 8, "sequence too short for the synthetic layout"
    lo = max(video_start + F + 7, int(0.7 * seqlen))
    ids = {k: torch.zeros(bsz, n_options, seqlen, dtype=torch.int64) for k in ("vqa", "vaq", "qav")}
    labels = {"vqa": torch.zeros(bsz, n_options, seqlen, dtype=torch.int64),
              "vaq": torch.zeros(bsz, n_options, seqlen, dtype=torch.int64),
              "qav": torch.full((bsz, n_options, seqlen), -1, dtype=torch.int64)}
    label_mask = {k: torch.zeros(bsz, n_options, seqlen) for k in ("vqa", "vaq", "qav")}
    qav_index = torch.zeros(bsz, F, dtype=torch.int64)
    qav_prefix, vqa_prefix, vaq_prefix = [], [], []
    for b in range(bsz):
        ln = seqlen if full_length else int(torch.randint(lo, seqlen + 1, (1,), generator=g))
        base = torch.randint(3, vocab, (seqlen,), generator=g)
        base[0] = 1                      # BOS
        base[ln - 1] = 2                 # EOS
        base[ln:] = 0                    # pad -> 0 (`base_dataset.py:99-104`)
        # VQA: answer span = last 4 real tokens ("(X)" + EOS, `tokenizer.py:69`)
        for o in range(n_options):
            t = base.clone()
            t[video_start:video_start + F] = 0          # video placeholders (-2 -> 0)
            if n_options > 1:                           # options differ only in the answer span
                t[ln - 4:ln - 1] = torch.randint(3, vocab, (3,), generator=g)
            ids["vqa"][b, o] = t
            labels["vqa"][b, o, ln - 4:ln] = t[ln - 4:ln]
            label_mask["vqa"][b, o, ln - 4:ln] = 1
        vqa_prefix.append(ln - 4)
        # VAQ: question span = last U[span] real tokens
        nq = int(torch.randint(vaq_label_span[0], vaq_label_span[1] + 1, (1,), generator=g))
        nq = min(nq, ln - (video_start + F + 2))
        t = torch.randint(3, vocab, (seqlen,), generator=g)
        t[0] = 1; t[ln - 1] = 2; t[ln:] = 0
        t[video_start:video_start + F] = 0
        for o in range(n_options):
            ids["vaq"][b, o] = t
            labels["vaq"][b, o, ln - nq:ln] = t[ln - nq:ln]
            label_mask["vaq"][b, o, ln - nq:ln] = 1
        vaq_prefix.append(ln - nq)
        # QAV: video slots just before the final token
        t = torch.randint(3, vocab, (seqlen,), generator=g)
        t[0] = 1; t[ln - 1] = 2; t[ln:] = 0
        p = ln - F - 1
        t[p:p + F] = 0
        for o in range(n_options):
            ids["qav"][b, o] = t
            labels["qav"][b, o, p:p + F] = torch.arange(F)
            label_mask["qav"][b, o, p] = 1
        qav_index[b] = torch.arange(p, p + F)
        qav_prefix.append(p)
    video = torch.randn(bsz, F, video_dim, generator=g)
    data = {
        "video": video,
        "video_len": torch.full((bsz,), F, dtype=torch.long),
        "text_id": ids,
        "label": labels,
        "video_start": {"vqa": [video_start] * bsz, "vaq": [video_start] * bsz, "qav": qav_prefix},
        "video_index": {"vqa": torch.stack([torch.arange(p, p + F) for p in vqa_prefix]),
                        "vaq": torch.stack([torch.arange(p, p + F) for p in vaq_prefix]),
                        "qav": qav_index},
        "label_mask": label_mask,
        "prefix_index": {"vqa": vqa_prefix, "vaq": vaq_prefix, "qav": qav_prefix},
        "answer": torch.randint(0, max(n_options, 1), (bsz,), generator=g),
        "qtype": torch.zeros(bsz, dtype=torch.long),
        "vid": [f"synthetic{b}" for b in range(bsz)],
        "qid": [f"q{b}" for b in range(bsz)],
        "text": [{} for _ in range(bsz)],
    }
    return data


GEN_QUESTION_MARKER = 894        # the id `llama/model.py:520` searches for to locate the question


def synthetic_generation_batch(bsz: int, seqlen: int, vocab: int, a_token_id: int, max_feats: int = 10, seed: int = 0,
                               video_start: int = 12, n_options: int = 4, video_dim: int = 768) -> Dict:
    """Validation batch for the generation evaluator (`llama/model.py:367-546`, `engine.py:78-85`): per sample `n_options` VQA
    sequences  [BOS, instr..] | F video slots | nl | 894 x question.. | a_token + 4 tokens | answer_o .. EOS | pad(0)  that differ
    only in the answer span; `prefix_index` = first answer position (= index(a_token) + 5, `model.py:551`); labels = answer span
    incl. EOS, 0 elsewhere (`base_dataset.py:65-70`). Needs vocab > 894 and prefix + 30 <= seqlen - 1."""
    assert vocab > GEN_QUESTION_MARKER and 2 < a_token_id < vocab and a_token_id != GEN_QUESTION_MARKER
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    F = max_feats

    def rand_tokens(n):
        t = torch.randint(3, vocab, (n,), generator=g)
        bad = (t == GEN_QUESTION_MARKER) | (t == a_token_id) | (t == 13)
        t[bad] = 7
        return t

    ids = torch.zeros(bsz, n_options, seqlen, dtype=torch.int64)
    labels = torch.zeros(bsz, n_options, seqlen, dtype=torch.int64)
    prefix = []
    for b in range(bsz):
        qlen = int(torch.randint(5, 10, (1,), generator=g))
        head = torch.cat([torch.tensor([1]), rand_tokens(video_start - 1), torch.zeros(F, dtype=torch.int64), torch.tensor([13]),
                          torch.tensor([GEN_QUESTION_MARKER]), rand_tokens(1 + qlen), torch.tensor([a_token_id]), rand_tokens(4)])
        p = head.numel()
        assert p + 30 <= seqlen - 1, "sequence too short for 31 generation steps"
        prefix.append(p)
        for o in range(n_options):
            alen = int(torch.randint(1, 5, (1,), generator=g))
            ans = torch.cat([rand_tokens(alen), torch.tensor([2])])
            ids[b, o, :p] = head
            ids[b, o, p:p + alen + 1] = ans
            labels[b, o, p:p + alen + 1] = ans
    return {
        "video": torch.randn(bsz, F, video_dim, generator=g),
        "text_id": {"vqa": ids}, "label": {"vqa": labels},
        "video_start": {"vqa": [video_start] * bsz},
        "prefix_index": {"vqa": prefix},
        "answer": torch.randint(0, n_options, (bsz,), generator=g),
        "qtype": torch.ones(bsz, dtype=torch.long),
        "vid": [f"synthetic{b}" for b in range(bsz)], "qid": [f"q{b}" for b in range(bsz)],
        "text": [{"options": [f"option{o}" for o in range(n_options)]} for _ in range(bsz)],
    }


def synthetic_dialogue_texts(n: int, n_options: int = 5, seed: int = 0, long_every: int = 2):
    """TVQA-style prompt pieces with a subtitle dialogue (`dataloader/tvqa.py:36-58`): every `long_every`-th sample has a dialogue long
    enough to overflow a 128-token sequence (the dialogue-aware truncation path), one sample has no dialogue at all."""
    import random
    rng = random.Random(seed)
    words = ["sheldon", "door", "coffee", "why", "because", "leonard", "angry", "phone", "left", "room", "said", "never", "again", "ok"]
    mapping = {i: f"({chr(65 + i)})" for i in range(n_options)}
    out = []
    for k in range(n):
        question = " ".join(rng.choice(words) for _ in range(rng.randint(4, 9))).capitalize() + "?"
        options = [" ".join(rng.choice(words) for _ in range(rng.randint(1, 4))) for _ in range(n_options)]
        n_d = 0 if k == n - 1 else (rng.randint(120, 200) if k % long_every == 0 else rng.randint(5, 30))
        d_text = ("Dialogue: " + " ".join(rng.choice(words) for _ in range(n_d)) + "\n") if n_d else ""
        o_text = "Choices: \n" + "".join(f"{mapping[i]} {options[i]}\n" for i in range(n_options))
        out.append(dict(text={"q_text": f"Question: {question}\n", "o_text": o_text, "a_text": "Answer: The answer is ", "d_text": d_text,
                              "options": options}, answer=rng.randrange(n_options), options=options))
    return out, mapping


class HashSentencePiece:
    """Deterministic stand-in for `SentencePieceProcessor.encode` (no `tokenizer.model` exists offline): words and
    punctuation marks hash to ids in [100, n_words); the pieces 'Video', 'Question', 'Answer' map to the ids the
    reference hard-codes (`llama/tokenizer.py:28-31`) so the prompt builders can locate them."""

    SPECIAL = {"Video": 15167, "Question": 16492, "Answer": 22550}

    def __init__(self, n_words: int = 32000):
        self.n_words = n_words

    def encode(self, s: str):
        import re
        import zlib
        out = []
        for piece in re.findall(r"\w+|[^\w\s]|\n", s):
            out.append(self.SPECIAL.get(piece, 100 + zlib.crc32(piece.encode()) % (self.n_words - 100)))
        return out


def hash_tokenizer(cls, n_words: int = 32000, is_generation_task: bool = False):
    """Instance of a tokenizer class (ours or the reference's) wired to `HashSentencePiece` without a model file."""
    import argparse
    t = object.__new__(cls)
    t.args = argparse.Namespace(is_generation_task=is_generation_task, debug=False)
    t.sp_model = HashSentencePiece(n_words)
    t.n_words, t.bos_id, t.eos_id, t.pad_id = n_words, 1, 2, -1
    t.v_token_id, t.q_token_id, t.a_token_id, t.nl_id = 15167, 16492, 22550, 13
    return t


def synthetic_qa_texts(n: int, n_options: int = 5, seed: int = 0):
    """NExT-QA-style prompt pieces (`dataloader/nextqa.py:24-39` layout) with random words."""
    import random
    rng = random.Random(seed)
    words = ["man", "dog", "ball", "red", "jump", "why", "after", "child", "table", "run", "holding", "water", "car", "before", "smile"]
    mapping = {i: f"({chr(65 + i)})" for i in range(n_options)}
    out = []
    for _ in range(n):
        question = " ".join(rng.choice(words) for _ in range(rng.randint(4, 10))).capitalize() + "?"
        options = [" ".join(rng.choice(words) for _ in range(rng.randint(1, 4))) for _ in range(n_options)]
        o_text = "Choices: \n" + "".join(f"{mapping[i]} {options[i]}\n" for i in range(n_options))
        out.append(dict(text={"q_text": f"Question: {question}\n", "o_text": o_text, "a_text": "Answer: The answer is ", "options": options},
                        answer=rng.randrange(n_options), options=options))
    return out, mapping
