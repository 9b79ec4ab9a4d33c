"""ORACLE support (test infrastructure): import the UNMODIFIED reference from /root/reference.

Only usable in the build container (the GPU box has no /root/reference). Nothing under
/root/reference is edited or copied; three monkey-patches make its hard-coded CUDA/fp16/tokenizer
assumptions runnable on a GPU-less host:

1. ``torch.Tensor.cuda`` -> identity when no GPU is visible (`llama/model.py:82-83,255-264,302`);
2. ``Tokenizer`` -> stub exposing the three ids `llama/model.py:201-204` reads
   (no ``tokenizer.model`` exists offline, `llama/tokenizer.py:18` asserts the file);
3. optionally ``torch.Tensor.half`` -> ``.to(dtype)`` so the same code runs in fp32 ("gold") or
   bf16 instead of its native fp16 (`llama/model.py:115,119,324,339-345`).
"""
from __future__ import annotations

import argparse
import contextlib
import importlib
import os
import sys

import torch

REFERENCE_ROOT = os.environ.get("FVQA_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "llama", "model.py"))


class StubTokenizer:
    def __init__(self, model_path=None, args=None, n_words: int = 512):
        self.n_words, self.bos_id, self.eos_id, self.pad_id = n_words, 1, 2, -1
        self.v_token_id, self.q_token_id, self.a_token_id, self.nl_id = 15167, 16492, 22550, 13

    def decode(self, t):
        return ""


@contextlib.contextmanager
def patched_torch(dtype: torch.dtype):
    """Apply shims 1 and 3 for the duration of a reference call."""
    orig_cuda, orig_half = torch.Tensor.cuda, torch.Tensor.half
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
    if dtype != torch.float16:
        torch.Tensor.half = lambda self, *a, **k: self.to(dtype)
    try:
        yield
    finally:
        torch.Tensor.cuda, torch.Tensor.half = orig_cuda, orig_half


def import_reference(variant: str = "model"):
    """variant: 'model' (training oracle, HEAD) or 'model_my_original_mod' (option scoring)."""
    assert reference_available(), f"reference tree not found at {REFERENCE_ROOT}"
    sys.dont_write_bytecode = True            # the reference tree is read-only
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    tok = importlib.import_module("llama.tokenizer")
    tok.Tokenizer = StubTokenizer
    mod = importlib.import_module(f"llama.{variant}")
    mod.Tokenizer = StubTokenizer
    return mod


def reference_args(max_feats=10, bias=3.5, tau=100.0, vaq=True, qav=True, audio_mode=None):
    """audio_mode: None | 'audio_only' | 'concat' | 'sum' | 'attention' -> the reference's --audio / --audio_only /
    --audio_merge flags (`train.py`, `llama/model.py:209-227`)."""
    return argparse.Namespace(max_feats=max_feats, bias=bias, tau=tau, llama_model_path="x/",
                              audio=audio_mode is not None, audio_only=audio_mode == "audio_only",
                              audio_merge=audio_mode if audio_mode in ("concat", "sum", "attention") else "none", debug=False,
                              vaq=vaq, qav=qav, is_generation_task=False)


def build_reference_model(mod, params_kwargs: dict, args, state_dict, dtype: torch.dtype):
    """Construct the reference Transformer under ``dtype`` (stand-in for `llama_vqa.py:63`), load
    ``state_dict`` and apply the freeze rule of `llama_vqa.py:71-76`."""
    params = mod.ModelArgs(**params_kwargs)
    with patched_torch(dtype):
        torch.set_default_dtype(dtype)
        try:
            model = mod.Transformer(params, args)
        finally:
            torch.set_default_dtype(torch.float32)
    missing, unexpected = model.load_state_dict(state_dict, strict=False)
    assert not unexpected, unexpected
    for name, p in model.named_parameters():
        if any(s in name for s in ("gate", "adapter", "temporal_emb", "visual_proj")):
            p.requires_grad = True
            p.data = state_dict[name].detach().clone().float()
        else:
            p.requires_grad = False
            p.data = state_dict[name].detach().to(dtype)
    return model
