"""Shared helpers for the model-level parity tests (CUDA path vs oracle / golden fixtures)."""
import argparse
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

# must match oracle/make_golden.py
GOLDEN = dict(dim=128, n_layers=3, n_heads=2, vocab_size=256, multiple_of=64, norm_eps=1e-6,
              max_batch_size=32, max_seq_len=48, adapter_len=10, adapter_layer=2)
GOLDEN_RUN = dict(bsz=3, seqlen=48, max_feats=10, bias=3.5, tau=100.0, video_start=12, seed=7)

LOSS_RTOL = 1e-2        # BASELINE.json north_star: per-objective losses within 1e-2 relative
GRAD_RTOL = 2e-2        # trainable-parameter gradients within 2e-2 relative L2


def golden_inputs(n_options=1):
    from flipped_vqa_b200.synthetic import synthetic_batch, synthetic_state_dict
    params = SimpleNamespace(**GOLDEN)
    r = GOLDEN_RUN
    sd = synthetic_state_dict(params, seed=r["seed"], max_feats=r["max_feats"], bias=r["bias"])
    data = synthetic_batch(r["bsz"], r["seqlen"], GOLDEN["vocab_size"], max_feats=r["max_feats"], seed=r["seed"],
                           video_start=r["video_start"], n_options=n_options, vaq_label_span=(5, 9))
    return params, sd, data


AUDIO_MODES = ("audio_only", "concat", "sum", "attention")


def golden_audio_inputs(mode):
    """must match oracle/make_golden.py::golden_audio_inputs"""
    from flipped_vqa_b200.synthetic import synthetic_audio, synthetic_audio_state
    params, sd, data = golden_inputs()
    sd, data = dict(sd), dict(data)
    if mode == "audio_only":
        sd.pop("visual_proj.weight")
        data.pop("video")
    sd.update(synthetic_audio_state(params, mode, seed=GOLDEN_RUN["seed"]))
    data["audio"] = synthetic_audio(GOLDEN_RUN["bsz"], 1 if mode == "attention" else GOLDEN_RUN["max_feats"], seed=GOLDEN_RUN["seed"])
    return params, sd, data


def make_args(max_feats=10, bias=3.5, tau=100.0, vaq=True, qav=True, audio_mode=None):
    return argparse.Namespace(max_feats=max_feats, bias=bias, tau=tau, llama_model_path="x/", audio=audio_mode is not None,
                              audio_only=audio_mode == "audio_only",
                              audio_merge=audio_mode if audio_mode in ("concat", "sum", "attention") else "none",
                              debug=False, vaq=vaq, qav=qav, is_generation_task=False)


def build_product_model(params_dict, sd, args):
    """The product model (CUDA path) with the given state dict; freeze rule of llama_vqa.py:71-76."""
    from flipped_vqa_b200.llama import ModelArgs, SyntheticTokenizer, Transformer
    model = Transformer(ModelArgs(**params_dict), args, tokenizer=SyntheticTokenizer(params_dict["vocab_size"]))
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and not missing, (missing, unexpected)
    return model


def rel_l2(a, b):
    a = torch.as_tensor(a).float().cpu()
    b = torch.as_tensor(b).float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def product_grads(model):
    return {n: p.grad.detach().float().cpu() for n, p in model.named_parameters() if p.requires_grad and p.grad is not None}
