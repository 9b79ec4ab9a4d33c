"""ORACLE support: generate tests/golden/*.npz by running the UNMODIFIED reference here.

    python oracle/make_golden.py            # writes tests/golden/{train_small,options_small}.npz

Requires /root/reference (build container only). The fixtures hold only *results* of the reference
(losses, trainable-parameter gradients, per-token option losses, predictions); the inputs are
regenerated bit-exactly from `flipped_vqa_b200.synthetic` (integer-hash weights, seeded batch).

Two reference runs are stored per case:
  * ``gold``  – the reference with its ``.half()`` calls remapped to fp32 (SURVEY.md §8(c) shim 3);
  * ``fp16``  – the reference exactly as shipped (fp16 frozen weights), to show its own distance
                from gold so tolerance failures can be attributed.
"""
from __future__ import annotations

import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from flipped_vqa_b200.synthetic import synthetic_audio, synthetic_audio_state, synthetic_batch, synthetic_state_dict  # noqa: E402
from oracle import ref_shims  # noqa: E402

# The golden config: small enough for a <1 MB fixture, big enough for the CUDA path
# (head_dim 64, dims multiple of 64).
GOLDEN = dict(dim=128, n_layers=3, n_heads=2, vocab_size=256, multiple_of=64, norm_eps=1e-6,
              max_batch_size=32, max_seq_len=48, adapter_len=10, adapter_layer=2)
GOLDEN_RUN = dict(bsz=3, seqlen=48, max_feats=10, bias=3.5, tau=100.0, video_start=12, seed=7)


def golden_inputs(n_options: int = 1):
    params = SimpleNamespace(**GOLDEN)
    r = GOLDEN_RUN
    sd = synthetic_state_dict(params, seed=r["seed"], max_feats=r["max_feats"], bias=r["bias"])
    data = synthetic_batch(r["bsz"], r["seqlen"], GOLDEN["vocab_size"], max_feats=r["max_feats"],
                           seed=r["seed"], video_start=r["video_start"], n_options=n_options,
                           vaq_label_span=(5, 9))
    return params, sd, data


def run_reference_train(dtype):
    params, sd, data = golden_inputs()
    mod = ref_shims.import_reference("model")
    r = GOLDEN_RUN
    args = ref_shims.reference_args(max_feats=r["max_feats"], bias=r["bias"], tau=r["tau"])
    model = ref_shims.build_reference_model(mod, GOLDEN, args, sd, dtype)
    with ref_shims.patched_torch(dtype):
        vqa, vaq, qav = model(data)
        (vqa + vaq + qav).backward()
    out = {"loss": np.array([float(vqa.detach()), float(vaq.detach()), float(qav.detach())], dtype=np.float64)}
    for n, p in model.named_parameters():
        if p.requires_grad and p.grad is not None:   # layers skipped by `model.py:338` get no grad
            out["grad/" + n] = p.grad.detach().float().numpy()
    return out


AUDIO_MODES = ("audio_only", "concat", "sum", "attention")


def golden_audio_inputs(mode: str):
    """Golden inputs of an audio-fusion variant: the video-only inputs + the variant's parameters and audio features
    ([B, F, 1024] frames; one ImageBind vector per sample for 'attention', `dataloader/nextqa.py:16-18`)."""
    params, sd, data = golden_inputs()
    sd = dict(sd)
    if mode == "audio_only":
        sd.pop("visual_proj.weight")
    sd.update(synthetic_audio_state(params, mode, seed=GOLDEN_RUN["seed"]))
    data = dict(data)
    data["audio"] = synthetic_audio(GOLDEN_RUN["bsz"], 1 if mode == "attention" else GOLDEN_RUN["max_feats"], seed=GOLDEN_RUN["seed"])
    if mode == "audio_only":
        data.pop("video")
    return params, sd, data


def run_reference_audio(mode: str, dtype=torch.float32):
    params, sd, data = golden_audio_inputs(mode)
    mod = ref_shims.import_reference("model")
    r = GOLDEN_RUN
    args = ref_shims.reference_args(max_feats=r["max_feats"], bias=r["bias"], tau=r["tau"], audio_mode=mode)
    model = ref_shims.build_reference_model(mod, GOLDEN, args, sd, dtype)
    if mode == "attention":                       # `CrossAttentionModule(768).float()` (`model.py:227`) stays fp32 in every build
        for n, p in model.video_audio_cross_attn.named_parameters():
            p.data = sd["video_audio_cross_attn." + n].detach().clone().float()
    with ref_shims.patched_torch(dtype):
        vqa, vaq, qav = model(data)
        (vqa + vaq + qav).backward()
    out = {"loss": np.array([float(vqa.detach()), float(vaq.detach()), float(qav.detach())], dtype=np.float64)}
    for n, p in model.named_parameters():
        if p.requires_grad and p.grad is not None:
            out["grad/" + n] = p.grad.detach().float().numpy()
    return out


def run_reference_options(dtype, n_options=5):
    params, sd, data = golden_inputs(n_options)
    mod = ref_shims.import_reference("model_my_original_mod")
    r = GOLDEN_RUN
    args = ref_shims.reference_args(max_feats=r["max_feats"], bias=r["bias"], tau=r["tau"])
    model = ref_shims.build_reference_model(mod, GOLDEN, args, sd, dtype)
    with ref_shims.patched_torch(dtype), torch.no_grad():
        tok = model(data, inference=True)
    # engine.py:88-93 applied verbatim to the reference's output
    count = (tok != 0).sum(-1)
    pred = (tok.sum(-1) / count).argmin(-1)
    return {"token_losses": tok.float().numpy(), "prediction": pred.numpy()}


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 1)
    gdir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(gdir, exist_ok=True)
    tr = {}
    for tag, dt in (("gold", torch.float32), ("fp16", torch.float16)):
        for k, v in run_reference_train(dt).items():
            tr[f"{tag}/{k}"] = v
    np.savez_compressed(os.path.join(gdir, "train_small.npz"), **tr)
    op = {}
    for tag, dt in (("gold", torch.float32), ("fp16", torch.float16)):
        for k, v in run_reference_options(dt).items():
            op[f"{tag}/{k}"] = v
    np.savez_compressed(os.path.join(gdir, "options_small.npz"), **op)
    au = {}
    for mode in AUDIO_MODES:
        for k, v in run_reference_audio(mode).items():
            au[f"{mode}/gold/{k}"] = v
        print(mode, "gold losses", au[f"{mode}/gold/loss"])
    np.savez_compressed(os.path.join(gdir, "train_audio_small.npz"), **au)
    print("gold losses", tr["gold/loss"], "fp16 losses", tr["fp16/loss"])
    print("gold pred", op["gold/prediction"], "fp16 pred", op["fp16/prediction"])
    for f in ("train_small.npz", "options_small.npz", "train_audio_small.npz"):
        print(f, os.path.getsize(os.path.join(gdir, f)), "bytes")


if __name__ == "__main__":
    main()
