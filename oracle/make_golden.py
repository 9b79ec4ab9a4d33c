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

from flipped_vqa_b200.synthetic import (synthetic_audio, synthetic_audio_state, synthetic_batch, synthetic_generation_batch,  # noqa: E402
                                        synthetic_state_dict)
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


# ------------------------------------------------------------------------------------------------
# multi-step training trajectory through the reference's OWN loop
# ------------------------------------------------------------------------------------------------
TRAJ = dict(n_batches=8, accum_iter=2, epochs=2, lr=5e-3, min_lr=5e-4, warmup_epochs=1, weight_decay=0.14, batch_seed0=900)


def trajectory_batches():
    r = GOLDEN_RUN
    return [synthetic_batch(r["bsz"], r["seqlen"], GOLDEN["vocab_size"], max_feats=r["max_feats"], seed=TRAJ["batch_seed0"] + i,
                            video_start=r["video_start"], vaq_label_span=(5, 9)) for i in range(TRAJ["n_batches"])]


def trajectory_args():
    import argparse
    return argparse.Namespace(accum_iter=TRAJ["accum_iter"], lr=TRAJ["lr"], min_lr=TRAJ["min_lr"], warmup_epochs=TRAJ["warmup_epochs"],
                              epochs=TRAJ["epochs"], debug=False)


def run_reference_trajectory(dtype=torch.float32):
    """The UNMODIFIED `engine.train_one_epoch` (`engine.py:10-56`) + `util.misc.NativeScalerWithGradNormCount` (`util/misc.py:253-279`)
    + `util.lr_sched.adjust_learning_rate` + AdamW(betas=(0.9, 0.95)) as `train.py:120-121` builds it (timm's weight-decay grouping
    puts every trainable of this model - all >= 2-D, none named *.bias - into the decayed group) driving the reference model for
    TRAJ['epochs'] epochs over the same 8 micro-batches (accum_iter 2 -> 4 optimizer steps per epoch). CPU run: the reference's
    GradScaler disables itself without CUDA (scale 1) and `torch.cuda.synchronize` (`engine.py:43`) is stubbed."""
    import importlib
    params, sd, _ = golden_inputs()
    mod = ref_shims.import_reference("model")
    engine = importlib.import_module("engine")
    misc = importlib.import_module("util.misc")
    r = GOLDEN_RUN
    model = ref_shims.build_reference_model(mod, GOLDEN, ref_shims.reference_args(max_feats=r["max_feats"], bias=r["bias"], tau=r["tau"]), sd, dtype)
    trainables = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.AdamW([{"params": trainables, "weight_decay": TRAJ["weight_decay"]}], lr=TRAJ["lr"], betas=(0.9, 0.95))
    step_losses = []

    class Recorder(torch.nn.Module):                     # records what model(data) returned; the loop itself is the reference's
        def __init__(self, inner):
            super().__init__()
            self.inner = inner

        def forward(self, data):
            out = self.inner(data)
            step_losses.append([float(x.detach()) for x in out])
            return out

    wrapped = Recorder(model)
    batches = trajectory_batches()
    targs = trajectory_args()
    out = {}
    orig_sync = torch.cuda.synchronize
    torch.cuda.synchronize = lambda *a, **k: None
    try:
        with ref_shims.patched_torch(dtype):
            scaler = misc.NativeScalerWithGradNormCount()
            for epoch in range(TRAJ["epochs"]):
                stats = engine.train_one_epoch(wrapped, batches, opt, epoch, scaler, args=targs)
                for k, v in stats.items():
                    out[f"epoch{epoch}/stats/{k}"] = np.float64(v)
                for n, p in model.named_parameters():
                    if p.requires_grad:
                        out[f"epoch{epoch}/param/{n}"] = p.detach().float().numpy().copy()
    finally:
        torch.cuda.synchronize = orig_sync
    out["step_losses"] = np.array(step_losses, dtype=np.float64)          # [epochs * n_batches, 3]
    return out


# ------------------------------------------------------------------------------------------------
# generation evaluator (HEAD `llama/model.py:367-623`): greedy decoding + cosine matching
# ------------------------------------------------------------------------------------------------
GEN = dict(dim=128, n_layers=3, n_heads=2, vocab_size=1024, multiple_of=64, norm_eps=1e-6,
           max_batch_size=32, max_seq_len=96, adapter_len=10, adapter_layer=2)
GEN_RUN = dict(bsz=3, seqlen=96, n_options=4, max_feats=10, bias=3.5, tau=100.0, video_start=12, seed=23, a_token_id=900, logit_gain=12.0)   # seed: smallest arg-max margin of the reference run 0.017, similarity gaps > 0.13


def generation_inputs():
    """Weights: the usual synthetic state dict, with the output projection scaled by `logit_gain` so that the arg-max margins of
    the random model sit well above 16-bit rounding noise (the fixture records the reference's own margins)."""
    params = SimpleNamespace(**GEN)
    r = GEN_RUN
    sd = synthetic_state_dict(params, seed=r["seed"], max_feats=r["max_feats"], bias=r["bias"])
    sd["output.weight"] = (sd["output.weight"] * r["logit_gain"]).to(torch.bfloat16).float()
    data = synthetic_generation_batch(r["bsz"], r["seqlen"], GEN["vocab_size"], r["a_token_id"], max_feats=r["max_feats"], seed=r["seed"],
                                      video_start=r["video_start"], n_options=r["n_options"])
    return params, sd, data


def run_reference_generation(dtype=torch.float32):
    params, sd, data = generation_inputs()
    mod = ref_shims.import_reference("model")
    r = GEN_RUN
    args = ref_shims.reference_args(max_feats=r["max_feats"], bias=r["bias"], tau=r["tau"])
    args.is_generation_task = True
    model = ref_shims.build_reference_model(mod, GEN, args, sd, dtype)
    model.answer_token_id = r["a_token_id"]              # the stub tokenizer's id (22550) lies outside this small vocabulary
    rec = {"logits": [], "final_ids": None, "sims": None}
    hook = model.output.register_forward_hook(lambda m, i, o: rec["logits"].append(o.detach().float()))
    orig_filter, orig_sim = model.filter_and_process_output_tokens, model.find_most_similar

    def filt(tokens, mask):
        rec["final_ids"] = tokens.detach().clone()
        return orig_filter(tokens, mask)

    def sim(out_emb, choice_emb):
        res = orig_sim(out_emb, choice_emb)
        rec["sims"] = res[1].detach().float()
        return res
    model.filter_and_process_output_tokens, model.find_most_similar = filt, sim
    with ref_shims.patched_torch(dtype), torch.no_grad():
        most_similar, extracted = model(data, inference=True)
    hook.remove()
    prefix, steps = data["prefix_index"]["vqa"], 31
    assert len(rec["logits"]) == r["bsz"] * steps
    margins = np.zeros((r["bsz"], steps))
    tokens = np.zeros((r["bsz"], steps), dtype=np.int64)
    for b in range(r["bsz"]):
        for t in range(steps):
            row = rec["logits"][b * steps + t][0, prefix[b] - 1 + t]
            top = torch.topk(row, 2)
            margins[b, t] = float(top.values[0] - top.values[1])
            tokens[b, t] = int(top.indices[0])
    return {"most_similar": most_similar.numpy(), "final_ids": rec["final_ids"].numpy(), "similarities": rec["sims"].numpy(),
            "tokens": tokens, "margins": margins,
            "generated_answer_lengths": np.array([len(e["generated_answer"]) for e in extracted])}


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
    gen = run_reference_generation()
    np.savez_compressed(os.path.join(gdir, "generation_small.npz"), **gen)
    print("generation: most similar", gen["most_similar"], "min top-2 margin", gen["margins"].min(), "tokens[0,:8]", gen["tokens"][0, :8])
    tj = run_reference_trajectory()
    np.savez_compressed(os.path.join(gdir, "trajectory_small.npz"), **tj)
    print("trajectory step losses (sum):", tj["step_losses"].sum(1))
    print("gold losses", tr["gold/loss"], "fp16 losses", tr["fp16/loss"])
    print("gold pred", op["gold/prediction"], "fp16 pred", op["fp16/prediction"])
    for f in ("train_small.npz", "options_small.npz", "train_audio_small.npz", "trajectory_small.npz", "generation_small.npz"):
        print(f, os.path.getsize(os.path.join(gdir, f)), "bytes")


if __name__ == "__main__":
    main()
