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
# gate1 / gate2 as parameter groups (stacked over layers). The default fp16-operand build meets north_star's 2e-2; the bf16-operand
# build (FVQA_DTYPE=bf16) cannot: bf16 operand rounding ALONE - everything else exact fp32 - already gives 6-8e-2 at 32 layers and
# up to 4e-2 on a handful of gate numbers at 2 layers (profiles/r2_numerics_ablation.txt, tests/gate_noise_probe.py).
GATE_STACK_RTOL = 2e-2 if os.environ.get("FVQA_DTYPE", "fp16").lower() == "fp16" else 6e-2


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


def big_state_dict(pd, seed=0, device="cuda"):
    """Random-init weights of a full-size model generated ON THE GPU (a CPU generator needs minutes at 7B), SURVEY 8(d) recipe:
    frozen >= 2-D N(0, 0.02^2), norms 1 + N(0, 0.1^2), adapter / temporal N(0, 1), gate1 N(0, 0.5^2) (not zero), gate2 -bias +
    N(0, 0.1^2). Values are rounded to be exactly representable in bf16 AND fp16 so every consumer sees the same numbers."""
    d, L, H, V = pd["dim"], pd["n_layers"], pd["n_heads"], pd["vocab_size"]
    from flipped_vqa_b200.synthetic import ffn_hidden_dim
    hid = ffn_hidden_dim(d, pd["multiple_of"])
    A = pd["adapter_len"]
    g = torch.Generator(device=device).manual_seed(seed)
    rn = lambda *s, std=0.02, mean=0.0: (torch.randn(*s, device=device, generator=g) * std + mean).to(torch.bfloat16).to(torch.float16).float()
    sd = {"tok_embeddings.weight": rn(V, d), "output.weight": rn(V, d), "norm.weight": rn(d, std=0.1, mean=1.0),
          "adapter_query.weight": rn(A * pd["adapter_layer"], d, std=1.0), "visual_proj.weight": rn(d, 768, std=0.036),
          "temporal_emb.weight": rn(10, d, std=1.0)}
    for i in range(L):
        p = f"layers.{i}."
        for nm in ("wq", "wk", "wv", "wo"):
            sd[p + f"attention.{nm}.weight"] = rn(d, d)
        sd[p + "feed_forward.w1.weight"] = rn(hid, d)
        sd[p + "feed_forward.w2.weight"] = rn(d, hid)
        sd[p + "feed_forward.w3.weight"] = rn(hid, d)
        sd[p + "attention_norm.weight"] = rn(d, std=0.1, mean=1.0)
        sd[p + "ffn_norm.weight"] = rn(d, std=0.1, mean=1.0)
        sd[p + "attention.gate1"] = rn(1, H, 1, 1, std=0.5)
        sd[p + "attention.gate2"] = rn(1, H, 1, 1, std=0.1, mean=-3.5)
    return sd


def rel_l2(a, b):
    a = torch.as_tensor(a).float().cpu()
    b = torch.as_tensor(b).float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def product_grads(model):
    return {n: p.grad.detach().float().cpu() for n, p in model.named_parameters() if p.requires_grad and p.grad is not None}


def full_depth_argmax_report(items=256, layers=32, per=8, dim=4096, heads=32, seqlen=128):
    """north_star's third criterion at FULL depth: loss-based option scoring (`llama/model_my_original_mod.py:375-377` +
    `engine.py:88-93`) of the product vs the fp32 oracle (TF32 off) on the same random-init weights, `items` x 5 options. Returns
    the agreement, the per-option normalised-loss error and the oracle's best-vs-runner-up margins (an item whose margin is below
    the loss error can flip legitimately). Used by tests/test_model_gpu.py and tools/argmax_full_depth.py."""
    from oracle import llama_vqa_oracle as O
    from flipped_vqa_b200 import _lib
    from flipped_vqa_b200.synthetic import synthetic_batch
    torch.backends.cuda.matmul.allow_tf32 = False
    pd = dict(dim=dim, n_layers=layers, n_heads=heads, vocab_size=32000, multiple_of=256, norm_eps=1e-6, max_batch_size=64,
              max_seq_len=seqlen, adapter_len=10, adapter_layer=layers)
    args = make_args()
    sd = big_state_dict(pd, seed=0)
    model = build_product_model(pd, sd, args)
    st = O.prepare_state(sd, frozen_dtype=torch.float32, device="cuda", requires_grad=False)
    del sd
    agree = total = 0
    errs, margins, flipped = [], [], []
    for b in range(items // per):
        data = synthetic_batch(per, seqlen, pd["vocab_size"], max_feats=args.max_feats, seed=1000 + b, n_options=5)
        with torch.no_grad():
            tok = model(data, inference=True)
            pred = model.predict_options(tok).cpu()
            ref_tok = O.option_token_losses(st, SimpleNamespace(**pd), data, max_feats=args.max_feats)
        ref_pred = O.option_predict(ref_tok).cpu()
        cnt = (ref_tok != 0).sum(-1).clamp(min=1)
        score = (tok.float() * (ref_tok != 0)).sum(-1) / cnt                       # normalised loss per (item, option)
        ref_score = ref_tok.sum(-1) / cnt
        errs.append((score - ref_score).abs().max(-1).values.cpu())
        top2 = ref_score.topk(2, dim=-1, largest=False).values
        m = (top2[:, 1] - top2[:, 0]).cpu()
        margins.append(m)
        flipped += m[pred != ref_pred].tolist()
        agree += int((pred == ref_pred).sum())
        total += pred.numel()
    errs, margins = torch.cat(errs), torch.cat(margins)
    return {"task": "option-scoring argmax agreement at full depth", "operand_dtype": _lib.DTYPE_NAME, "dim": dim, "layers": layers,
            "items": total, "options": 5, "seqlen": seqlen, "agree": agree, "agreement": agree / total,
            "normalised_loss_abs_err_max": float(errs.max()), "normalised_loss_abs_err_median": float(errs.median()),
            "oracle_margin_min": float(margins.min()), "oracle_margin_median": float(margins.median()),
            "items_with_margin_below_2x_max_err": int((margins < 2 * errs.max()).sum()),
            "oracle_margins_of_flipped_items": flipped}
