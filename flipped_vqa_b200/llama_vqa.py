"""Model factory with the contract of `/root/reference/llama_vqa.py:6-77`: read `params.json`, load and
merge the `*.pth` shards (Meta model-parallel shards concatenated by their split dimension, `:25-58`),
build the Transformer with frozen bf16 base weights and fp32 trainables (`:61-76`)."""
from __future__ import annotations

import json
from pathlib import Path

import torch

from .llama import ModelArgs, Tokenizer, Transformer

_COLUMN_PARALLEL = ("attention.wq.weight", "attention.wk.weight", "attention.wv.weight", "feed_forward.w1.weight", "feed_forward.w3.weight")
_ROW_PARALLEL = ("attention.wo.weight", "feed_forward.w2.weight")


def merge_shards(loaded, n_layers: int):
    """Concatenate tensor-parallel checkpoint shards (`llama_vqa.py:25-58`): column-parallel weights on
    dim 0 (wq/wk/wv/w1/w3/output), row-parallel on dim 1 (wo/w2/tok_embeddings), norms replicated."""
    if len(loaded) == 1:
        return loaded[0]
    full = {}

    def take(name, dim):
        full[name] = loaded[0][name].clone() if dim < 0 else torch.cat([x[name] for x in loaded], dim=dim)

    take("tok_embeddings.weight", 1)
    take("norm.weight", -1)
    take("output.weight", 0)
    for i in range(n_layers):
        p = f"layers.{i}."
        for k in ("attention_norm.weight", "ffn_norm.weight"):
            take(p + k, -1)
        for k in _COLUMN_PARALLEL:
            take(p + k, 0)
        for k in _ROW_PARALLEL:
            take(p + k, 1)
    return full


def apply_freeze_rule(model):
    """`llama_vqa.py:71-76`: trainable iff the name contains gate/adapter/temporal_emb/visual_proj (fp32);
    everything else frozen (bf16 here, fp16 in the reference)."""
    for name, param in model.named_parameters():
        if ("gate" in name) or ("adapter" in name) or ("temporal_emb" in name) or ("visual_proj" in name):
            param.requires_grad = True
            param.data = param.data.float()
        else:
            param.requires_grad = False
    return model


def LLaMA_VQA(args, **kwargs):
    with open(f"{args.llama_model_path}{args.model}/params.json", "r") as f:
        params = json.loads(f.read())
    tokenizer = Tokenizer(model_path=f"{args.llama_model_path}/tokenizer.model")
    print(f"Using model: {args.model}")
    checkpoints = sorted((Path(args.llama_model_path) / args.model).glob("*.pth"))
    loaded = []
    for x in checkpoints:
        print("loading from", x)
        loaded.append(torch.load(x, map_location="cpu"))
    full_state_dict = merge_shards(loaded, params["n_layers"])
    model_args = ModelArgs(max_seq_len=args.max_seq_len, max_batch_size=32, adapter_len=args.adapter_len,
                           adapter_layer=args.adapter_layer, **params)
    model_args.vocab_size = tokenizer.n_words
    model = Transformer(model_args, args, tokenizer=tokenizer)
    model.load_state_dict(full_state_dict, strict=False)      # Meta's extra rope.freqs is tolerated (`:68`)
    return apply_freeze_rule(model)
