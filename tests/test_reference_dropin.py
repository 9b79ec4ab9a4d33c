"""The reference-side binding of INTEGRATION.md, EXECUTED: the reference's own `llama_vqa.py` (model factory: params.json, shard
merge, `Transformer(model_args, args)`, `load_state_dict(strict=False)`, freeze rule, `llama_vqa.py:6-77`) runs unmodified with its
`from llama import ModelArgs, Tokenizer, Transformer` resolved to `flipped_vqa_b200.llama` - the three-line patch a maintainer applies.
Needs the reference tree (build container only; skipped on the GPU box) and builds a throw-away checkpoint directory: two
model-parallel shards, params.json and a SentencePiece model trained on the spot."""
import argparse
import importlib.util
import json
import os
import sys

import pytest
import torch

from oracle import ref_shims
from tests.util_parity import make_args

pytestmark = pytest.mark.skipif(not ref_shims.reference_available(), reason="reference tree not present")


def _make_checkpoint_dir(root, params, sd):
    import sentencepiece as spm
    model_dir = root / "7B"
    model_dir.mkdir(parents=True)
    (model_dir / "params.json").write_text(json.dumps(params))
    corpus = root / "corpus.txt"
    words = "instruction predict the answer based on video and question choices man dog ball red jump why after child table run".split()
    corpus.write_text("\n".join(" ".join(words[(i * 7 + j) % len(words)] for j in range(12)) for i in range(200)))
    spm.SentencePieceTrainer.train(input=str(corpus), model_prefix=str(root / "tokenizer"), vocab_size=64, model_type="bpe",
                                   bos_id=1, eos_id=2, unk_id=0, pad_id=-1, minloglevel=2)
    # two Meta-style model-parallel shards (`llama_vqa.py:25-58`): column-parallel on dim 0, row-parallel / embeddings on dim 1
    col = ("attention.wq.weight", "attention.wk.weight", "attention.wv.weight", "feed_forward.w1.weight", "feed_forward.w3.weight", "output.weight")
    row = ("attention.wo.weight", "feed_forward.w2.weight", "tok_embeddings.weight")
    shards = [{}, {}]
    for name, t in sd.items():
        dim = 0 if name.endswith(col) else 1 if name.endswith(row) else -1
        for r in range(2):
            shards[r][name] = t.half() if dim < 0 else t.half().chunk(2, dim=dim)[r].clone()
    for r in range(2):
        shards[r]["rope.freqs"] = torch.zeros(8)                       # Meta's extra key, tolerated by strict=False (`:68`)
        torch.save(shards[r], model_dir / f"consolidated.0{r}.pth")


def test_reference_model_factory_builds_the_product_model(tmp_path, monkeypatch):
    import flipped_vqa_b200.llama as our_llama
    from flipped_vqa_b200 import _lib
    from flipped_vqa_b200.synthetic import synthetic_state_dict
    from types import SimpleNamespace
    params = dict(dim=128, n_layers=2, n_heads=2, multiple_of=64, norm_eps=1e-6, vocab_size=-1)
    full = synthetic_state_dict(SimpleNamespace(adapter_len=10, adapter_layer=2, **{**params, "vocab_size": 64}), seed=3)
    frozen = {k: v for k, v in full.items() if not any(s in k for s in ("gate", "adapter", "temporal_emb", "visual_proj"))}
    _make_checkpoint_dir(tmp_path, params, frozen)
    # --- the patch of INTEGRATION.md: `llama` resolves to this package; the file itself is the reference's, unmodified
    monkeypatch.setitem(sys.modules, "llama", our_llama)
    monkeypatch.setattr(torch, "set_default_tensor_type", lambda t: None)     # `llama_vqa.py:63-65` needs a GPU; dtype/device are the model's own
    spec = importlib.util.spec_from_file_location("ref_llama_vqa", os.path.join(ref_shims.REFERENCE_ROOT, "llama_vqa.py"))
    ref_factory = importlib.util.module_from_spec(spec)
    sys.dont_write_bytecode = True
    spec.loader.exec_module(ref_factory)
    args = make_args()
    args.llama_model_path, args.model, args.max_seq_len, args.adapter_len, args.adapter_layer = str(tmp_path) + "/", "7B", 64, 10, 2
    model = ref_factory.LLaMA_VQA(args)                                       # the reference's code path end to end
    assert isinstance(model, our_llama.Transformer) and model.vocab_size == 64 and model.tokenizer.n_words == 64
    names = dict(model.named_parameters())
    for k, v in frozen.items():                                               # merged shards landed in the packed layout, in the operand dtype
        assert names[k].dtype == _lib.H16 and not names[k].requires_grad, k
        assert torch.equal(names[k].detach().float().cpu(), v.half().float()), k
    trainable = {n for n, p in names.items() if p.requires_grad}
    assert trainable == {n for n in names if any(s in n for s in ("gate", "adapter", "temporal_emb", "visual_proj"))}
    assert all(names[n].dtype == torch.float32 for n in trainable)
    # the packed views survived load_state_dict: wq / wk / wv alias ONE [3d, d] buffer (one GEMM), no repack copy needed later
    blk = model.layers[0]
    assert blk.attention.wk.weight.data_ptr() == blk._wqkv[128:].data_ptr()
    # and our own factory (same contract) builds the same model from the same directory
    from flipped_vqa_b200.llama_vqa import LLaMA_VQA
    ours = LLaMA_VQA(args)
    for (n, a), (_, b) in zip(model.named_parameters(), ours.named_parameters()):
        if not a.requires_grad:
            assert torch.equal(a, b), n
