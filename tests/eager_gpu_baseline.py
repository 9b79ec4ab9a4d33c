"""GPU diagnostic (checker side, lives under tests/ because it executes the oracle): wall time of the ORACLE — the
reference's op sequence (llama/model.py:250-365: three streams run one after another, unfused attention with
materialised scores, full-vocabulary logits, autograd) — in eager PyTorch on the SAME B200, at the 7B NExT-QA bench shape.
This is the "stock eager PyTorch" comparator SURVEY.md 8(d) asks for; the reference itself cannot travel to the GPU box.

    python tests/eager_gpu_baseline.py [bf16|fp16] [layers]
"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from types import SimpleNamespace
import torch
from oracle import llama_vqa_oracle as O
from flipped_vqa_b200.synthetic import synthetic_batch


def main():
    dt = {"bf16": torch.bfloat16, "fp16": torch.float16}[sys.argv[1] if len(sys.argv) > 1 else "bf16"]
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    d, hid, V, H, B, S = 4096, 11008, 32000, 32, 8, 128
    pd = dict(dim=d, n_layers=L, n_heads=H, vocab_size=V, multiple_of=256, norm_eps=1e-6, max_batch_size=32, max_seq_len=S,
              adapter_len=10, adapter_layer=L)
    g = torch.Generator(device="cuda").manual_seed(0)
    rn = lambda *s, std=0.02, mean=0.0: (torch.randn(*s, device="cuda", generator=g) * std + mean)
    sd = {"tok_embeddings.weight": rn(V, d), "output.weight": rn(V, d), "norm.weight": rn(d, std=0.1, mean=1.0),
          "adapter_query.weight": rn(10 * L, d, std=1.0), "visual_proj.weight": rn(d, 768, std=0.036), "temporal_emb.weight": rn(10, d, std=1.0)}
    for i in range(L):
        p = f"layers.{i}."
        for nm in ("wq", "wk", "wv", "wo"):
            sd[p + f"attention.{nm}.weight"] = rn(d, d).to(dt)
        sd[p + "feed_forward.w1.weight"] = rn(hid, d).to(dt); sd[p + "feed_forward.w2.weight"] = rn(d, hid).to(dt); sd[p + "feed_forward.w3.weight"] = rn(hid, d).to(dt)
        sd[p + "attention_norm.weight"] = rn(d, std=0.1, mean=1.0); sd[p + "ffn_norm.weight"] = rn(d, std=0.1, mean=1.0)
        sd[p + "attention.gate1"] = rn(1, H, 1, 1, std=0.5); sd[p + "attention.gate2"] = rn(1, H, 1, 1, std=0.1, mean=-3.5)
    st = O.prepare_state(sd, frozen_dtype=dt, device="cuda")
    del sd
    trainables = [st[n] for n in O.trainable_names(st)]
    opt = torch.optim.AdamW(trainables, lr=1e-4, betas=(0.9, 0.95), weight_decay=0.05, fused=True)
    batches = [synthetic_batch(B, S, V, seed=i) for i in range(2)]

    def step(i):
        losses = O.forward_losses(st, SimpleNamespace(**pd), batches[i % 2], max_feats=10, tau=100.0)
        sum(losses).backward()
        opt.step(); opt.zero_grad(set_to_none=True)

    for i in range(3):
        step(i)
    torch.cuda.synchronize()
    n = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        step(i)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(json.dumps({"what": "oracle (reference op sequence) in eager PyTorch on this GPU", "dtype": str(dt), "layers": L,
                      "ms_per_step": ms, "samples_per_s": B / (ms * 1e-3), "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}))


if __name__ == "__main__":
    main()
