"""GPU diagnostic (checker, lives under tests/ because it runs the oracle): per-parameter parity report of the CUDA path vs the fp32 oracle (run on the GPU)
at a chosen depth of 7B-shaped layers. Not a test; prints a table."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from types import SimpleNamespace
import torch
from oracle import llama_vqa_oracle as O
from tests.util_parity import build_product_model, make_args, product_grads, rel_l2

def main():
    L = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    S = int(sys.argv[3]) if len(sys.argv) > 3 else 128
    pd = dict(dim=4096, n_layers=L, n_heads=32, vocab_size=32000, multiple_of=256, norm_eps=1e-6, max_batch_size=32,
              max_seq_len=S, adapter_len=10, adapter_layer=L)
    from flipped_vqa_b200.synthetic import synthetic_batch
    g = torch.Generator(device="cuda").manual_seed(0)
    rn = lambda *s, std=0.02, mean=0.0: (torch.randn(*s, device="cuda", generator=g) * std + mean).to(torch.bfloat16).float()
    d, hid, V, H = 4096, 11008, 32000, 32
    sd = {"tok_embeddings.weight": rn(V, d), "output.weight": rn(V, d), "norm.weight": rn(d, std=0.1, mean=1.0),
          "adapter_query.weight": rn(10 * L, d, std=1.0), "visual_proj.weight": rn(d, 768, std=0.036), "temporal_emb.weight": rn(10, d, std=1.0)}
    for i in range(L):
        p = f"layers.{i}."
        for nm in ("wq", "wk", "wv", "wo"):
            sd[p + f"attention.{nm}.weight"] = rn(d, d)
        sd[p + "feed_forward.w1.weight"] = rn(hid, d); sd[p + "feed_forward.w2.weight"] = rn(d, hid); sd[p + "feed_forward.w3.weight"] = rn(hid, d)
        sd[p + "attention_norm.weight"] = rn(d, std=0.1, mean=1.0); sd[p + "ffn_norm.weight"] = rn(d, std=0.1, mean=1.0)
        sd[p + "attention.gate1"] = rn(1, H, 1, 1, std=0.5); sd[p + "attention.gate2"] = rn(1, H, 1, 1, std=0.1, mean=-3.5)
    data = synthetic_batch(B, S, V, seed=5, full_length=(S > 400))
    args = make_args()
    model = build_product_model(pd, sd, args)
    t0 = time.time()
    vqa, vaq, qav = model(data); (vqa + vaq + qav).backward(); torch.cuda.synchronize()
    print("product step", time.time() - t0, "s")
    losses = [float(vqa), float(vaq), float(qav)]
    grads = product_grads(model)
    del model; torch.cuda.empty_cache()
    st = O.prepare_state(sd, frozen_dtype=torch.float32, device="cuda")
    t0 = time.time()
    ref = O.forward_losses(st, SimpleNamespace(**pd), data, max_feats=10, tau=100.0)
    sum(ref).backward(); torch.cuda.synchronize()
    print("oracle fp32 step", time.time() - t0, "s")
    for n, a, b in zip(("vqa", "vaq", "qav"), losses, ref):
        print(f"loss {n}: ours {a:.5f} oracle {float(b):.5f} rel {abs(a - float(b)) / abs(float(b)):.2e}")
    for grp in ("gate1", "gate2"):
        names = sorted((n for n in grads if n.endswith(grp)), key=lambda s: int(s.split(".")[1]))
        a = torch.cat([grads[n].flatten() for n in names]); b = torch.cat([st[n].grad.flatten().cpu() for n in names])
        print(f"{grp} stacked rel L2 {rel_l2(a, b):.3e}; per layer:", " ".join(f"{rel_l2(grads[n], st[n].grad):.3f}" for n in names))
    for n in ("adapter_query.weight", "visual_proj.weight", "temporal_emb.weight"):
        print(f"{n}: rel L2 {rel_l2(grads[n], st[n].grad):.3e}")
    aq, bq = grads["adapter_query.weight"].view(L, 10, d), st["adapter_query.weight"].grad.cpu().view(L, 10, d)
    print("adapter per layer:", " ".join(f"{rel_l2(aq[l], bq[l]):.3f}" for l in range(L)))
    # noise floor: the SAME oracle (= the reference's op sequence under autograd) run in bf16 / fp16
    gold = {n: st[n].grad.detach().cpu() for n in O.trainable_names(st) if st[n].grad is not None}
    gold_loss = [float(x) for x in ref]
    del st, ref; torch.cuda.empty_cache()
    for dt in (torch.bfloat16, torch.float16):
        st2 = O.prepare_state(sd, frozen_dtype=dt, device="cuda")
        r2 = O.forward_losses(st2, SimpleNamespace(**pd), data, max_feats=10, tau=100.0)
        sum(r2).backward(); torch.cuda.synchronize()
        print(f"--- oracle in {dt} vs fp32 oracle: loss rel", [f"{abs(float(a) - b) / abs(b):.2e}" for a, b in zip(r2, gold_loss)])
        for grp in ("gate1", "gate2"):
            names = sorted((n for n in gold if n.endswith(grp)), key=lambda s: int(s.split(".")[1]))
            a = torch.cat([st2[n].grad.flatten().cpu() for n in names]); b = torch.cat([gold[n].flatten() for n in names])
            print(f"    {grp} stacked rel L2 {rel_l2(a, b):.3e}")
        for n in ("adapter_query.weight", "visual_proj.weight", "temporal_emb.weight"):
            print(f"    {n}: rel L2 {rel_l2(st2[n].grad, gold[n]):.3e}")
        del st2, r2; torch.cuda.empty_cache()

if __name__ == "__main__":
    main()
