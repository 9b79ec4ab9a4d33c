"""GPU diagnostic (checker; lives under tests/ because it runs the oracle): WHERE does the full-depth gradient error of
the CUDA path come from?  The step's op sequence is restated in fp32 PyTorch (autograd) with a fake-quantiser `Q` at every
point where the product rounds an operand to 16 bits, forward and backward separately:

    act   forward GEMM operands / saved activations (xn, qkv, o, [a|b], c, hn)
    p     probabilities fed to the P.V tensor-core product
    grad  backward GEMM operands (d xn, d qkv, d o, d[a|b], dc, dx, dlogits, dhn)
    ds    dS fed to the dQ / dK tensor-core products
    adp   adapter prompt, adapter K|V and their gradients

Each class is switched between fp32 (off), bf16 and fp16 and the trainable gradients are compared with the fp32 oracle.
Usage: python tests/numerics_ablation.py [L=32] [B=8] [S=128]   (prints a table; ~6 s per variant at 7B/32 layers).
"""
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from types import SimpleNamespace

import torch
import torch.nn.functional as Fn

from oracle import llama_vqa_oracle as O
from tests.util_parity import rel_l2

DT = {"bf16": torch.bfloat16, "fp16": torch.float16, None: None}


class _Q(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, f, b):
        ctx.b = b
        return x.to(f).float() if f is not None else x

    @staticmethod
    def backward(ctx, g):
        return (g.to(ctx.b).float() if ctx.b is not None else g), None, None


def Q(x, f=None, b=None):
    return _Q.apply(x, DT[f], DT[b]) if (f is not None or b is not None) else x


def rms(x, w, eps):
    return x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps) * w


def rope(x, cos, sin):
    xf = x.reshape(*x.shape[:-1], -1, 2)
    a, b = xf[..., 0], xf[..., 1]
    c, s = cos[None, :, None, :], sin[None, :, None, :]
    return torch.stack((a * c - b * s, a * s + b * c), dim=-1).flatten(3)


def block(x, sd, p, adapter, cos, sin, H, eps, vs, Fv, c):
    B, S, d = x.shape
    hd = d // H
    g = lambda n: sd[p + n]
    xn = Q(rms(x, g("attention_norm.weight"), eps), c.act, c.grad)
    q = rope((xn @ g("attention.wq.weight").t()).view(B, S, H, hd), cos[:S], sin[:S])
    k = rope((xn @ g("attention.wk.weight").t()).view(B, S, H, hd), cos[:S], sin[:S])
    v = (xn @ g("attention.wv.weight").t()).view(B, S, H, hd)
    q, k, v = (Q(t, c.act, c.grad).transpose(1, 2) for t in (q, k, v))           # stored qkv / dqkv
    ad = Q(adapter, c.adp, None)
    A = ad.shape[0]
    ak = Q(ad @ g("attention.wk.weight").t(), c.adp, c.adp).view(A, H, hd).transpose(0, 1)[None]
    av = Q(ad @ g("attention.wv.weight").t(), c.adp, c.adp).view(A, H, hd).transpose(0, 1)[None]
    sc = Q((q @ k.transpose(2, 3)) / math.sqrt(hd), None, c.ds)
    sa = Q((q @ ak.transpose(2, 3)) / math.sqrt(hd), None, c.ds)
    mask = torch.triu(torch.full((S, S), float("-inf"), device=x.device), diagonal=1)
    sc = sc + mask
    if vs is not None:
        bias = torch.zeros(S, S, device=x.device)
        bias[vs + Fv:, vs:vs + Fv] = 1.0
        sc = sc + bias[None, None] * g("attention.gate2")
    pt = Q(Fn.softmax(sc, -1), c.p, None)
    pa = Q(Fn.softmax(sa, -1) * g("attention.gate1").tanh(), c.p, None)
    o = (pt @ v + pa @ av).transpose(1, 2).reshape(B, S, d)
    o = Q(o, c.act, c.grad)
    h = x + o @ g("attention.wo.weight").t()
    xn2 = Q(rms(h, g("ffn_norm.weight"), eps), c.act, c.grad)
    a = Q(xn2 @ g("feed_forward.w1.weight").t(), c.act, c.grad)
    b = Q(xn2 @ g("feed_forward.w3.weight").t(), c.act, c.grad)
    cc = Q(Fn.silu(a) * b, c.act, c.grad)
    return h + cc @ g("feed_forward.w2.weight").t()


def emulated_losses(sd, params, data, c, max_feats=10, tau=100.0):
    dev = sd["tok_embeddings.weight"].device
    ids = {k: data["text_id"][k].to(dev) for k in ("vqa", "vaq", "qav")}
    lab = {k: data["label"][k].to(dev) for k in ("vqa", "vaq", "qav")}
    vs_vqa, vs_vaq = int(data["video_start"]["vqa"][0]), int(data["video_start"]["vaq"][0])
    qav_index = data["video_index"]["qav"].to(dev)
    bsz, n_opt, S = ids["vqa"].shape
    d, H, L = params.dim, params.n_heads, params.n_layers
    cos, sin = O.rope_table(d // H, params.max_seq_len * 2)
    cos, sin = cos.to(dev), sin.to(dev)
    emb = sd["tok_embeddings.weight"]
    adapter = sd["adapter_query.weight"].reshape(-1, params.adapter_len, d)
    _vf = data["video"].to(dev) @ sd["visual_proj.weight"].t()
    vf = _vf + sd["temporal_emb.weight"][None]

    def run(h, vs):
        for i in range(L):
            h = block(h, sd, f"layers.{i}.", adapter[i], cos, sin, H, params.norm_eps, vs, max_feats, c)
        return h

    def ce(h, label):
        rows = (label != 0).nonzero().flatten()
        hn = Q(rms(h[:, :-1].reshape(-1, d)[rows], sd["norm.weight"], params.norm_eps), c.act, c.grad)
        logits = Q(hn @ sd["output.weight"].t(), None, c.grad)
        return Fn.cross_entropy(logits, label[rows])

    out = []
    for k, vs in (("vqa", vs_vqa), ("vaq", vs_vaq)):
        h = emb[ids[k].reshape(-1, S)].detach().clone()
        h[:, vs:vs + max_feats] = vf
        out.append(ce(run(h, vs), lab[k].reshape(-1, S)[:, 1:].flatten()))
    qfull = lab["qav"].reshape(-1, S)
    h = emb[ids["qav"].reshape(-1, S)].detach() * (~qfull.ge(0))[..., None]
    h = h.scatter_add(1, qav_index[..., None].expand(-1, -1, d), vf)
    h = run(h, None)
    hn = Q(rms(h, sd["norm.weight"], params.norm_eps), c.act, None)
    lg = torch.bmm(hn[:, :-1], _vf.transpose(1, 2))
    out.append(Fn.cross_entropy(lg.reshape(-1, max_feats) / tau, qfull[:, 1:].flatten(), ignore_index=-1))
    return out


def cfg(act=None, p=None, grad=None, ds=None, adp=None):
    return SimpleNamespace(act=act, p=p, grad=grad, ds=ds, adp=adp)


VARIANTS = [
    ("fp32 (sanity: emulation == oracle)", cfg(), 1.0),
    ("all bf16  (= product today)", cfg("bf16", "bf16", "bf16", "bf16", "bf16"), 1.0),
    ("fwd bf16 only (act, p, adp)", cfg("bf16", "bf16", None, None, "bf16"), 1.0),
    ("bwd bf16 only (grad, ds)", cfg(None, None, "bf16", "bf16", None), 1.0),
    ("act bf16 only", cfg("bf16"), 1.0),
    ("p bf16 only", cfg(None, "bf16"), 1.0),
    ("grad bf16 only", cfg(None, None, "bf16"), 1.0),
    ("ds bf16 only", cfg(None, None, None, "bf16"), 1.0),
    ("adp bf16 only", cfg(None, None, None, None, "bf16"), 1.0),
    ("all fp16, grad scale 2^12", cfg("fp16", "fp16", "fp16", "fp16", "fp16"), 4096.0),
    ("all fp16, grad scale 1 (underflow check)", cfg("fp16", "fp16", "fp16", "fp16", "fp16"), 1.0),
    ("all fp16, grad scale 2^20 (overflow check)", cfg("fp16", "fp16", "fp16", "fp16", "fp16"), 2.0 ** 20),
    ("fwd fp16 (act, p, adp), bwd bf16", cfg("fp16", "fp16", "bf16", "bf16", "fp16"), 1.0),
    ("fwd fp16, bwd fp32", cfg("fp16", "fp16", None, None, "fp16"), 1.0),
    ("act fp16, p bf16, bwd fp16 2^12", cfg("fp16", "bf16", "fp16", "fp16", "fp16"), 4096.0),
    ("act fp16, p fp16, grad fp16 2^12, ds bf16", cfg("fp16", "fp16", "fp16", "bf16", "fp16"), 4096.0),
]


def main():
    L = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    S = int(sys.argv[3]) if len(sys.argv) > 3 else 128
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    d, hid, V, H = (4096, 11008, 32000, 32) if dev == "cuda" else (256, 768, 512, 4)     # CPU: tiny dims, script self-check only
    pd = dict(dim=d, n_layers=L, n_heads=H, vocab_size=V, multiple_of=256, norm_eps=1e-6, max_batch_size=32,
              max_seq_len=S, adapter_len=10, adapter_layer=L)
    from flipped_vqa_b200.synthetic import synthetic_batch
    g = torch.Generator(device=dev).manual_seed(0)
    rn = lambda *s, std=0.02, mean=0.0: (torch.randn(*s, device=dev, generator=g) * std + mean).to(torch.bfloat16).float()
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
    params = SimpleNamespace(**pd)
    st = O.prepare_state(sd, frozen_dtype=torch.float32, device=dev)
    names = O.trainable_names(st)
    ref = O.forward_losses(st, params, data, max_feats=10, tau=100.0)
    sum(ref).backward()
    gold = {n: st[n].grad.detach().clone() for n in names}
    gold_loss = [float(x) for x in ref]
    groups = ("adapter_query.weight", "visual_proj.weight", "temporal_emb.weight", "gate1", "gate2")

    def stacked(gr, grp):
        ns = sorted((n for n in names if n.endswith(grp)), key=lambda s: (len(s), s))
        return torch.cat([gr[n].flatten() for n in ns])

    print(f"# numerics ablation: 7B-shaped, L={L} B={B} S={S}; rel L2 of the trainable gradients vs the fp32 oracle")
    print(f"{'variant':46s} {'loss rel (max)':>14s} " + " ".join(f"{g_[:12]:>12s}" for g_ in groups))
    for name, c, scale in VARIANTS:
        for n in names:
            st[n].grad = None
        t0 = time.time()
        ls = emulated_losses(st, params, data, c)
        (sum(ls) * scale).backward()
        if dev == "cuda":
            torch.cuda.synchronize()
        gr = {n: st[n].grad.detach() / scale for n in names}
        lrel = max(abs(float(a) - b) / abs(b) for a, b in zip(ls, gold_loss))
        errs = [rel_l2(stacked(gr, grp), stacked(gold, grp)) for grp in groups]
        print(f"{name:46s} {lrel:14.2e} " + " ".join(f"{e:12.3e}" for e in errs) + f"   ({time.time() - t0:.1f}s)", flush=True)


if __name__ == "__main__":
    main()
