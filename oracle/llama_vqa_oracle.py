"""ORACLE (test infrastructure only — never imported by the product path).

A plain-PyTorch restatement of the reference's LLaMA-VQA training step and loss-based option
scoring, written functionally over a state dict so that it can run in fp32 ("gold") on CPU or on
whatever device its inputs live on. Gradients come from autograd. Each function cites the
reference lines (relative to /root/reference) it restates.

Pinned against the reference itself: `oracle/make_golden.py` imports the real
`llama/model.py` / `llama/model_my_original_mod.py` (via `oracle/ref_shims.py`) in this container,
runs them on the deterministic inputs of `flipped_vqa_b200.synthetic` and commits the resulting
losses / gradients / option scores under `tests/golden/`; `tests/test_oracle_golden.py` checks
this file against those fixtures (the reference ships no tests or golden vectors of its own —
SURVEY.md §4, §8(c)).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` leg may
import this module.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------
# building blocks
# ----------------------------------------------------------------------------------------------
def rmsnorm(x: torch.Tensor, weight: torch.Tensor, eps: float) -> torch.Tensor:
    """`llama/model.py:37-42`: normalise in fp32, cast back to x's dtype, THEN multiply by weight."""
    xf = x.float()
    normed = (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)).to(x.dtype)
    return normed * weight


def rope_table(head_dim: int, end: int, theta: float = 10000.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """`llama/model.py:45-50`: angle[pos, i] = pos * theta^(-2i/head_dim); returned as (cos, sin)."""
    inv = 1.0 / (theta ** (torch.arange(0, head_dim, 2)[: head_dim // 2].float() / head_dim))
    ang = torch.outer(torch.arange(end).float(), inv)
    return torch.cos(ang), torch.sin(ang)


def apply_rope(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """`llama/model.py:61-67`: rotate interleaved pairs (2i, 2i+1) in fp32, cast back.
    x: [B, S, H, hd]; cos/sin: [S, hd/2]."""
    xf = x.float().reshape(*x.shape[:-1], -1, 2)
    a, b = xf[..., 0], xf[..., 1]
    c = cos.to(x.device)[None, :, None, :]
    s = sin.to(x.device)[None, :, None, :]
    out = torch.stack((a * c - b * s, a * s + b * c), dim=-1).flatten(3)
    return out.to(x.dtype)


def attention_core(q, k, v, ak, av, gate1, gate2, cos, sin, video_start: Optional[int], max_feats: int) -> torch.Tensor:
    """`llama/model.py:96-126` after the projections. q,k,v [B,S,H,hd] (pre-RoPE); ak,av [A,H,hd]
    adapter keys/values (no RoPE, shared over the batch, `:99-100`). Returns [B,S,H*hd].
    Separate softmax over the A adapter keys scaled by tanh(gate1) (`:115`); gate2 added to text
    scores of rows >= vs+F, columns vs..vs+F (`:116-119`); video_start=None (QAV) -> plain softmax
    in the activation dtype (`:121-122`)."""
    B, S, H, hd = q.shape
    dt = q.dtype
    A = ak.shape[0]
    q, k = apply_rope(q, cos[:S], sin[:S]), apply_rope(k, cos[:S], sin[:S])
    akb = ak[None].expand(B, -1, -1, -1)
    avb = av[None].expand(B, -1, -1, -1)
    q = q.transpose(1, 2)                                    # [B,H,S,hd]
    keys = torch.cat([akb, k], dim=1).transpose(1, 2)        # [B,H,A+S,hd]
    vals = torch.cat([avb, v], dim=1).transpose(1, 2)
    scores = (q @ keys.transpose(2, 3)) / math.sqrt(hd)
    causal = torch.triu(torch.full((S, S), float("-inf"), device=q.device), diagonal=1).to(dt)
    mask = torch.cat([torch.zeros(S, A, device=q.device, dtype=dt), causal], dim=-1)
    scores = scores + mask[None, None]
    p_adapter = F.softmax(scores[..., :A].float(), dim=-1).to(dt) * gate1.tanh().to(dt)
    text = scores[..., A:]
    if video_start is not None:
        vs, Fv = video_start, max_feats
        text = text.clone()
        text[:, :, vs + Fv:, vs:vs + Fv] = text[:, :, vs + Fv:, vs:vs + Fv] + gate2.to(dt)
        p_text = F.softmax(text.float(), dim=-1).to(dt)
    else:
        p_text = F.softmax(text, dim=-1)                     # QAV: softmax in the activation dtype
    out = torch.cat([p_adapter, p_text], dim=-1) @ vals      # [B,H,S,hd]
    return out.transpose(1, 2).contiguous().view(B, S, H * hd)


def attention(x, wq, wk, wv, wo, gate1, gate2, adapter, cos, sin, n_heads: int,
              video_start: Optional[int], max_feats: int) -> torch.Tensor:
    """`llama/model.py:87-128`. x [B,S,d]; adapter [A,d]."""
    B, S, d = x.shape
    hd = d // n_heads
    q = (x @ wq.t()).view(B, S, n_heads, hd)
    k = (x @ wk.t()).view(B, S, n_heads, hd)
    v = (x @ wv.t()).view(B, S, n_heads, hd)
    A = adapter.shape[0]
    ak = (adapter @ wk.t()).view(A, n_heads, hd)
    av = (adapter @ wv.t()).view(A, n_heads, hd)
    out = attention_core(q, k, v, ak, av, gate1, gate2, cos, sin, video_start, max_feats)
    return out @ wo.t()


def feed_forward(x, w1, w2, w3) -> torch.Tensor:
    """`llama/model.py:141-142`: w2( silu(w1 x) * w3 x )."""
    return (F.silu(x @ w1.t()) * (x @ w3.t())) @ w2.t()


def block(x, sd, prefix: str, adapter, cos, sin, n_heads, eps, video_start, max_feats):
    """`llama/model.py:184-187`: pre-norm residual block."""
    g = lambda n: sd[prefix + n]
    h = x + attention(rmsnorm(x, g("attention_norm.weight"), eps),
                      g("attention.wq.weight"), g("attention.wk.weight"), g("attention.wv.weight"),
                      g("attention.wo.weight"), g("attention.gate1"), g("attention.gate2"),
                      adapter, cos, sin, n_heads, video_start, max_feats)
    return h + feed_forward(rmsnorm(h, g("ffn_norm.weight"), eps),
                            g("feed_forward.w1.weight"), g("feed_forward.w2.weight"),
                            g("feed_forward.w3.weight"))


# ----------------------------------------------------------------------------------------------
# the training step and option scoring
# ----------------------------------------------------------------------------------------------
def prepare_state(sd: Dict[str, torch.Tensor], frozen_dtype=torch.float32, device="cpu",
                  requires_grad: bool = True) -> Dict[str, torch.Tensor]:
    """Freeze/dtype rule of `llama_vqa.py:71-76`: names containing gate/adapter/temporal_emb/
    visual_proj are trainable fp32 leaves; everything else is frozen in ``frozen_dtype``."""
    out = {}
    for n, t in sd.items():
        if any(s in n for s in ("gate", "adapter", "temporal_emb", "visual_proj")):
            p = t.detach().clone().float().to(device)
            p.requires_grad_(requires_grad)
        else:
            p = t.detach().to(frozen_dtype).to(device)
        out[n] = p
    return out


def trainable_names(sd) -> list:
    return [n for n in sd if any(s in n for s in ("gate", "adapter", "temporal_emb", "visual_proj"))]


def _run_layers(h, sd, params, adapter, cos, sin, video_start, max_feats):
    L, AL = params.n_layers, params.adapter_layer
    # only the last `adapter_layer` layers run (`llama/model.py:338`)
    for i, li in enumerate(range(L)[-AL:]):
        h = block(h, sd, f"layers.{li}.", adapter[i].to(h.dtype), cos, sin, params.n_heads,
                  params.norm_eps, video_start, max_feats)
    return h


def fused_video_feature(sd, data, audio_mode: Optional[str], dev, dt):
    """`_video_feature` of the five input-fusion branches (`llama/model.py:306-322`; `CrossAttentionModule` `:145-169`)."""
    video = data["video"].to(dev) if audio_mode != "audio_only" else None
    audio = data["audio"].to(dev).to(dt) if audio_mode is not None else None          # `.cuda().half()`, `:258-261`
    if audio_mode is None:
        return video @ sd["visual_proj.weight"].t()                                   # fp32, `:322`
    if audio_mode == "audio_only":
        return audio @ sd["audio_proj.weight"].t()                                    # `:307`
    if audio_mode == "concat":
        return torch.cat([video, audio], dim=-1) @ sd["visual_proj.weight"].t()       # `:310-311`
    if audio_mode == "sum":
        return (audio @ sd["audio_proj.weight"].t()).to(dt) + (video @ sd["visual_proj.weight"].t()).to(dt)   # `:314`
    if audio_mode == "attention":
        af = (audio @ sd["audio_proj.weight"].t()).float()                            # [B, Fa, 768], `:317`, `:158`
        lin = lambda n, x: x @ sd[f"video_audio_cross_attn.{n}.weight"].float().t() + sd[f"video_audio_cross_attn.{n}.bias"].float()
        q, k, v = lin("query", video), lin("key", af), lin("value", af)
        p = torch.softmax(q @ k.transpose(-2, -1) / math.sqrt(q.shape[-1]), dim=-1)  # `:165-168`
        return ((p @ v) @ sd["visual_proj.weight"].t()).to(dt)                         # `:318-320`
    raise ValueError(audio_mode)


def forward_losses(sd, params, data, max_feats: int = 10, tau: float = 100.0,
                   vaq: bool = True, qav: bool = True, audio_mode: Optional[str] = None):
    """`llama/model.py:250-365`. Returns (vqa_loss, vaq_loss, qav_loss). `audio_mode`: see `fused_video_feature`."""
    dev = sd["tok_embeddings.weight"].device
    dt = sd["tok_embeddings.weight"].dtype
    ids = {k: data["text_id"][k].to(dev) for k in ("vqa", "vaq", "qav")}
    lab = {k: data["label"][k].to(dev) for k in ("vqa", "vaq", "qav")}
    vs_vqa, vs_vaq = int(data["video_start"]["vqa"][0]), int(data["video_start"]["vaq"][0])  # sample 0 only, `:264`
    qav_index = data["video_index"]["qav"].to(dev)
    bsz, n_opt, S = ids["vqa"].shape
    d = params.dim
    cos, sin = rope_table(d // params.n_heads, params.max_seq_len * 2)
    cos, sin = cos.to(dev), sin.to(dev)

    vqa_id, vaq_id, qav_id = (ids[k].reshape(-1, S) for k in ("vqa", "vaq", "qav"))
    vqa_label = lab["vqa"].reshape(-1, S)[:, 1:].flatten()
    vaq_label = lab["vaq"].reshape(-1, S)[:, 1:].flatten()
    qav_full = lab["qav"].reshape(-1, S)
    qav_video_mask = qav_full.ge(0)
    qav_label = qav_full[:, 1:].flatten()

    emb = sd["tok_embeddings.weight"]
    adapter = sd["adapter_query.weight"].reshape(-1, params.adapter_len, d)       # `:304`
    _video_feature = fused_video_feature(sd, data, audio_mode, dev, dt)            # `:306-322`
    video_feature = (_video_feature + sd["temporal_emb.weight"][None]).to(dt)     # `:324`

    def inject(idmat, vs):
        h = emb[idmat].detach().clone()
        h[:, vs:vs + max_feats] = video_feature                                    # `:326-332`
        return h

    zero = torch.zeros(1, dtype=torch.int64, device=dev)
    vqa_h = _run_layers(inject(vqa_id, vs_vqa), sd, params, adapter, cos, sin, vs_vqa, max_feats)
    vqa_h = rmsnorm(vqa_h, sd["norm.weight"], params.norm_eps)
    logits = (vqa_h @ sd["output.weight"].t())[:, :-1].reshape(-1, params.vocab_size)
    vqa_loss = F.cross_entropy(logits, vqa_label, ignore_index=0)                 # `:347-350`
    vaq_loss, qav_loss = zero, zero
    if vaq:
        vaq_h = _run_layers(inject(vaq_id, vs_vaq), sd, params, adapter, cos, sin, vs_vaq, max_feats)
        vaq_h = rmsnorm(vaq_h, sd["norm.weight"], params.norm_eps)
        logits = (vaq_h @ sd["output.weight"].t())[:, :-1].reshape(-1, params.vocab_size)
        vaq_loss = F.cross_entropy(logits, vaq_label, ignore_index=0)             # `:352-356`
    if qav:
        h = emb[qav_id].detach() * (~qav_video_mask)[..., None]                    # `:335`
        h = h.scatter_add(1, qav_index[..., None].expand(-1, -1, d), video_feature)  # `:336`
        qav_h = _run_layers(h, sd, params, adapter, cos, sin, None, max_feats)
        qav_h = rmsnorm(qav_h, sd["norm.weight"], params.norm_eps)
        out = torch.bmm(qav_h[:, :-1].float(), _video_feature.transpose(1, 2).float())
        qav_loss = F.cross_entropy(out.reshape(-1, max_feats) / tau, qav_label, ignore_index=-1)  # `:358-361`
    return vqa_loss, vaq_loss, qav_loss


def option_token_losses(sd, params, data, max_feats: int = 10) -> torch.Tensor:
    """`llama/model_my_original_mod.py:281,332-333,348-360,375-377,506`: VQA stream only over
    bsz*n_options sequences, video features repeated per option, per-token CE (ignore_index=0,
    reduction='none') reshaped to [bsz, n_options, S-1]."""
    dev = sd["tok_embeddings.weight"].device
    dt = sd["tok_embeddings.weight"].dtype
    video = data["video"].to(dev)
    ids = data["text_id"]["vqa"].to(dev)
    lab = data["label"]["vqa"].to(dev)
    vs = int(data["video_start"]["vqa"][0])
    bsz, n_opt, S = ids.shape
    d = params.dim
    cos, sin = rope_table(d // params.n_heads, params.max_seq_len * 2)
    cos, sin = cos.to(dev), sin.to(dev)
    adapter = sd["adapter_query.weight"].reshape(-1, params.adapter_len, d)
    vf = video @ sd["visual_proj.weight"].t()
    vf = vf.unsqueeze(1).repeat(1, n_opt, 1, 1).view(-1, vf.shape[-2], vf.shape[-1])
    video_feature = (vf + sd["temporal_emb.weight"][None]).to(dt)
    h = sd["tok_embeddings.weight"][ids.reshape(-1, S)].detach().clone()
    h[:, vs:vs + max_feats] = video_feature
    h = _run_layers(h, sd, params, adapter, cos, sin, vs, max_feats)
    h = rmsnorm(h, sd["norm.weight"], params.norm_eps)
    logits = (h @ sd["output.weight"].t())[:, :-1].reshape(-1, params.vocab_size)
    label = lab.reshape(-1, S)[:, 1:].flatten()
    tok = F.cross_entropy(logits, label, ignore_index=0, reduction="none")
    return tok.reshape(bsz, n_opt, -1)


def option_predict(token_losses: torch.Tensor) -> torch.Tensor:
    """`engine.py:88-93`: count = (loss != 0) per option; prediction = argmin(sum/count)."""
    count = (token_losses != 0).sum(-1)
    return (token_losses.sum(-1) / count).argmin(-1)


# ----------------------------------------------------------------------------------------------
# op-level restatements used by the kernel unit tests (fp32 math on bf16 data)
# ----------------------------------------------------------------------------------------------
def swiglu(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """`llama/model.py:142` inner product: silu(a) * b."""
    return F.silu(a) * b


def flops_per_step(dim, n_layers_run, hidden, vocab, bsz, seqlen, adapter_len, max_feats,
                   n_streams=3, n_ce=2, n_labelled=None) -> float:
    """Algorithmic FLOPs of one training step, SURVEY.md §8(d) (dense-position head convention
    unless ``n_labelled`` = (labelled rows per CE stream) is given)."""
    T = bsz * seqlen
    body = n_layers_run * 2 * T * (4 * dim * dim + 3 * dim * hidden)
    attn = n_layers_run * 4 * bsz * dim * (seqlen * (seqlen + 1) / 2 + seqlen * adapter_len)
    if n_labelled is None:
        head = [2 * bsz * (seqlen - 1) * dim * vocab] * n_ce
    else:
        head = [2 * n * dim * vocab for n in n_labelled]
    small = n_layers_run * 2 * adapter_len * 2 * dim * dim + 2 * bsz * max_feats * 768 * dim
    return n_streams * (2 * body + 3.5 * attn) + 2 * sum(head) + small
