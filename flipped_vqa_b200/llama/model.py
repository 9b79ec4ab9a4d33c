"""LLaMA-VQA model with the reference's Python surface, running on the B200 kernels.

Drop-in for `/root/reference/llama/model.py`: same `ModelArgs`, same `Transformer(params, args)`
constructor attributes, same parameter / state-dict names (so `llama_vqa.py`'s freeze rule and
`util/misc.py`'s trainable-only checkpoints keep working), same `forward(data, inference=False)`
contract returning `(vqa_loss, vaq_loss, qav_loss)` (`model.py:250-365`) or, with
``inference=True``, the per-token option losses `[bsz, n_options, S-1]` of
`llama/model_my_original_mod.py:375-377,506`.

Underneath, nothing of the reference's op sequence is kept: the three objective streams are
concatenated into one token matrix, every frozen Linear is the tcgen05 GEMM, attention / norms /
SwiGLU / heads are the fused kernels of libfvqa.so, and backward is hand-written (dX-only through
the frozen base). The nn.Module tree below is only a *named parameter container*.
There is no CPU path: calling forward without an sm_100 GPU raises.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import List, Optional

import torch
from torch import nn

from ..step import BatchPlan, GradBuffers, LayerWeights, OptionPlan, PinnedPool, StepEngine
from .. import ops
from ..synthetic import AUDIO_DIM, audio_mode, ffn_hidden_dim
from .tokenizer import Tokenizer

H16 = ops.H16


@dataclass
class ModelArgs:
    """Same fields and defaults as `llama/model.py:17-29` (max_feats / bias are injected by
    Transformer.__init__, `:193-194`)."""
    dim: int = 512
    n_layers: int = 8
    n_heads: int = 8
    vocab_size: int = -1
    multiple_of: int = 256
    norm_eps: float = 1e-5

    max_batch_size: int = 32
    max_seq_len: int = 2048
    adapter_len: int = 10
    adapter_layer: int = 30


class _FrozenLinear(nn.Module):
    """Named holder of a frozen [out, in] weight (a view into the packed device layout)."""

    def __init__(self, weight_view: torch.Tensor):
        super().__init__()
        self.weight = nn.Parameter(weight_view, requires_grad=False)


class RMSNorm(nn.Module):
    def __init__(self, dim: int, eps: float, device, dtype):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(dim, device=device, dtype=dtype), requires_grad=False)


class Attention(nn.Module):
    """Parameter names of `llama/model.py:71-85`: wq/wk/wv/wo (frozen) + gate1 (zeros) and
    gate2 (-bias) [1,H,1,1] fp32 trainables. The reference's dead KV caches are not allocated."""

    def __init__(self, wqkv: torch.Tensor, wo: torch.Tensor, n_heads: int, bias: float, device):
        super().__init__()
        d = wo.shape[0]
        self.wq = _FrozenLinear(wqkv[0:d])
        self.wk = _FrozenLinear(wqkv[d:2 * d])
        self.wv = _FrozenLinear(wqkv[2 * d:3 * d])
        self.wo = _FrozenLinear(wo)
        self.gate1 = nn.Parameter(torch.zeros(1, n_heads, 1, 1, device=device))
        self.gate2 = nn.Parameter(torch.ones(1, n_heads, 1, 1, device=device) * -bias)


class FeedForward(nn.Module):
    def __init__(self, w13: torch.Tensor, w2: torch.Tensor):
        super().__init__()
        hid = w2.shape[1]
        self.w1 = _FrozenLinear(w13[0:hid])
        self.w2 = _FrozenLinear(w2)
        self.w3 = _FrozenLinear(w13[hid:2 * hid])


class TransformerBlock(nn.Module):
    def __init__(self, layer_id: int, args: ModelArgs, hidden: int, device, dtype, init_std: float):
        super().__init__()
        d = args.dim
        self.layer_id = layer_id
        # packed device layout: one GEMM for Wq|Wk|Wv, one for W1|W3
        self._wqkv = torch.empty(3 * d, d, device=device, dtype=dtype).normal_(0, init_std)
        self._wo = torch.empty(d, d, device=device, dtype=dtype).normal_(0, init_std)
        self._w13 = torch.empty(2 * hidden, d, device=device, dtype=dtype).normal_(0, init_std)
        self._w2 = torch.empty(d, hidden, device=device, dtype=dtype).normal_(0, init_std)
        self.attention = Attention(self._wqkv, self._wo, args.n_heads, args.bias, device)
        self.feed_forward = FeedForward(self._w13, self._w2)
        self.attention_norm = RMSNorm(d, args.norm_eps, device, dtype)
        self.ffn_norm = RMSNorm(d, args.norm_eps, device, dtype)


class _Embedding(nn.Module):
    def __init__(self, n: int, dim: int, device, dtype, requires_grad: bool, std: float = 1.0):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(n, dim, device=device, dtype=dtype).normal_(0, std), requires_grad=requires_grad)


class _TrainableLinear(nn.Module):
    def __init__(self, in_f: int, out_f: int, device):
        super().__init__()
        bound = 1.0 / math.sqrt(in_f)                      # nn.Linear default init
        self.weight = nn.Parameter(torch.empty(out_f, in_f, device=device).uniform_(-bound, bound))


class _FrozenAffine(nn.Module):
    def __init__(self, dim: int, device):
        super().__init__()
        bound = 1.0 / math.sqrt(dim)                        # nn.Linear default init
        self.weight = nn.Parameter(torch.empty(dim, dim, device=device).uniform_(-bound, bound), requires_grad=False)
        self.bias = nn.Parameter(torch.empty(dim, device=device).uniform_(-bound, bound), requires_grad=False)


class _CrossAttention(nn.Module):
    """Parameter container of `CrossAttentionModule` (`model.py:145-150`): query / key / value Linear(768, 768) with
    bias, fp32 (`.float()`, `:227`), frozen by the substring rule. The math runs in `Transformer._fuse_inputs`."""

    def __init__(self, dim: int, device):
        super().__init__()
        self.query, self.key, self.value = _FrozenAffine(dim, device), _FrozenAffine(dim, device), _FrozenAffine(dim, device)


class _StepFn(torch.autograd.Function):
    """The whole training step as ONE autograd node: forward = fused kernels (+ saved activations),
    backward = hand-written dX-only pass. Upstream gradients (e.g. GradScaler's scale,
    `util/misc.py:260`) are honoured: every loss-gradient kernel multiplies by grad_output."""

    @staticmethod
    def forward(ctx, model, plan, n_run, *trainables):
        adapter_w, visual_w, temporal_w = trainables[:3]
        gate1 = [g.view(-1) for g in trainables[3:3 + n_run]]
        gate2 = [g.view(-1) for g in trainables[3 + n_run:3 + 2 * n_run]]
        akv_pre = model._take_adapter_kv()
        losses, sv = model._engine.forward(plan, model._run_weights, model.tok_embeddings.weight, model.output.weight,
                                           model.norm.weight, adapter_w, visual_w, temporal_w, gate1, gate2, save=True, akv_pre=akv_pre)
        ctx.model, ctx.sv, ctx.n_run = model, sv, n_run
        ctx.gates = (gate1, gate2)
        ctx.streams = list(plan.streams)
        return tuple(losses[k] for k in plan.streams)

    @staticmethod
    def backward(ctx, *grad_out):
        model, sv, n_run = ctx.model, ctx.sv, ctx.n_run
        dev = model._engine.device
        gscale = torch.zeros(3, dtype=torch.float32, device=dev)
        for k, g in zip(ctx.streams, grad_out):
            if g is not None:
                gscale[{"vqa": 0, "vaq": 1, "qav": 2}[k]] = g.float()
        gb = model._grad_buffers
        sync = model.grad_sync
        model._engine.backward(sv, gscale, model._run_weights, model._output_t, model.norm.weight, ctx.gates[0], ctx.gates[1], gb,
                               on_layer_done=(sync.layer_done if sync is not None else None), out_w=model.output.weight.data)
        if sync is not None:
            sync.finish()
        flat = gb.flat.clone()                              # autograd may keep what we return: never alias the work buffer
        view = lambda k, shape: flat[gb.offsets[k]:gb.offsets[k] + gb.sizes[k]].view(shape)
        H = model.params.n_heads
        out = [None, None, None, view("adapter", gb.adapter.shape), view("visual", gb.visual.shape) if gb.sizes["visual"] else None,
               view("temporal", gb.temporal.shape)]
        g1, g2 = view("gate1", gb.gate1.shape), view("gate2", gb.gate2.shape)
        out += [g1[l].view(1, H, 1, 1) for l in range(n_run)]
        out += [g2[l].view(1, H, 1, 1) for l in range(n_run)]
        ctx.sv = None
        return tuple(out)


class Transformer(nn.Module):
    def __init__(self, params: ModelArgs, args, tokenizer=None, device=None, init_std: float = 0.02):
        super().__init__()
        params.max_feats = args.max_feats                  # `model.py:193-194`
        params.bias = args.bias
        self.args = args
        self.params = params
        self.vocab_size = params.vocab_size
        self.n_layers = params.n_layers
        self.max_feats = args.max_feats
        # input-fusion variant (`model.py:209-227,306-322`): None | 'audio_only' | 'concat' | 'sum' | 'attention'
        self.audio_mode = audio_mode(args)
        self.tokenizer = tokenizer if tokenizer is not None else Tokenizer(model_path=f"{args.llama_model_path}./tokenizer.model", args=args)
        self.eos_id = self.tokenizer.eos_id
        self.answer_token_id = self.tokenizer.a_token_id
        self.q_token_id = self.tokenizer.q_token_id
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
        self._device = torch.device(device)
        dt = H16
        d = params.dim
        self.hidden_dim = ffn_hidden_dim(d, params.multiple_of)
        self.feature_dim = 768                              # CLIP ViT-L/14 frame features
        # input width of the TRAINABLE projection: video, [video | audio] (`model.py:213`) or none (`model.py:209-210`)
        self.video_dim = {None: 768, "audio_only": 0, "concat": 768 + AUDIO_DIM, "sum": 768, "attention": 768}[self.audio_mode]

        self.tok_embeddings = _Embedding(params.vocab_size, d, device, dt, requires_grad=False, std=init_std)
        self.adapter_query = _Embedding(params.adapter_len * params.adapter_layer, d, device, torch.float32, requires_grad=True)
        if self.audio_mode != "audio_only":
            self.visual_proj = _TrainableLinear(self.video_dim, d, device)
        # frozen extras of the audio variants (the freeze rule `llama_vqa.py:72` does not match their names)
        if self.audio_mode in ("audio_only", "sum"):
            self.audio_proj = _FrozenLinear(torch.empty(d, AUDIO_DIM, device=device, dtype=dt).normal_(0, 1.0 / math.sqrt(AUDIO_DIM)))
        elif self.audio_mode == "attention":
            self.audio_proj = _FrozenLinear(torch.empty(self.feature_dim, AUDIO_DIM, device=device, dtype=dt).normal_(0, 1.0 / math.sqrt(AUDIO_DIM)))
            self.video_audio_cross_attn = _CrossAttention(self.feature_dim, device)
        self._audio_f32 = {}                                # fp32 compute copies of the frozen audio-side weights (repack)
        self.temporal_emb = _Embedding(self.max_feats, d, device, torch.float32, requires_grad=True)
        self.adapter_len = params.adapter_len
        self.adapter_layer = params.adapter_layer
        self.layers = nn.ModuleList([TransformerBlock(i, params, self.hidden_dim, device, dt, init_std) for i in range(params.n_layers)])
        self.norm = RMSNorm(d, params.norm_eps, device, dt)
        self.output = _FrozenLinear(torch.empty(params.vocab_size, d, device=device, dtype=dt).normal_(0, init_std))
        self.tau = args.tau

        self.grad_sync = None                               # set by flipped_vqa_b200.dp.DataParallel
        self._engine: Optional[StepEngine] = None
        self._run_weights: Optional[List[LayerWeights]] = None
        self._output_t = None
        self._grad_buffers: Optional[GradBuffers] = None
        self._pack_token = None
        self.last_plan: Optional[BatchPlan] = None
        self._pinned: Optional[PinnedPool] = None
        self.share_option_prefix = True                     # validation: shared-prefix option scoring (step.OptionPlan)
        # frozen weights held ONCE: the dX-only backward reads W [out, in] itself as an MN-major tcgen05 operand (ops.gemm_nn) instead of
        # a load-time transposed copy (-13.5 GB at 7B, -26 GB at 13B, half the load time). False = the round-1 layout (A/B, tests).
        self.weights_once = os.environ.get("FVQA_WEIGHTS_ONCE", "1") != "0"
        self._akv_pre = None                                # adapter K|V enqueued ahead of host planning (forward(data))
        # False (default) = the reference: video_start of SAMPLE 0 places the video span and the gate2 bias block of every sample
        # (`model.py:264`); True = every sample's own value (the kernels take one video_start per sequence)
        self.per_sample_video_start = False

    # ------------------------------------------------------------------ weight layout
    def run_layers(self):
        """Only the last `adapter_layer` layers run (`model.py:338`)."""
        return list(self.layers[-1 * self.adapter_layer:])

    def _token(self):
        ps = [self.layers[0].attention.wq.weight, self.layers[-1].feed_forward.w2.weight, self.output.weight, self.tok_embeddings.weight]
        return tuple((p.data_ptr(), p._version, p.dtype, str(p.device)) for p in ps)

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._pack_token = None                             # .to()/.cuda()/.half() replace parameter storage
        dev = self.tok_embeddings.weight.device             # a model built on the CPU and moved with .cuda() runs on that device
        if dev.type == "cuda":
            self._device = dev
        return out

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        res = super().load_state_dict(state_dict, strict=strict, **kw)
        self._pack_token = None
        return res

    def repack(self):
        """(Re)build the packed / transposed frozen-weight layout from the named parameters and apply
        the dtype contract of `llama_vqa.py:71-76` on the device (frozen h16 = the library's 16-bit operand format, trainables fp32)."""
        dev = self._device
        if dev.type != "cuda":
            raise RuntimeError("flipped_vqa_b200 runs only on a CUDA sm_100a device (no CPU path)")
        d, hid = self.params.dim, self.hidden_dim
        for name, p in self.named_parameters():
            trainable = any(s in name for s in ("gate", "adapter", "temporal_emb", "visual_proj"))
            want = torch.float32 if (trainable or name.startswith("video_audio_cross_attn")) else H16   # `.float()`, `model.py:227`
            if p.dtype != want or p.device != dev:
                p.data = p.data.to(device=dev, dtype=want)
        for blk in self.layers:
            at, ff = blk.attention, blk.feed_forward
            if not (at.wq.weight.data_ptr() == blk._wqkv.data_ptr() and blk._wqkv.device == dev and blk._wqkv.dtype == H16
                    and at.wk.weight.data_ptr() == blk._wqkv[d:].data_ptr() and at.wv.weight.data_ptr() == blk._wqkv[2 * d:].data_ptr()):
                blk._wqkv = torch.cat([at.wq.weight.data, at.wk.weight.data, at.wv.weight.data], 0).contiguous()
                at.wq.weight.data, at.wk.weight.data, at.wv.weight.data = blk._wqkv[0:d], blk._wqkv[d:2 * d], blk._wqkv[2 * d:]
            if not (ff.w1.weight.data_ptr() == blk._w13.data_ptr() and blk._w13.device == dev and blk._w13.dtype == H16
                    and ff.w3.weight.data_ptr() == blk._w13[hid:].data_ptr()):
                blk._w13 = torch.cat([ff.w1.weight.data, ff.w3.weight.data], 0).contiguous()
                ff.w1.weight.data, ff.w3.weight.data = blk._w13[0:hid], blk._w13[hid:]
            blk._wo = at.wo.weight.data
            blk._w2 = ff.w2.weight.data
        run = []
        for blk in self.run_layers():
            lw = LayerWeights(wqkv=blk._wqkv, wo=blk._wo, w13=blk._w13, w2=blk._w2,
                              attn_norm=blk.attention_norm.weight.data, ffn_norm=blk.ffn_norm.weight.data)
            if self.weights_once:          # dX GEMMs read the forward weights as MN-major operands; only [Wk; Wv]^T is copied (adapter gradient)
                lw.wkv_t = blk._wqkv[d:].t().contiguous()
            else:                          # round-1 layout: a transposed copy of every frozen weight (2 x 13.5 GB at 7B)
                lw.wqkv_t, lw.wo_t = blk._wqkv.t().contiguous(), blk._wo.t().contiguous()
                lw.w13_t, lw.w2_t = blk._w13.t().contiguous(), blk._w2.t().contiguous()
            run.append(lw)
        self._run_weights = run
        self._output_t = None if self.weights_once else self.output.weight.data.t().contiguous()
        self._audio_f32 = {"audio_proj": self.audio_proj.weight.data.float().contiguous()} if hasattr(self, "audio_proj") else {}
        if self._engine is None:
            self._engine = StepEngine(d, self.params.n_heads, hid, self.params.vocab_size, self.adapter_len, self.max_feats,
                                      self.params.norm_eps, self.tau, self.params.max_seq_len, dev)
        n_run = len(run)
        self._grad_buffers = GradBuffers(n_run, self.adapter_len, d, self.params.n_heads, self.video_dim, self.max_feats, dev)
        self._pack_token = self._token()

    def _ensure_packed(self):
        if self._pack_token is None or self._pack_token != self._token():
            self.repack()

    def trainable_parameters(self):
        n_run = len(self.run_layers())
        # audio only: no trainable projection (`model.py:209-210`); the frozen audio_proj takes its place in the kernels
        visual = self.visual_proj.weight if self.audio_mode != "audio_only" else self._audio_f32["audio_proj"]
        ps = [self.adapter_query.weight, visual, self.temporal_emb.weight]
        ps += [blk.attention.gate1 for blk in self.run_layers()]
        ps += [blk.attention.gate2 for blk in self.run_layers()]
        return ps, n_run

    # ------------------------------------------------------------------ forward
    def streams(self):
        return ["vqa"] + (["vaq"] if self.args.vaq else []) + (["qav"] if self.args.qav else [])

    def plan_batch(self, data, inference: bool = False) -> BatchPlan:
        """Host side of a step: flatten the batch dict (`dataloader/__init__.py:28-90`) into the int32
        arrays the kernels read and start the (single) async H2D copy from pinned memory."""
        if self._pinned is None:
            self._pinned = PinnedPool()
        streams = ["vqa"] if inference else self.streams()
        data, post = self._fuse_inputs(data)
        compact = self._engine is not None and self._engine.skip_pad_rows        # padding-free row maps only when they are used
        return post(BatchPlan(data, streams, self.max_feats, inference=inference, pool=self._pinned, compact=compact,
                              per_sample_video_start=self.per_sample_video_start).to_device(self._device))

    def _fuse_inputs(self, data):
        """The input-fusion branches of `model.py:306-322` reduced to what the step kernels see: a feature matrix
        `plan.video` [B*F, video_dim] for the (trainable) projection and an optional frozen additive term `plan.vf_extra`
        [B*F, d]. Returns (batch dict whose 'video' is that feature matrix where the host can build it, fix-up run on the
        device-resident plan)."""
        mode = self.audio_mode
        if mode is None:
            return data, (lambda plan: plan)
        self._ensure_packed()
        dev, F = self._device, self.max_feats
        audio = data["audio"].float()                       # `.cuda().half()` in the reference (`model.py:258-261`); kept fp32 here
        B, Fa = audio.shape[0], audio.shape[1]
        d2 = dict(data)
        if mode == "audio_only":                            # `_video_feature = audio_proj(audio)`, `:307`
            d2["video"] = audio
            return d2, (lambda plan: plan)
        if mode == "concat":                                # `visual_proj(cat([video, audio]))`, `:310-311`
            d2["video"] = torch.cat([data["video"].float(), audio], dim=-1)
            return d2, (lambda plan: plan)
        if mode == "sum" and Fa != F:                       # `:314` adds [B, Fa, d] to [B, F, d]: torch broadcasting needs Fa in {1, F}
            if Fa != 1:
                raise ValueError(f"audio_merge='sum': audio has {Fa} frames per sample, video features have {F} (need {F} or 1)")
            audio, Fa = audio.expand(B, F, audio.shape[-1]), F
        audio_dev = audio.reshape(B * Fa, AUDIO_DIM).to(dev, non_blocking=True)
        if mode == "sum":                                   # `audio_proj(audio) + visual_proj(video)`, `:314`

            def post(plan):
                plan.vf_extra = ops.linear_f32(audio_dev, self._audio_f32["audio_proj"])
                return plan
            return d2, post

        def post(plan):                                     # 'attention', `:317-320` + CrossAttentionModule `:153-169`
            ca = self.video_audio_cross_attn
            af = ops.linear_f32(audio_dev, self._audio_f32["audio_proj"])                        # [B*Fa, 768]
            q = ops.linear_f32(plan.video, ca.query.weight.data, bias=ca.query.bias.data)
            k = ops.linear_f32(af, ca.key.weight.data, bias=ca.key.bias.data)
            v = ops.linear_f32(af, ca.value.weight.data, bias=ca.value.bias.data)
            plan.video = ops.cross_attn_fwd(q, k, v, B, F, Fa)                                   # input of visual_proj
            return plan
        return d2, post

    def forward(self, data, inference: bool = False):
        if inference:
            # HEAD `llama/model.py:251-252,367` generates and matches; the loss-based scorer is `model_my_original_mod.py` (`engine.py:78-93`)
            return self.generate_answers(data) if getattr(self.args, "is_generation_task", False) else self.inference(data)
        self._ensure_packed()
        self._prefetch_adapter_kv()                         # GPU work that does not need the batch, enqueued before host planning
        return self.forward_plan(self.plan_batch(data))

    def _prefetch_adapter_kv(self):
        w = self.adapter_query.weight
        self._akv_pre = ((w.data_ptr(), w._version), self._engine.adapter_kv(self._run_weights, w.data))

    def _take_adapter_kv(self):
        """The prefetched adapter K|V if it was computed from the current adapter weights, else None (computed in-loop)."""
        pre, self._akv_pre = self._akv_pre, None
        w = self.adapter_query.weight
        return pre[1] if pre is not None and pre[0] == (w.data_ptr(), w._version) else None

    def forward_plan(self, plan: BatchPlan):
        """The device side of the training step for an already device-resident batch plan."""
        self._ensure_packed()
        dev = self._device
        self.last_plan = plan
        trainables, n_run = self.trainable_parameters()
        if torch.is_grad_enabled() and any(p.requires_grad for p in trainables):
            outs = _StepFn.apply(self, plan, n_run, *trainables)
            losses = dict(zip(plan.streams, outs))
        else:
            g1 = [p.data.view(-1) for p in trainables[3:3 + n_run]]
            g2 = [p.data.view(-1) for p in trainables[3 + n_run:]]
            akv_pre = self._take_adapter_kv()
            losses, _ = self._engine.forward(plan, self._run_weights, self.tok_embeddings.weight.data, self.output.weight.data,
                                             self.norm.weight.data, trainables[0].data, trainables[1].data, trainables[2].data,
                                             g1, g2, save=False, akv_pre=akv_pre)
        # disabled objectives return `torch.tensor([0]).cuda()` (`model.py:302`). Built on the device, and only when needed: a
        # torch.tensor(list, device=cuda) is a pageable-memory H2D copy, i.e. a host sync on everything the step has enqueued so far
        zero = lambda: torch.zeros(1, dtype=torch.int64, device=dev)
        return losses["vqa"], (losses["vaq"] if "vaq" in losses else zero()), (losses["qav"] if "qav" in losses else zero())

    def plan_options(self, data) -> OptionPlan:
        """Host side of shared-prefix option scoring (`step.OptionPlan`) + its async H2D copy."""
        if self._pinned is None:
            self._pinned = PinnedPool()
        data, post = self._fuse_inputs(data)
        return post(OptionPlan(data, self.max_feats, pool=self._pinned,
                               per_sample_video_start=self.per_sample_video_start).to_device(self._device))

    @torch.no_grad()
    def inference(self, data):
        """Loss-based option scoring: VQA stream only over bsz*n_options sequences
        (`model_my_original_mod.py:281,332-333,348-360,375-377,506`). With `share_option_prefix` (default) the
        option-invariant prefix of each sample is evaluated once (`step.OptionPlan`); the per-token losses are the same."""
        self._ensure_packed()
        self._prefetch_adapter_kv()
        return self.inference_plan(self.plan_options(data) if self.share_option_prefix else self.plan_batch(data, inference=True))

    @torch.no_grad()
    def inference_plan(self, plan):
        """Device side of option scoring for an already device-resident `OptionPlan` (or dense inference `BatchPlan`)."""
        self._ensure_packed()
        self.last_plan = plan
        trainables, n_run = self.trainable_parameters()
        g1 = [p.data.view(-1) for p in trainables[3:3 + n_run]]
        g2 = [p.data.view(-1) for p in trainables[3 + n_run:]]
        w = (self._run_weights, self.tok_embeddings.weight.data, self.output.weight.data, self.norm.weight.data,
             trainables[0].data, trainables[1].data, trainables[2].data, g1, g2)
        akv_pre = self._take_adapter_kv()
        if isinstance(plan, OptionPlan):
            return self._engine.forward_options(plan, *w, akv_pre=akv_pre)
        tok, _ = self._engine.forward(plan, *w, save=False, token_losses=True, akv_pre=akv_pre)
        return tok

    # ------------------------------------------------------------------ generation evaluator (`llama/model.py:367-623`)
    GENERATION_STEPS = 31                                   # `range(prefix - 1, prefix + 30)`, `model.py:433`
    QUESTION_MARKER_ID = 894                                # hard-coded in `model.py:520`

    @torch.no_grad()
    def generate_answers(self, data, want_margin: bool = False):
        """`Transformer.inference` of HEAD (`llama/model.py:367-546`): greedy-decode 31 tokens from `prefix_index - 1` on option
        0's sequence of every sample, embed the generated answer (token-embedding mean up to the first EOS, restricted to the
        positions of option 0's answer span) and every option's answer, and pick the option with the highest cosine similarity.
        Returns (most_similar_indices [bsz] int64, extracted_answers: list of {'video_id', 'question', 'generated_answer'}), as
        `engine.py:78-85,99-121` consumes them. Decoding is KV-cached (`StepEngine.generate`): every position is evaluated
        once instead of 31 x bsz full-stack re-runs."""
        self._ensure_packed()
        self._prefetch_adapter_kv()
        ids_all = data["text_id"]["vqa"]
        bsz, n_options, S = ids_all.shape
        d0 = dict(data)
        d0["text_id"] = {"vqa": ids_all[:, 0:1].contiguous()}                 # `vqa_id = vqa_id[:, 0:1, :]`, `model.py:385-387`
        d0["label"] = {"vqa": data["label"]["vqa"][:, 0:1].contiguous()}
        plan = self.plan_batch(d0, inference=True)
        self.last_plan = plan
        trainables, n_run = self.trainable_parameters()
        g1 = [p.data.view(-1) for p in trainables[3:3 + n_run]]
        g2 = [p.data.view(-1) for p in trainables[3 + n_run:]]
        prefix = [int(p) for p in data["prefix_index"]["vqa"]]
        tokens, margin = self._engine.generate(plan, self._run_weights, self.tok_embeddings.weight.data, self.output.weight.data,
                                               self.norm.weight.data, trainables[0].data, trainables[1].data, trainables[2].data, g1, g2,
                                               prefix, n_steps=self.GENERATION_STEPS, akv_pre=self._take_adapter_kv(), want_margin=want_margin)
        vqa_id = plan.ids.view(bsz, S).long().cpu()                           # option 0's sequences with the generated tokens written in
        self.last_generation = dict(tokens=tokens, margin=margin, ids=vqa_id)
        # ---- cosine matching (`model.py:478-512,548-623`), token-embedding means in fp32 on the device
        emb = self.tok_embeddings.weight.data
        label0 = data["label"]["vqa"][:, 0, 1:]                               # `vqa_label[:, 1:]` of option 0
        placeholder = label0 != 0                                             # positions of the answer span ([bsz, S-1])
        choice_emb = []
        for b in range(bsz):                                                  # extract_answers + embed_and_aggregate_answers
            row0 = ids_all[b, 0].tolist()
            start = row0.index(self.answer_token_id) + 5
            answers = []
            for o in range(n_options):
                tail = ids_all[b, o, start:].tolist()
                end = start + tail.index(self.eos_id) if self.eos_id in tail else S
                answers.append(ids_all[b, o, start:end])
            padded = torch.nn.utils.rnn.pad_sequence(answers, batch_first=True, padding_value=0).to(emb.device)
            choice_emb.append(emb[padded].float().mean(dim=1))               # pad id 0 is embedded too, as in the reference
        choice_emb = torch.stack(choice_emb)                                  # [bsz, n_options, d]
        out_emb = []
        for b in range(bsz):                                                  # filter_and_process_output_tokens + aggregate
            toks = vqa_id[b, 1:][placeholder[b]]
            eos = (toks == self.eos_id).nonzero(as_tuple=True)[0]
            if eos.numel() > 0:
                toks = toks[:eos[0]]
            out_emb.append(emb[toks.to(emb.device)].float().mean(dim=0) if toks.numel() > 0 else torch.zeros(emb.shape[1], device=emb.device))
        out_emb = torch.stack(out_emb)
        sims = torch.bmm(torch.nn.functional.normalize(choice_emb, p=2, dim=2),
                         torch.nn.functional.normalize(out_emb, p=2, dim=1).unsqueeze(-1)).squeeze(-1)          # find_most_similar
        most_similar = sims.argmax(dim=1)
        self.last_generation["similarities"] = sims
        extracted = []
        for b in range(bsz):                                                  # `model.py:514-544`
            row = vqa_id[b].tolist()
            q_start = row.index(self.QUESTION_MARKER_ID) + 2
            q_end = row.index(self.answer_token_id)
            a_tokens = row[q_end + 5:]
            try:
                a_end = a_tokens.index(self.eos_id)
            except ValueError:
                a_end = next((i for i, t in enumerate(a_tokens) if t == 0), len(a_tokens))
            extracted.append({"video_id": data["vid"][b], "question": self.tokenizer.decode(row[q_start:q_end]),
                              "generated_answer": self.tokenizer.decode(a_tokens[:a_end])})
        return most_similar, extracted

    @staticmethod
    def predict_options(token_losses: torch.Tensor) -> torch.Tensor:
        """`engine.py:88-93` as one kernel: argmin over options of sum / count(loss != 0)."""
        from .. import ops
        pred, _ = ops.option_score(token_losses.contiguous())
        return pred.long()
