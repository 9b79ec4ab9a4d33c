"""Host-side orchestration of one LLaMA-VQA step on the CUDA kernels (no torch math on the hot path).

`BatchPlan`   – turns the reference's batch dict (`dataloader/__init__.py:28-90`) into flat int32
                arrays: concatenated objective streams, labelled-row lists for the fused heads.
`PackedWeights` – the frozen base in its B200 layout: Wq|Wk|Wv and W1|W3 concatenated so each is one
                GEMM, plus load-time TRANSPOSED copies so the dX-only backward is the same K-major
                tcgen05 GEMM (HBM is plentiful: 2x13.5 GB for 7B of 180 GB).
`StepEngine`  – workspace + forward / backward kernel sequences (`llama/model.py:286-361` forward;
                backward derived in SURVEY.md §8(a) addendum: dX only through the frozen base,
                weight gradients only for adapter prompts, gates, visual_proj, temporal_emb).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch

from . import ops

H16 = ops.H16
I32 = torch.int32


# ------------------------------------------------------------------------------------------------
# batch -> flat device arrays
# ------------------------------------------------------------------------------------------------
class BatchPlan:
    """Host-side (CPU) preparation of one batch. Everything the kernels need is packed into ONE
    pinned int32 buffer (one H2D copy) plus the fp32 video features."""

    def __init__(self, data: Dict, streams: List[str], max_feats: int, inference: bool = False, pool: "PinnedPool" = None,
                 compact: bool = True, per_sample_video_start: bool = False):
        F = max_feats
        ids = {k: _cpu(data["text_id"][k]) for k in streams}
        lab = {k: _cpu(data["label"][k]) for k in streams}
        B, n_opt, S = ids[streams[0]].shape
        self.B, self.n_opt, self.S, self.F = B, n_opt, S, F
        self.streams = streams
        per = B * n_opt                       # sequences per stream
        self.n_seq = per * len(streams)
        self.T = self.n_seq * S
        self.n_video = B
        ids_all = torch.cat([ids[k].reshape(per, S) for k in streams], 0).to(I32)
        labels_all = torch.zeros(self.n_seq, S, dtype=I32)
        vstart = torch.empty(self.n_seq, dtype=I32)
        seq_video = (torch.arange(per) // n_opt).repeat(len(streams)).to(I32)
        # the reference reads video_start of sample 0 only (`llama/model.py:264`); the kernels take one value per sequence, so
        # `per_sample_video_start` (opt-in) gives every sample its own video span / gate2 column block
        for si, k in enumerate(streams):
            if k == "qav":
                vstart[si * per:(si + 1) * per] = -1
            elif per_sample_video_start:
                vs_b = torch.as_tensor([int(v) for v in data["video_start"][k]], dtype=I32)
                assert vs_b.numel() == B, f"video_start[{k}] has {vs_b.numel()} entries for a batch of {B}"
                vstart[si * per:(si + 1) * per] = vs_b.repeat_interleave(n_opt)
            else:
                vstart[si * per:(si + 1) * per] = int(data["video_start"][k][0])
        qav_index = torch.zeros(B, F, dtype=I32)
        self.ce_counts: Dict[str, int] = {}
        ce_rows, ce_tgt, ce_dst = [], [], []
        self.q_count = 0
        q_rows = q_tgt = q_vid = torch.zeros(0, dtype=I32)
        for si, k in enumerate(streams):
            l2 = lab[k].reshape(per, S)
            if k == "qav":
                labels_all[si * per:(si + 1) * per] = l2.to(I32)
                qav_index = _cpu(data["video_index"]["qav"]).reshape(B, F).to(I32)
                nz = (l2[:, 1:] >= 0).nonzero()                       # ignore_index=-1 (`model.py:235`)
                q_rows = ((si * per + nz[:, 0]) * S + nz[:, 1]).to(I32)
                q_tgt = l2[:, 1:][nz[:, 0], nz[:, 1]].to(I32)
                q_vid = (nz[:, 0] // n_opt).to(I32)
                self.q_count = int(nz.shape[0])
            else:
                nz = (l2[:, 1:] != 0).nonzero()                       # ignore_index=0 (`model.py:233-234`)
                ce_rows.append(((si * per + nz[:, 0]) * S + nz[:, 1]).to(I32))
                ce_tgt.append(l2[:, 1:][nz[:, 0], nz[:, 1]].to(I32))
                ce_dst.append((nz[:, 0] * (S - 1) + nz[:, 1]).to(I32))   # index into [B, n_opt, S-1]
                self.ce_counts[k] = int(nz.shape[0])
        self.ce_total = sum(self.ce_counts.values())
        ce_rows = torch.cat(ce_rows) if ce_rows else torch.zeros(0, dtype=I32)
        ce_tgt = torch.cat(ce_tgt) if ce_tgt else torch.zeros(0, dtype=I32)
        ce_dst = torch.cat(ce_dst) if ce_dst else torch.zeros(0, dtype=I32)
        # rows whose final hidden state any loss reads: the LAST layer's wo / FFN only run on these (sorted, unique);
        # ce_rows_c / q_rows_c index into that compact row set
        live = torch.unique(torch.cat([ce_rows.long(), q_rows.long()]), sorted=True)
        self.n_live = int(live.numel())
        ce_rows_c = torch.searchsorted(live, ce_rows.long()).to(I32)
        q_rows_c = torch.searchsorted(live, q_rows.long()).to(I32)
        parts = [ids_all.flatten(), labels_all.flatten(), vstart, seq_video, qav_index.flatten(),
                 ce_rows, ce_tgt, ce_dst, q_rows, q_tgt, q_vid, live.to(I32), ce_rows_c, q_rows_c]
        # padding-free row set (only built when the engine will use it: `StepEngine.skip_pad_rows`): rows after the last
        # loss-relevant position of a sequence cannot reach any loss through the causal mask (`model.py:298-299`), so the
        # row-wise ops only need rows [0, end_n) of sequence n ("k" = kept rows).
        self.T_c = None
        if compact:
            loss_rows = torch.cat([ce_rows.long(), q_rows.long()])
            end = torch.zeros(self.n_seq, dtype=torch.long)
            if loss_rows.numel():
                end.scatter_reduce_(0, loss_rows // S, loss_rows % S + 1, reduce="amax")
            keep = (torch.arange(S).view(1, S) < end.view(-1, 1)).reshape(-1)
            c2f = keep.nonzero().flatten()
            self.T_c = int(c2f.numel())
            f2c = torch.full((self.T,), -1, dtype=torch.long)
            f2c[c2f] = torch.arange(self.T_c)
            ce_rows_k, q_rows_k = f2c[ce_rows.long()], f2c[q_rows.long()]
            live_k = torch.unique(torch.cat([ce_rows_k, q_rows_k]), sorted=True)
            self.n_live_k = int(live_k.numel())
            parts += [c2f % S, c2f, f2c, ce_rows_k, q_rows_k, live_k, torch.searchsorted(live_k, ce_rows_k), torch.searchsorted(live_k, q_rows_k)]
        sizes = [p.numel() for p in parts]
        total = max(sum(sizes), 1)
        video = _cpu(data["video"]).reshape(B * F, -1).float()
        slot = pool.acquire(total, video.numel()) if pool is not None else None
        if slot is not None:
            host, hvideo = slot.ints[:total], slot.video[:video.numel()].view(video.shape)
        else:
            host, hvideo = torch.empty(total, dtype=I32), torch.empty(video.shape, dtype=torch.float32)
        off = 0
        self._slices = []
        for p, n in zip(parts, sizes):
            host[off:off + n] = p
            self._slices.append((off, n))
            off += n
        hvideo.copy_(video)
        self.host_ints, self.host_video, self._slot = host, hvideo, slot
        self.h2d_bytes = host.numel() * 4 + video.numel() * 4

    def to_device(self, device):
        dev = self.host_ints.to(device, non_blocking=True)
        names = ["ids", "labels", "vstart", "seq_video", "qav_index", "ce_rows", "ce_tgt", "ce_dst", "q_rows", "q_tgt", "q_vid",
                 "live_rows", "ce_rows_c", "q_rows_c",
                 "pos_ids", "c2f", "f2c", "ce_rows_k", "q_rows_k", "live_rows_k", "ce_rows_kc", "q_rows_kc"]
        for nm, (off, n) in zip(names, self._slices):
            setattr(self, nm, dev[off:off + n])
        self.video = self.host_video.to(device, non_blocking=True)
        self._dev = dev
        if self._slot is not None:
            self._slot.event.record()          # the pinned slot may be rewritten once these copies are done
        return self

    def record_stream(self, stream):
        """The plan was copied on another stream (dataloader.PlannedLoader): tell the allocator who uses it."""
        self._dev.record_stream(stream)
        self.video.record_stream(stream)
        extra = getattr(self, "vf_extra", None)           # audio 'sum' / 'attention': computed on the copy stream by Transformer._fuse_inputs
        if extra is not None:
            extra.record_stream(stream)


class OptionPlan:
    """Host-side plan of loss-based option scoring (`llama/model_my_original_mod.py:332-377`, `engine.py:87-93`) that
    evaluates the OPTION-INVARIANT PREFIX of every sample once.

    The n_opt sequences of a sample are identical up to the answer span, and under the causal mask the hidden state of
    a position depends only on the tokens at and before it, so rows [0, P_b) of all options of sample b are the same
    numbers; rows after the last labelled position E_b of a sample feed no loss at all. The layer GEMMs, norms and the
    FFN therefore run on a COMPACT, ragged row set

        [ prefix rows of sample 0 | ... | prefix rows of sample B-1 | suffix rows (b, o, P_b <= pos < E_b) ... ]

    (T_c = sum_b P_b + n_opt * sum_b (E_b - P_b) instead of B * n_opt * S rows); only attention sees the full
    [B * n_opt, S] layout, rebuilt per layer by a row gather of q|k|v. Results equal the dense evaluation's.
      pos_ids[T_c]      position of each compact row (RoPE angle)
      c2f[T_c]          full row (sequence * S + pos) that compact row r stands for (option 0 for prefix rows)
      f2c[n_seq * S]    compact row holding full row's values, -1 for rows past E_b (never read: causal)
    """

    def __init__(self, data: Dict, max_feats: int, pool: "PinnedPool" = None, per_sample_video_start: bool = False):
        ids = _cpu(data["text_id"]["vqa"]).long()
        lab = _cpu(data["label"]["vqa"]).long()
        B, n_opt, S = ids.shape
        F = max_feats
        self.B, self.n_opt, self.S, self.F = B, n_opt, S, F
        self.streams = ["vqa"]
        self.n_seq = B * n_opt
        self.T = self.n_seq * S
        self.n_video = B
        neq = (ids != ids[:, :1]).any(1)                                     # [B, S] some option differs here
        cp = torch.where(neq.any(1), neq.int().argmax(1), torch.full((B,), S))  # common-prefix length per sample
        labelled = lab[:, :, 1:] != 0                                         # row t predicts token t+1 (`model.py:350`)
        any_row = labelled.any(1)                                             # [B, S-1]
        last = (S - 2) - any_row.flip(1).int().argmax(1)
        E = torch.where(any_row.any(1), last + 1, torch.zeros(B, dtype=torch.long))   # rows needed: [0, E_b)
        P = torch.minimum(cp, E)
        self.prefix_len, self.end = P, E
        pos = torch.arange(S).view(1, 1, S).expand(B, n_opt, S)
        opt = torch.arange(n_opt).view(1, n_opt, 1).expand(B, n_opt, S)
        Pb, Eb = P.view(B, 1, 1), E.view(B, 1, 1)
        m_pre = ((opt == 0) & (pos < Pb)).reshape(-1)
        m_suf = ((pos >= Pb) & (pos < Eb)).reshape(-1)
        c2f = torch.cat([m_pre.nonzero().flatten(), m_suf.nonzero().flatten()])
        self.n_prefix_rows = int(m_pre.sum())
        self.T_c = int(c2f.numel())
        pre_off = torch.cumsum(P, 0) - P                                      # first compact row of each sample's prefix
        f2c = torch.full((self.T,), -1, dtype=torch.long)
        in_pre = (pos < Pb).reshape(-1)
        f2c[in_pre] = (pre_off.view(B, 1, 1) + pos).reshape(-1)[in_pre]
        f2c[m_suf] = self.n_prefix_rows + torch.arange(int(m_suf.sum()))
        pos_ids = pos.reshape(-1)[c2f]
        nz = labelled.reshape(self.n_seq, S - 1).nonzero()
        ce_full = nz[:, 0] * S + nz[:, 1]
        ce_rows = f2c[ce_full]
        assert bool((ce_rows >= 0).all())
        ce_tgt = lab.reshape(self.n_seq, S)[:, 1:][nz[:, 0], nz[:, 1]]
        ce_dst = nz[:, 0] * (S - 1) + nz[:, 1]
        self.ce_total = int(nz.shape[0])
        self.ce_counts = {"vqa": self.ce_total}
        live = torch.unique(ce_rows, sorted=True)                             # last layer: wo / FFN on these only
        self.n_live = int(live.numel())
        ce_rows_c = torch.searchsorted(live, ce_rows)
        if per_sample_video_start:                                            # opt-in, see BatchPlan
            vstart = torch.as_tensor([int(v) for v in data["video_start"]["vqa"]], dtype=I32).repeat_interleave(n_opt)
            assert vstart.numel() == self.n_seq
        else:
            vs = int(data["video_start"]["vqa"][0])                           # sample 0's, `model_my_original_mod.py:264`
            vstart = torch.full((self.n_seq,), vs, dtype=I32)
        seq_video = (torch.arange(self.n_seq) // n_opt).to(I32)
        parts = [ids.reshape(-1), torch.zeros(self.T, dtype=I32), vstart, seq_video, torch.zeros(B * F, dtype=I32),
                 pos_ids, c2f, f2c, ce_rows, ce_tgt, ce_dst, live, ce_rows_c]
        self._names = ["ids", "labels", "vstart", "seq_video", "qav_index", "pos_ids", "c2f", "f2c", "ce_rows", "ce_tgt", "ce_dst",
                       "live_rows", "ce_rows_c"]
        sizes = [p.numel() for p in parts]
        total = max(sum(sizes), 1)
        video = _cpu(data["video"]).reshape(B * F, -1).float()
        slot = pool.acquire(total, video.numel()) if pool is not None else None
        if slot is not None:
            host, hvideo = slot.ints[:total], slot.video[:video.numel()].view(video.shape)
        else:
            host, hvideo = torch.empty(total, dtype=I32), torch.empty(video.shape, dtype=torch.float32)
        off = 0
        self._slices = []
        for p, n in zip(parts, sizes):
            host[off:off + n] = p
            self._slices.append((off, n))
            off += n
        hvideo.copy_(video)
        self.host_ints, self.host_video, self._slot = host, hvideo, slot
        self.h2d_bytes = host.numel() * 4 + video.numel() * 4

    def host(self, name: str) -> torch.Tensor:
        off, n = self._slices[self._names.index(name)]
        return self.host_ints[off:off + n]

    def to_device(self, device):
        dev = self.host_ints.to(device, non_blocking=True)
        for nm, (off, n) in zip(self._names, self._slices):
            setattr(self, nm, dev[off:off + n])
        self.video = self.host_video.to(device, non_blocking=True)
        self._dev = dev
        if self._slot is not None:
            self._slot.event.record()
        return self

    def record_stream(self, stream):
        self._dev.record_stream(stream)
        self.video.record_stream(stream)
        extra = getattr(self, "vf_extra", None)
        if extra is not None:
            extra.record_stream(stream)


class PinnedPool:
    """Small ring of pinned host staging buffers for the per-step H2D copy (ids/labels/row lists in
    one int32 buffer + the fp32 video features). A slot is reused only after the CUDA event recorded
    behind its last async copy has completed."""

    class _Slot:
        def __init__(self, n_int, n_f32):
            self.ints = torch.empty(n_int, dtype=I32, pin_memory=True)
            self.video = torch.empty(n_f32, dtype=torch.float32, pin_memory=True)
            self.event = torch.cuda.Event()
            self.used = False

    def __init__(self, slots: int = 4):
        self.n, self.slots, self.i = slots, [], 0

    def acquire(self, n_int: int, n_f32: int):
        if not torch.cuda.is_available():
            return None
        if len(self.slots) < self.n:
            self.slots.append(self._Slot(max(n_int, 1) * 2, max(n_f32, 1)))
        s = self.slots[self.i % len(self.slots)]
        self.i += 1
        if s.ints.numel() < n_int or s.video.numel() < n_f32:
            s.event.synchronize() if s.used else None
            s.__init__(max(n_int, s.ints.numel()) * 2, max(n_f32, s.video.numel()))
        elif s.used:
            s.event.synchronize()
        s.used = True
        return s


def _cpu(t):
    return t.detach().cpu() if isinstance(t, torch.Tensor) else torch.as_tensor(t)


# ------------------------------------------------------------------------------------------------
# frozen weights in their B200 layout
# ------------------------------------------------------------------------------------------------
@dataclass
class LayerWeights:
    """Frozen weights of one layer. `*_t` are load-time transposed copies for the K-major dX GEMMs; with `weights_once` (default) only
    `wkv_t` exists - the [Wk; Wv] block the skinny adapter-gradient GEMM streams - and the dX GEMMs read `wqkv` / `wo` / `w13` / `w2`
    themselves as MN-major operands (`ops.gemm_nn`)."""
    wqkv: torch.Tensor      # [3d, d]   rows: Wq | Wk | Wv
    wo: torch.Tensor        # [d, d]
    w13: torch.Tensor       # [2*hid, d] rows: W1 | W3
    w2: torch.Tensor        # [d, hid]
    attn_norm: torch.Tensor
    ffn_norm: torch.Tensor
    wkv_t: Optional[torch.Tensor] = None      # [d, 2d] = [Wk; Wv]^T (weights_once)
    wqkv_t: Optional[torch.Tensor] = None     # [d, 3d]
    wo_t: Optional[torch.Tensor] = None
    w13_t: Optional[torch.Tensor] = None      # [d, 2*hid]
    w2_t: Optional[torch.Tensor] = None       # [hid, d]

    def kv_t(self, d: int) -> torch.Tensor:
        return self.wkv_t if self.wkv_t is not None else self.wqkv_t[:, d:]


# ------------------------------------------------------------------------------------------------
# the step
# ------------------------------------------------------------------------------------------------
@dataclass
class SavedStep:
    plan: BatchPlan
    x: List[torch.Tensor] = field(default_factory=list)      # layer inputs  (L+1 entries; last = final hidden)
    qkv: List[torch.Tensor] = field(default_factory=list)
    akv: List[torch.Tensor] = field(default_factory=list)
    o: List[torch.Tensor] = field(default_factory=list)
    lse: List[torch.Tensor] = field(default_factory=list)
    h: List[torch.Tensor] = field(default_factory=list)
    g: List[torch.Tensor] = field(default_factory=list)
    rstd1: List[torch.Tensor] = field(default_factory=list)
    rstd2: List[torch.Tensor] = field(default_factory=list)
    vf32: Optional[torch.Tensor] = None
    pruned: bool = False        # last layer ran on plan.live_rows only (h / g / rstd2 / final x of that layer are compact)
    compact: bool = False       # row-wise tensors (x, h, g, rstd) hold the padding-free rows plan.c2f; qkv / o / lse are full
    ce: Optional[dict] = None
    qav: Optional[dict] = None


class StepEngine:
    """Runs the kernels. Stateless w.r.t. parameters: callers pass the packed frozen weights and the
    current trainable tensors each step."""

    def __init__(self, dim, n_heads, hidden, vocab, adapter_len, max_feats, norm_eps, tau, max_seq_len, device):
        self.d, self.H, self.hid, self.V = dim, n_heads, hidden, vocab
        self.hd = dim // n_heads
        self.A, self.F, self.eps, self.tau = adapter_len, max_feats, norm_eps, tau
        self.device = device
        # RoPE table `llama/model.py:45-50,245`: resident on the device (the reference re-uploads it every step)
        inv = 1.0 / (10000.0 ** (torch.arange(0, self.hd, 2)[: self.hd // 2].float() / self.hd))
        ang = torch.outer(torch.arange(max_seq_len * 2).float(), inv)
        self.cos = torch.cos(ang).to(device).contiguous()
        self.sin = torch.sin(ang).to(device).contiguous()
        self._attn_ws = None
        self._tables = None
        self.adapter_grad_chunk = 8  # layers per grouped d-adapter GEMM = layers per early all-reduce message (dp.GradSync)
        self.sample_layers = ()      # layers whose GEMM launches bench.py's GemmTimer samples
        # wo / FFN of the last layer only on the rows the losses read: -2.1 % FLOPs, -1.8 % step time in a same-box A/B
        # (tools/ab_step.py); results identical (tests/test_model_gpu.py::test_last_layer_live_row_pruning_is_equivalent)
        self.prune_last_layer = True
        # padding-free rows (BatchPlan.c2f), opt-in: the row-wise ops skip the rows after each sequence's last loss-relevant
        # position. Same losses and gradients (tests/test_model_gpu.py::test_padding_free_rows_are_equivalent); costs four row
        # copies per layer around attention, so it is only used when it drops at least `skip_pad_min_saving` of the rows.
        # Off by default: the default step computes the reference's row set (every position of every sequence).
        self.skip_pad_rows = False
        self.skip_pad_min_saving = 0.06
        # fp16 operands: backward runs on upstream gradients normalised by ONE power of two so that the 16-bit gradient tensors sit
        # mid-range for any caller-side loss scale / accum_iter (fvqa_grad_scale_prepare); the trainable gradients are multiplied by
        # its exact inverse when they are final. 2^14: dlogits <= 2^14 / n_labelled, d hidden ~ O(1) (fp16 normal range 6e-5 .. 65504).
        # bf16 operands have fp32's exponent range: no scaling (target 0).
        self.grad_scale_target = 2.0 ** 14 if ops.H16 == torch.float16 else 0.0
        self._gen_ws = None          # persistent workspace (+ captured decode graph) of StepEngine.generate
        self.decode_graph = True     # generation evaluator: replay the single-row decode step as a CUDA graph (StepEngine.generate)

    # -------------------------------------------------------------------------------- batch-independent prologue
    def adapter_kv(self, layers: List[LayerWeights], adapter_w):
        """Adapter K|V of every layer, `wk/wv(adapter)` (`model.py:99-100`): the only part of the step that does not depend
        on the batch. `Transformer.forward(data)` enqueues it BEFORE it builds the batch plan on the host, so the GPU has
        ~1 ms of work while the host flattens ids / labels and starts the H2D copy (instead of idling after the previous
        step's loss read). Same kernels and values as the in-loop computation."""
        A, d, L = self.A, self.d, len(layers)
        adapter_h16 = ops.f32_to_h16(adapter_w)                                  # `adapter[i].half()`, `model.py:339`
        tables = self._weight_tables(layers)
        if tables is None:                                                          # small / odd dims: one skinny GEMM per layer
            return [ops.gemm_nt(adapter_h16[l * A:(l + 1) * A], w.wqkv[d:]) for l, w in enumerate(layers)]
        akv_all = torch.empty(L, A, 2 * d, dtype=H16, device=self.device)
        ops.gemm_skinny_grouped(adapter_h16.view(L, A, d), tables[0], d, 2 * d, akv_all)   # all layers, one launch (2.1 GB of Wk|Wv at 7B)
        return [akv_all[l] for l in range(L)]

    def _weight_tables(self, layers: List[LayerWeights]):
        """Device pointer tables of the per-layer [Wk; Wv] blocks and their transposes for the grouped skinny GEMM
        (None when the shape is outside its range). Cached per packed-weight list; the weights are frozen."""
        if self.A > 16 or self.d % 256 != 0 or not layers:
            return None
        d = self.d
        key = (id(layers), layers[0].wqkv.data_ptr(), layers[-1].kv_t(d).data_ptr())
        if self._tables is None or self._tables[0] != key:
            kv = torch.tensor([w.wqkv[d:].data_ptr() for w in layers], dtype=torch.int64, device=self.device)
            kv_t = torch.tensor([w.kv_t(d).data_ptr() for w in layers], dtype=torch.int64, device=self.device)
            self._tables = (key, (kv, kv_t, layers[0].kv_t(d).stride(0)))
        return self._tables[1]

    # -------------------------------------------------------------------------------- forward
    def forward(self, plan: BatchPlan, layers: List[LayerWeights], tok_emb, out_w, norm_w, adapter_w, visual_w, temporal_w,
                gate1: List[torch.Tensor], gate2: List[torch.Tensor], save: bool, token_losses: bool = False, akv_pre=None):
        """Returns (losses dict of 0-dim fp32 device tensors | per-token loss tensor, SavedStep|None)."""
        d, H, hd, hid, A, F, S = self.d, self.H, self.hd, self.hid, self.A, self.F, plan.S
        T, n_seq, dev = plan.T, plan.n_seq, self.device
        sv = SavedStep(plan) if save else None
        L = len(layers)
        # --- inputs (`model.py:286-336`)
        # [B*F, d] fp32: visual_proj(features) (+ the frozen audio term of the 'sum' variant, `model.py:306-322`)
        vf32 = ops.linear_f32(plan.video, visual_w, add=getattr(plan, "vf_extra", None))
        x = ops.build_h0_fwd(tok_emb, plan.ids, plan.labels, plan.vstart, plan.seq_video, plan.qav_index, vf32, temporal_w,
                             n_seq, S, F)
        # Row set of the row-wise ops: all T rows, or the padding-free compact rows (attention always sees all T rows)
        compact = self.skip_pad_rows and plan.T_c is not None and 0 < plan.T_c <= (1.0 - self.skip_pad_min_saving) * T
        R = plan.T_c if compact else T
        if compact:
            x = ops.gather_rows(x, plan.c2f)
            ce_full, q_full, live_rows, ce_live, q_live, n_live = (plan.ce_rows_k, plan.q_rows_k, plan.live_rows_k, plan.ce_rows_kc,
                                                                   plan.q_rows_kc, plan.n_live_k)
        else:
            ce_full, q_full, live_rows, ce_live, q_live, n_live = (plan.ce_rows, plan.q_rows, plan.live_rows, plan.ce_rows_c,
                                                                   plan.q_rows_c, plan.n_live)
        akv_all = akv_pre if akv_pre is not None else self.adapter_kv(layers, adapter_w)   # adapter K|V, no RoPE (`model.py:99-100`)
        xn = torch.empty(R, d, dtype=H16, device=dev)
        c = torch.empty(R, hid, dtype=H16, device=dev)
        qkv_b = o_b = g_b = None
        qkv_c = torch.empty(R, 3 * d, dtype=H16, device=dev) if compact else None
        o_c = torch.empty(R, d, dtype=H16, device=dev) if compact else None
        # The last layer's wo / FFN outputs are read only at the rows the losses use (labelled positions): run them on
        # those rows alone, exactly like the vocabulary projection (heads below). Everything upstream needs all rows (K/V).
        prune = self.prune_last_layer and 0 < n_live < R
        ce_idx, q_idx = (ce_live, q_live) if prune else (ce_full, q_full)
        for l, w in enumerate(layers):
            if ops.GEMM_TIMER is not None:
                ops.GEMM_TIMER.active, ops.GEMM_TIMER.tag = l in self.sample_layers, l
            _, rstd1 = ops.rmsnorm_fwd(x, w.attn_norm, self.eps, y=xn)
            # Wq|Wk|Wv in one GEMM, RoPE applied to q|k in its epilogue (`model.py:89,96`)
            if compact:
                ops.gemm_nt_rope(xn, w.wqkv, self.cos, self.sin, 2 * d, hd, S, out=qkv_c, pos_ids=plan.pos_ids)
                qkv = ops.expand_rows(qkv_c, plan.f2c, dst=None if save else qkv_b)  # full layout (zero rows past the end)
            else:
                qkv = ops.gemm_nt_rope(xn, w.wqkv, self.cos, self.sin, 2 * d, hd, S, out=None if save else qkv_b)
            akv = akv_all[l]
            o_full, lse = ops.attn_fwd(qkv, akv, self.cos, self.sin, gate1[l], gate2[l], plan.vstart, n_seq, S, H, hd, A, F,
                                       out=None if save else o_b)
            o = ops.gather_rows(o_full, plan.c2f, dst=o_c) if compact else o_full
            if prune and l == L - 1:
                o_g = ops.gather_rows(o, live_rows)
                x_g = ops.gather_rows(x, live_rows)
                h = ops.gemm_nt(o_g, w.wo, residual=x_g, out_fp32=True)            # [n_live, d]
                xn_g, rstd2 = ops.rmsnorm_fwd(h, w.ffn_norm, self.eps)
                g, c_g = ops.gemm_swiglu_fwd(xn_g, w.w13)
                x_next = ops.gemm_nt(c_g, w.w2, residual=h, out_fp32=True)
            else:
                h = ops.gemm_nt(o, w.wo, residual=x, out_fp32=True)                # h = x + attn  (`model.py:185`), fp32 stream
                _, rstd2 = ops.rmsnorm_fwd(h, w.ffn_norm, self.eps, y=xn)
                g, _ = ops.gemm_swiglu_fwd(xn, w.w13, g=None if save else g_b, c=c)  # W1|W3 GEMM, SwiGLU in its epilogue
                x_next = ops.gemm_nt(c, w.w2, residual=h, out_fp32=True)           # out = h + ffn (`model.py:186`)
            if save:
                sv.x.append(x); sv.qkv.append(qkv); sv.akv.append(akv); sv.o.append(o_full); sv.lse.append(lse)
                sv.h.append(h); sv.g.append(g); sv.rstd1.append(rstd1); sv.rstd2.append(rstd2)
            elif not (prune and l == L - 1):
                qkv_b, o_b, g_b = qkv, o_full, g
            x = x_next
        if save:
            sv.x.append(x)
            sv.vf32 = vf32
            sv.pruned = prune
            sv.compact = compact
        if ops.GEMM_TIMER is not None:
            ops.GEMM_TIMER.active = False
        # --- heads
        losses = {}
        if plan.ce_total > 0:
            hn, rstd = ops.rmsnorm_gather_fwd(x, ce_idx, norm_w, self.eps)          # final norm on labelled rows only
            logits = ops.gemm_nt(hn, out_w, out_fp32=True)                          # [rows, V] fp32, never [B*S, V]
            row_loss, row_lse = ops.ce_fwd(logits, plan.ce_tgt)
            if token_losses:
                tl = torch.zeros(plan.B * plan.n_opt * (S - 1), dtype=torch.float32, device=dev)
                ops.scatter_rows(row_loss, plan.ce_dst, tl)
                return tl.view(plan.B, plan.n_opt, S - 1), None
            off = 0
            for k in plan.streams:
                if k == "qav":
                    continue
                n = plan.ce_counts[k]
                out = torch.empty((), dtype=torch.float32, device=dev)
                if n > 0:
                    ops.sum_scale(row_loss[off:off + n], n, 1.0 / n, out)           # mean over ALL labelled tokens of the batch
                else:
                    out.fill_(float("nan"))
                losses[k] = out
                off += n
            if save:
                sv.ce = dict(hn=hn, rstd=rstd, logits=logits, row_lse=row_lse)
        else:
            if token_losses:
                return torch.zeros(plan.B, plan.n_opt, S - 1, dtype=torch.float32, device=dev), None
            for k in plan.streams:
                if k != "qav":
                    losses[k] = torch.full((), float("nan"), dtype=torch.float32, device=dev)
        if "qav" in plan.streams:
            out = torch.empty((), dtype=torch.float32, device=dev)
            if plan.q_count > 0:
                hnq, rstdq = ops.rmsnorm_gather_fwd(x, q_idx, norm_w, self.eps)
                row_loss, prob = ops.qav_loss_fwd(hnq, vf32, plan.q_vid, plan.q_tgt, self.tau, F)
                ops.sum_scale(row_loss, plan.q_count, 1.0 / plan.q_count, out)
                if save:
                    sv.qav = dict(hn=hnq, rstd=rstdq, prob=prob)
            else:
                out.fill_(float("nan"))
            losses["qav"] = out
        return losses, sv

    # -------------------------------------------------------------------------------- option scoring, shared prefix
    def forward_options(self, plan: "OptionPlan", layers: List[LayerWeights], tok_emb, out_w, norm_w, adapter_w, visual_w, temporal_w,
                        gate1: List[torch.Tensor], gate2: List[torch.Tensor], akv_pre=None):
        """Per-token VQA losses [B, n_opt, S-1] (`model_my_original_mod.py:375-377`) with the option-invariant prefix of
        each sample evaluated once: every row-wise op (norms, the 7 frozen GEMMs per layer, SwiGLU, residuals) runs on
        `plan`'s compact ragged rows; attention runs on the full [B * n_opt, S] layout rebuilt by a q|k|v row gather."""
        d, H, hd, hid, A, F, S = self.d, self.H, self.hd, self.hid, self.A, self.F, plan.S
        n_seq, dev, Tc = plan.n_seq, self.device, plan.T_c
        tl = torch.zeros(plan.B * plan.n_opt * (S - 1), dtype=torch.float32, device=dev)
        if plan.ce_total == 0 or Tc == 0:
            return tl.view(plan.B, plan.n_opt, S - 1)
        L = len(layers)
        vf32 = ops.linear_f32(plan.video, visual_w, add=getattr(plan, "vf_extra", None))
        x_full = ops.build_h0_fwd(tok_emb, plan.ids, plan.labels, plan.vstart, plan.seq_video, plan.qav_index, vf32, temporal_w,
                                  n_seq, S, F)
        x = ops.gather_rows(x_full, plan.c2f)                                       # [Tc, d] fp32 residual stream, compact
        del x_full
        akv_all = akv_pre if akv_pre is not None else self.adapter_kv(layers, adapter_w)
        xn = torch.empty(Tc, d, dtype=H16, device=dev)
        qkv_c = torch.empty(Tc, 3 * d, dtype=H16, device=dev)
        qkv_f = torch.empty(plan.T, 3 * d, dtype=H16, device=dev)
        o_f = torch.empty(plan.T, d, dtype=H16, device=dev)
        lse = torch.empty(n_seq, H, S, dtype=torch.float32, device=dev)
        o_c = torch.empty(Tc, d, dtype=H16, device=dev)
        g = torch.empty(Tc, 2 * hid, dtype=H16, device=dev)
        c = torch.empty(Tc, hid, dtype=H16, device=dev)
        prune = self.prune_last_layer and 0 < plan.n_live < Tc
        ce_idx = plan.ce_rows_c if prune else plan.ce_rows
        for l, w in enumerate(layers):
            ops.rmsnorm_fwd(x, w.attn_norm, self.eps, y=xn)
            ops.gemm_nt_rope(xn, w.wqkv, self.cos, self.sin, 2 * d, hd, S, out=qkv_c, pos_ids=plan.pos_ids)
            ops.expand_rows(qkv_c, plan.f2c, dst=qkv_f)                             # compact -> every option's sequence, zero rows past E_b
            akv = akv_all[l]
            ops.attn_fwd(qkv_f, akv, self.cos, self.sin, gate1[l], gate2[l], plan.vstart, n_seq, S, H, hd, A, F, out=o_f, lse=lse)
            if prune and l == L - 1:
                o_g = ops.gather_rows(ops.gather_rows(o_f, plan.c2f, dst=o_c), plan.live_rows)
                x_g = ops.gather_rows(x, plan.live_rows)
                h = ops.gemm_nt(o_g, w.wo, residual=x_g, out_fp32=True)
                xn_g, _ = ops.rmsnorm_fwd(h, w.ffn_norm, self.eps)
                _, c_g = ops.gemm_swiglu_fwd(xn_g, w.w13)
                x = ops.gemm_nt(c_g, w.w2, residual=h, out_fp32=True)
            else:
                ops.gather_rows(o_f, plan.c2f, dst=o_c)
                h = ops.gemm_nt(o_c, w.wo, residual=x, out_fp32=True)
                ops.rmsnorm_fwd(h, w.ffn_norm, self.eps, y=xn)
                ops.gemm_swiglu_fwd(xn, w.w13, g=g, c=c)
                x = ops.gemm_nt(c, w.w2, residual=h, out_fp32=True)
        hn, _ = ops.rmsnorm_gather_fwd(x, ce_idx, norm_w, self.eps)
        logits = ops.gemm_nt(hn, out_w, out_fp32=True)
        row_loss, _ = ops.ce_fwd(logits, plan.ce_tgt)
        ops.scatter_rows(row_loss, plan.ce_dst, tl)
        return tl.view(plan.B, plan.n_opt, S - 1)

    # -------------------------------------------------------------------------------- greedy generation, KV-cached
    def generate(self, plan: BatchPlan, layers: List[LayerWeights], tok_emb, out_w, norm_w, adapter_w, visual_w, temporal_w,
                 gate1: List[torch.Tensor], gate2: List[torch.Tensor], prefix_index: List[int], n_steps: int = 31, akv_pre=None,
                 want_margin: bool = False):
        """The generation evaluator's decoding loop (`llama/model.py:429-467`): for every sample, `n_steps` greedy steps starting at
        position prefix - 1; step t reads the logits of position p = prefix - 1 + t and writes the arg-max token to position p + 1.

        The reference re-runs the WHOLE layer stack over the whole sequence for every step of every sample (31 x bsz stack
        evaluations). Under the causal mask the hidden state of position p only depends on tokens <= p, and those never change
        once written, so this evaluates every position ONCE: one prefill over all sequences (the training forward's kernels; the
        per-layer q|k|v buffers are kept as the K/V cache) and n_steps - 1 single-row decode steps for all samples together
        (M = bsz rows through the GEMMs; the row's k|v is scattered into the cache and the fused attention kernel - adapter
        branch, gate2 video bias, causal mask - runs over the cached layout).

        `plan`: a dense inference plan of option 0's sequences (one per sample). Returns (tokens [bsz, n_steps] int32,
        top-2 logit margins [bsz, n_steps] fp32 or None); plan.ids holds the sequences with the generated tokens written in."""
        d, H, hd, hid, A, F, S = self.d, self.H, self.hd, self.hid, self.A, self.F, plan.S
        B, T, dev = plan.n_seq, plan.T, self.device
        assert len(prefix_index) == B
        if min(prefix_index) < 1 or max(prefix_index) + n_steps - 1 > S - 1:
            raise IndexError(f"generation needs 1 <= prefix_index and prefix_index + {n_steps - 1} <= {S - 1} (max_seq_len - 1); "
                             f"got prefix_index in [{min(prefix_index)}, {max(prefix_index)}]")
        L = len(layers)
        # positions / cache rows of every step, known up front (decoding never stops early, `model.py:433`)
        pos_host = torch.tensor([[p - 1 + t for p in prefix_index] for t in range(n_steps)], dtype=I32)
        rows_host = pos_host + (torch.arange(B, dtype=I32) * S).view(1, B)
        sched = torch.cat([pos_host.flatten(), rows_host.flatten()]).to(dev, non_blocking=True)
        pos_all, rows_all = sched[: n_steps * B].view(n_steps, B), sched[n_steps * B:].view(n_steps, B)
        tokens = torch.zeros(B, n_steps, dtype=I32, device=dev)
        margin = torch.zeros(B, n_steps, dtype=torch.float32, device=dev) if want_margin else None
        # Persistent workspace per (batch, sequence length, depth): the decode step below is captured ONCE as a CUDA graph and replayed
        # for every later step of every later batch, so every buffer it touches must keep its address across calls - the batch's ids /
        # video_start / adapter K|V are copied in.
        key = (B, S, L, bool(want_margin), tuple(w.wqkv.data_ptr() for w in layers[:2]), tok_emb.data_ptr(), gate1[0].data_ptr())
        ws = self._gen_ws if self._gen_ws is not None and self._gen_ws["key"] == key else None
        if ws is None:
            e = lambda *shape, dt=H16: torch.empty(*shape, dtype=dt, device=dev)
            f32 = torch.float32
            ws = dict(key=key, graph=None, cache=[e(T, 3 * d) for _ in range(L)], o=e(T, d), lse=e(B, H, S, dt=f32), ids=e(T, dt=I32),
                      vstart=e(B, dt=I32), akv=e(L, A, 2 * d), xs=e(B, d, dt=f32), xa=e(B, d, dt=f32), xb=e(B, d, dt=f32), h=e(B, d, dt=f32),
                      xn=e(B, d), qkv=e(B, 3 * d), o_s=e(B, d), g=e(B, 2 * hid), c=e(B, hid), rstd=e(B, dt=f32), logits=e(B, self.V, dt=f32),
                      pos=e(B, dt=I32), rows=e(B, dt=I32), tok=torch.zeros(B, 1, dtype=I32, device=dev),
                      mar=(torch.zeros(B, 1, dtype=f32, device=dev) if want_margin else None))
            self._gen_ws = ws
        cache, o, lse, ids_w, vstart_w, xs, logits = ws["cache"], ws["o"], ws["lse"], ws["ids"], ws["vstart"], ws["xs"], ws["logits"]
        ids_w.copy_(plan.ids)
        vstart_w.copy_(plan.vstart)
        akv_src = akv_pre if akv_pre is not None else self.adapter_kv(layers, adapter_w)
        for l in range(L):
            ws["akv"][l].copy_(akv_src[l][:A])
        akv_all = [ws["akv"][l] for l in range(L)]
        # ---- prefill: every sequence, all positions (option 0's original tokens; positions >= prefix are overwritten later)
        vf32 = ops.linear_f32(plan.video, visual_w, add=getattr(plan, "vf_extra", None))
        x = ops.build_h0_fwd(tok_emb, ids_w, plan.labels, vstart_w, plan.seq_video, plan.qav_index, vf32, temporal_w, B, S, F)
        xn = torch.empty(T, d, dtype=H16, device=dev)
        c = torch.empty(T, hid, dtype=H16, device=dev)
        g = torch.empty(T, 2 * hid, dtype=H16, device=dev)
        for l, w in enumerate(layers):
            ops.rmsnorm_fwd(x, w.attn_norm, self.eps, y=xn)
            ops.gemm_nt_rope(xn, w.wqkv, self.cos, self.sin, 2 * d, hd, S, out=cache[l])
            ops.attn_fwd(cache[l], akv_all[l], self.cos, self.sin, gate1[l], gate2[l], vstart_w, B, S, H, hd, A, F, out=o, lse=lse)
            h = ops.gemm_nt(o, w.wo, residual=x, out_fp32=True)
            ops.rmsnorm_fwd(h, w.ffn_norm, self.eps, y=xn)
            ops.gemm_swiglu_fwd(xn, w.w13, g=g, c=c)
            x = ops.gemm_nt(c, w.w2, residual=h, out_fp32=True)
        hn, _ = ops.rmsnorm_gather_fwd(x, rows_all[0], norm_w, self.eps)      # position prefix - 1 of every sample
        ops.gemm_nt(hn, out_w, out_fp32=True, out=logits)
        ops.greedy_next(logits, tok_emb, ids_w, S, pos_all[0], tokens, 0, xs, margin)
        del x, xn, c, g
        # ---- decode: one row per sample and step. The step is launch-bound (~10 small launches per layer around 13.5 GB of weight
        # streaming): after one eager step it is captured in a CUDA graph (kept in the workspace) and replayed; positions / cache rows
        # of the step are read from fixed device buffers that a copy refreshes between replays.
        xn_s, qkv_s, o_s, g_s, c_s, h_s, xa, xb, rstd_s = (ws[k] for k in ("xn", "qkv", "o_s", "g", "c", "h", "xa", "xb", "rstd"))
        pos_c, rows_c, tok_c, mar_c = ws["pos"], ws["rows"], ws["tok"], ws["mar"]

        def decode_step():
            """x (embedding of the last token) in `xs` -> logits of the current position -> next token / embedding back into `xs`."""
            x_in, x_out = xs, xa
            for l, w in enumerate(layers):
                ops.rmsnorm_fwd(x_in, w.attn_norm, self.eps, y=xn_s, rstd=rstd_s)
                ops.gemm_nt_rope(xn_s, w.wqkv, self.cos, self.sin, 2 * d, hd, S, out=qkv_s, pos_ids=pos_c)
                ops.scatter_row_vectors(qkv_s, rows_c, cache[l])                # the new position's q | k | v into the cached layout
                ops.attn_fwd(cache[l], akv_all[l], self.cos, self.sin, gate1[l], gate2[l], vstart_w, B, S, H, hd, A, F, out=o, lse=lse)
                ops.gather_rows(o, rows_c, dst=o_s)
                ops.gemm_nt(o_s, w.wo, residual=x_in, out_fp32=True, out=h_s)
                ops.rmsnorm_fwd(h_s, w.ffn_norm, self.eps, y=xn_s, rstd=rstd_s)
                ops.gemm_nt(xn_s, w.w13, out=g_s)
                ops.swiglu_fwd(g_s, c=c_s)
                ops.gemm_nt(c_s, w.w2, residual=h_s, out_fp32=True, out=x_out)
                x_in, x_out = x_out, (xb if x_out is xa else xa)
            ops.rmsnorm_fwd(x_in, norm_w, self.eps, y=xn_s, rstd=rstd_s)
            ops.gemm_nt(xn_s, out_w, out_fp32=True, out=logits)
            ops.greedy_next(logits, tok_emb, ids_w, S, pos_c, tok_c, 0, xs, mar_c)

        for t in range(1, n_steps):
            pos_c.copy_(pos_all[t])
            rows_c.copy_(rows_all[t])
            if self.decode_graph and ws["graph"] is not None:
                ws["graph"].replay()
            else:
                decode_step()
                if self.decode_graph and ws["graph"] is None and n_steps > 2:
                    try:
                        torch.cuda.synchronize()
                        graph = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(graph):
                            decode_step()
                        ws["graph"] = graph
                    except Exception:                                           # capture unsupported here: keep launching eagerly
                        self.decode_graph = False
            tokens[:, t].copy_(tok_c[:, 0])
            if want_margin:
                margin[:, t].copy_(mar_c[:, 0])
        plan.ids.copy_(ids_w)                                                   # the sequences with the generated tokens written in
        return tokens, margin

    # -------------------------------------------------------------------------------- backward
    def backward(self, sv: SavedStep, gscale: torch.Tensor, layers: List[LayerWeights], out_w_t, norm_w, gate1, gate2,
                 grads: "GradBuffers", on_layer_done=None, out_w=None):
        """gscale: fp32 device tensor [3] = upstream gradients of (vqa, vaq, qav) losses.
        Fills `grads` (fp32): adapter [L*A, d], gate1/gate2 [L, H], visual [d, vdim], temporal [F, d]."""
        plan = sv.plan
        d, H, hd, hid, A, F, S = self.d, self.H, self.hd, self.hid, self.A, self.F, plan.S
        T, n_seq, dev = plan.T, plan.n_seq, self.device
        L = len(layers)
        x_final = sv.x[L]
        pruned, compact = sv.pruned, sv.compact
        if compact:
            ce_full, q_full, live_rows, ce_live, q_live, n_live = (plan.ce_rows_k, plan.q_rows_k, plan.live_rows_k, plan.ce_rows_kc,
                                                                   plan.q_rows_kc, plan.n_live_k)
        else:
            ce_full, q_full, live_rows, ce_live, q_live, n_live = (plan.ce_rows, plan.q_rows, plan.live_rows, plan.ce_rows_c,
                                                                   plan.q_rows_c, plan.n_live)
        Tr = plan.T_c if compact else T                       # rows of the row-wise tensors
        R = n_live if pruned else Tr                          # rows of the final hidden state that exist
        ce_idx, q_idx = (ce_live, q_live) if pruned else (ce_full, q_full)
        # gradient of the residual stream: fp32 master + h16 copy (A operand of the next dX GEMM)
        dx = torch.zeros(R, d, dtype=torch.float32, device=dev)
        dx_h = torch.zeros(R, d, dtype=H16, device=dev)
        gidx = {"vqa": 0, "vaq": 1, "qav": 2}
        # dX = dY . W: from the forward weight itself (MN-major operand) when no transposed copy exists (`weights_once`)
        def dgemm(dy, w_fwd, w_t, out=None):
            return ops.gemm_nt(dy, w_t, out=out) if w_t is not None else ops.gemm_nn(dy, w_fwd, out=out)

        def dswiglu(dy, w, gsaved, dg=None):
            return ops.gemm_swiglu_bwd(dy, w.w2_t, gsaved, dg=dg) if w.w2_t is not None else ops.gemm_swiglu_bwd(dy, w.w2, gsaved, dg=dg, nn=True)
        inv_k = None
        if self.grad_scale_target > 0:
            gscale, inv_k = ops.grad_scale_prepare(gscale, self.grad_scale_target)
        unscale = (lambda t: ops.scale_f32(t, inv_k)) if inv_k is not None else (lambda t: None)
        # --- heads backward
        if sv.ce is not None:
            ce = sv.ce
            dlogits = torch.empty(plan.ce_total, self.V, dtype=H16, device=dev)
            off = 0
            for k in plan.streams:
                if k == "qav":
                    continue
                n = plan.ce_counts[k]
                if n > 0:
                    ops.ce_bwd(ce["logits"][off:off + n], plan.ce_tgt[off:off + n], ce["row_lse"][off:off + n],
                               gscale[gidx[k]:gidx[k] + 1], 1.0 / n, dlogits=dlogits[off:off + n])
                off += n
            dhn = dgemm(dlogits, out_w, out_w_t)                                    # dH = dlogits . W_out
            ops.rmsnorm_scatter_bwd(dhn, x_final, ce_idx, norm_w, ce["rstd"], dx, dx_h)
        dvf_qav = None
        if sv.qav is not None:
            q = sv.qav
            dhnq, dvf_qav = ops.qav_loss_bwd(q["hn"], sv.vf32, plan.q_vid, plan.q_tgt, q["prob"], gscale[2:3],
                                             1.0 / plan.q_count, self.tau, plan.n_video, F)
            ops.rmsnorm_scatter_bwd(dhnq, x_final, q_idx, norm_w, q["rstd"], dx, dx_h)
        # --- layers, last to first
        if self._attn_ws is None or self._attn_ws.numel() < ops.attn_bwd_ws_bytes(n_seq, S, H, hd, A):
            self._attn_ws = torch.empty(ops.attn_bwd_ws_bytes(n_seq, S, H, hd, A), dtype=torch.uint8, device=dev)
        dg = torch.empty(Tr, 2 * hid, dtype=H16, device=dev)
        dtmp = torch.empty(Tr, d, dtype=H16, device=dev)
        dh = torch.empty(Tr, d, dtype=torch.float32, device=dev)
        dh_h = torch.empty(Tr, d, dtype=H16, device=dev)
        dqkv = torch.empty(T, 3 * d, dtype=H16, device=dev)
        dakv = torch.empty(A, 2 * d, dtype=torch.float32, device=dev)
        dakv_h = torch.empty(A, 2 * d, dtype=H16, device=dev)
        tables, chunk = self._weight_tables(layers), max(1, self.adapter_grad_chunk)
        dakv_h_all = torch.empty(L, A, 2 * d, dtype=H16, device=dev) if tables is not None else None
        dx_next = torch.empty(Tr, d, dtype=torch.float32, device=dev)
        dx_next_h = torch.empty(Tr, d, dtype=H16, device=dev)
        do_full = torch.empty(T, d, dtype=H16, device=dev) if compact else None
        dqkv_c = torch.empty(Tr, 3 * d, dtype=H16, device=dev) if compact else None
        for l in range(L - 1, -1, -1):
            w = layers[l]
            if ops.GEMM_TIMER is not None:
                ops.GEMM_TIMER.active, ops.GEMM_TIMER.tag = l in self.sample_layers, l
            if pruned and l == L - 1:
                # compact rows through the FFN and wo of the last layer, then scatter d(attn out) and dh back to all rows
                dg_c = dswiglu(dx_h, w, sv.g[l])
                dtmp_c = dgemm(dg_c, w.w13, w.w13_t)
                dh_c, dh_c_h = ops.rmsnorm_bwd(dtmp_c, sv.h[l], w.ffn_norm, sv.rstd2[l], dres=dx,
                                                dx_h16=torch.empty(R, d, dtype=H16, device=dev))
                do_c = dgemm(dh_c_h, w.wo, w.wo_t)
                dtmp.zero_()
                ops.scatter_row_vectors(do_c, live_rows, dtmp)
                dh.zero_()
                ops.scatter_row_vectors(dh_c, live_rows, dh)
                dx = torch.empty(Tr, d, dtype=torch.float32, device=dev)          # full-size stream from here down
                dx_h = torch.empty(Tr, d, dtype=H16, device=dev)
            else:
                dswiglu(dx_h, w, sv.g[l], dg=dg)                                   # d[a|b] = swiglu'(g) . (dout . W2), dc stays on chip
                dgemm(dg, w.w13, w.w13_t, out=dtmp)                                # d(ffn_norm out) = [da|db] . [W1;W3]
                ops.rmsnorm_bwd(dtmp, sv.h[l], w.ffn_norm, sv.rstd2[l], dres=dx, dx=dh, dx_h16=dh_h)
                dgemm(dh_h, w.wo, w.wo_t, out=dtmp)                               # d(attn out) = dh . Wo
            # attention sees the full layout: d(attn out) of the rows that were never computed is zero
            dout = ops.expand_rows(dtmp, plan.f2c, dst=do_full) if compact else dtmp
            ops.attn_bwd(sv.qkv[l], sv.akv[l], self.cos, self.sin, gate1[l], gate2[l], plan.vstart, sv.o[l], sv.lse[l], dout,
                         n_seq, S, H, hd, A, F, dqkv=dqkv, dakv=dakv, dgate1=grads.gate1[l], dgate2=grads.gate2[l], ws=self._attn_ws)
            dqkv_r = ops.gather_rows(dqkv, plan.c2f, dst=dqkv_c) if compact else dqkv
            dgemm(dqkv_r, w.wqkv, w.wqkv_t, out=dtmp)                              # d(attn_norm out) = dqkv . [Wq;Wk;Wv]
            ops.rmsnorm_bwd(dtmp, sv.x[l], w.attn_norm, sv.rstd1[l], dres=dh, dx=dx_next, dx_h16=dx_next_h)
            # d adapter_l = dK_a . Wk + dV_a . Wv   (fp32 out, straight into the gradient rows of layer l): one grouped launch
            # per chunk of layers (their rows then become final together, which is what dp.GradSync reduces early)
            if tables is None:
                ops.f32_to_h16(dakv, dakv_h)
                ops.gemm_nt(dakv_h, w.kv_t(d), out=grads.adapter[l * A:(l + 1) * A], out_fp32=True)
                unscale(grads.adapter[l * A:(l + 1) * A])
            else:
                ops.f32_to_h16(dakv, dakv_h_all[l])
                if l % chunk == 0:
                    hi = min(l + chunk, L)
                    ops.gemm_skinny_grouped(dakv_h_all[l:hi], tables[1][l:hi], tables[2], d, grads.adapter.view(L, A, d)[l:hi])
                    unscale(grads.adapter.view(L, A, d)[l:hi])                       # final rows of this chunk, before dp.GradSync reduces them
            dx, dx_next = dx_next, dx
            dx_h, dx_next_h = dx_next_h, dx_h
            if on_layer_done is not None:
                on_layer_done(l)
        if ops.GEMM_TIMER is not None:
            ops.GEMM_TIMER.active = False
        # --- input side (`model.py:322-336` backward)
        if compact:
            dx = ops.expand_rows(dx, plan.f2c)                                     # [T, d] fp32, zero at the skipped rows
        dvf, _ = ops.video_grad(dx, plan.vstart, plan.seq_video, plan.qav_index, dvf_qav, n_seq, plan.n_video, S, F, dtemporal=grads.temporal)
        if grads.sizes["visual"]:                              # audio only: the projection is frozen (`model.py:209-210`)
            ops.visual_proj_bwd(dvf, plan.video, dwv=grads.visual)
        unscale(grads.flat[grads.late_offset:])                # gates | visual_proj | temporal_emb: one contiguous tail
        return grads


class GradBuffers:
    """One flat fp32 buffer for every trainable gradient, laid out so that the part that is final
    early in backward (adapter prompts, last layer first) is contiguous and the late part
    (gates, visual_proj, temporal_emb) is one contiguous tail -> few, large all-reduce messages."""

    def __init__(self, n_layers_run: int, A: int, d: int, H: int, vdim: int, F: int, device):
        self.sizes = dict(adapter=n_layers_run * A * d, gate1=n_layers_run * H, gate2=n_layers_run * H, visual=d * vdim, temporal=F * d)
        total = sum(self.sizes.values())
        self.flat = torch.zeros(total, dtype=torch.float32, device=device)
        off = 0
        self.offsets = {}
        for k, n in self.sizes.items():
            self.offsets[k] = off
            off += n
        v = lambda k: self.flat[self.offsets[k]:self.offsets[k] + self.sizes[k]]
        self.adapter = v("adapter").view(n_layers_run * A, d)
        self.gate1 = v("gate1").view(n_layers_run, H)
        self.gate2 = v("gate2").view(n_layers_run, H)
        self.visual = v("visual").view(d, vdim)
        self.temporal = v("temporal").view(F, d)
        self.late_offset = self.offsets["gate1"]
