"""Tensor-level wrappers around the C ABI (include/fvqa.h). PyTorch only supplies device memory
and the current stream; every computation happens in libfvqa.so. Outputs may be passed in
(pre-allocated workspace) or are allocated with torch.empty."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import check as _check_rc, ptr, stream

H16 = _lib.H16

# kernels launched through the C ABI since import (bench.py reports the count inside its timed region)
LAUNCHES = 0


class GemmTimer:
    """Optional CUDA-event timing of individual GEMM launches on the launching stream (bench.py's
    roofline leg). `active` is toggled by the caller to sample a subset of launches."""

    def __init__(self):
        self.active = False
        self.tag = None            # set by the caller (step.py: the layer index) and stored with every record
        self.records = []          # (start_event, end_event, flops, tag)

    def summary(self, tags=None):
        recs = [r for r in self.records if tags is None or r[3] in tags]
        if not recs:
            return None
        ms = [a.elapsed_time(b) for a, b, _, _ in recs]
        fl = [f for _, _, f, _ in recs]
        return dict(launches=len(ms), total_ms=sum(ms), total_flops=sum(fl), avg_ms=sum(ms) / len(ms),
                    avg_flops=sum(fl) / len(fl), tflops=sum(fl) / (sum(ms) * 1e-3) / 1e12)


GEMM_TIMER: Optional[GemmTimer] = None
FUSE_SWIGLU = True      # A/B switch (tools/ab_step.py): False = plain GEMM + stand-alone SwiGLU kernels


class OpTimer:
    """Optional CUDA-event timing of EVERY op wrapper below (tools/step_breakdown.py): per-op in-step device time
    including the launch gap before it. Off (None) on the product path."""

    def __init__(self):
        self.records = []          # (name, start_event, end_event)

    def summary(self):
        agg = {}
        for name, a, b in self.records:
            t = agg.setdefault(name, [0, 0.0])
            t[0] += 1
            t[1] += a.elapsed_time(b)
        return agg


OP_TIMER: Optional[OpTimer] = None


def _timed(fn):
    import functools

    @functools.wraps(fn)
    def wrapper(*a, **k):
        tm = OP_TIMER
        if tm is None:
            return fn(*a, **k)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn(*a, **k)
        e1.record()
        name = fn.__name__
        if name == "gemm_nt":
            M = k.get("M") or a[0].shape[0]
            name = f"gemm_nt[{M}x{a[1].shape[0]}x{a[0].shape[1]}]"
        tm.records.append((name, e0, e1))
        return out
    return wrapper


def _count(n: int = 1):
    global LAUNCHES
    LAUNCHES += n


_KERNELS_PER_CALL = {"attn_bwd": 3, "attn_bwd_tc": 2, "qav_loss_bwd": 2}


def check(rc: int, what: str):
    _check_rc(rc, what)
    _count(_KERNELS_PER_CALL.get(what, 1))


def _chk(t: torch.Tensor, dtype, name: str):
    assert t.is_cuda and t.dtype == dtype and t.is_contiguous(), f"{name}: need contiguous cuda {dtype}, got {t.dtype} {t.device} contiguous={t.is_contiguous()}"


# ------------------------------------------------------------------ RMSNorm (fp32 residual stream in, h16 GEMM operand out)
F32 = torch.float32


@_timed
def rmsnorm_fwd(x, w, eps: float, y=None, rstd=None):
    _chk(x, F32, "x"); _chk(w, H16, "w")
    rows, dim = x.shape
    y = torch.empty(rows, dim, dtype=H16, device=x.device) if y is None else y
    rstd = torch.empty(rows, dtype=torch.float32, device=x.device) if rstd is None else rstd
    check(_lib.lib().fvqa_rmsnorm_fwd(ptr(x), ptr(w), ptr(y), ptr(rstd), rows, dim, eps, stream()), "rmsnorm_fwd")
    return y, rstd


@_timed
def rmsnorm_bwd(dy, x, w, rstd, dres=None, dx=None, dx_h16=None):
    """Returns (dx fp32, dx_h16). dx = dres + rmsnorm'(x).dy"""
    _chk(dy, H16, "dy"); _chk(x, F32, "x")
    rows, dim = x.shape
    dx = torch.empty(rows, dim, dtype=F32, device=x.device) if dx is None else dx
    check(_lib.lib().fvqa_rmsnorm_bwd(ptr(dy), ptr(x), ptr(w), ptr(rstd), ptr(dres), ptr(dx), ptr(dx_h16), rows, dim, stream()), "rmsnorm_bwd")
    return dx, dx_h16


@_timed
def rmsnorm_gather_fwd(x, idx, w, eps: float, y=None, rstd=None):
    _chk(x, F32, "x"); _chk(idx, torch.int32, "idx")
    rows, dim = idx.numel(), x.shape[-1]
    y = torch.empty(rows, dim, dtype=H16, device=x.device) if y is None else y
    rstd = torch.empty(rows, dtype=torch.float32, device=x.device) if rstd is None else rstd
    check(_lib.lib().fvqa_rmsnorm_gather_fwd(ptr(x), ptr(idx), ptr(w), ptr(y), ptr(rstd), rows, dim, eps, stream()), "rmsnorm_gather_fwd")
    return y, rstd


@_timed
def rmsnorm_scatter_bwd(dy, x, idx, w, rstd, dx, dx_h16=None):
    rows, dim = idx.numel(), x.shape[-1]
    check(_lib.lib().fvqa_rmsnorm_scatter_bwd(ptr(dy), ptr(x), ptr(idx), ptr(w), ptr(rstd), ptr(dx), ptr(dx_h16), rows, dim, stream()), "rmsnorm_scatter_bwd")
    return dx


# ------------------------------------------------------------------ SwiGLU
@_timed
def swiglu_fwd(g, c=None):
    _chk(g, H16, "g")
    rows, two_hid = g.shape
    hid = two_hid // 2
    c = torch.empty(rows, hid, dtype=H16, device=g.device) if c is None else c
    check(_lib.lib().fvqa_swiglu_fwd(ptr(g), ptr(c), rows, hid, stream()), "swiglu_fwd")
    return c


@_timed
def swiglu_bwd(dc, g, dg=None):
    _chk(dc, H16, "dc"); _chk(g, H16, "g")
    rows, hid = dc.shape
    dg = torch.empty_like(g) if dg is None else dg
    check(_lib.lib().fvqa_swiglu_bwd(ptr(dc), ptr(g), ptr(dg), rows, hid, stream()), "swiglu_bwd")
    return dg


# ------------------------------------------------------------------ GEMM
@_timed
def gemm_nt(a: torch.Tensor, b: torch.Tensor, out: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None,
            out_fp32: bool = False, M: Optional[int] = None) -> torch.Tensor:
    """out[M,N] = a[M,K] @ b[N,K]^T (+ residual). a/b may be row-strided views (last dim contiguous)."""
    assert a.dtype == H16 and b.dtype == H16 and a.stride(-1) == 1 and b.stride(-1) == 1
    Mfull, K = a.shape
    M = Mfull if M is None else M
    N, Kb = b.shape
    assert K == Kb, (a.shape, b.shape)
    if out is None:
        out = torch.empty(M, N, dtype=torch.float32 if out_fp32 else H16, device=a.device)
    assert out.stride(-1) == 1 and out.dtype == (torch.float32 if out_fp32 else H16)
    assert residual is None or (residual.dtype == out.dtype and residual.stride(-1) == 1)
    ldr = residual.stride(0) if residual is not None else 0
    tm = GEMM_TIMER
    if tm is not None and tm.active:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    check(_lib.lib().fvqa_gemm_nt(ptr(a), a.stride(0), ptr(b), b.stride(0), ptr(out), out.stride(0), ptr(residual), ldr,
                                       M, N, K, 1 if out_fp32 else 0, stream()), "gemm_h16_nt")
    if tm is not None and tm.active:
        e1.record()
        tm.records.append((e0, e1, 2.0 * M * N * K, tm.tag))
    return out


@_timed
def gemm_nt_rope(a: torch.Tensor, b: torch.Tensor, cos, sin, rope_cols: int, hd: int, S: int, out: Optional[torch.Tensor] = None,
                 pos_ids: Optional[torch.Tensor] = None) -> torch.Tensor:
    """QKV projection with RoPE applied to the q|k columns in the GEMM epilogue (h16 out). Position of row r is
    r % S, or pos_ids[r] (int32) for ragged / compacted token layouts."""
    assert a.dtype == H16 and b.dtype == H16 and a.stride(-1) == 1 and b.stride(-1) == 1
    M, K = a.shape
    N = b.shape[0]
    out = torch.empty(M, N, dtype=H16, device=a.device) if out is None else out
    tm = GEMM_TIMER
    if tm is not None and tm.active:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    if pos_ids is not None:
        _chk(pos_ids, torch.int32, "pos_ids")
        assert pos_ids.numel() >= M
        check(_lib.lib().fvqa_gemm_nt_rope_pos(ptr(a), a.stride(0), ptr(b), b.stride(0), ptr(out), out.stride(0), M, N, K,
                                                    ptr(cos), ptr(sin), rope_cols, hd, ptr(pos_ids), stream()), "gemm_h16_nt_rope_pos")
    else:
        check(_lib.lib().fvqa_gemm_nt_rope(ptr(a), a.stride(0), ptr(b), b.stride(0), ptr(out), out.stride(0), M, N, K,
                                                ptr(cos), ptr(sin), rope_cols, hd, S, stream()), "gemm_h16_nt_rope")
    if tm is not None and tm.active:
        e1.record()
        tm.records.append((e0, e1, 2.0 * M * N * K, tm.tag))
    return out


@_timed
def gemm_swiglu_fwd(x: torch.Tensor, w13: torch.Tensor, g: Optional[torch.Tensor] = None, c: Optional[torch.Tensor] = None):
    """g = x @ [W1;W3]^T (saved for backward) and c = silu(g[:, :hid]) * g[:, hid:] with the SwiGLU in the GEMM epilogue
    (`llama/model.py:142`). Falls back to GEMM + swiglu kernel when hid is not a multiple of 128. Returns (g, c)."""
    M, K = x.shape
    hid = w13.shape[0] // 2
    g = torch.empty(M, 2 * hid, dtype=H16, device=x.device) if g is None else g
    c = torch.empty(M, hid, dtype=H16, device=x.device) if c is None else c
    if hid % 128 != 0 or not FUSE_SWIGLU:
        gemm_nt(x, w13, out=g)
        swiglu_fwd(g, c)
        return g, c
    tm = GEMM_TIMER
    if tm is not None and tm.active:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    check(_lib.lib().fvqa_gemm_swiglu_fwd(ptr(x), x.stride(0), ptr(w13), w13.stride(0), ptr(g), g.stride(0), ptr(c), c.stride(0),
                                          M, hid, K, stream()), "gemm_swiglu_fwd")
    if tm is not None and tm.active:
        e1.record()
        tm.records.append((e0, e1, 2.0 * M * 2 * hid * K, tm.tag))
    return g, c


@_timed
def gemm_nn(a: torch.Tensor, b: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[M,N] = a[M,K] @ b[K,N] with b ROW-MAJOR [K, N] (a forward weight [out = K, in = N]): dX = dY . W without a transposed
    weight copy (MN-major tcgen05 B operand). h16 in / out."""
    assert a.dtype == H16 and b.dtype == H16 and a.stride(-1) == 1 and b.stride(-1) == 1
    M, K = a.shape
    Kb, N = b.shape
    assert K == Kb, (a.shape, b.shape)
    out = torch.empty(M, N, dtype=H16, device=a.device) if out is None else out
    assert out.dtype == H16 and out.stride(-1) == 1
    tm = GEMM_TIMER
    if tm is not None and tm.active:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    check(_lib.lib().fvqa_gemm_nn(ptr(a), a.stride(0), ptr(b), b.stride(0), ptr(out), out.stride(0), M, N, K, stream()), "gemm_nn")
    if tm is not None and tm.active:
        e1.record()
        tm.records.append((e0, e1, 2.0 * M * N * K, tm.tag))
    return out


@_timed
def gemm_swiglu_bwd(dy: torch.Tensor, w2t: torch.Tensor, g: torch.Tensor, dg: Optional[torch.Tensor] = None, nn: bool = False):
    """dg = swiglu'(g) . (dy @ W2t^T): backward through w2 and the SwiGLU in one GEMM (dc never materialised).
    nn=True: `w2t` is the forward weight W2 [d, hid] itself (row-major [K, N]) instead of its transposed copy [hid, d]."""
    M, K = dy.shape
    hid = w2t.shape[1] if nn else w2t.shape[0]
    dg = torch.empty(M, 2 * hid, dtype=H16, device=dy.device) if dg is None else dg
    if hid % 32 != 0 or not FUSE_SWIGLU:
        dc = gemm_nn(dy, w2t) if nn else gemm_nt(dy, w2t)
        return swiglu_bwd(dc, g, dg)
    if nn:
        tm = GEMM_TIMER
        if tm is not None and tm.active:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        check(_lib.lib().fvqa_gemm_swiglu_bwd_nn(ptr(dy), dy.stride(0), ptr(w2t), w2t.stride(0), ptr(g), g.stride(0), ptr(dg), dg.stride(0),
                                                 M, hid, K, stream()), "gemm_swiglu_bwd_nn")
        if tm is not None and tm.active:
            e1.record()
            tm.records.append((e0, e1, 2.0 * M * hid * K, tm.tag))
        return dg
    tm = GEMM_TIMER
    if tm is not None and tm.active:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    check(_lib.lib().fvqa_gemm_swiglu_bwd(ptr(dy), dy.stride(0), ptr(w2t), w2t.stride(0), ptr(g), g.stride(0), ptr(dg), dg.stride(0),
                                          M, hid, K, stream()), "gemm_swiglu_bwd")
    if tm is not None and tm.active:
        e1.record()
        tm.records.append((e0, e1, 2.0 * M * hid * K, tm.tag))
    return dg


# ------------------------------------------------------------------ attention
@_timed
def attn_fwd(qkv, akv, cos, sin, gate1, gate2, vstart, n_seq, S, H, hd, A, F, out=None, lse=None):
    _chk(qkv, H16, "qkv"); _chk(vstart, torch.int32, "vstart")
    assert akv.dtype == H16 and akv.stride(-1) == 1
    out = torch.empty(n_seq * S, H * hd, dtype=H16, device=qkv.device) if out is None else out
    lse = torch.empty(n_seq, H, S, dtype=torch.float32, device=qkv.device) if lse is None else lse
    check(_lib.lib().fvqa_attn_fwd(ptr(qkv), ptr(akv), akv.stride(0), ptr(cos), ptr(sin), ptr(gate1), ptr(gate2), ptr(vstart),
                                   ptr(out), ptr(lse), n_seq, S, H, hd, A, F, stream()), "attn_fwd")
    return out, lse


def attn_bwd_ws_bytes(n_seq, S, H, hd, A) -> int:
    return int(_lib.load().fvqa_attn_bwd_ws_bytes(n_seq, S, H, hd, A))


@_timed
def attn_bwd(qkv, akv, cos, sin, gate1, gate2, vstart, out, lse, dout, n_seq, S, H, hd, A, F,
             dqkv=None, dakv=None, dgate1=None, dgate2=None, ws=None):
    dev = qkv.device
    dqkv = torch.empty_like(qkv) if dqkv is None else dqkv
    dakv = torch.empty(A, 2 * H * hd, dtype=torch.float32, device=dev) if dakv is None else dakv
    dgate1 = torch.empty(H, dtype=torch.float32, device=dev) if dgate1 is None else dgate1
    dgate2 = torch.empty(H, dtype=torch.float32, device=dev) if dgate2 is None else dgate2
    if ws is None:
        ws = torch.empty(attn_bwd_ws_bytes(n_seq, S, H, hd, A), dtype=torch.uint8, device=dev)
    check(_lib.lib().fvqa_attn_bwd(ptr(qkv), ptr(akv), akv.stride(0), ptr(cos), ptr(sin), ptr(gate1), ptr(gate2), ptr(vstart),
                                   ptr(out), ptr(lse), ptr(dout), ptr(dqkv), ptr(dakv), ptr(dgate1), ptr(dgate2), ptr(ws),
                                   n_seq, S, H, hd, A, F, stream()),
          "attn_bwd_tc" if _lib.lib().fvqa_attn_uses_tc(S, hd, A) else "attn_bwd")
    return dqkv, dakv, dgate1, dgate2


# ------------------------------------------------------------------ input side
@_timed
def visual_proj_fwd(video2d, wv, out=None):
    _chk(video2d, torch.float32, "video"); _chk(wv, torch.float32, "wv")
    rows, vdim = video2d.shape
    dim = wv.shape[0]
    out = torch.empty(rows, dim, dtype=torch.float32, device=wv.device) if out is None else out
    check(_lib.lib().fvqa_visual_proj_fwd(ptr(video2d), ptr(wv), ptr(out), rows, dim, vdim, stream()), "visual_proj_fwd")
    return out


@_timed
def visual_proj_bwd(dvf2d, video2d, dwv=None):
    rows, vdim = video2d.shape
    dim = dvf2d.shape[1]
    dwv = torch.empty(dim, vdim, dtype=torch.float32, device=dvf2d.device) if dwv is None else dwv
    check(_lib.lib().fvqa_visual_proj_bwd(ptr(dvf2d), ptr(video2d), ptr(dwv), rows, dim, vdim, stream()), "visual_proj_bwd")
    return dwv


@_timed
def build_h0_fwd(tok_emb, ids, labels, vstart, seq_video, qav_index, vf32, temporal, n_seq, S, F, h0=None):
    dim = tok_emb.shape[1]
    h0 = torch.empty(n_seq * S, dim, dtype=torch.float32, device=tok_emb.device) if h0 is None else h0
    check(_lib.lib().fvqa_build_h0_fwd(ptr(tok_emb), ptr(ids), ptr(labels), ptr(vstart), ptr(seq_video), ptr(qav_index), ptr(vf32),
                                       ptr(temporal), ptr(h0), n_seq, S, dim, F, stream()), "build_h0_fwd")
    return h0


@_timed
def build_h0_bwd(dh0, vstart, seq_video, qav_index, n_seq, n_video, S, F, dvf=None):
    dim = dh0.shape[1]
    dvf = torch.empty(n_video * F, dim, dtype=torch.float32, device=dh0.device) if dvf is None else dvf
    check(_lib.lib().fvqa_build_h0_bwd(ptr(dh0), ptr(vstart), ptr(seq_video), ptr(qav_index), ptr(dvf), n_seq, n_video, S, dim, F, stream()), "build_h0_bwd")
    return dvf


@_timed
def video_grad_finish(dvf, dvf_qav, n_video, F, dtemporal=None):
    dim = dvf.shape[1]
    dtemporal = torch.empty(F, dim, dtype=torch.float32, device=dvf.device) if dtemporal is None else dtemporal
    check(_lib.lib().fvqa_video_grad_finish(ptr(dvf), ptr(dvf_qav), ptr(dtemporal), n_video, dim, F, stream()), "video_grad_finish")
    return dtemporal


# ------------------------------------------------------------------ heads
@_timed
def video_grad(dh0, vstart, seq_video, qav_index, dvf_qav, n_seq, n_video, S, F, dvf=None, dtemporal=None):
    """build_h0_bwd + video_grad_finish in one launch: returns (dvf [n_video*F, d] incl. the QAV term, dtemporal [F, d])."""
    dim = dh0.shape[-1]
    dvf = torch.empty(n_video * F, dim, dtype=F32, device=dh0.device) if dvf is None else dvf
    dtemporal = torch.empty(F, dim, dtype=F32, device=dh0.device) if dtemporal is None else dtemporal
    check(_lib.lib().fvqa_video_grad(ptr(dh0), ptr(vstart), ptr(seq_video), ptr(qav_index), ptr(dvf_qav), ptr(dvf), ptr(dtemporal),
                                     n_seq, n_video, S, dim, F, stream()), "video_grad")
    return dvf, dtemporal


@_timed
def ce_fwd(logits, target, row_loss=None, row_lse=None):
    _chk(target, torch.int32, "target")
    assert logits.dtype == torch.float32 and logits.stride(-1) == 1
    rows, V = target.numel(), logits.shape[1]
    dev = logits.device
    row_loss = torch.empty(rows, dtype=torch.float32, device=dev) if row_loss is None else row_loss
    row_lse = torch.empty(rows, dtype=torch.float32, device=dev) if row_lse is None else row_lse
    check(_lib.lib().fvqa_ce_fwd(ptr(logits), logits.stride(0), ptr(target), ptr(row_loss), ptr(row_lse), rows, V, stream()), "ce_fwd")
    return row_loss, row_lse


@_timed
def ce_bwd(logits, target, row_lse, gscale, inv_count: float, dlogits=None):
    rows, V = target.numel(), logits.shape[1]
    dlogits = torch.empty(rows, V, dtype=H16, device=logits.device) if dlogits is None else dlogits
    check(_lib.lib().fvqa_ce_bwd(ptr(logits), logits.stride(0), ptr(target), ptr(row_lse), ptr(gscale), inv_count, ptr(dlogits),
                                 dlogits.stride(0), rows, V, stream()), "ce_bwd")
    return dlogits


@_timed
def sum_scale(v, rows: int, scale: float, out):
    check(_lib.lib().fvqa_sum_scale(ptr(v), rows, scale, ptr(out), stream()), "sum_scale")
    return out


@_timed
def qav_loss_fwd(hn, vf32, row_video, target, tau: float, F: int, row_loss=None, prob=None):
    rows, dim = hn.shape
    dev = hn.device
    row_loss = torch.empty(rows, dtype=torch.float32, device=dev) if row_loss is None else row_loss
    prob = torch.empty(rows, F, dtype=torch.float32, device=dev) if prob is None else prob
    check(_lib.lib().fvqa_qav_loss_fwd(ptr(hn), ptr(vf32), ptr(row_video), ptr(target), tau, ptr(row_loss), ptr(prob), rows, dim, F, stream()), "qav_loss_fwd")
    return row_loss, prob


@_timed
def qav_loss_bwd(hn, vf32, row_video, target, prob, gscale, inv_count: float, tau: float, n_video: int, F: int, dhn=None, dvf_qav=None):
    rows, dim = hn.shape
    dev = hn.device
    dhn = torch.empty_like(hn) if dhn is None else dhn
    dvf_qav = torch.empty(n_video * F, dim, dtype=torch.float32, device=dev) if dvf_qav is None else dvf_qav
    check(_lib.lib().fvqa_qav_loss_bwd(ptr(hn), ptr(vf32), ptr(row_video), ptr(target), ptr(prob), ptr(gscale), inv_count, tau,
                                       ptr(dhn), ptr(dvf_qav), rows, n_video, dim, F, stream()), "qav_loss_bwd")
    return dhn, dvf_qav


@_timed
def scatter_rows(row_val, dst_index, dst):
    check(_lib.lib().fvqa_scatter_rows(ptr(row_val), ptr(dst_index), ptr(dst), dst_index.numel(), stream()), "scatter_rows")
    return dst


@_timed
def option_score(token_loss: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    _chk(token_loss, torch.float32, "token_loss")
    n_items, n_opt, ln = token_loss.shape
    pred = torch.empty(n_items, dtype=torch.int32, device=token_loss.device)
    mean = torch.empty(n_items, n_opt, dtype=torch.float32, device=token_loss.device)
    check(_lib.lib().fvqa_option_score(ptr(token_loss), ptr(pred), ptr(mean), n_items, n_opt, ln, stream()), "option_score")
    return pred, mean


@_timed
@_timed
def greedy_next(logits, tok_emb, ids, S: int, pos, out_tokens, step: int, x_next, margin=None):
    """One greedy decoding step (see include/fvqa.h): argmax per row -> ids[b, pos[b] + 1] / out_tokens[b, step] / next embedding."""
    _chk(logits, F32, "logits"); _chk(ids, torch.int32, "ids"); _chk(pos, torch.int32, "pos"); _chk(out_tokens, torch.int32, "out_tokens")
    rows, V = logits.shape
    check(_lib.lib().fvqa_greedy_next(ptr(logits), logits.stride(0), V, ptr(tok_emb), tok_emb.shape[1], ptr(ids), S, ptr(pos), ptr(out_tokens),
                                      out_tokens.stride(0), step, ptr(x_next), ptr(margin), rows, stream()), "greedy_next")
    return x_next


def grad_scale_prepare(gscale: torch.Tensor, target: float):
    """(gscale * k, 1 / k) with k = 2^round(log2(target / max|gscale|)) computed on the device (no host sync)."""
    _chk(gscale, F32, "gscale")
    assert gscale.numel() == 3
    gs = torch.empty(3, dtype=F32, device=gscale.device)
    inv_k = torch.empty(1, dtype=F32, device=gscale.device)
    check(_lib.lib().fvqa_grad_scale_prepare(ptr(gscale), float(target), ptr(gs), ptr(inv_k), stream()), "grad_scale_prepare")
    return gs, inv_k


def scale_f32(x: torch.Tensor, factor: torch.Tensor) -> torch.Tensor:
    """x *= factor[0] in place (factor: 1-element fp32 device tensor)."""
    _chk(x, F32, "x"); _chk(factor, F32, "factor")
    check(_lib.lib().fvqa_scale_f32(ptr(x), ptr(factor), x.numel(), stream()), "scale_f32")
    return x


def f32_to_h16(src, dst=None):
    _chk(src, torch.float32, "src")
    dst = torch.empty(src.shape, dtype=H16, device=src.device) if dst is None else dst
    check(_lib.lib().fvqa_f32_to_h16(ptr(src), ptr(dst), src.numel(), stream()), "f32_to_h16")
    return dst


@_timed
def gather_rows(src: torch.Tensor, idx: torch.Tensor, dst: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dst[i] = src[idx[i]] for 2-D contiguous src (any dtype whose rows are a multiple of 16 bytes)."""
    assert src.is_contiguous() and idx.dtype == torch.int32
    rows, row_bytes = idx.numel(), src.shape[1] * src.element_size()
    dst = torch.empty(rows, src.shape[1], dtype=src.dtype, device=src.device) if dst is None else dst
    check(_lib.lib().fvqa_gather_rows(ptr(src), ptr(idx), ptr(dst), rows, row_bytes, stream()), "gather_rows")
    return dst


@_timed
def scatter_row_vectors(src: torch.Tensor, idx: torch.Tensor, dst: torch.Tensor) -> torch.Tensor:
    """dst[idx[i]] = src[i] (dst pre-initialised by the caller; indices unique)."""
    assert src.is_contiguous() and dst.is_contiguous() and idx.dtype == torch.int32 and src.dtype == dst.dtype
    rows, row_bytes = idx.numel(), src.shape[1] * src.element_size()
    check(_lib.lib().fvqa_scatter_row_vectors(ptr(src), ptr(idx), ptr(dst), rows, row_bytes, stream()), "scatter_row_vectors")
    return dst


@_timed
def expand_rows(src: torch.Tensor, idx: torch.Tensor, dst: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dst[r] = src[idx[r]] where idx[r] >= 0, zeros elsewhere (compact row set -> full [n_seq * S] token layout)."""
    assert src.is_contiguous() and idx.dtype == torch.int32
    rows, row_bytes = idx.numel(), src.shape[1] * src.element_size()
    dst = torch.empty(rows, src.shape[1], dtype=src.dtype, device=src.device) if dst is None else dst
    assert dst.is_contiguous() and dst.shape[0] == rows
    check(_lib.lib().fvqa_expand_rows(ptr(src), ptr(idx), ptr(dst), rows, row_bytes, stream()), "expand_rows")
    return dst


@_timed
def linear_f32(x2d, w, bias=None, add=None, out=None):
    """fp32 y = x @ w.T (+ bias) (+ add): the small input-side Linears of the fusion variants (`llama/model.py:306-322`)."""
    _chk(x2d, torch.float32, "x"); _chk(w, torch.float32, "w")
    rows, in_dim = x2d.shape
    dim = w.shape[0]
    assert w.shape[1] == in_dim
    out = torch.empty(rows, dim, dtype=torch.float32, device=w.device) if out is None else out
    check(_lib.lib().fvqa_linear_f32(ptr(x2d), ptr(w), ptr(bias), ptr(add), ptr(out), rows, dim, in_dim, stream()), "linear_f32")
    return out


@_timed
def cross_attn_fwd(q, k, v, n_samples: int, frames: int, tokens: int):
    """`CrossAttentionModule.forward` after its Linears (`llama/model.py:161-169`): softmax(q k^T / sqrt(D)) v per sample."""
    _chk(q, torch.float32, "q"); _chk(k, torch.float32, "k"); _chk(v, torch.float32, "v")
    out = torch.empty_like(q)
    check(_lib.lib().fvqa_cross_attn_fwd(ptr(q), ptr(k), ptr(v), ptr(out), n_samples, frames, tokens, q.shape[1], stream()), "cross_attn_fwd")
    return out


@_timed
def gemm_skinny_grouped(a: torch.Tensor, b_ptrs: torch.Tensor, ldb: int, N: int, out: torch.Tensor) -> torch.Tensor:
    """out[g] = a[g] @ B_g^T for every group in one launch. a [G, M<=16, K] h16 contiguous; b_ptrs int64 device tensor of G
    weight-block pointers ([N, K] h16 each, leading dimension ldb); out [G, M, N] h16 or fp32 contiguous."""
    _chk(a, H16, "a"); _chk(b_ptrs, torch.int64, "b_ptrs")
    G, M, K = a.shape
    assert out.is_contiguous() and tuple(out.shape) == (G, M, N) and out.dtype in (H16, torch.float32) and b_ptrs.numel() >= G
    check(_lib.lib().fvqa_gemm_skinny_grouped(ptr(a), M * K, K, ptr(b_ptrs), ldb, ptr(out), M * N, N, M, N, K, G,
                                              1 if out.dtype == torch.float32 else 0, stream()), "gemm_skinny_grouped")
    return out
