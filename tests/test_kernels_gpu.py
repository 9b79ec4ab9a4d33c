"""Parity of every CUDA kernel (called through the C ABI) against the oracle's PyTorch restatement
of the same reference op, on seeded inputs. 16-bit I/O in the library's operand format H16 (fp16 by default, bf16 with
FVQA_DTYPE=bf16), fp32 math: tolerances are stated per test (sized for the coarser of the two formats)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from flipped_vqa_b200._lib import H16  # noqa: E402
from oracle import llama_vqa_oracle as O  # noqa: E402  (checker only)


def relerr(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / (b.norm() + 1e-30))


def h16_randn(*shape, std=1.0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, device="cuda", generator=g) * std).to(H16)


# ------------------------------------------------------------------ RMSNorm / SwiGLU
@pytest.mark.parametrize("rows,dim", [(7, 256), (1024, 4096), (33, 5120)])
def test_rmsnorm_fwd_bwd(fvqa_lib, rows, dim):
    """fp32 residual stream in, bf16 GEMM operand out; backward returns fp32 dx (+ bf16 copy)."""
    from flipped_vqa_b200 import ops
    x = h16_randn(rows, dim, seed=1).float() + 1e-3 * torch.randn(rows, dim, device="cuda")
    w = (1 + 0.1 * torch.randn(dim, device="cuda")).to(H16)
    dy = h16_randn(rows, dim, seed=2)
    res = torch.randn(rows, dim, device="cuda")
    y, rstd = ops.rmsnorm_fwd(x, w, 1e-6)
    xr = x.clone().requires_grad_(True)
    yr = O.rmsnorm(xr, w.float(), 1e-6)
    assert relerr(y, yr) < 3e-3                 # one bf16 rounding of the output
    (yr * dy.float()).sum().backward()
    dxb = torch.empty(rows, dim, dtype=H16, device="cuda")
    dx, _ = ops.rmsnorm_bwd(dy, x, w, rstd, dres=res, dx_h16=dxb)
    assert relerr(dx, xr.grad + res) < 1e-5
    assert relerr(dxb, xr.grad + res) < 3e-3
    dx2, _ = ops.rmsnorm_bwd(dy, x, w, rstd)
    assert relerr(dx2, xr.grad) < 1e-5


def test_rmsnorm_gather_scatter(fvqa_lib):
    from flipped_vqa_b200 import ops
    rows, dim = 300, 512
    x = torch.randn(rows, dim, device="cuda")
    w = (1 + 0.1 * torch.randn(dim, device="cuda")).to(H16)
    idx = torch.tensor([5, 17, -1, 299, 0, -1, 42], dtype=torch.int32, device="cuda")
    y, rstd = ops.rmsnorm_gather_fwd(x, idx, w, 1e-6)
    valid = idx >= 0
    yr = O.rmsnorm(x[idx[valid].long()], w.float(), 1e-6)
    assert relerr(y[valid], yr) < 3e-3
    assert float(y[~valid].float().abs().max()) == 0.0
    dy = h16_randn(idx.numel(), dim, seed=5)
    dx = torch.zeros_like(x)
    dxb = torch.zeros(rows, dim, dtype=H16, device="cuda")
    ops.rmsnorm_scatter_bwd(dy, x, idx, w, rstd, dx, dxb)
    xr = x.clone().requires_grad_(True)
    (O.rmsnorm(xr[idx[valid].long()], w.float(), 1e-6) * dy[valid].float()).sum().backward()
    assert relerr(dx, xr.grad) < 1e-5
    assert relerr(dxb, xr.grad) < 3e-3


@pytest.mark.parametrize("rows,hid", [(5, 768), (1024, 11008)])
def test_swiglu(fvqa_lib, rows, hid):
    from flipped_vqa_b200 import ops
    g = h16_randn(rows, 2 * hid, seed=6)
    dc = h16_randn(rows, hid, seed=7)
    c = ops.swiglu_fwd(g)
    gr = g.float().requires_grad_(True)
    cr = O.swiglu(gr[:, :hid], gr[:, hid:])
    assert relerr(c, cr) < 3e-3
    (cr * dc.float()).sum().backward()
    dg = ops.swiglu_bwd(dc, g)
    assert relerr(dg, gr.grad) < 4e-3


# ------------------------------------------------------------------ tcgen05 GEMM
GEMM_SHAPES = [
    (128, 256, 64), (128, 128, 128), (256, 512, 256), (200, 384, 128), (10, 256, 512),
    (384, 1536, 256), (1024, 4096, 4096), (3072, 11008, 4096), (3072, 4096, 11008), (96, 32000, 4096),
]


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_gemm_nt(fvqa_lib, M, N, K):
    from flipped_vqa_b200 import ops
    a = h16_randn(M, K, seed=10)
    b = h16_randn(N, K, std=0.05, seed=11)
    ref = a.float() @ b.float().t()
    c = ops.gemm_nt(a, b)
    assert relerr(c, ref) < 5e-3, f"bf16 out relerr {relerr(c, ref)}"
    c32 = ops.gemm_nt(a, b, out_fp32=True)
    assert relerr(c32, ref) < 2e-4, f"fp32 out relerr {relerr(c32, ref)}"
    r = h16_randn(M, N, seed=12)
    cr = ops.gemm_nt(a, b, residual=r)
    assert relerr(cr, ref + r.float()) < 5e-3
    r32 = torch.randn(M, N, device="cuda")            # fp32 residual stream (h = x + attn, out = h + ffn)
    cr32 = ops.gemm_nt(a, b, residual=r32, out_fp32=True)
    assert relerr(cr32, ref + r32) < 2e-4


@pytest.mark.parametrize("G,M,N,K,f32", [(5, 10, 512, 256, False), (32, 10, 1024, 512, False), (8, 10, 512, 1024, True), (3, 16, 264, 256, True)])
def test_gemm_skinny_grouped(fvqa_lib, G, M, N, K, f32):
    """One launch for the adapter projections of all layers: C_g = A_g B_g^T with the weight blocks B_g in separate
    allocations (device pointer table) and strided views (ldb > K, as the [Wk; Wv] rows of the packed Wqkv^T)."""
    from flipped_vqa_b200 import ops
    a = h16_randn(G * M, K, seed=70).view(G, M, K)
    ldb = K + 64
    blocks = [h16_randn(N, ldb, std=0.05, seed=71 + g) for g in range(G)]
    ptrs = torch.tensor([b.data_ptr() for b in blocks], dtype=torch.int64, device="cuda")
    out = torch.full((G + 1, M, N), 7.0, device="cuda", dtype=torch.float32 if f32 else H16)
    ops.gemm_skinny_grouped(a, ptrs, ldb, N, out[:G])
    assert torch.all(out[G].float() == 7.0)
    for g in range(G):
        ref = a[g].float() @ blocks[g][:, :K].float().t()
        assert relerr(out[g], ref) < (2e-4 if f32 else 5e-3), g
        single = ops.gemm_nt(a[g], blocks[g][:, :K], out_fp32=f32)            # the per-layer launch it replaces
        assert torch.equal(single, out[g]), g


@pytest.mark.parametrize("M,N,K", [(256, 512, 64), (700, 1536, 384), (129, 512, 128), (3072, 4096, 4096), (2535, 4096, 1024)])
def test_gemm_quad_cluster_multicast(fvqa_lib, M, N, K):
    """2x2-cluster variant (two CTA pairs share an A slice through TMA multicast): bit-identical to the CTA-pair kernel —
    same tiles, same k order — for bf16 and fp32 + residual outputs, ragged M included, nothing written out of bounds."""
    from flipped_vqa_b200 import ops
    if fvqa_lib.fvqa_gemm_quad_clusters() <= 0:
        pytest.skip("no 4-CTA clusters on this device")
    a = h16_randn(M, K, seed=80)
    b = h16_randn(N, K, std=0.05, seed=81)
    r32 = torch.randn(M, N, device="cuda")
    prev = fvqa_lib.fvqa_gemm_debug_quad(0)
    try:
        p16, p32 = ops.gemm_nt(a, b), ops.gemm_nt(a, b, residual=r32, out_fp32=True)
        fvqa_lib.fvqa_gemm_debug_quad(2)
        guard = torch.full((M + 2, N), 7.0, device="cuda")
        q32 = ops.gemm_nt(a, b, residual=r32, out_fp32=True, out=guard[1:M + 1])
        q16 = ops.gemm_nt(a, b)
    finally:
        fvqa_lib.fvqa_gemm_debug_quad(prev)
    assert torch.all(guard[0] == 7.0) and torch.all(guard[M + 1] == 7.0)
    assert torch.equal(q16, p16) and torch.equal(q32, p32)
    assert relerr(q32, a.float() @ b.float().t() + r32) < 2e-4


@pytest.mark.parametrize("bn", [-1, 64, 128, 144, 176, 208, 240, 256])
def test_gemm_pair_tile_widths(fvqa_lib, bn):
    """The CTA-pair (cta_group::2) kernel for every runtime tile width (and the single-CTA kernel, bn=-1)
    on ragged shapes: M not a multiple of 256 (peer CTA partly / fully out of range), N not a multiple of bn."""
    from flipped_vqa_b200 import ops
    for (M, N, K) in [(200, 384, 128), (384, 1536, 256), (650, 1000, 192), (129, 72, 64)]:
        a = h16_randn(M, K, seed=20)
        b = h16_randn(N, K, std=0.05, seed=21)
        r32 = torch.randn(M, N, device="cuda")
        ref = a.float() @ b.float().t()
        prev = fvqa_lib.fvqa_gemm_debug_force_bn(bn)
        try:
            c32 = ops.gemm_nt(a, b, residual=r32, out_fp32=True)
            c16 = ops.gemm_nt(a, b)
        finally:
            fvqa_lib.fvqa_gemm_debug_force_bn(prev)
        assert relerr(c32, ref + r32) < 2e-4, (M, N, K, bn)
        assert relerr(c16, ref) < 5e-3, (M, N, K, bn)


@pytest.mark.parametrize("M,hid,K", [(200, 384, 128), (1024, 768, 256), (3072, 11008, 4096)])
def test_gemm_swiglu_fused_matches_unfused(fvqa_lib, M, hid, K):
    """SwiGLU fused into the W1|W3 GEMM epilogue (fwd) and into the W2^T GEMM epilogue (bwd) must be
    BIT-identical to GEMM + swiglu kernel (llama/model.py:142), and close to an fp32 restatement."""
    from flipped_vqa_b200 import ops
    x = h16_randn(M, K, seed=30)
    w13 = h16_randn(2 * hid, K, std=0.05, seed=31)
    g_ref = ops.gemm_nt(x, w13)
    c_ref = ops.swiglu_fwd(g_ref)
    g, c = ops.gemm_swiglu_fwd(x, w13)
    assert torch.equal(g, g_ref) and torch.equal(c, c_ref)
    gf = x.float() @ w13.float().t()
    assert relerr(c, torch.nn.functional.silu(gf[:, :hid]) * gf[:, hid:]) < 1e-2
    d = 4 * K if K <= 256 else K
    dy = h16_randn(M, d, seed=32)
    w2t = h16_randn(hid, d, std=0.05, seed=33)
    dc_ref = ops.gemm_nt(dy, w2t)
    dg_ref = ops.swiglu_bwd(dc_ref, g_ref)
    dg = ops.gemm_swiglu_bwd(dy, w2t, g_ref)
    assert torch.equal(dg, dg_ref)


@pytest.mark.parametrize("S,H,hd,B", [(48, 2, 64, 3), (128, 4, 128, 2)])
def test_gemm_rope_epilogue(fvqa_lib, S, H, hd, B):
    """QKV projection with RoPE folded into the epilogue == plain projection followed by the oracle's
    apply_rope on the q|k parts (llama/model.py:61-67,89-96)."""
    from flipped_vqa_b200 import ops
    d = H * hd
    x = h16_randn(B * S, d, seed=17)
    w = h16_randn(3 * d, d, std=0.05, seed=18)
    cos, sin = O.rope_table(hd, S)
    cos, sin = cos.cuda().contiguous(), sin.cuda().contiguous()
    out = ops.gemm_nt_rope(x, w, cos, sin, 2 * d, hd, S)
    ref = (x.float() @ w.float().t()).view(B, S, 3, H, hd)
    q = O.apply_rope(ref[:, :, 0], cos, sin)
    k = O.apply_rope(ref[:, :, 1], cos, sin)
    ref = torch.stack([q, k, ref[:, :, 2]], dim=2).reshape(B * S, 3 * d)
    assert relerr(out, ref) < 5e-3


def test_gemm_rope_epilogue_position_table(fvqa_lib):
    """fvqa_gemm_nt_rope_pos: the row's position comes from an int32 table (ragged / compacted token layouts of
    shared-prefix option scoring). Must be bit-identical to the row % S variant on the same (row, position) pairs."""
    from flipped_vqa_b200 import ops
    S, H, hd, B = 128, 4, 128, 3
    d = H * hd
    x = h16_randn(B * S, d, seed=27)
    w = h16_randn(3 * d, d, std=0.05, seed=28)
    cos, sin = O.rope_table(hd, S)
    cos, sin = cos.cuda().contiguous(), sin.cuda().contiguous()
    dense = ops.gemm_nt_rope(x, w, cos, sin, 2 * d, hd, S)
    g = torch.Generator().manual_seed(5)
    rows = torch.randperm(B * S, generator=g)[:301].sort().values                 # a ragged subset of the rows
    pos = (rows % S).to(torch.int32).cuda()
    xc = x[rows.cuda()].contiguous()
    out = ops.gemm_nt_rope(xc, w, cos, sin, 2 * d, hd, S, pos_ids=pos)
    assert torch.equal(out, dense[rows.cuda()])


@pytest.mark.parametrize("M,H,hd,hid", [(8, 32, 128, 11008), (3, 2, 64, 256), (16, 4, 128, 768)])
def test_decode_step_projections(fvqa_lib, M, H, hd, hid):
    """The M = bsz <= 16 rows of a KV-cached decode step (StepEngine.generate) take the HBM-bound skinny kernel with the same
    epilogues as the tcgen05 GEMM: RoPE by a per-row position table (QKV) and the fp32 residual add (wo, w2). Checked against
    the tcgen05 kernels on the same inputs (debug hook: force the single-CTA tile) and against torch."""
    from flipped_vqa_b200 import ops
    d, S = H * hd, 128
    x = h16_randn(M, d, seed=91)
    wqkv = h16_randn(3 * d, d, std=0.05, seed=92)
    w2 = h16_randn(d, hid, std=0.05, seed=93)
    cact = h16_randn(M, hid, seed=94)
    res = torch.randn(M, d, device="cuda")
    cos, sin = O.rope_table(hd, 2 * S)
    cos, sin = cos.cuda().contiguous(), sin.cuda().contiguous()
    pos = torch.randint(0, S, (M,), dtype=torch.int32).cuda()
    qkv = ops.gemm_nt_rope(x, wqkv, cos, sin, 2 * d, hd, S, pos_ids=pos)
    y = ops.gemm_nt(cact, w2, residual=res, out_fp32=True)
    prev = fvqa_lib.fvqa_gemm_debug_force_bn(-1)
    try:
        qkv_tc = ops.gemm_nt_rope(x, wqkv, cos, sin, 2 * d, hd, S, pos_ids=pos)
        y_tc = ops.gemm_nt(cact, w2, residual=res, out_fp32=True)
    finally:
        fvqa_lib.fvqa_gemm_debug_force_bn(prev)
    ref = (x.float() @ wqkv.float().t()).view(M, 1, 3, H, hd)
    c, s_ = cos[pos.long()][:, None], sin[pos.long()][:, None]                       # [M, 1, hd/2]

    def rot(t):                                                                      # t [M, 1, H, hd]
        a, b = t[..., 0::2], t[..., 1::2]
        return torch.stack((a * c[:, :, None] - b * s_[:, :, None], a * s_[:, :, None] + b * c[:, :, None]), dim=-1).flatten(3)
    ref = torch.stack([rot(ref[:, :, 0]), rot(ref[:, :, 1]), ref[:, :, 2]], dim=2).reshape(M, 3 * d)
    assert relerr(qkv, ref) < 5e-3 and relerr(qkv, qkv_tc) < 2e-3
    assert relerr(y, cact.float() @ w2.float().t() + res) < 2e-4 and relerr(y, y_tc) < 1e-4


@pytest.mark.parametrize("M,N,K", [(3072, 4096, 12288), (3072, 4096, 4096), (2304, 4096, 22016), (260, 4096, 4096), (180, 4096, 32000),
                                   (700, 1536, 384), (129, 520, 128), (40, 264, 64)])
def test_gemm_nn_equals_nt_on_transposed_copy(fvqa_lib, M, N, K):
    """dX = dY . W from the forward's [out, in] weight (MN-major tcgen05 B operand, `fvqa_gemm_nn`) must be BIT-identical to the
    K-major kernel on a transposed copy of W (same tiles, same k order) - incl. the 2x2-cluster variant on the N = 4096 shapes, ragged
    M / N, and nothing written out of bounds."""
    from flipped_vqa_b200 import ops
    a = h16_randn(M, K, seed=101)
    w = h16_randn(K, N, std=0.05, seed=102)                      # forward weight [out = K, in = N]
    buf = torch.full((M + 2, N), 7.0, device="cuda", dtype=H16)
    out = ops.gemm_nn(a, w, out=buf[1:M + 1])
    prev = fvqa_lib.fvqa_gemm_debug_force_bn(256)                 # the NN kernel always uses 256-wide tiles
    try:
        ref = ops.gemm_nt(a, w.t().contiguous())
    finally:
        fvqa_lib.fvqa_gemm_debug_force_bn(prev)
    assert relerr(out, a.float() @ w.float()) < 5e-3
    assert torch.equal(out, ref) or relerr(out, ref) < 1e-6       # quad / pair scheduling differs for some shapes: same math
    assert torch.all(buf[0].float() == 7.0) and torch.all(buf[M + 1].float() == 7.0)


def test_gemm_swiglu_bwd_nn(fvqa_lib):
    """SwiGLU backward through W2 read as [d, hid] (no transposed copy) == the K-major kernel on W2^T, bit for bit."""
    from flipped_vqa_b200 import ops
    M, d, hid = 700, 256, 768
    dy = h16_randn(M, d, seed=111)
    w2 = h16_randn(d, hid, std=0.05, seed=112)
    g = h16_randn(M, 2 * hid, seed=113)
    a = ops.gemm_swiglu_bwd(dy, w2, g, nn=True)
    b = ops.gemm_swiglu_bwd(dy, w2.t().contiguous(), g)
    assert torch.equal(a, b)


def test_gemm_strided_views(fvqa_lib):
    """Sub-blocks of larger matrices (used for Wk|Wv of the fused QKV weight and its transpose)."""
    from flipped_vqa_b200 import ops
    d = 256
    a_full = h16_randn(16, 2 * d, seed=13)
    wT = h16_randn(d, 3 * d, std=0.05, seed=14)          # [d, 3d] ; use columns d..3d  -> B[N=d, K=2d], ldb=3d
    b = wT[:, d:]
    ref = a_full[:10].float() @ b.float().t()
    out = ops.gemm_nt(a_full, b, out_fp32=True, M=10)
    assert out.shape == (10, d)
    assert relerr(out, ref) < 2e-4
    w = h16_randn(3 * d, d, std=0.05, seed=15)           # rows d..3d of [3d, d]
    x = h16_randn(10, d, seed=16)
    out2 = ops.gemm_nt(x, w[d:])
    assert relerr(out2, x.float() @ w[d:].float().t()) < 5e-3


def test_gemm_rejects_bad_shapes(fvqa_lib):
    from flipped_vqa_b200 import ops, _lib
    a = h16_randn(128, 72)
    b = h16_randn(128, 72)
    with pytest.raises(_lib.FvqaError):
        ops.gemm_nt(a, b)           # K not a multiple of 64


# ------------------------------------------------------------------ attention
def _attn_case(n_seq, S, H, hd, A, F, vstarts, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    D = H * hd
    qkv = (torch.randn(n_seq * S, 3 * D, device="cuda", generator=g)).to(H16)
    akv = (torch.randn(16, 2 * D, device="cuda", generator=g)).to(H16)
    akv[A:] = 0
    gate1 = torch.randn(H, device="cuda", generator=g) * 0.5
    gate2 = -3.5 + 0.1 * torch.randn(H, device="cuda", generator=g)
    cos, sin = O.rope_table(hd, S)
    cos, sin = cos.cuda().contiguous(), sin.cuda().contiguous()
    vstart = torch.tensor(vstarts, dtype=torch.int32, device="cuda")
    dout = (torch.randn(n_seq * S, D, device="cuda", generator=g)).to(H16)
    return qkv, akv, gate1, gate2, cos, sin, vstart, dout


def _rotate_qk(qkv, cos, sin, n_seq, S, H, hd):
    x = qkv.float().view(n_seq, S, 3, H, hd)
    q = O.apply_rope(x[:, :, 0], cos, sin)
    k = O.apply_rope(x[:, :, 1], cos, sin)
    return torch.stack([q, k, x[:, :, 2]], dim=2).reshape(n_seq * S, 3 * H * hd).to(H16).contiguous()


def _attn_ref(qkv, akv, gate1, gate2, cos, sin, vstarts, dout, n_seq, S, H, hd, A, F):
    D = H * hd
    qkv_r = qkv.float().requires_grad_(True)
    akv_r = akv[:A].float().requires_grad_(True)
    g1 = gate1.clone().requires_grad_(True)
    g2 = gate2.clone().requires_grad_(True)
    outs = []
    x = qkv_r.view(n_seq, S, 3, H, hd)
    ak = akv_r[:, :D].view(A, H, hd)
    av = akv_r[:, D:].view(A, H, hd)
    for n in range(n_seq):
        vs = vstarts[n]
        o = O.attention_core(x[n:n + 1, :, 0], x[n:n + 1, :, 1], x[n:n + 1, :, 2], ak, av, g1.view(1, H, 1, 1), g2.view(1, H, 1, 1),
                             cos, sin, None if vs < 0 else vs, F)
        outs.append(o)
    out = torch.cat(outs, 0).view(n_seq * S, D)
    (out * dout.float()).sum().backward()
    return out.detach(), qkv_r.grad, akv_r.grad, g1.grad, g2.grad


@pytest.mark.parametrize("n_seq,S,H,hd,vstarts,use_tc", [
    (3, 48, 2, 64, [12, 12, -1], 1),          # hd 64: mma.sync kernels
    (2, 130, 2, 64, [100, -1], 1),
    (3, 128, 2, 128, [18, -1, 18], 1),        # tcgen05, one tile (7B / 13B NExT-QA shape)
    (2, 100, 2, 128, [18, -1], 1),            # tcgen05, partial tile
    (2, 200, 3, 128, [18, -1], 1),            # tcgen05 tiled (S > 128), ragged last tile
    (2, 384, 2, 128, [18, -1], 1),            # DramaQA-shaped
    (1, 650, 2, 128, [18], 1),                # TVQA-shaped
    (3, 128, 2, 128, [18, -1, 18], 0),        # the same shapes on the mma.sync kernels (hook)
    (2, 200, 3, 128, [18, -1], 0),
])
def test_attention_fwd_bwd(fvqa_lib, n_seq, S, H, hd, vstarts, use_tc):
    from flipped_vqa_b200 import ops
    A, F = 10, 10
    prev = fvqa_lib.fvqa_attn_debug_use_tc(use_tc)
    try:
        assert fvqa_lib.fvqa_attn_uses_tc(S, hd, A) == (1 if (use_tc and hd == 128 and S <= 128) else 0)
        _attention_fwd_bwd_case(n_seq, S, H, hd, vstarts, A, F)
    finally:
        fvqa_lib.fvqa_attn_debug_use_tc(prev)


def _attention_fwd_bwd_case(n_seq, S, H, hd, vstarts, A, F):
    from flipped_vqa_b200 import ops
    qkv, akv, gate1, gate2, cos, sin, vstart, dout = _attn_case(n_seq, S, H, hd, A, F, vstarts, seed=20)
    qkv_rot = _rotate_qk(qkv, cos, sin, n_seq, S, H, hd)          # the kernel receives rotated q|k (GEMM epilogue)
    out, lse = ops.attn_fwd(qkv_rot, akv, cos, sin, gate1, gate2, vstart, n_seq, S, H, hd, A, F)
    ref_out, ref_dqkv, ref_dakv, ref_dg1, ref_dg2 = _attn_ref(qkv, akv, gate1, gate2, cos, sin, vstarts, dout, n_seq, S, H, hd, A, F)
    assert relerr(out, ref_out) < 1e-2, f"out {relerr(out, ref_out)}"
    dqkv, dakv, dg1, dg2 = ops.attn_bwd(qkv_rot, akv, cos, sin, gate1, gate2, vstart, out, lse, dout, n_seq, S, H, hd, A, F)
    D = H * hd                                                   # dq|dk come back inverse-rotated: gradients of the raw projections
    for name, sl in (("dq", slice(0, D)), ("dk", slice(D, 2 * D)), ("dv", slice(2 * D, 3 * D))):
        e = relerr(dqkv[:, sl], ref_dqkv[:, sl])
        assert e < 2e-2, f"{name} relerr {e}"
    assert relerr(dakv, ref_dakv) < 2e-2, f"dakv {relerr(dakv, ref_dakv)}"
    assert relerr(dg1, ref_dg1) < 2e-2, f"dgate1 {relerr(dg1, ref_dg1)} {dg1} {ref_dg1}"
    assert relerr(dg2, ref_dg2) < 2e-2, f"dgate2 {relerr(dg2, ref_dg2)} {dg2} {ref_dg2}"


@pytest.mark.parametrize("A,F", [(1, 1), (4, 6), (7, 3), (16, 12)])
@pytest.mark.parametrize("S,hd,use_tc", [(48, 64, 1), (128, 128, 1), (100, 128, 1), (200, 128, 1), (128, 128, 0)])
def test_attention_adapter_len_and_max_feats(fvqa_lib, A, F, S, hd, use_tc):
    """`--adapter_len` / `--max_feats` are command-line arguments of the reference (`train.py`), not constants: every attention path
    (mma.sync, tcgen05 one tile / partial tile / tiled) with 1..16 adapter prompts and 1..12 video columns under the gate2 bias."""
    prev = fvqa_lib.fvqa_attn_debug_use_tc(use_tc)
    try:
        _attention_fwd_bwd_case(2, S, 2, hd, [18, -1], A, F)
    finally:
        fvqa_lib.fvqa_attn_debug_use_tc(prev)


@pytest.mark.parametrize("S", [128, 300])
def test_attention_deterministic(fvqa_lib, S):
    from flipped_vqa_b200 import ops
    n_seq, H, hd, A, F = 3, 2, 128, 10, 10
    qkv, akv, gate1, gate2, cos, sin, vstart, dout = _attn_case(n_seq, S, H, hd, A, F, [18, 18, -1], seed=21)
    qkv = _rotate_qk(qkv, cos, sin, n_seq, S, H, hd)
    out, lse = ops.attn_fwd(qkv, akv, cos, sin, gate1, gate2, vstart, n_seq, S, H, hd, A, F)
    r1 = ops.attn_bwd(qkv, akv, cos, sin, gate1, gate2, vstart, out, lse, dout, n_seq, S, H, hd, A, F)
    r2 = ops.attn_bwd(qkv, akv, cos, sin, gate1, gate2, vstart, out, lse, dout, n_seq, S, H, hd, A, F)
    for a, b in zip(r1, r2):
        assert torch.equal(a, b)


# ------------------------------------------------------------------ input side + heads
def test_visual_proj_and_h0(fvqa_lib):
    from flipped_vqa_b200 import ops
    B, F, d, V, S, vdim = 3, 10, 256, 500, 40, 768
    g = torch.Generator(device="cuda").manual_seed(30)
    video = torch.randn(B * F, vdim, device="cuda", generator=g)
    wv = torch.randn(d, vdim, device="cuda", generator=g) / math.sqrt(vdim)
    temporal = torch.randn(F, d, device="cuda", generator=g)
    vf = ops.visual_proj_fwd(video, wv)
    assert relerr(vf, video @ wv.t()) < 1e-5
    dvf_in = torch.randn(B * F, d, device="cuda", generator=g)
    dwv = ops.visual_proj_bwd(dvf_in, video)
    assert relerr(dwv, dvf_in.t() @ video) < 1e-5
    emb = h16_randn(V, d, seed=31)
    n_seq = 3 * B
    ids = torch.randint(0, V, (n_seq, S), device="cuda", generator=g, dtype=torch.int32)
    labels = torch.zeros(n_seq, S, dtype=torch.int32, device="cuda")
    qav_index = torch.stack([torch.arange(p, p + F) for p in (20, 25, 29)]).to(torch.int32).cuda()
    labels[2 * B:] = -1
    for b in range(B):
        labels[2 * B + b, qav_index[b].long()] = torch.arange(F, dtype=torch.int32, device="cuda")
    vstart = torch.tensor([12] * (2 * B) + [-1] * B, dtype=torch.int32, device="cuda")
    seq_video = torch.tensor(list(range(B)) * 3, dtype=torch.int32, device="cuda")
    h0 = ops.build_h0_fwd(emb, ids, labels, vstart, seq_video, qav_index, vf, temporal, n_seq, S, F)
    # reference (llama/model.py:324-336)
    video_feature = (vf.view(B, F, d) + temporal[None]).to(H16)
    ref = emb[ids.long()].clone()
    ref[:2 * B, 12:12 + F] = video_feature.repeat(2, 1, 1)
    q = ref[2 * B:] * (~(labels[2 * B:] >= 0))[..., None]
    q = q.scatter_add(1, qav_index.long()[..., None].expand(-1, -1, d), video_feature)
    ref[2 * B:] = q
    assert h0.dtype == torch.float32 and torch.equal(h0.view(n_seq, S, d), ref.float())
    dh0 = torch.randn(n_seq * S, d, device="cuda", generator=g)
    dvf = ops.build_h0_bwd(dh0, vstart, seq_video, qav_index, n_seq, B, S, F)
    dh = dh0.view(n_seq, S, d)
    ref_dvf = dh[:B, 12:12 + F] + dh[B:2 * B, 12:12 + F] + torch.gather(dh[2 * B:], 1, qav_index.long()[..., None].expand(-1, -1, d))
    assert relerr(dvf, ref_dvf.reshape(B * F, d)) < 1e-6
    dq = torch.randn(B * F, d, device="cuda", generator=g)
    dvf_copy = dvf.clone()
    dtemp = ops.video_grad_finish(dvf, dq, B, F)
    assert relerr(dtemp, dvf_copy.view(B, F, d).sum(0)) < 1e-6
    assert relerr(dvf, dvf_copy + dq) < 1e-6
    # the fused full-grid kernel the step uses = the two calls above
    dvf2, dtemp2 = ops.video_grad(dh0, vstart, seq_video, qav_index, dq, n_seq, B, S, F)
    assert torch.equal(dvf2, dvf) and torch.equal(dtemp2, dtemp)
    dvf3, dtemp3 = ops.video_grad(dh0, vstart, seq_video, qav_index, None, n_seq, B, S, F)
    assert torch.equal(dvf3, dvf_copy) and torch.equal(dtemp3, dtemp)


@pytest.mark.parametrize("rows,dim,in_dim,bias,add", [(80, 4096, 768, False, False), (30, 256, 1792, True, True), (300, 520, 100, True, False),
                                                      (1, 32, 1, False, True)])
def test_linear_f32_shapes(fvqa_lib, rows, dim, in_dim, bias, add):
    """fp32 Linear (visual_proj `model.py:322`, the audio-fusion projections `:307-320`, CrossAttentionModule q/k/v `:148-163`) and its
    weight gradient on ragged shapes: row passes > 1, k tails, column tails, bias / additive term."""
    from flipped_vqa_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(rows + dim)
    x = torch.randn(rows, in_dim, device="cuda", generator=g)
    w = torch.randn(dim, in_dim, device="cuda", generator=g) / math.sqrt(in_dim)
    b = torch.randn(dim, device="cuda", generator=g) if bias else None
    a = torch.randn(rows, dim, device="cuda", generator=g) if add else None
    buf = torch.full((rows + 2, dim), 7.0, device="cuda")
    y = ops.linear_f32(x, w, bias=b, add=a, out=buf[1:rows + 1])
    ref = x.double() @ w.double().t() + (b.double() if bias else 0) + (a.double() if add else 0)
    assert relerr(y, ref.float()) < 1e-5
    assert float((buf[0] - 7).abs().max()) == 0 and float((buf[-1] - 7).abs().max()) == 0
    dy = torch.randn(rows, dim, device="cuda", generator=g)
    dw = ops.visual_proj_bwd(dy, x)
    assert relerr(dw, (dy.double().t() @ x.double()).float()) < 1e-5


def test_ce_fwd_bwd(fvqa_lib):
    from flipped_vqa_b200 import ops
    rows, V = 37, 32000
    g = torch.Generator(device="cuda").manual_seed(40)
    logits = torch.randn(rows, V, device="cuda", generator=g) * 3
    target = torch.randint(1, V, (rows,), device="cuda", generator=g, dtype=torch.int32)
    target[5] = -1
    target[20] = -1
    row_loss, row_lse = ops.ce_fwd(logits, target)
    valid = target >= 0
    lr = logits.clone().requires_grad_(True)
    ref = torch.nn.functional.cross_entropy(lr[valid], target[valid].long(), reduction="none")
    assert relerr(row_loss[valid], ref) < 1e-5
    assert float(row_loss[~valid].abs().max()) == 0.0
    n = int(valid.sum())
    out = torch.zeros(1, device="cuda")
    ops.sum_scale(row_loss, rows, 1.0 / n, out)
    assert abs(float(out) - float(ref.mean())) < 1e-4
    gs = torch.tensor([2.5], device="cuda")
    dl = ops.ce_bwd(logits, target, row_lse, gs, 1.0 / n)
    (ref.mean() * 2.5).backward()
    assert relerr(dl[valid], lr.grad[valid]) < 5e-3
    assert float(dl[~valid].float().abs().max()) == 0.0


def test_qav_loss(fvqa_lib):
    from flipped_vqa_b200 import ops
    B, F, d, tau = 4, 10, 512, 100.0
    g = torch.Generator(device="cuda").manual_seed(41)
    rows = B * F + 3
    hn = h16_randn(rows, d, seed=42)
    vf = torch.randn(B * F, d, device="cuda", generator=g)
    row_video = torch.tensor([b for b in range(B) for _ in range(F)] + [-1, -1, -1], dtype=torch.int32, device="cuda")
    target = torch.tensor(list(range(F)) * B + [0, 0, 0], dtype=torch.int32, device="cuda")
    row_loss, prob = ops.qav_loss_fwd(hn, vf, row_video, target, tau, F)
    hr = hn.float().requires_grad_(True)
    vr = vf.clone().requires_grad_(True)
    valid = row_video >= 0
    logits = torch.einsum("rd,rfd->rf", hr[valid], vr.view(B, F, d)[row_video[valid].long()]) / tau
    ref = torch.nn.functional.cross_entropy(logits, target[valid].long(), reduction="none")
    assert relerr(row_loss[valid], ref) < 1e-5
    gs = torch.tensor([0.7], device="cuda")
    n = int(valid.sum())
    dhn, dvfq = ops.qav_loss_bwd(hn, vf, row_video, target, prob, gs, 1.0 / n, tau, B, F)
    (ref.mean() * 0.7).backward()
    assert relerr(dhn[valid], hr.grad[valid]) < 5e-3
    assert relerr(dvfq, vr.grad) < 1e-5


def test_option_score(fvqa_lib):
    from flipped_vqa_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(43)
    tok = torch.zeros(6, 5, 47, device="cuda")
    tok[:, :, 40:44] = torch.rand(6, 5, 4, device="cuda", generator=g) * 5
    tok[2, 3, 41] = 0.0      # an exactly-zero labelled loss is NOT counted (engine.py:88 quirk)
    pred, mean = ops.option_score(tok)
    ref = O.option_predict(tok)
    assert torch.equal(pred.long(), ref)
    vals = torch.tensor([1.0, 2.0, 3.0], device="cuda")
    dst = torch.zeros(10, device="cuda")
    ops.scatter_rows(vals, torch.tensor([7, -1, 2], dtype=torch.int32, device="cuda"), dst)
    assert dst.tolist() == [0, 0, 3.0, 0, 0, 0, 0, 1.0, 0, 0]


# ------------------------------------------------------------------ canaries (compute-sanitizer is not available on the pool)
def _with_canary(shape, dtype, pad_elems=4096, value=7.0):
    n = 1
    for s in shape:
        n *= s
    buf = torch.full((n + 2 * pad_elems,), value, dtype=dtype, device="cuda")
    return buf, buf[pad_elems:pad_elems + n].view(*shape)


def _canary_intact(buf, n_inner, pad_elems=4096, value=7.0):
    return bool((buf[:pad_elems] == value).all()) and bool((buf[pad_elems + n_inner:] == value).all())


@pytest.mark.parametrize("n_seq,S,H", [(2, 100, 2), (3, 128, 2), (2, 200, 2), (1, 650, 2)])
def test_attention_writes_stay_in_bounds(fvqa_lib, n_seq, S, H):
    """Every output of the tcgen05 attention kernels (TMA stores clipped at the sequence end, per-row stores, workspace)
    lands inside its buffer: ragged last tiles must not spill into the next sequence / past the allocation."""
    from flipped_vqa_b200 import ops
    hd, A, F = 128, 10, 10
    D = H * hd
    qkv, akv, gate1, gate2, cos, sin, vstart, dout = _attn_case(n_seq, S, H, hd, A, F, [18] * n_seq, seed=40)
    out_buf, out = _with_canary((n_seq * S, D), H16)
    lse_buf, lse = _with_canary((n_seq, H, S), torch.float32)
    ops.attn_fwd(qkv, akv, cos, sin, gate1, gate2, vstart, n_seq, S, H, hd, A, F, out=out, lse=lse)
    torch.cuda.synchronize()
    assert _canary_intact(out_buf, out.numel()) and _canary_intact(lse_buf, lse.numel())
    assert not bool((out == 7.0).all(dim=1).any()), "a row of the output was never written"
    dq_buf, dqkv = _with_canary((n_seq * S, 3 * D), H16)
    dakv_buf, dakv = _with_canary((A, 2 * D), torch.float32)
    g1_buf, dg1 = _with_canary((H,), torch.float32)
    g2_buf, dg2 = _with_canary((H,), torch.float32)
    nbytes = ops.attn_bwd_ws_bytes(n_seq, S, H, hd, A)
    ws_buf, ws = _with_canary((nbytes,), torch.uint8, value=7)
    ops.attn_bwd(qkv, akv, cos, sin, gate1, gate2, vstart, out, lse, dout, n_seq, S, H, hd, A, F,
                 dqkv=dqkv, dakv=dakv, dgate1=dg1, dgate2=dg2, ws=ws)
    torch.cuda.synchronize()
    for buf, inner in ((dq_buf, dqkv), (dakv_buf, dakv), (g1_buf, dg1), (g2_buf, dg2)):
        assert _canary_intact(buf, inner.numel())
    assert _canary_intact(ws_buf, nbytes, value=7)
    assert not bool((dqkv == 7.0).all(dim=1).any()), "a row of dqkv was never written"


@pytest.mark.parametrize("M,N,K", [(200, 384, 128), (650, 1000, 192), (3072, 4096, 256)])
def test_gemm_writes_stay_in_bounds(fvqa_lib, M, N, K):
    from flipped_vqa_b200 import ops
    a = h16_randn(M, K, seed=41)
    b = h16_randn(N, K, std=0.05, seed=42)
    for f32 in (False, True):
        buf, c = _with_canary((M, N), torch.float32 if f32 else H16)
        ops.gemm_nt(a, b, out=c, out_fp32=f32)
        torch.cuda.synchronize()
        assert _canary_intact(buf, c.numel())
    if N % 256 == 0:
        hid = N // 2
        gbuf, g = _with_canary((M, N), H16)
        cbuf, c = _with_canary((M, hid), H16)
        ops.gemm_swiglu_fwd(a, b, g=g, c=c)
        dbuf, dg = _with_canary((M, N), H16)
        ops.gemm_swiglu_bwd(a, h16_randn(hid, K, std=0.05, seed=43), g, dg=dg)
        torch.cuda.synchronize()
        assert _canary_intact(gbuf, g.numel()) and _canary_intact(cbuf, c.numel()) and _canary_intact(dbuf, dg.numel())
