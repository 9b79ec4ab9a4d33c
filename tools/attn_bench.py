"""GPU micro-benchmark of the attention kernels and one GEMM at 7B NExT-QA shapes (ncu target)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flipped_vqa_b200 import ops, _lib
from flipped_vqa_b200._lib import H16

def main():
    _lib.lib()
    n_seq, S, H, hd, A, F = 24, 128, 32, 128, 10, 10
    if len(sys.argv) > 2:
        n_seq, S = int(sys.argv[1]), int(sys.argv[2])
    D = H * hd
    g = torch.Generator(device="cuda").manual_seed(0)
    qkv = torch.randn(n_seq * S, 3 * D, device="cuda", generator=g).to(H16)
    akv = torch.randn(A, 2 * D, device="cuda", generator=g).to(H16)
    gate1 = torch.randn(H, device="cuda", generator=g) * 0.5
    gate2 = torch.full((H,), -3.5, device="cuda")
    ang = torch.outer(torch.arange(S).float(), 1.0 / (10000.0 ** (torch.arange(0, hd, 2).float() / hd)))
    cos, sin = torch.cos(ang), torch.sin(ang); cos, sin = cos.cuda().contiguous(), sin.cuda().contiguous()
    vstart = torch.tensor([18] * (2 * n_seq // 3) + [-1] * (n_seq - 2 * n_seq // 3), dtype=torch.int32, device="cuda")
    dout = torch.randn(n_seq * S, D, device="cuda", generator=g).to(H16)
    a = torch.randn(n_seq * S, 4096, device="cuda", generator=g).to(H16)
    w = (torch.randn(11008, 4096, device="cuda", generator=g) * 0.02).to(H16)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    xr = torch.randn(n_seq * S, 4096, device="cuda", generator=g)
    wn = torch.ones(4096, device="cuda", dtype=H16)
    rstd = torch.rand(n_seq * S, device="cuda", generator=g) + 0.5
    dres = torch.randn(n_seq * S, 4096, device="cuda", generator=g)
    dxo = torch.empty_like(xr); dxb = torch.empty(n_seq * S, 4096, device="cuda", dtype=H16)
    def run():
        ops.rmsnorm_bwd(a, xr, wn, rstd, dres=dres, dx=dxo, dx_h16=dxb)
        out, lse = ops.attn_fwd(qkv, akv, cos, sin, gate1, gate2, vstart, n_seq, S, H, hd, A, F)
        ops.attn_bwd(qkv, akv, cos, sin, gate1, gate2, vstart, out, lse, dout, n_seq, S, H, hd, A, F)
        ops.gemm_nt(a, w)
    for _ in range(3): run()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    tot = [0.0] * 4
    for _ in range(10):
        flush.zero_()
        ev[0].record()
        out, lse = ops.attn_fwd(qkv, akv, cos, sin, gate1, gate2, vstart, n_seq, S, H, hd, A, F)
        ev[1].record()
        ops.attn_bwd(qkv, akv, cos, sin, gate1, gate2, vstart, out, lse, dout, n_seq, S, H, hd, A, F)
        ev[2].record()
        ops.gemm_nt(a, w)
        ev[3].record()
        torch.cuda.synchronize()
        for i in range(3): tot[i] += ev[i].elapsed_time(ev[i + 1])
    print(f"attn fwd {tot[0] / 10 * 1e3:.1f} us | attn bwd (3 kernels) {tot[1] / 10 * 1e3:.1f} us | gemm 3072x11008x4096 {tot[2] / 10 * 1e3:.1f} us")

if __name__ == "__main__":
    main()
