"""2x2-cluster multicast GEMM (fvqa_gemm_debug_quad) vs the CTA-pair kernel: correctness on small and step shapes, burst time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flipped_vqa_b200 import _lib, ops
from flipped_vqa_b200._lib import H16
from tools.gemm_step_shapes import timeit

lib = _lib.lib()
print("quad clusters on this device:", lib.fvqa_gemm_quad_clusters(), flush=True)
shapes = [(256, 512, 64), (256, 512, 256), (512, 1024, 512), (700, 1536, 384), (3072, 4096, 4096), (3072, 4096, 11008), (2535, 4096, 4096)]
if len(sys.argv) > 1 and sys.argv[1] == "small":
    shapes = shapes[:2]
for (M, N, K) in shapes:
    a = torch.randn(M, K, device="cuda").to(H16)
    b = (torch.randn(N, K, device="cuda") * 0.05).to(H16)
    r = torch.randn(M, N, device="cuda")
    ref = a.float() @ b.float().t()
    lib.fvqa_gemm_debug_quad(0)
    p16 = ops.gemm_nt(a, b); p32 = ops.gemm_nt(a, b, residual=r, out_fp32=True)
    lib.fvqa_gemm_debug_quad(2)
    q16 = ops.gemm_nt(a, b); q32 = ops.gemm_nt(a, b, residual=r, out_fp32=True)
    torch.cuda.synchronize()
    e16 = float((q16.float() - ref).norm() / ref.norm()); e32 = float((q32 - ref - r).norm() / (ref + r).norm())
    print(f"{M}x{N}x{K}: quad relerr bf16 {e16:.2e} f32+res {e32:.2e}; identical to pair kernel: {torch.equal(p16, q16)} {torch.equal(p32, q32)}", flush=True)
    assert e16 < 5e-3 and e32 < 2e-4
    if M >= 2000:
        c16 = torch.empty(M, N, device="cuda", dtype=H16); c32 = torch.empty(M, N, device="cuda")
        res = []
        for mode in (0, 2):
            lib.fvqa_gemm_debug_quad(mode)
            t16 = timeit(lambda: ops.gemm_nt(a, b, out=c16)); t32 = timeit(lambda: ops.gemm_nt(a, b, out=c32, residual=r, out_fp32=True))
            res.append(f"{'quad' if mode else 'pair'}: bf16 {t16:6.1f} us ({2.0*M*N*K/t16/1e6:5.0f} TF/s) f32+res {t32:6.1f} us")
        print("   " + " | ".join(res), flush=True)
lib.fvqa_gemm_debug_quad(1)
print("ok")
