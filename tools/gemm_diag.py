"""GPU diagnostic for the tcgen05 GEMMs (not a test): correctness of the CTA-pair kernel for every tile
width, then timing of the 7B/13B layer shapes per tile width against the heuristic's choice, the
single-CTA kernel and cuBLAS (torch.matmul)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flipped_vqa_b200 import _lib, ops
from flipped_vqa_b200._lib import H16


def force(bn):
    return _lib.lib().fvqa_gemm_debug_force_bn(bn)


def check(M, N, K, bn, out_fp32=True, residual=False):
    torch.manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda").to(H16)
    b = (torch.randn(N, K, device="cuda") * 0.05).to(H16)
    ref = a.float() @ b.float().t()
    r = torch.randn(M, N, device="cuda") if residual else None
    if residual:
        ref = ref + r
    force(bn)
    try:
        c = ops.gemm_nt(a, b, out_fp32=out_fp32, residual=(r if out_fp32 or r is None else r.to(H16)))
        torch.cuda.synchronize()
    finally:
        force(0)
    rel = float((c.float() - ref).norm() / ref.norm())
    return rel, int(torch.isnan(c.float()).sum())


def time_gemm(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    _lib.lib()
    mode = sys.argv[1] if len(sys.argv) > 1 else "all"
    if mode in ("all", "check"):
        bad = 0
        for (M, N, K) in [(256, 256, 64), (256, 512, 256), (200, 384, 128), (384, 1536, 256), (1950, 4096, 512), (3072, 4096, 4096), (180, 32000, 256)]:
            for bn in (-1, 64, 128, 144, 176, 192, 240, 256):
                if bn > 0 and M <= 128:
                    continue
                rel, nan = check(M, N, K, bn, out_fp32=True, residual=(bn in (176, 256)))
                tol = 2e-4
                flag = "" if (rel < tol and nan == 0) else "   <<<<<< BAD"
                bad += bool(flag)
                print(f"CHECK {M}x{N}x{K} bn={bn:4d}: relerr {rel:.3e} nan {nan}{flag}", flush=True)
        print("check failures:", bad, flush=True)
    if mode in ("all", "time"):
        shapes = [(3072, 12288, 4096), (3072, 4096, 4096), (3072, 22016, 4096), (3072, 4096, 11008), (3072, 11008, 4096),
                  (3072, 4096, 22016), (3072, 4096, 12288), (3072, 5120, 5120), (3072, 5120, 13824), (2304, 4096, 11008), (1950, 4096, 11008)]
        for (M, N, K) in shapes:
            a = torch.randn(M, K, device="cuda").to(H16)
            b = (torch.randn(N, K, device="cuda") * 0.05).to(H16)
            c = torch.empty(M, N, device="cuda", dtype=H16)
            fl = 2.0 * M * N * K
            res = []
            for bn in (0, -1, 256, 240, 224, 208, 192, 176, 144, 128):
                force(bn)
                ms = time_gemm(lambda: ops.gemm_nt(a, b, out=c))
                res.append((bn, fl / ms / 1e9))
            force(0)
            ms2 = time_gemm(lambda: torch.matmul(a, b.t(), out=c))
            txt = " ".join(f"{('auto' if bn == 0 else '1cta' if bn < 0 else bn)}:{tf:.0f}" for bn, tf in res)
            print(f"TIME {M}x{N}x{K}: {txt} | cublas {fl / ms2 / 1e9:.0f}  (TFLOP/s)", flush=True)


def skinny():
    d = 4096
    for (M, N, K, f32, ldb) in [(10, 2 * d, d, False, d), (10, d, 2 * d, True, 3 * d), (10, 2 * 5120, 5120, False, 5120)]:
        a = torch.randn(M, K, device="cuda").to(H16)
        wfull = (torch.randn(N, ldb, device="cuda") * 0.05).to(H16)
        b = wfull[:, ldb - K:]
        ref = a.float() @ b.float().t()
        res = {}
        for bn in (0, -1):
            force(bn)
            c = ops.gemm_nt(a, b, out_fp32=f32)
            torch.cuda.synchronize()
            rel = float((c.float() - ref).norm() / ref.norm())
            ms = time_gemm(lambda: ops.gemm_nt(a, b, out_fp32=f32), iters=50)
            res[bn] = (rel, ms)
        force(0)
        print(f"SKINNY {M}x{N}x{K} ldb={ldb}: skinny rel {res[0][0]:.2e} {res[0][1] * 1e3:.1f} us | tcgen05 1cta rel {res[-1][0]:.2e} {res[-1][1] * 1e3:.1f} us", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "skinny":
        _lib.lib(); skinny(); sys.exit(0)
    main()
