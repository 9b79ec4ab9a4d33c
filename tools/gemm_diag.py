"""GPU diagnostic for the tcgen05 GEMM: error structure per shape + quick timing (not a test)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flipped_vqa_b200 import ops, _lib

def main():
    _lib.lib()
    torch.manual_seed(0)
    for (M, N, K) in [(128, 256, 64), (128, 256, 128), (128, 256, 256), (128, 128, 64), (256, 512, 256), (200, 384, 128), (3072, 4096, 4096)]:
        a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
        b = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
        ref = a.float() @ b.float().t()
        try:
            c = ops.gemm_nt(a, b, out_fp32=True)
            torch.cuda.synchronize()
        except Exception as e:
            print("GEMM", M, N, K, "EXC", e); continue
        err = (c - ref).abs()
        rel = float((c - ref).norm() / ref.norm())
        print(f"GEMM {M}x{N}x{K}: relerr {rel:.3e} maxabs {float(err.max()):.3e} nan {int(torch.isnan(c).sum())}")
        if rel > 1e-3:
            bad = err > (1e-2 * ref.abs().max())
            rows = bad.any(1).nonzero().flatten()
            cols = bad.any(0).nonzero().flatten()
            print("  bad rows:", rows[:16].tolist(), "... n=", rows.numel(), " bad cols:", cols[:16].tolist(), "... n=", cols.numel())
            print("  c[0,:8]", c[0, :8].tolist()); print("  r[0,:8]", ref[0, :8].tolist())
            # is it a K-permutation problem? compare with partial-K references
            for kk in range(64, K + 1, 64):
                rk = a[:, :kk].float() @ b[:, :kk].float().t()
                print(f"   vs K[:{kk}] relerr {float((c - rk).norm() / rk.norm()):.3e}")
                if kk >= 256: break
    # timing
    for (M, N, K) in [(3072, 12288, 4096), (3072, 4096, 4096), (3072, 22016, 4096), (3072, 4096, 11008), (3072, 11008, 4096), (3072, 4096, 22016), (3072, 4096, 12288)]:
        a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
        b = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
        c = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        for _ in range(3): ops.gemm_nt(a, b, out=c)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): ops.gemm_nt(a, b, out=c)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        for _ in range(3): torch.matmul(a, b.t(), out=c)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10): torch.matmul(a, b.t(), out=c)
        e1.record(); torch.cuda.synchronize()
        ms2 = e0.elapsed_time(e1) / 10
        fl = 2.0 * M * N * K
        print(f"TIME {M}x{N}x{K}: ours {ms:.3f} ms {fl / ms / 1e9:.0f} TFLOP/s | cublas {ms2:.3f} ms {fl / ms2 / 1e9:.0f} TFLOP/s")

if __name__ == "__main__":
    main()
