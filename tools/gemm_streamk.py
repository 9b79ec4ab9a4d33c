"""Stream-K tail of the CTA-pair GEMM vs the data-parallel schedule on the step's N = 4096 GEMMs (each with the epilogue it
has in the step): burst (CUDA events, 30 launches after a pause) and sustained (2.5 s back to back, power-capped)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flipped_vqa_b200 import _lib, ops
import pynvml
from tools.gemm_sustained import sustained
from tools.gemm_step_shapes import timeit


def main():
    pynvml.nvmlInit()
    lib = _lib.lib()
    ops.ensure_gemm_workspace()
    do_sustained = "--burst-only" not in sys.argv
    shapes = [(3072, 4096, 4096, True), (3072, 4096, 4096, False), (3072, 4096, 11008, True), (3072, 4096, 22016, False),
              (3072, 4096, 12288, False), (1950, 4096, 4096, True), (3072, 5120, 5120, True)]
    for (M, N, K, f32) in shapes:
        a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
        b = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
        r = torch.randn(M, N, device="cuda") if f32 else None
        c = torch.empty(M, N, device="cuda", dtype=torch.float32 if f32 else torch.bfloat16)
        fl = 2.0 * M * N * K
        fn = lambda: ops.gemm_nt(a, b, out=c, residual=r, out_fp32=f32)
        res = []
        for name, sk in (("dp", 0), ("streamk", 1)):
            lib.fvqa_gemm_debug_stream_k(sk)
            us = timeit(fn)
            res.append(f"{name}: burst {us:6.1f} us {fl / us / 1e6:5.0f} TF/s")
            if do_sustained:
                tf, pw, ck = sustained(fn, fl)
                res[-1] += f", sustained {tf:5.0f} TF/s {pw:4.0f} W {ck:4.0f} MHz {pw / tf * 1e3:4.0f} mJ/TF"
                time.sleep(1.0)
        lib.fvqa_gemm_debug_stream_k(1)
        print(f"{M}x{N}x{K} {'f32+res' if f32 else 'bf16   '}: " + " | ".join(res), flush=True)


if __name__ == "__main__":
    main()
