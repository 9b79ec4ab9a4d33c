import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch, pynvml
from flipped_vqa_b200 import _lib, ops
from flipped_vqa_b200._lib import H16
from gemm_sustained import sustained
pynvml.nvmlInit()
lib = _lib.lib()
for (M, N, K) in [(3072, 4096, 11008), (3072, 4096, 4096), (3072, 4096, 22016)]:
    a = torch.randn(M, K, device="cuda").to(H16)
    b = (torch.randn(N, K, device="cuda") * 0.05).to(H16)
    c = torch.empty(M, N, device="cuda", dtype=H16)
    fl = 2.0 * M * N * K
    out = []
    for bn in (256, 240, 208, 176, 128):
        lib.fvqa_gemm_debug_force_bn(bn)
        tf, pw, ck = sustained(lambda: ops.gemm_nt(a, b, out=c), fl, secs=1.5)
        out.append(f"bn{bn}: {tf:.0f} TF/s {ck:.0f} MHz")
    lib.fvqa_gemm_debug_force_bn(0)
    print(f"{M}x{N}x{K}: " + " | ".join(out), flush=True)
