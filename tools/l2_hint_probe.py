"""Does an L2 eviction hint on the pair GEMM's TMA loads help? Step time and sustained GEMM rate with hints off / on."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch, pynvml
from flipped_vqa_b200 import _lib, ops
from flipped_vqa_b200._lib import H16
from gemm_sustained import sustained
pynvml.nvmlInit()
lib = _lib.lib()
for (M, N, K) in [(3072, 4096, 22016), (3072, 4096, 12288), (3072, 4096, 11008), (3072, 22016, 4096)]:
    a = torch.randn(M, K, device="cuda").to(H16)
    b = (torch.randn(N, K, device="cuda") * 0.05).to(H16)
    c = torch.empty(M, N, device="cuda", dtype=H16)
    out = []
    for on in (0, 1, 0, 1):
        lib.fvqa_gemm_debug_l2_hints(on)
        tf, pw, ck = sustained(lambda: ops.gemm_nt(a, b, out=c), 2.0 * M * N * K, secs=1.5)
        out.append(f"hints={on}: {tf:.0f} TF/s {ck:.0f} MHz")
    lib.fvqa_gemm_debug_l2_hints(0)
    print(f"{M}x{N}x{K}: " + " | ".join(out), flush=True)
