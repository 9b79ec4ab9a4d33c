"""ncu target: one data-parallel and one stream-K launch of the 3072 x 4096 x K bf16 GEMM (K from argv, default 4096)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flipped_vqa_b200 import _lib, ops

K = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
lib = _lib.lib()
ops.ensure_gemm_workspace()
a = torch.randn(3072, K, device="cuda").to(torch.bfloat16)
b = (torch.randn(4096, K, device="cuda") * 0.05).to(torch.bfloat16)
c = torch.empty(3072, 4096, device="cuda", dtype=torch.bfloat16)
for sk in (0, 1, 0, 1):
    lib.fvqa_gemm_debug_stream_k(sk)
    ops.gemm_nt(a, b, out=c)
torch.cuda.synchronize()
print("ok")
