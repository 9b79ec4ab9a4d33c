"""Grouped skinny GEMM (adapter projections of all layers): achieved HBM bandwidth per column-block width nt."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flipped_vqa_b200 import _lib, ops
from flipped_vqa_b200._lib import H16
from tools.gemm_step_shapes import timeit

lib = _lib.lib()
L, A, d = 32, 10, 4096
w = [(torch.randn(3 * d, d, device="cuda") * 0.02).to(H16) for _ in range(L)]
wt = [x.t().contiguous() for x in w]
kv = torch.tensor([x[d:].data_ptr() for x in w], dtype=torch.int64, device="cuda")
kvt = torch.tensor([x[:, d:].data_ptr() for x in wt], dtype=torch.int64, device="cuda")
a = torch.randn(L, A, d, device="cuda").to(H16)
da = torch.randn(L, A, 2 * d, device="cuda").to(H16)
out = torch.empty(L, A, 2 * d, device="cuda", dtype=H16)
g = torch.empty(L, A, d, device="cuda")
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for nt in (0, 1, 2, 4):
    lib.fvqa_gemm_debug_skinny_nt(nt)
    def fwd():
        flush.zero_()                               # the weights (2.1 GB) exceed the L2 anyway; keep the output / A cold too
        ops.gemm_skinny_grouped(a, kv, d, 2 * d, out)
    def bwd():
        flush.zero_()
        ops.gemm_skinny_grouped(da[:8], kvt[:8], 3 * d, d, g[:8])
    def base():
        flush.zero_()
    tb = timeit(base)
    tf, tw = timeit(fwd) - tb, timeit(bwd) - tb
    print(f"nt={nt}: forward all 32 layers {tf:7.1f} us = {L * 2 * d * d * 2 / tf / 1e6:5.2f} TB/s | backward chunk of 8 {tw:7.1f} us = {8 * 2 * d * d * 2 / tw / 1e6:5.2f} TB/s", flush=True)
lib.fvqa_gemm_debug_skinny_nt(0)
