"""ncu target: which kernels / tile shapes cuBLAS picks for the step's GEMM shapes (plain bf16 matmul)."""
import sys
import torch
shapes = [(3072, 4096, 11008), (3072, 4096, 4096), (3072, 12288, 4096), (2535, 4096, 11008)]
for M, N, K in shapes:
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    b = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    c = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for _ in range(2):
        torch.matmul(a, b.t(), out=c)
torch.cuda.synchronize()
print("ok")
