"""Burst time of the eight GEMM launches of one 7B NExT-QA layer (forward + dX backward), each with the epilogue it
has in the step; cuBLAS (plain bf16 out, no epilogue work) beside it."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flipped_vqa_b200 import _lib, ops
from flipped_vqa_b200._lib import H16

BF = H16


MARKER = None


def timeit(fn, n=None):
    n = n or int(os.environ.get("ITERS", 30))                  # ITERS=1 WARM=1 under ncu (tools/gemm_shapes_ncu.py reads the launch list)
    MARKER.fill_(1.0)                                          # one fill kernel in front of every timed case: the segment separator of that list
    for _ in range(int(os.environ.get("WARM", 3))):
        fn()
    torch.cuda.synchronize()
    time.sleep(0.5)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def main():
    global MARKER
    lib = _lib.lib()
    MARKER = torch.empty(1, device="cuda")
    T, d, hid, S, hd = int(os.environ.get("ROWS", 3072)), 4096, 11008, 128, 128
    dev = "cuda"
    rn = lambda *s, std=1.0: (torch.randn(*s, device=dev) * std).to(BF)
    x, xh = rn(T, d), rn(T, hid)
    wqkv, wo, w13, w2 = rn(3 * d, d, std=.02), rn(d, d, std=.02), rn(2 * hid, d, std=.02), rn(d, hid, std=.02)
    wqkv_t, w13_t, w2_t = rn(d, 3 * d, std=.02), rn(d, 2 * hid, std=.02), rn(hid, d, std=.02)
    res = torch.randn(T, d, device=dev)
    dqkv, dg = rn(T, 3 * d), rn(T, 2 * hid)
    inv = 1.0 / (10000.0 ** (torch.arange(0, hd, 2).float() / hd))
    ang = torch.outer(torch.arange(2 * S).float(), inv)
    cos, sin = torch.cos(ang).to(dev).contiguous(), torch.sin(ang).to(dev).contiguous()
    o32 = torch.empty(T, d, device=dev)
    o16 = torch.empty(T, d, device=dev, dtype=BF)
    qkv = torch.empty(T, 3 * d, device=dev, dtype=BF)
    g = rn(T, 2 * hid); c = torch.empty(T, hid, device=dev, dtype=BF); dgo = torch.empty(T, 2 * hid, device=dev, dtype=BF)
    cases = [
        ("qkv+rope      N=12288 K=4096 ", lambda: ops.gemm_nt_rope(x, wqkv, cos, sin, 2 * d, hd, S, out=qkv), 2. * T * 3 * d * d, (x, wqkv)),
        ("wo f32+res    N=4096  K=4096 ", lambda: ops.gemm_nt(x, wo, out=o32, residual=res, out_fp32=True), 2. * T * d * d, (x, wo)),
        ("w13+swiglu    N=22016 K=4096 ", lambda: ops.gemm_swiglu_fwd(x, w13, g=g, c=c), 2. * T * 2 * hid * d, (x, w13)),
        ("w2 f32+res    N=4096  K=11008", lambda: ops.gemm_nt(xh, w2, out=o32, residual=res, out_fp32=True), 2. * T * d * hid, (xh, w2)),
        ("w2t+swiglu'   N=11008 K=4096 ", lambda: ops.gemm_swiglu_bwd(x, w2_t, g, dg=dgo), 2. * T * hid * d, (x, w2_t)),
        ("w13t dX       N=4096  K=22016", lambda: ops.gemm_nt(dg, w13_t, out=o16), 2. * T * d * 2 * hid, (dg, w13_t)),
        ("wot dX        N=4096  K=4096 ", lambda: ops.gemm_nt(x, wo, out=o16), 2. * T * d * d, (x, wo)),
        ("wqkvt dX      N=4096  K=12288", lambda: ops.gemm_nt(dqkv, wqkv_t, out=o16), 2. * T * d * 3 * d, (dqkv, wqkv_t)),
    ]
    # the four dX GEMMs again, reading the FORWARD weight as an MN-major operand (no transposed copy: the product's default)
    nn_cases = [
        ("w2+swiglu' NN N=11008 K=4096 ", lambda: ops.gemm_swiglu_bwd(x, w2, g, dg=dgo, nn=True), 2. * T * hid * d),
        ("w13 dX NN     N=4096  K=22016", lambda: ops.gemm_nn(dg, w13, out=o16), 2. * T * d * 2 * hid),
        ("wo dX NN      N=4096  K=4096 ", lambda: ops.gemm_nn(x, wo, out=o16), 2. * T * d * d),
        ("wqkv dX NN    N=4096  K=12288", lambda: ops.gemm_nn(dqkv, wqkv, out=o16), 2. * T * d * 3 * d),
    ]
    tot = {"pair": 0.0, "cublas": 0.0}
    for name, fn, fl, (a, b) in cases:
        row = []
        us = timeit(fn)
        tot["pair"] += us
        row.append(f"pair kernel {us:6.1f} us {fl / us / 1e6:5.0f} TF/s")
        cb = torch.empty(a.shape[0], b.shape[0], device=dev, dtype=BF)
        us = timeit(lambda: torch.matmul(a, b.t(), out=cb))
        tot["cublas"] += us
        row.append(f"cublas plain {us:6.1f} us {fl / us / 1e6:5.0f} TF/s")
        print(name, " | ".join(row), flush=True)
    print("layer total: " + " | ".join(f"{k}: {v:.0f} us" for k, v in tot.items()))
    nn_tot = 0.0
    for name, fn, fl in nn_cases:
        us = timeit(fn)
        nn_tot += us
        print(name, f"pair kernel, MN-major B {us:6.1f} us {fl / us / 1e6:5.0f} TF/s", flush=True)
    print(f"dX GEMMs without transposed copies: {nn_tot:.0f} us per layer")


if __name__ == "__main__":
    main()
