"""Sustained (power-capped) throughput of the tcgen05 GEMMs vs cuBLAS: each shape back to back for ~2.5 s,
with NVML power / SM clock sampled over the last second. Energy per FLOP = power / throughput."""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flipped_vqa_b200 import _lib, ops
from flipped_vqa_b200._lib import H16
import pynvml


def sample(h, stop, out):
    while not stop.is_set():
        out.append((pynvml.nvmlDeviceGetPowerUsage(h) / 1e3, pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
        time.sleep(0.05)


def sustained(fn, flops, secs=2.5):
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    t_end = time.time() + secs
    stop, samples = threading.Event(), []
    th = threading.Thread(target=sample, args=(h, stop, samples)); th.start()
    n_total, t0 = 0, time.time()
    last = None
    while time.time() < t_end:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            fn()
        e1.record()
        torch.cuda.synchronize()
        last = e0.elapsed_time(e1) / 50
        n_total += 50
    stop.set(); th.join()
    tail = samples[len(samples) // 2:]
    pw = sum(s[0] for s in tail) / len(tail); ck = sum(s[1] for s in tail) / len(tail)
    return flops / last / 1e9, pw, ck


def main():
    pynvml.nvmlInit()
    lib = _lib.lib()
    lib.fvqa_gemm_debug_quad(0)                                  # "pair" below = the CTA-pair kernel
    shapes = [(3072, 4096, 11008), (3072, 4096, 22016), (3072, 4096, 4096), (3072, 12288, 4096), (3072, 22016, 4096), (8192, 8192, 8192)]
    if len(sys.argv) > 1:
        shapes = shapes[:int(sys.argv[1])]
    for (M, N, K) in shapes:
        a = torch.randn(M, K, device="cuda").to(H16)
        b = (torch.randn(N, K, device="cuda") * 0.05).to(H16)
        c = torch.empty(M, N, device="cuda", dtype=H16)
        fl = 2.0 * M * N * K
        res = []
        for name, fn in [("pair", lambda: ops.gemm_nt(a, b, out=c)), ("cublas", lambda: torch.matmul(a, b.t(), out=c))]:
            tf, pw, ck = sustained(fn, fl)
            res.append(f"{name}: {tf:6.0f} TF/s {pw:5.0f} W {ck:5.0f} MHz {pw / tf * 1e3:5.0f} mJ/TFLOP")
            time.sleep(1.0)
        if N % 512 == 0 and lib.fvqa_gemm_quad_clusters() > 0:
            lib.fvqa_gemm_debug_quad(2)
            tf, pw, ck = sustained(lambda: ops.gemm_nt(a, b, out=c), fl)
            lib.fvqa_gemm_debug_quad(0)
            res.append(f"quad: {tf:6.0f} TF/s {pw:5.0f} W {ck:5.0f} MHz {pw / tf * 1e3:5.0f} mJ/TFLOP")
            time.sleep(1.0)
        lib.fvqa_gemm_debug_force_bn(-1)
        tf, pw, ck = sustained(lambda: ops.gemm_nt(a, b, out=c), fl)
        lib.fvqa_gemm_debug_force_bn(0)
        res.append(f"1cta: {tf:6.0f} TF/s {pw:5.0f} W {ck:5.0f} MHz {pw / tf * 1e3:5.0f} mJ/TFLOP")
        print(f"{M}x{N}x{K}: " + " | ".join(res), flush=True)


if __name__ == "__main__":
    main()
