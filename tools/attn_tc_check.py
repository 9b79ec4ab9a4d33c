"""GPU check of the tcgen05 attention kernels against the mma.sync kernels (same C ABI, debug hook)
and timing of both at the 7B NExT-QA shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flipped_vqa_b200 import ops, _lib
from flipped_vqa_b200._lib import H16


def case(n_seq, S, H, A=10, F=10, seed=0):
    hd = 128
    D = H * hd
    g = torch.Generator(device="cuda").manual_seed(seed)
    qkv = torch.randn(n_seq * S, 3 * D, device="cuda", generator=g).to(H16)
    akv = torch.randn(A, 2 * D, device="cuda", generator=g).to(H16)
    gate1 = torch.randn(H, device="cuda", generator=g) * 0.5
    gate2 = torch.full((H,), -3.5, device="cuda") + 0.1 * torch.randn(H, device="cuda", generator=g)
    ang = torch.outer(torch.arange(S).float(), 1.0 / (10000.0 ** (torch.arange(0, hd, 2).float() / hd)))
    cos, sin = torch.cos(ang), torch.sin(ang)
    cos, sin = cos.cuda().contiguous(), sin.cuda().contiguous()
    vs = [min(18, max(S - 12, 0)) if i % 3 != 2 else -1 for i in range(n_seq)]
    vstart = torch.tensor(vs, dtype=torch.int32, device="cuda")
    dout = torch.randn(n_seq * S, D, device="cuda", generator=g).to(H16)
    return dict(qkv=qkv, akv=akv, gate1=gate1, gate2=gate2, cos=cos, sin=sin, vstart=vstart, dout=dout, n_seq=n_seq, S=S, H=H, hd=hd, A=A, F=F)


def run(c, tc, bwd=True):
    lib = _lib.lib()
    prev = lib.fvqa_attn_debug_use_tc(1 if tc else 0)
    try:
        out, lse = ops.attn_fwd(c["qkv"], c["akv"], c["cos"], c["sin"], c["gate1"], c["gate2"], c["vstart"], c["n_seq"], c["S"], c["H"], c["hd"], c["A"], c["F"])
        res = [out, lse]
        if bwd:
            res += list(ops.attn_bwd(c["qkv"], c["akv"], c["cos"], c["sin"], c["gate1"], c["gate2"], c["vstart"], out, lse, c["dout"],
                                     c["n_seq"], c["S"], c["H"], c["hd"], c["A"], c["F"]))
        torch.cuda.synchronize()
    finally:
        lib.fvqa_attn_debug_use_tc(prev)
    return res


def rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / (b.norm() + 1e-30))


def main():
    bwd = "--fwd-only" not in sys.argv
    names = ["out", "lse", "dqkv", "dakv", "dgate1", "dgate2"]
    long_mode = "--long" in sys.argv
    shapes = [(2, 130, 2), (2, 200, 3), (2, 256, 2), (2, 384, 4), (3, 650, 32)] if long_mode else [(3, 128, 2), (2, 48, 2), (5, 100, 3), (24, 128, 32)]
    for (n_seq, S, H) in shapes:
        c = case(n_seq, S, H)
        ref = run(c, False, bwd)
        got = run(c, True, bwd)
        txt = " ".join(f"{nm}:{rel(g, r):.2e}" for nm, g, r in zip(names, got, ref))
        nan = sum(int(torch.isnan(g.float()).sum()) for g in got)
        print(f"n_seq={n_seq} S={S} H={H}: tc vs mma.sync  {txt}  nan={nan}", flush=True)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    lib = _lib.lib()
    for (tn, tS) in ([(6, 384), (3, 650)] if long_mode else [(24, 128)]):
      c = case(tn, tS, 32)
      for tc in (0, 1):
          lib.fvqa_attn_debug_use_tc(tc)
          tf = tb = 0.0
          for it in range(13):
              flush.zero_()
              e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
              e[0].record()
              out, lse = ops.attn_fwd(c["qkv"], c["akv"], c["cos"], c["sin"], c["gate1"], c["gate2"], c["vstart"], tn, tS, 32, 128, 10, 10)
              e[1].record()
              if bwd:
                  ops.attn_bwd(c["qkv"], c["akv"], c["cos"], c["sin"], c["gate1"], c["gate2"], c["vstart"], out, lse, c["dout"], tn, tS, 32, 128, 10, 10)
              e[2].record()
              torch.cuda.synchronize()
              if it >= 3:
                  tf += e[0].elapsed_time(e[1]); tb += e[1].elapsed_time(e[2])
          print(f"n_seq={tn} S={tS} tc={tc}: fwd {tf / 10 * 1e3:.1f} us  bwd {tb / 10 * 1e3:.1f} us", flush=True)
    lib.fvqa_attn_debug_use_tc(1)


if __name__ == "__main__":
    main()
