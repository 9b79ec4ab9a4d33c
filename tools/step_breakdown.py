"""In-step device time per op kind (CUDA events around every C-ABI call, real clocks / warm L2), 7B NExT-QA step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as B
from flipped_vqa_b200 import ops
from flipped_vqa_b200.llama import ModelArgs, SyntheticTokenizer, Transformer
from flipped_vqa_b200.synthetic import synthetic_batch


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "7b-nextqa"
    cfg = dict(B.CONFIGS[name], name=name)
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    params = ModelArgs(dim=cfg["dim"], n_layers=cfg["n_layers"], n_heads=cfg["n_heads"], vocab_size=cfg["vocab_size"], multiple_of=cfg["multiple_of"],
                       norm_eps=1e-6, max_batch_size=32, max_seq_len=cfg["seqlen"], adapter_len=B.ADAPTER_LEN, adapter_layer=cfg["adapter_layer"])
    model = Transformer(params, B.make_args(), tokenizer=SyntheticTokenizer(cfg["vocab_size"]), device=dev)
    with torch.no_grad():
        for blk in model.layers:
            blk.attention.gate1.normal_(0, 0.5)
    model.repack()
    opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-4, betas=(0.9, 0.95), weight_decay=0.05, fused=True)
    plans = [model.plan_batch(synthetic_batch(cfg["bsz"], cfg["seqlen"], cfg["vocab_size"], max_feats=B.MAX_FEATS, seed=i)) for i in range(2)]

    def step(i):
        vqa, vaq, qav = model.forward_plan(plans[i % 2])
        (vqa + vaq + qav).backward()
        opt.step(); opt.zero_grad(set_to_none=True)

    for i in range(4):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 5
    e0.record()
    for i in range(n):
        step(i)
    e1.record(); torch.cuda.synchronize()
    plain = e0.elapsed_time(e1) / n
    ops.OP_TIMER = ops.OpTimer()
    e0.record()
    for i in range(n):
        step(i)
    e1.record(); torch.cuda.synchronize()
    timed = e0.elapsed_time(e1) / n
    agg = ops.OP_TIMER.summary(); ops.OP_TIMER = None
    tot = sum(v[1] for v in agg.values()) / n
    print(f"{name}: step {plain:.2f} ms untimed, {timed:.2f} ms with per-op events; sum of ops {tot:.2f} ms")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"  {v[1] / n:8.3f} ms {100 * v[1] / n / tot:5.1f}%  n={v[0] // n:4d}  avg {1e3 * v[1] / v[0]:8.1f} us  {k}")


if __name__ == "__main__":
    main()
