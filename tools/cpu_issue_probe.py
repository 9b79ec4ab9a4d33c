"""Is the step GPU-bound or host-bound? Host time to ENQUEUE one step (no sync) vs device time per step."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as B
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

    for i in range(5):
        step(i)
    torch.cuda.synchronize()
    n = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(n):
        step(i)
    e1.record()
    t_issue = time.perf_counter() - t0
    torch.cuda.synchronize()
    t_total = time.perf_counter() - t0
    print(f"{name}: host enqueue {t_issue / n * 1e3:.2f} ms/step, device {e0.elapsed_time(e1) / n:.2f} ms/step, wall {t_total / n * 1e3:.2f} ms/step")
    # isolated single steps: host time with an idle GPU queue
    ts = []
    for i in range(5):
        torch.cuda.synchronize()
        t0 = time.perf_counter(); step(i); ts.append(time.perf_counter() - t0)
    print("   host enqueue of a single step (GPU queue empty):", " ".join(f"{t * 1e3:.1f}" for t in ts), "ms")


if __name__ == "__main__":
    main()
