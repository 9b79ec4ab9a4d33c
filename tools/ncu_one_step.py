"""One training step between cudaProfilerStart / Stop, for an UNFILTERED ncu launch list (every kernel of the step, ATen's included):

    ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \\
        --csv --log-file gpurun_out/step_all.csv python tools/ncu_one_step.py [config] [--e2e]
    python tools/launch_summary.py gpurun_out/step_all.csv
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as B
from flipped_vqa_b200.llama import ModelArgs, SyntheticTokenizer, Transformer
from flipped_vqa_b200.synthetic import synthetic_batch


def main():
    name = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else "7b-nextqa"
    e2e = "--e2e" in sys.argv
    cfg = dict(B.CONFIGS[name], name=name)
    dev = torch.device("cuda", 0)
    params = ModelArgs(dim=cfg["dim"], n_layers=cfg["n_layers"], n_heads=cfg["n_heads"], vocab_size=cfg["vocab_size"], multiple_of=cfg["multiple_of"],
                       norm_eps=1e-6, max_batch_size=32, max_seq_len=cfg["seqlen"], adapter_len=B.ADAPTER_LEN, adapter_layer=cfg["adapter_layer"])
    model = Transformer(params, B.make_args(), tokenizer=SyntheticTokenizer(cfg["vocab_size"]), device=dev)
    with torch.no_grad():
        for blk in model.layers:
            blk.attention.gate1.normal_(0, 0.5)
    model.repack()
    opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-4, betas=(0.9, 0.95), weight_decay=0.05, fused=True)
    batches = [synthetic_batch(cfg["bsz"], cfg["seqlen"], cfg["vocab_size"], max_feats=B.MAX_FEATS, seed=i) for i in range(2)]
    plans = [model.plan_batch(b) for b in batches]

    def step(i):
        vqa, vaq, qav = model(batches[i % 2]) if e2e else model.forward_plan(plans[i % 2])
        loss = vqa + vaq + qav
        loss.backward()
        opt.step(); opt.zero_grad(set_to_none=True)
        if e2e:
            loss.item()

    for i in range(3):
        step(i)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    step(3)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
