"""cProfile of the host side of the training step (where do the ~40 ms of enqueue time per 7B step go?).
    python tools/host_profile.py [config] [steps]"""
import cProfile
import os
import pstats
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as B
from flipped_vqa_b200.llama import ModelArgs, SyntheticTokenizer, Transformer
from flipped_vqa_b200.synthetic import synthetic_batch


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "7b-nextqa"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 5
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
    plans = [model.plan_batch(synthetic_batch(cfg["bsz"], cfg["seqlen"], cfg["vocab_size"], max_feats=B.MAX_FEATS, seed=i)) for i in range(2)]

    e2e = "--e2e" in sys.argv                    # through model(data) with host tensors + a loss read, like bench.py's e2e leg
    batches = [synthetic_batch(cfg["bsz"], cfg["seqlen"], cfg["vocab_size"], max_feats=B.MAX_FEATS, seed=i) for i in range(2)]

    def step(i):
        vqa, vaq, qav = model(batches[i % 2]) if e2e else model.forward_plan(plans[i % 2])
        loss = vqa + vaq + qav
        loss.backward()
        opt.step(); opt.zero_grad(set_to_none=True)
        if e2e:
            loss.item()

    if "--val" in sys.argv:                      # validation: loss-based option scoring through model(data, inference=True) + predict_options
        vb = [synthetic_batch(cfg["bsz"], cfg["seqlen"], cfg["vocab_size"], max_feats=B.MAX_FEATS, seed=i, n_options=5) for i in range(2)]

        def step(i):                             # noqa: F811
            with torch.no_grad():
                tok = model(vb[i % 2], inference=True)
                return model.predict_options(tok).cpu()

    for i in range(3):
        step(i)
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    for i in range(steps):
        torch.cuda.synchronize()                 # empty queue: pure enqueue cost, never blocked on a full launch queue
        pr.enable()
        step(i)
        pr.disable()
    torch.cuda.synchronize()
    st = pstats.Stats(pr)
    st.sort_stats("tottime").print_stats(28)
    st.sort_stats("cumulative").print_stats(22)


if __name__ == "__main__":
    main()
