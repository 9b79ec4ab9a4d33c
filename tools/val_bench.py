"""Validation-path throughput: loss-based option scoring, engine.py:87-93, at the 7B NExT-QA
shape — bs items x 5 options x S tokens, VQA stream only, forward only, through `model(data, inference=True)` +
`predict_options` with host batches (pinned H2D in the timed region, prediction read back every step)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as B
from flipped_vqa_b200.llama import ModelArgs, SyntheticTokenizer, Transformer
from flipped_vqa_b200.synthetic import synthetic_batch


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "7b-nextqa"
    bs = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    cfg = dict(B.CONFIGS[name], name=name)
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    params = ModelArgs(dim=cfg["dim"], n_layers=cfg["n_layers"], n_heads=cfg["n_heads"], vocab_size=cfg["vocab_size"], multiple_of=cfg["multiple_of"],
                       norm_eps=1e-6, max_batch_size=32, max_seq_len=cfg["seqlen"], adapter_len=B.ADAPTER_LEN, adapter_layer=cfg["adapter_layer"])
    model = Transformer(params, B.make_args(), tokenizer=SyntheticTokenizer(cfg["vocab_size"]), device=dev)
    model.repack()
    batches = [synthetic_batch(bs, cfg["seqlen"], cfg["vocab_size"], max_feats=B.MAX_FEATS, seed=i, n_options=5) for i in range(3)]

    def step(i):
        tok = model(batches[i % 3], inference=True)
        return model.predict_options(tok).cpu()

    d, L, S, hid = cfg["dim"], cfg["adapter_layer"], cfg["seqlen"], 11008 if cfg["dim"] == 4096 else 13824
    out = {"task": "option scoring (validation)", "config": name, "items_per_step": bs, "options": 5}
    preds = {}
    for mode in ("dense", "shared_prefix"):
        model.share_option_prefix = mode == "shared_prefix"
        preds[mode] = [step(i).tolist() for i in range(3)]          # warm-up; also the agreement check
        rows = model.last_plan.T_c if mode == "shared_prefix" else model.last_plan.T
        torch.cuda.synchronize()
        n = 10
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            step(i)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        # FLOPs actually executed: row-wise GEMMs on `rows`, attention on the full [bs*5, S] layout
        flops = L * 2 * rows * (4 * d * d + 3 * d * hid) + L * 4 * bs * 5 * d * (S * (S + 1) / 2 + S * 10)
        out[mode] = {"ms_per_step": ms, "items_per_s": bs / (ms * 1e-3), "sequences_per_s": bs * 5 / (ms * 1e-3),
                     "gemm_rows_per_step": rows, "tflops_executed": flops / (ms * 1e-3) / 1e12}
    out["predictions_agree"] = preds["dense"] == preds["shared_prefix"]
    out["speedup"] = out["dense"]["ms_per_step"] / out["shared_prefix"]["ms_per_step"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
