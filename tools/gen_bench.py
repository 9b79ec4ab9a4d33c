"""Generation-evaluator throughput (`llama/model.py:367-546`) at a bench config: bs samples x 4 options,
31 greedy steps each, through `model(data, inference=True)` with host batches.
  kv_cached_graph : the product path (prefill once + 30 single-row decode steps replayed as a CUDA graph)
  kv_cached_eager : same, decode steps launched eagerly
  full_rerun      : the reference's schedule on the same kernels - the whole stack over the whole sequence for every step of
                    every sample (`model.py:429-467`), timed on a SAMPLE of steps and scaled (31 * bs stack evaluations)
    python tools/gen_bench.py [config] [bs]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as B
from flipped_vqa_b200.llama import ModelArgs, SyntheticTokenizer, Transformer
from flipped_vqa_b200.synthetic import synthetic_batch, synthetic_generation_batch


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "7b-nextqa"
    bs = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    cfg = dict(B.CONFIGS[name], name=name)
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    params = ModelArgs(dim=cfg["dim"], n_layers=cfg["n_layers"], n_heads=cfg["n_heads"], vocab_size=cfg["vocab_size"], multiple_of=cfg["multiple_of"],
                       norm_eps=1e-6, max_batch_size=32, max_seq_len=cfg["seqlen"], adapter_len=B.ADAPTER_LEN, adapter_layer=cfg["adapter_layer"])
    args = B.make_args()
    args.is_generation_task = True
    a_tok = 22550
    model = Transformer(params, args, tokenizer=SyntheticTokenizer(cfg["vocab_size"], a_token_id=a_tok), device=dev)
    model.repack()
    batches = [synthetic_generation_batch(bs, cfg["seqlen"], cfg["vocab_size"], a_tok, max_feats=B.MAX_FEATS, seed=i, video_start=18) for i in range(2)]
    out = {"task": "generation evaluator (validation)", "config": name, "samples_per_step": bs, "greedy_steps": model.GENERATION_STEPS}

    def timed(fn, n):
        fn(0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(i)
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    toks = {}
    for mode in ("kv_cached_graph", "kv_cached_eager"):
        model._engine.decode_graph = mode == "kv_cached_graph"

        def step(i):
            ms_idx, _ = model(batches[i % 2], inference=True)
            return ms_idx.cpu()
        ms = timed(step, 5)
        model(batches[0], inference=True)
        toks[mode] = model.last_generation["tokens"].cpu().tolist()
        out[mode] = {"ms_per_batch": ms, "samples_per_s": bs / (ms * 1e-3)}
    out["graph_and_eager_tokens_agree"] = toks["kv_cached_graph"] == toks["kv_cached_eager"]
    # the reference's schedule: one dense forward over ONE sequence per step per sample (here: the product's own dense scorer on a
    # single-sequence batch = the same kernels without a KV cache), 31 * bs of them per batch
    one = synthetic_batch(1, cfg["seqlen"], cfg["vocab_size"], max_feats=B.MAX_FEATS, seed=9, n_options=1)
    model.args.is_generation_task = False
    model.share_option_prefix = False
    ms1 = timed(lambda i: model(one, inference=True), 20)
    out["full_rerun"] = {"ms_per_stack_evaluation": ms1, "ms_per_batch_scaled": ms1 * model.GENERATION_STEPS * bs,
                         "samples_per_s": bs / (ms1 * model.GENERATION_STEPS * bs * 1e-3), "note": "scaled from 20 timed single-sequence stack evaluations"}
    out["speedup_vs_full_rerun"] = out["full_rerun"]["ms_per_batch_scaled"] / out["kv_cached_graph"]["ms_per_batch"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
