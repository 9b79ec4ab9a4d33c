"""Same-box A/B of step-level switches: every variant runs interleaved in ONE process on ONE GPU (box-to-box variation
of the power-capped step is ~1.5 %, larger than most effects)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as B
from flipped_vqa_b200 import _lib, ops
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
    lib = _lib.lib()
    opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-4, betas=(0.9, 0.95), weight_decay=0.05, fused=True)
    plans = [model.plan_batch(synthetic_batch(cfg["bsz"], cfg["seqlen"], cfg["vocab_size"], max_feats=B.MAX_FEATS, seed=i)) for i in range(2)]

    def step(i):
        vqa, vaq, qav = model.forward_plan(plans[i % 2])
        (vqa + vaq + qav).backward()
        opt.step(); opt.zero_grad(set_to_none=True)

    def reset():
        model._engine.prune_last_layer = False
        ops.FUSE_SWIGLU = True
        lib.fvqa_attn_debug_use_tc(1); lib.fvqa_gemm_debug_force_bn(0); lib.fvqa_gemm_debug_l2_hints(0); lib.fvqa_gemm_debug_quad(1)

    variants = {
        "default": lambda: None,
        "prune_last_layer": lambda: setattr(model._engine, "prune_last_layer", True),
        "unfused_swiglu": lambda: setattr(ops, "FUSE_SWIGLU", False),
        "mma_sync_attention": lambda: lib.fvqa_attn_debug_use_tc(0),
        "bn240": lambda: lib.fvqa_gemm_debug_force_bn(240),
        "single_cta_gemm": lambda: lib.fvqa_gemm_debug_force_bn(-1),
        "no_quad_cluster_gemm": lambda: lib.fvqa_gemm_debug_quad(1),
        "quad_cluster_gemm_everywhere": lambda: lib.fvqa_gemm_debug_quad(2),
    }
    only = os.environ.get("AB_VARIANTS")
    if only:
        variants = {k: v for k, v in variants.items() if k in only.split(",")}
    for i in range(10):
        step(i)
    torch.cuda.synchronize()
    res = {k: [] for k in variants}
    for rnd in range(3):
        for k, setup in variants.items():
            reset(); setup()
            for i in range(3):
                step(i)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(12):
                step(i)
            e1.record(); torch.cuda.synchronize()
            res[k].append(e0.elapsed_time(e1) / 12)
    reset()
    base = sum(res["default"]) / len(res["default"])
    for k, v in res.items():
        m = sum(v) / len(v)
        print(f"{name} {k:22s} {m:7.2f} ms/step ({(m / base - 1) * 100:+5.1f} %)  rounds: " + " ".join(f"{x:.2f}" for x in v), flush=True)


if __name__ == "__main__":
    main()
