"""GPU diagnostic (checker): gate1/gate2 gradients of the S=650 narrow-model case from the fp32 oracle, the bf16 oracle
(= the reference's op sequence under autograd in bf16), the mma.sync attention path and the tcgen05 path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from types import SimpleNamespace
import torch
from oracle import llama_vqa_oracle as O
from flipped_vqa_b200 import _lib
from flipped_vqa_b200.synthetic import synthetic_batch, synthetic_state_dict
from tests.util_parity import build_product_model, make_args, product_grads, rel_l2


def main():
    seeds = [13, 14, 15] if len(sys.argv) < 2 else [int(x) for x in sys.argv[1:]]
    pd = dict(dim=256, n_layers=2, n_heads=2, vocab_size=512, multiple_of=256, norm_eps=1e-6, max_batch_size=32,
              max_seq_len=650, adapter_len=10, adapter_layer=2)
    args = make_args()
    lib = _lib.lib()
    for seed in seeds:
        sd = synthetic_state_dict(SimpleNamespace(**pd), seed=seed, max_feats=args.max_feats, bias=args.bias)
        data = synthetic_batch(1, 650, 512, max_feats=args.max_feats, seed=seed, video_start=18, full_length=True)
        res = {}
        for name, dt in (("oracle_fp32", torch.float32), ("oracle_bf16", torch.bfloat16)):
            st = O.prepare_state(sd, frozen_dtype=dt, device="cuda")
            losses = O.forward_losses(st, SimpleNamespace(**pd), data, max_feats=args.max_feats, tau=args.tau)
            sum(losses).backward()
            res[name] = {n: st[n].grad.detach().float().cpu() for n in O.trainable_names(st) if st[n].grad is not None}
        for name, tc in (("mma_sync", 0), ("tcgen05", 1)):
            lib.fvqa_attn_debug_use_tc(tc)
            model = build_product_model(pd, sd, args)
            vqa, vaq, qav = model(data)
            (vqa + vaq + qav).backward()
            torch.cuda.synchronize()
            res[name] = product_grads(model)
        lib.fvqa_attn_debug_use_tc(1)
        gold = res["oracle_fp32"]
        for grp in ("gate1", "gate2", "adapter_query.weight", "visual_proj.weight", "temporal_emb.weight"):
            names = sorted(n for n in gold if n.endswith(grp))
            b = torch.cat([gold[n].flatten() for n in names])
            txt = []
            for k in ("oracle_bf16", "mma_sync", "tcgen05"):
                a = torch.cat([res[k][n].flatten() for n in names])
                txt.append(f"{k} {rel_l2(a, b):.3e}")
            print(f"seed {seed} {grp:22s} rel L2 vs fp32 oracle: " + "  ".join(txt), flush=True)
        names = sorted(n for n in gold if n.endswith("gate1"))
        for k in ("oracle_fp32", "oracle_bf16", "mma_sync", "tcgen05"):
            print(f"   gate1 {k:12s}", [round(float(x), 5) for n in names for x in res[k][n].flatten()])


if __name__ == "__main__":
    main()
