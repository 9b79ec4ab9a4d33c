"""ORACLE / golden generator (test infrastructure, build container only): runs the UNMODIFIED reference input pipeline
— `llama/tokenizer.py` prompt builders (`encode_vqa/vaq/qav`), `dataloader/base_dataset.py` `_get_text_token` and
`dataloader/__init__.py` `batch_collate` — on synthetic NExT-QA-style texts with a hash tokenizer standing in for
SentencePiece, and commits inputs' seeds + outputs as tests/golden/collate_small.npz.

    python oracle/make_golden_collate.py
"""
import argparse
import importlib
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flipped_vqa_b200.synthetic import hash_tokenizer, synthetic_dialogue_texts, synthetic_qa_texts  # noqa: E402

REF = os.environ.get("FVQA_REFERENCE_ROOT", "/root/reference")
CASES = [("train", False), ("val", False), ("train", True), ("val", True)]
N, S, F = 3, 96, 10
SUB_CASES, SUB_N, SUB_S = ("train", "val"), 4, 128


def main():
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)
    sys.modules.setdefault("pysrt", types.ModuleType("pysrt"))          # dataloader/tvqa.py imports it; unused here
    ref_tok = importlib.import_module("llama.tokenizer")
    ref_dl = importlib.import_module("dataloader")
    base = importlib.import_module("dataloader.base_dataset")
    out = {}
    for ci, (split, gen) in enumerate(CASES):
        tok = hash_tokenizer(ref_tok.Tokenizer, is_generation_task=gen)
        args = argparse.Namespace(max_feats=F, max_seq_len=S, debug=False)
        ds = base.BaseDataset(args, tok, split)
        samples, mapping = synthetic_qa_texts(N, seed=10 + ci)
        ds.answer_mapping = mapping
        items = []
        for i, smp in enumerate(samples):
            text_id, label, video_start, video_index, label_mask, prefix_index = ds._get_text_token(smp["text"], smp["answer"], smp["options"])
            g = torch.Generator().manual_seed(100 * ci + i)
            items.append({"vid": f"v{i}", "video": torch.randn(F, 768, generator=g), "video_len": F, "text": smp["text"], "text_id": text_id,
                          "label": label, "video_start": video_start, "video_index": video_index, "label_mask": label_mask,
                          "prefix_index": prefix_index, "qid": i, "answer": smp["answer"], "qtype": i % 3})
        batch = ref_dl.batch_collate(items)
        k = f"c{ci}"
        for name in ("text_id", "label", "video_index", "label_mask"):
            for t in ("vqa", "vaq", "qav"):
                out[f"{k}/{name}/{t}"] = batch[name][t].numpy()
        for name in ("video_start", "prefix_index"):
            for t in ("vqa", "vaq", "qav"):
                out[f"{k}/{name}/{t}"] = np.asarray(batch[name][t])
        out[f"{k}/video_sum"] = batch["video"].sum((1, 2)).numpy()       # the items are regenerated from their seeds by the test
        out[f"{k}/video_len"] = batch["video_len"].numpy()
        out[f"{k}/answer"] = batch["answer"].numpy()
        out[f"{k}/qtype"] = batch["qtype"].numpy()
    # --sub (TVQA / VLEP): dialogue prompt builders + dialogue-aware overflow handling, `dataloader/tvqa.py:75-160`
    tvqa = importlib.import_module("dataloader.tvqa")
    for ci, split in enumerate(SUB_CASES):
        tok = hash_tokenizer(ref_tok.Tokenizer)
        ds = object.__new__(tvqa.TVQA)                       # no files: only the token path is exercised
        ds.max_seq_len, ds.max_feats, ds.sub, ds.split, ds.tokenizer = SUB_S, F, True, split, tok
        samples, mapping = synthetic_dialogue_texts(SUB_N, seed=40 + ci)
        ds.answer_mapping = mapping
        for i, smp in enumerate(samples):
            text_id, label, video_start, video_index, label_mask = ds._get_text_token(smp["text"], smp["answer"])
            for name, dct in (("text_id", text_id), ("label", label), ("video_index", video_index), ("label_mask", label_mask)):
                for t in ("vqa", "vaq", "qav"):
                    out[f"sub{ci}/{i}/{name}/{t}"] = dct[t].numpy()
            out[f"sub{ci}/{i}/video_start"] = np.asarray([video_start[t] for t in ("vqa", "vaq", "qav")])
    path = os.path.join(ROOT, "tests", "golden", "collate_small.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, len(out), "arrays")


if __name__ == "__main__":
    main()
