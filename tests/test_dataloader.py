"""Input pipeline against golden vectors produced by the UNMODIFIED reference
(`oracle/make_golden_collate.py`: llama/tokenizer.py prompt builders, dataloader/base_dataset.py `_get_text_token`,
dataloader/__init__.py `batch_collate`) on the same synthetic texts and hash tokenizer."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from flipped_vqa_b200 import dataloader as D
from flipped_vqa_b200.llama.tokenizer import Tokenizer
from flipped_vqa_b200.synthetic import hash_tokenizer, synthetic_qa_texts

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "collate_small.npz")
CASES = [("train", False), ("val", False), ("train", True), ("val", True)]
N, S, F = 3, 96, 10


def _items(ci, split, gen):
    tok = hash_tokenizer(Tokenizer, is_generation_task=gen)
    samples, mapping = synthetic_qa_texts(N, seed=10 + ci)
    items = []
    for i, smp in enumerate(samples):
        t = D.encode_sample(tok, smp["text"], smp["answer"], mapping, split, S, F, options=smp["options"])
        g = torch.Generator().manual_seed(100 * ci + i)
        items.append({"vid": f"v{i}", "video": torch.randn(F, 768, generator=g), "video_len": F, "text": smp["text"], "qid": i,
                      "answer": smp["answer"], "qtype": i % 3, **t})
    return items


@pytest.mark.parametrize("ci", range(len(CASES)))
def test_prompts_labels_and_collate_match_reference(ci):
    g = np.load(GOLDEN)
    split, gen = CASES[ci]
    batch = D.batch_collate(_items(ci, split, gen))
    k = f"c{ci}"
    for name in ("text_id", "label", "video_index", "label_mask"):
        for t in ("vqa", "vaq", "qav"):
            ref = g[f"{k}/{name}/{t}"]
            got = batch[name][t].numpy()
            assert got.shape == ref.shape and got.dtype == ref.dtype, (name, t, got.shape, ref.shape, got.dtype, ref.dtype)
            assert np.array_equal(got, ref), (name, t)
    for name in ("video_start", "prefix_index"):
        for t in ("vqa", "vaq", "qav"):
            assert list(batch[name][t]) == g[f"{k}/{name}/{t}"].tolist(), (name, t)
    assert np.allclose(batch["video"].sum((1, 2)).numpy(), g[f"{k}/video_sum"])
    assert np.array_equal(batch["video_len"].numpy(), g[f"{k}/video_len"])
    assert np.array_equal(batch["answer"].numpy(), g[f"{k}/answer"]) and np.array_equal(batch["qtype"].numpy(), g[f"{k}/qtype"])
    n_opt = 1 if split == "train" else 5
    assert batch["text_id"]["vqa"].shape == (N, n_opt, S)


def test_padding_truncates_and_qav_labels_clip():
    ids = D.pad_text_ids([[5, 6, 7], list(range(1, 20))], 8)
    assert ids.tolist() == [[5, 6, 7, -1, -1, -1, -1, -1], list(range(1, 9))]
    t = D.build_text_tensors({"vqa": [[1, 9, 9, 4, 2]], "vaq": [[1, 9, 4, 4, 2]], "qav": [[1, 3, 3, -2, -2, -2, 2]]},
                             {"vqa": 3, "vaq": 2, "qav": 5}, {"vqa": 1, "vaq": 1}, max_seq_len=8, max_feats=10)
    assert t["label"]["qav"][0].tolist() == [-1, -1, -1, -1, -1, 0, 1, 2]          # clipped to the sequence (base_dataset.py:84-91)
    assert t["label"]["vqa"][0].tolist() == [0, 0, 0, 4, 2, 0, 0, 0]
    assert t["text_id"]["qav"][0].tolist() == [1, 3, 3, 0, 0, 0, 2, 0]              # placeholders and padding -> 0


def test_planned_loader_feeds_plans_in_order():
    """PlannedLoader on a CPU-only host: same batches, same order, plans built ahead by the worker thread."""
    from flipped_vqa_b200.llama import ModelArgs, SyntheticTokenizer, Transformer
    from flipped_vqa_b200.synthetic import synthetic_batch
    from tests.util_parity import make_args
    params = ModelArgs(dim=64, n_layers=1, n_heads=1, vocab_size=128, multiple_of=64, norm_eps=1e-6, max_batch_size=8, max_seq_len=64,
                       adapter_len=10, adapter_layer=1)
    model = Transformer(params, make_args(), tokenizer=SyntheticTokenizer(128), device="cpu")
    batches = [synthetic_batch(2, 64, 128, seed=i) for i in range(5)]
    got = list(D.PlannedLoader(batches, model, depth=2))
    assert len(got) == 5
    for (data, plan), ref in zip(got, batches):
        assert data is ref and plan.T == 3 * 2 * 64 and plan.ce_total > 0
    val = [synthetic_batch(2, 64, 128, seed=10 + i, n_options=5) for i in range(3)]
    plans = [p for _, p in D.PlannedLoader(val, model, inference=True)]
    assert [p.n_opt for p in plans] == [5, 5, 5] and all(p.T_c < p.T for p in plans)

    def boom():
        yield batches[0]
        raise RuntimeError("loader failed")
    with pytest.raises(RuntimeError, match="loader failed"):
        list(D.PlannedLoader(boom(), model))


def test_planned_loader_worker_stops_when_the_consumer_leaves():
    """An abandoned iteration (`args.debug` break in `engine.py:52/81`, an exception, the non-finite-loss exit) must not leave a
    worker thread blocked on a full queue holding pinned slots and device plans."""
    import threading
    import time
    from flipped_vqa_b200.dataloader import PlannedLoader

    class _Model:
        _device = "cpu"

        def plan_batch(self, d):
            return ("plan", d)

        def plan_options(self, d):
            return ("options", d)

    before = threading.active_count()
    for i, (d, p) in enumerate(PlannedLoader(list(range(100)), _Model(), depth=2)):
        assert p == ("plan", d)
        if i == 3:
            break
    deadline = time.time() + 5
    while threading.active_count() > before and time.time() < deadline:
        time.sleep(0.05)
    assert threading.active_count() == before
    assert [d for d, _ in PlannedLoader([1, 2, 3], _Model(), inference=True)] == [1, 2, 3]

    class _Boom(list):
        def __iter__(self):
            yield 1
            raise RuntimeError("loader failed")

    import pytest
    with pytest.raises(RuntimeError, match="loader failed"):
        list(PlannedLoader(_Boom(), _Model()))


@pytest.mark.parametrize("ci,split", [(0, "train"), (1, "val")])
def test_sub_dialogue_prompts_and_overflow_match_reference(ci, split):
    """`--sub` (BASELINE.json configs[4]: TVQA with subtitles): `encode_dvqa / encode_dvaq / encode_dqav` (`llama/tokenizer.py:218-302`)
    and the dialogue-aware overflow handling + labels of `TVQA._get_text_token` (`dataloader/tvqa.py:75-160`) against golden vectors
    from the unmodified reference: samples with a short dialogue, a dialogue that overflows max_seq_len, and no dialogue."""
    from flipped_vqa_b200.synthetic import synthetic_dialogue_texts
    g = np.load(GOLDEN)
    tok = hash_tokenizer(Tokenizer)
    samples, mapping = synthetic_dialogue_texts(4, seed=40 + ci)
    overflowed = 0
    for i, smp in enumerate(samples):
        t = D.encode_sample_sub(tok, smp["text"], smp["answer"], mapping, split, 128, F)
        for name in ("text_id", "label", "video_index", "label_mask"):
            for task in ("vqa", "vaq", "qav"):
                ref = g[f"sub{ci}/{i}/{name}/{task}"]
                got = t[name][task].numpy()
                assert got.shape == ref.shape and got.dtype == ref.dtype and np.array_equal(got, ref), (i, name, task)
        assert [t["video_start"][k] for k in ("vqa", "vaq", "qav")] == g[f"sub{ci}/{i}/video_start"].tolist()
        overflowed += int(len(tok.encode_dvqa(text=smp["text"], max_feats=F, split=split, answer_mapping=mapping, answer=smp["answer"])[0][0]) > 128)
    assert overflowed >= 1                       # the dialogue-cutting branch really ran
