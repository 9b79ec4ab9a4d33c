"""SURVEY 8(f) rank 2, second half: the generation evaluator of HEAD (`llama/model.py:367-623`, consumed by `engine.py:78-85,
99-121`) through the KV-cached decoder (`StepEngine.generate`), pinned to `tests/golden/generation_small.npz` - what the UNMODIFIED
reference produced on the same weights / batch (oracle/make_golden.py::run_reference_generation, fp32-shimmed): the 31 greedy
tokens of every sample, the final sequences, the cosine similarities and the chosen options."""
import argparse
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle.make_golden import GEN, GEN_RUN, generation_inputs  # noqa: E402  (inputs only; the reference itself is not needed)
from tests.util_parity import GOLDEN_DIR, make_args  # noqa: E402


def _model():
    from flipped_vqa_b200.llama import ModelArgs, SyntheticTokenizer, Transformer
    params, sd, data = generation_inputs()
    r = GEN_RUN
    args = make_args(r["max_feats"], r["bias"], r["tau"])
    args.is_generation_task = True
    model = Transformer(ModelArgs(**GEN), args, tokenizer=SyntheticTokenizer(GEN["vocab_size"], a_token_id=r["a_token_id"]))
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not missing and not unexpected
    return model, data


def test_generation_matches_reference_golden(fvqa_lib):
    g = np.load(os.path.join(GOLDEN_DIR, "generation_small.npz"))
    model, data = _model()
    most_similar, extracted = model(data, inference=True)                      # HEAD `forward(data, inference=True)` -> inference()
    gen = model.last_generation
    assert g["margins"].min() > 0.01                                           # the fixture's arg-max decisions are not near-ties
    assert gen["tokens"].cpu().tolist() == g["tokens"].tolist()                # 3 x 31 greedy tokens, exact
    assert gen["ids"].tolist() == g["final_ids"].tolist()                      # sequences with the generated tokens written in
    assert np.abs(gen["similarities"].cpu().numpy() - g["similarities"]).max() < 5e-3
    assert most_similar.cpu().tolist() == g["most_similar"].tolist()
    assert [e["video_id"] for e in extracted] == data["vid"] and all(set(e) == {"video_id", "question", "generated_answer"} for e in extracted)
    # the decode step replayed as a CUDA graph (default) and launched eagerly give the same tokens
    assert model._engine.decode_graph
    model._engine.decode_graph = False
    model(data, inference=True)
    assert model.last_generation["tokens"].cpu().tolist() == g["tokens"].tolist()


def test_kv_cached_decode_equals_full_rerun(fvqa_lib):
    """The decoder evaluates every position once; the reference re-runs the whole stack per step. Equivalence inside the product:
    feeding the generated sequence back through the DENSE forward must reproduce every greedy choice (argmax of the dense
    logits at position p == generated token at p + 1) - checked through the per-token losses of the dense scorer."""
    model, data = _model()
    model.generate_answers(data, want_margin=True)
    gen = model.last_generation
    ids = gen["ids"]                                                           # [bsz, S] with generated tokens
    bsz, S = ids.shape
    prefix = data["prefix_index"]["vqa"]
    # dense evaluation of the final sequences: label every generated position, read the per-token loss of the generated token and
    # compare it with the loss of a few random alternatives: the generated token must be the arg-max (lowest loss)
    model.args.is_generation_task = False
    torch.manual_seed(0)
    n_alt = 6
    alt_ids = ids.view(bsz, 1, S).repeat(1, n_alt, 1)
    labels = torch.zeros(bsz, n_alt, S, dtype=torch.int64)
    for b in range(bsz):
        for t in (0, 7, 30):                                                   # first, middle and last decoding step
            p = prefix[b] + t
            labels[b, :, p] = alt_ids[b, :, p]
            alts = torch.randint(3, model.vocab_size, (n_alt - 1,))
            alts[alts == ids[b, p]] = 5
            labels[b, 1:, p] = alts                                            # same inputs, alternative TARGETS at position p
    d = dict(data)
    d["text_id"] = {"vqa": alt_ids}
    d["label"] = {"vqa": labels}
    model.share_option_prefix = False
    tok = model(d, inference=True).cpu()                                       # [bsz, n_alt, S-1] per-token CE of the labelled targets
    for b in range(bsz):
        for t in (0, 7, 30):
            p = prefix[b] + t
            assert int(tok[b, :, p - 1].argmin()) == 0, (b, t, tok[b, :, p - 1])
    assert float(gen["margin"].min()) > 0


def test_val_one_epoch_generation_branch(fvqa_lib, tmp_path):
    """`engine.py:78-85,99-121`: generation run -> extracted answers saved per epoch, accuracy = nearest option == answer, per-qtype
    meters; a model without a generator is refused, not silently scored by loss."""
    import json
    from flipped_vqa_b200 import engine
    model, data = _model()
    g = np.load(os.path.join(GOLDEN_DIR, "generation_small.npz"))
    args = argparse.Namespace(is_generation_task=True, output_dir=str(tmp_path), dataset="nextqa", debug=False)
    stats = engine.val_one_epoch(model, [data], None, 3, args=args)
    expect = float((torch.from_numpy(g["most_similar"]) == data["answer"]).float().mean())
    assert abs(stats["acc"] - expect) < 1e-9 and "Total" in stats and "C" in stats
    saved = json.load(open(tmp_path / "extracted_answers" / "extracted_answers_epoch3.json"))
    assert [e["video_id"] for e in saved] == data["vid"]

    class NoGenerator(torch.nn.Module):
        def forward(self, data, inference=False):
            raise AssertionError("must not be called")
    with pytest.raises(NotImplementedError):
        engine.val_one_epoch(NoGenerator(), [data], None, 0, args=args)
