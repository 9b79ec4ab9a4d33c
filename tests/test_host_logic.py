"""CPU: host-side logic of the drop-in boundary — parameter names / freeze rule, batch planning,
LR schedule, gradient-bucket layout. No kernels are launched."""
import math

import pytest
import torch

from flipped_vqa_b200 import _lib

from tests.util_parity import GOLDEN, golden_inputs, make_args


def _model():
    from flipped_vqa_b200.llama import ModelArgs, SyntheticTokenizer, Transformer
    return Transformer(ModelArgs(**GOLDEN), make_args(), tokenizer=SyntheticTokenizer(GOLDEN["vocab_size"]), device="cpu")


def test_parameter_names_match_reference_state_dict():
    m = _model()
    names = dict(m.named_parameters())
    d, hid = GOLDEN["dim"], 384
    expect = {"tok_embeddings.weight": (256, d), "output.weight": (256, d), "norm.weight": (d,),
              "adapter_query.weight": (GOLDEN["adapter_len"] * GOLDEN["adapter_layer"], d),
              "visual_proj.weight": (d, 768), "temporal_emb.weight": (10, d)}
    for i in range(GOLDEN["n_layers"]):
        p = f"layers.{i}."
        for w in ("wq", "wk", "wv", "wo"):
            expect[p + f"attention.{w}.weight"] = (d, d)
        expect[p + "attention.gate1"] = (1, 2, 1, 1)
        expect[p + "attention.gate2"] = (1, 2, 1, 1)
        expect[p + "feed_forward.w1.weight"] = (hid, d)
        expect[p + "feed_forward.w2.weight"] = (d, hid)
        expect[p + "feed_forward.w3.weight"] = (hid, d)
        expect[p + "attention_norm.weight"] = (d,)
        expect[p + "ffn_norm.weight"] = (d,)
    assert set(names) == set(expect)
    for n, s in expect.items():
        assert tuple(names[n].shape) == s, n
    assert set(m.state_dict()) == set(expect)


def test_freeze_rule_matches_llama_vqa():
    """llama_vqa.py:71-76: trainable <=> name contains gate/adapter/temporal_emb/visual_proj, fp32; rest frozen."""
    m = _model()
    for n, p in m.named_parameters():
        trainable = any(s in n for s in ("gate", "adapter", "temporal_emb", "visual_proj"))
        assert p.requires_grad == trainable, n
        assert p.dtype == (torch.float32 if trainable else _lib.H16), n
    assert float(m.layers[0].attention.gate1.abs().sum()) == 0.0                       # zero-init (model.py:84)
    assert torch.allclose(m.layers[0].attention.gate2, torch.full((1, 2, 1, 1), -3.5))  # -bias (model.py:85)


def test_load_state_dict_fills_packed_views():
    from flipped_vqa_b200.synthetic import synthetic_state_dict
    from types import SimpleNamespace
    m = _model()
    sd = synthetic_state_dict(SimpleNamespace(**GOLDEN), seed=1)
    m.load_state_dict(sd)
    d = GOLDEN["dim"]
    blk = m.layers[1]
    assert torch.equal(blk._wqkv[d:2 * d].float(), sd["layers.1.attention.wk.weight"])
    assert torch.equal(blk._w13[384:].float(), sd["layers.1.feed_forward.w3.weight"])
    assert m._pack_token is None            # transposed copies are rebuilt lazily on the next forward


@pytest.mark.parametrize("mode", ["audio_only", "concat", "sum", "attention"])
def test_audio_variant_parameters_match_reference_names(mode):
    """Input-fusion variants (`llama/model.py:209-227`): parameter names / shapes as in the reference, and only the names
    the substring rule of `llama_vqa.py:72` matches are trainable (audio_proj and the cross-attention stay frozen)."""
    from flipped_vqa_b200.llama import ModelArgs, SyntheticTokenizer, Transformer
    from tests.util_parity import golden_audio_inputs
    params, sd, data = golden_audio_inputs(mode)
    m = Transformer(ModelArgs(**GOLDEN), make_args(audio_mode=mode), tokenizer=SyntheticTokenizer(256), device="cpu")
    names = {n: tuple(p.shape) for n, p in m.named_parameters()}
    assert set(names) == set(sd), set(names) ^ set(sd)
    for n, t in sd.items():
        assert names[n] == tuple(t.shape), n
    for n, p in m.named_parameters():
        assert p.requires_grad == any(s in n for s in ("gate", "adapter", "temporal_emb", "visual_proj")), n
    assert ("visual_proj.weight" in names) == (mode != "audio_only")
    assert m.video_dim == {"audio_only": 0, "concat": 768 + 1024, "sum": 768, "attention": 768}[mode]


def test_batch_plan_rows_and_targets():
    from flipped_vqa_b200.step import BatchPlan
    params, sd, data = golden_inputs()
    plan = BatchPlan(data, ["vqa", "vaq", "qav"], 10)
    B, S = 3, 48
    assert (plan.n_seq, plan.T) == (9, 9 * S)
    names = ["ids", "labels", "vstart", "seq_video", "qav_index", "ce_rows", "ce_tgt", "ce_dst", "q_rows", "q_tgt", "q_vid"]
    arr = {n: plan.host_ints[o:o + k] for n, (o, k) in zip(names, plan._slices)}
    assert arr["vstart"].tolist() == [12] * 6 + [-1] * 3
    assert arr["seq_video"].tolist() == [0, 1, 2] * 3
    # labelled rows: position p predicts label[p+1] (model.py:278 shift), ignore_index 0
    lab = data["label"]["vqa"].reshape(B, S)
    n_vqa = int((lab[:, 1:] != 0).sum())
    assert plan.ce_counts["vqa"] == n_vqa == 12
    rows = arr["ce_rows"][:n_vqa].long()
    assert torch.equal(arr["ce_tgt"][:n_vqa].long(), lab.flatten()[rows + 1])
    # QAV rows precede each video slot; targets 0..F-1; ignore_index -1
    assert plan.q_count == 30
    assert arr["q_tgt"].tolist() == list(range(10)) * 3
    ql = data["label"]["qav"].reshape(B, S)
    qrows = arr["q_rows"].long() - 6 * S
    assert torch.equal(ql.flatten()[qrows + 1], arr["q_tgt"].long())


def test_plans_video_start_rule():
    """Default: sample 0's video_start for every sample of the batch (`llama/model.py:264`); `per_sample_video_start=True`: each
    sample's own, repeated over its options; the QAV stream never has one (-1). Host-side arrays only (no GPU)."""
    from flipped_vqa_b200.step import BatchPlan, OptionPlan
    from flipped_vqa_b200.synthetic import synthetic_batch
    data = synthetic_batch(3, 64, 256, seed=9, video_start=12, n_options=2)
    data["video_start"] = {"vqa": [12, 15, 9], "vaq": [11, 14, 8], "qav": data["video_start"]["qav"]}
    vstart = lambda plan: plan.host_ints[plan._slices[2][0]:plan._slices[2][0] + plan._slices[2][1]].tolist()
    assert vstart(BatchPlan(data, ["vqa", "vaq", "qav"], 10)) == [12] * 6 + [11] * 6 + [-1] * 6
    assert vstart(BatchPlan(data, ["vqa", "vaq", "qav"], 10, per_sample_video_start=True)) == \
        [12, 12, 15, 15, 9, 9] + [11, 11, 14, 14, 8, 8] + [-1] * 6
    assert OptionPlan(data, 10).host("vstart").tolist() == [12] * 6
    assert OptionPlan(data, 10, per_sample_video_start=True).host("vstart").tolist() == [12, 12, 15, 15, 9, 9]
    with pytest.raises(AssertionError):
        BatchPlan(dict(data, video_start={"vqa": [12, 15], "vaq": [11, 14], "qav": [0, 0]}), ["vqa"], 10, per_sample_video_start=True)


def test_batch_plan_padding_free_rows():
    """BatchPlan's compact row set: every sequence keeps exactly the rows [0, last loss-relevant position]; the loss row
    lists re-indexed into it point at the same (sequence, position) pairs."""
    from flipped_vqa_b200.step import BatchPlan
    from flipped_vqa_b200.synthetic import synthetic_batch
    data = synthetic_batch(4, 64, 512, seed=5)
    data["label"]["vaq"][2] = 0
    plan = BatchPlan(data, ["vqa", "vaq", "qav"], 10)
    names = ["ids", "labels", "vstart", "seq_video", "qav_index", "ce_rows", "ce_tgt", "ce_dst", "q_rows", "q_tgt", "q_vid",
             "live_rows", "ce_rows_c", "q_rows_c", "pos_ids", "c2f", "f2c", "ce_rows_k", "q_rows_k", "live_rows_k", "ce_rows_kc", "q_rows_kc"]
    arr = {n: plan.host_ints[o:o + k].long() for n, (o, k) in zip(names, plan._slices)}
    S, c2f, f2c = plan.S, arr["c2f"], arr["f2c"]
    assert plan.T_c == c2f.numel() < plan.T
    assert torch.equal(f2c[c2f], torch.arange(plan.T_c)) and torch.equal(arr["pos_ids"], c2f % S)
    loss_rows = torch.cat([arr["ce_rows"], arr["q_rows"]])
    for n in range(plan.n_seq):
        kept = (f2c[n * S:(n + 1) * S] >= 0)
        mine = loss_rows[(loss_rows // S) == n] % S
        end = int(mine.max()) + 1 if mine.numel() else 0
        assert int(kept.sum()) == end and bool(kept[:end].all())
    assert int((f2c[(4 + 2) * S:(4 + 3) * S] >= 0).sum()) == 0          # the VAQ sequence without labels keeps no rows
    assert torch.equal(c2f[arr["ce_rows_k"]], arr["ce_rows"]) and torch.equal(c2f[arr["q_rows_k"]], arr["q_rows"])
    assert torch.equal(arr["live_rows_k"][arr["ce_rows_kc"]], arr["ce_rows_k"])
    assert torch.equal(arr["live_rows_k"][arr["q_rows_kc"]], arr["q_rows_k"])


def test_batch_plan_option_layout():
    from flipped_vqa_b200.step import BatchPlan
    params, sd, data = golden_inputs(5)
    plan = BatchPlan(data, ["vqa"], 10, inference=True)
    assert plan.n_seq == 15 and plan.n_opt == 5
    names = ["ids", "labels", "vstart", "seq_video"]
    arr = {n: plan.host_ints[o:o + k] for n, (o, k) in zip(names, plan._slices)}
    assert arr["seq_video"].tolist() == [0] * 5 + [1] * 5 + [2] * 5      # video repeated per option (model_my_original_mod.py:332-333)


def test_option_plan_shared_prefix_layout():
    """step.OptionPlan: the compact ragged row set of shared-prefix option scoring. Every labelled row is present, every
    present row is backed by a compact row with the SAME token history (so its hidden state is the same number), the
    kept rows of a sequence are causally closed, and sequences without labels / with an early divergence are handled."""
    from flipped_vqa_b200.step import OptionPlan
    from flipped_vqa_b200.synthetic import synthetic_batch
    for seed, n_opt in ((0, 1), (1, 5), (3, 5)):
        data = synthetic_batch(4, 64, 512, seed=seed, n_options=n_opt)
        if seed == 3:
            data["label"]["vqa"][1] = 0                 # a sample no loss reads
            data["text_id"]["vqa"][2, 3, 5] = 7         # options diverge at position 5
        p = OptionPlan(data, 10)
        ids, lab = data["text_id"]["vqa"], data["label"]["vqa"]
        B, n, S = ids.shape
        c2f, f2c, pos = p.host("c2f").long(), p.host("f2c").long(), p.host("pos_ids").long()
        assert p.T_c == c2f.numel() < p.T
        assert torch.equal(c2f % S, pos)
        assert torch.equal(f2c[c2f], torch.arange(p.T_c))
        for b in range(B):
            for o in range(n):
                present = f2c[(b * n + o) * S:(b * n + o + 1) * S] >= 0
                k = int(present.sum())
                assert bool(present[:k].all()) and not bool(present[k:].any())          # causally closed
                assert k == int(p.end[b])
                for t in range(k):
                    src = int(c2f[f2c[(b * n + o) * S + t]])
                    sb, so, sp = src // (n * S), (src // S) % n, src % S
                    assert sb == b and sp == t and torch.equal(ids[b, so, :t + 1], ids[b, o, :t + 1])
                labelled = (lab[b, o, 1:] != 0).nonzero().flatten()
                assert bool((labelled < k).all())
        # the CE row list: targets and destinations of every labelled (sequence, position) pair
        nz = (lab.reshape(B * n, S)[:, 1:] != 0).nonzero()
        assert p.ce_total == nz.shape[0]
        assert torch.equal(p.host("ce_tgt").long(), lab.reshape(B * n, S)[:, 1:][nz[:, 0], nz[:, 1]])
        assert torch.equal(p.host("ce_dst").long(), nz[:, 0] * (S - 1) + nz[:, 1])
        assert torch.equal(c2f[p.host("ce_rows").long()] % S, nz[:, 1])
        assert torch.equal(p.host("live_rows").long()[p.host("ce_rows_c").long()], p.host("ce_rows").long())
        if seed == 3:
            assert int(p.end[1]) == 0 and int(p.prefix_len[2]) == 5
        if n_opt == 5 and seed == 1:
            assert p.T_c < p.T // 4                     # 5 options cost about one sequence per sample


def test_lr_schedule_matches_reference_formula():
    from flipped_vqa_b200.util import lr_sched
    from types import SimpleNamespace
    opt = SimpleNamespace(param_groups=[{"lr": 0.0}, {"lr": 0.0, "lr_scale": 0.5}])
    a = SimpleNamespace(lr=0.09, min_lr=0.0, warmup_epochs=2, epochs=5)
    assert lr_sched.adjust_learning_rate(opt, 1.0, a) == pytest.approx(0.045)
    assert opt.param_groups[1]["lr"] == pytest.approx(0.0225)
    lr = lr_sched.adjust_learning_rate(opt, 3.5, a)
    assert lr == pytest.approx(0.09 * 0.5 * (1 + math.cos(math.pi * 1.5 / 3)))


def test_grad_buffer_layout():
    from flipped_vqa_b200.step import GradBuffers
    gb = GradBuffers(32, 10, 4096, 32, 768, 10, "cpu")
    assert gb.flat.numel() == 4499456             # 7B trainables = 4 499 456 fp32 = 18.0 MB
    assert gb.late_offset == 32 * 10 * 4096
    gb.adapter[5, 7] = 3.0
    assert float(gb.flat[5 * 4096 + 7]) == 3.0


def test_merge_shards_reassembles_model_parallel_checkpoint():
    """llama_vqa.py:25-58: Meta's model-parallel shards are concatenated along their split dimension
    (column-parallel wq/wk/wv/w1/w3/output on dim 0, row-parallel wo/w2/tok_embeddings on dim 1, norms replicated)."""
    from types import SimpleNamespace
    from flipped_vqa_b200.llama_vqa import merge_shards
    from flipped_vqa_b200.synthetic import synthetic_state_dict
    full = {k: v for k, v in synthetic_state_dict(SimpleNamespace(**GOLDEN), seed=2).items()
            if not any(s in k for s in ("gate", "adapter", "temporal_emb", "visual_proj"))}
    col = ("wq.weight", "wk.weight", "wv.weight", "w1.weight", "w3.weight", "output.weight")
    row = ("wo.weight", "w2.weight", "tok_embeddings.weight")
    shards = [{}, {}]
    for k, v in full.items():
        for r in range(2):
            if k.endswith(col):
                shards[r][k] = v.chunk(2, dim=0)[r].clone()
            elif k.endswith(row):
                shards[r][k] = v.chunk(2, dim=1)[r].clone()
            else:
                shards[r][k] = v.clone()
    merged = merge_shards(shards, GOLDEN["n_layers"])
    assert set(merged) == set(full)
    for k in full:
        assert torch.equal(merged[k], full[k]), k
    assert merge_shards([full], GOLDEN["n_layers"]) is full


def test_checkpoint_save_resume_roundtrip(tmp_path):
    """util/misc.py:297-336: checkpoints hold only the trainables (+ optimizer, scaler, epoch, args) under the
    reference's parameter names; resume restores them into a fresh model with strict=False."""
    import argparse
    from flipped_vqa_b200.util import misc
    m = _model()
    with torch.no_grad():
        for n, p in m.named_parameters():
            if p.requires_grad:
                p.add_(torch.randn_like(p) * 0.1)
    opt = torch.optim.AdamW([p for p in m.parameters() if p.requires_grad], lr=1e-3)
    args = argparse.Namespace(output_dir=str(tmp_path), resume="")
    misc.save_model(args, 3, m, m, opt, misc.NativeScalerWithGradNormCount(), "checkpoint_best")   # the name `train.py:141` passes
    ck = torch.load(tmp_path / "checkpoint_best.pth", map_location="cpu", weights_only=False)
    names = set(ck["model"])
    assert names == {n for n, p in m.named_parameters() if p.requires_grad}          # trainables only, reference names
    assert "adapter_query.weight" in names and "layers.1.attention.gate2" in names and not any("wq" in n for n in names)
    m2 = _model()
    opt2 = torch.optim.AdamW([p for p in m2.parameters() if p.requires_grad], lr=1e-3)
    args.resume = str(tmp_path / "checkpoint_best.pth")
    misc.load_model(args, m2, opt2, misc.NativeScalerWithGradNormCount())
    assert args.start_epoch == 4
    for (n, a), (_, b) in zip(m.named_parameters(), m2.named_parameters()):
        if a.requires_grad:
            assert torch.equal(a, b), n
    # `util/misc.py:331`: an --eval run restores the weights only (no optimizer state, start_epoch untouched)
    m3 = _model()
    opt3 = torch.optim.AdamW([p for p in m3.parameters() if p.requires_grad], lr=1e-3)
    args3 = argparse.Namespace(output_dir=str(tmp_path), resume=args.resume, eval=True, start_epoch=0)
    misc.load_model(args3, m3, opt3, misc.NativeScalerWithGradNormCount())
    assert args3.start_epoch == 0 and not opt3.state_dict()["state"]
    assert torch.equal(m3.adapter_query.weight, m.adapter_query.weight)


def test_log_qtype_meters_match_reference_formulas():
    """`util/misc.py:361-532`: per-question-type meters of the validation loop; MetricLogger.update(count=1, **metrics) keeps the
    reference's quirk that a keyword n= is a meter named 'n'."""
    import argparse
    from flipped_vqa_b200.util import misc
    ml = misc.MetricLogger()
    data = {"qtype": torch.tensor([1, 2, 3, 6, 6, 8])}
    hit = torch.tensor([True, False, True, True, False, True])
    misc.log_qtype(data, hit, ml, argparse.Namespace(dataset="nextqa"))
    eps = 1e-10
    assert abs(ml.meters["C"].global_avg - 1 / (2 + eps)) < 1e-12
    assert abs(ml.meters["T"].global_avg - 1 / (1 + eps)) < 1e-12
    assert abs(ml.meters["D"].global_avg - 2 / (3 + eps)) < 1e-12
    assert abs(ml.meters["Total"].global_avg - 4 / 6) < 1e-12
    assert ml.meters["n"].count == 4                                   # four update(n=...) calls -> meter 'n'
    ml.update(n=6, acc=0.5)
    assert ml.meters["acc"].count == 1 and ml.meters["n"].count == 5
    ml2 = misc.MetricLogger()
    misc.log_qtype({"qtype": torch.tensor([1, 7, 12, 15])}, torch.tensor([1, 0, 1, 1]), ml2, argparse.Namespace(dataset="musicavqa"))
    assert abs(ml2.meters["audio"].global_avg - 1 / (1 + eps)) < 1e-12 and abs(ml2.meters["visual"].global_avg) < 1e-12
    assert abs(ml2.meters["audio_visual"].global_avg - 2 / (2 + eps)) < 1e-12
    assert abs(ml2.meters["existential"].global_avg - 1 / (2 + eps)) < 1e-12 and abs(ml2.meters["counting"].global_avg - 1 / (1 + eps)) < 1e-12
    assert misc.get_qtype_mapping("valor32k")["rel_pos_both"] == 18 and misc.get_qtype_mapping("star")["Feas"] == 4
    assert misc.get_qtype_mapping("musicavqa")["Audio-Visual_Counting"] == 15 and misc.get_qtype_mapping("tvqa") == {}
    ml3 = misc.MetricLogger()
    misc.log_qtype(data, hit, ml3, argparse.Namespace(dataset="tvqa"))
    assert not ml3.meters


def test_save_result_merges_rank_files(tmp_path):
    """`util/misc.py:570-610`: per-rank result files merged by the main process (single process here: rank 0 only)."""
    import json
    from flipped_vqa_b200.util import misc
    res = [{"video_id": "v0", "question": "q", "generated_answer": "a"}, {"video_id": "v1", "question": "q2", "generated_answer": "b"}]
    final = misc.save_result(res, str(tmp_path), "extracted_answers_epoch0")
    assert final.endswith("extracted_answers_epoch0.json")
    assert json.load(open(final)) == res and json.load(open(tmp_path / "extracted_answers_epoch0_rank0.json")) == res
    d = misc.save_result({"a": torch.tensor([1, 2])}, str(tmp_path), "blob", is_json=False, is_list=False)
    assert torch.equal(torch.load(d, weights_only=False)["a"], torch.tensor([1, 2]))


def test_generation_batch_layout():
    """The synthetic validation batch of the generation evaluator follows the layout `llama/model.py:367-546` relies on:
    prefix_index = index(answer token) + 5 = first answer position, labels = answer span incl. EOS, 31 steps fit the sequence."""
    from flipped_vqa_b200.synthetic import GEN_QUESTION_MARKER, synthetic_generation_batch
    d = synthetic_generation_batch(5, 96, 1024, a_token_id=900, seed=4, n_options=4)
    ids, lab = d["text_id"]["vqa"], d["label"]["vqa"]
    assert ids.shape == lab.shape == (5, 4, 96)
    for b, p in enumerate(d["prefix_index"]["vqa"]):
        row = ids[b, 0].tolist()
        assert row.index(900) + 5 == p and GEN_QUESTION_MARKER in row[:p] and p + 30 <= 95
        assert row[12:22] == [0] * 10                                         # video placeholders at video_start
        for o in range(4):
            assert torch.equal(ids[b, o, :p], ids[b, 0, :p])                  # options differ only in the answer span
            span = (lab[b, o] != 0).nonzero().flatten().tolist()
            assert span[0] == p and span == list(range(p, p + len(span))) and int(ids[b, o, span[-1]]) == 2   # ends with EOS
            assert torch.equal(lab[b, o, span], ids[b, o, span])
