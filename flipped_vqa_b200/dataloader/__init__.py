"""Input pipeline of the LLaMA-VQA step: from tokenised prompts to device-resident batch plans.

Reference: `dataloader/base_dataset.py:16-173` (`_get_padding_id`, `_get_text_token`: padding, labels, masks, video
indices), `dataloader/__init__.py:28-90` (`batch_collate`: the batch dict `Transformer.forward` consumes).
The dataset classes themselves (CSV / feature-file readers, per-dataset prompt text) are out of scope.

`PlannedLoader` is the B200 part: a worker thread turns each collated batch into a `BatchPlan` / `OptionPlan` (flat
int32 arrays in pinned memory) and issues its single H2D copy on a side stream one or more steps AHEAD of the compute
stream, so the step never waits for host-side planning or the copy (`model.forward_plan(plan)` /
`model.inference_plan(plan)`); the reference does the H2D synchronously inside `forward` (`llama/model.py:254-264`).
"""
from __future__ import annotations

import queue
import threading
from typing import Dict, Iterable, List, Optional, Sequence

import torch

_TASKS = ("vqa", "vaq", "qav")


def pad_text_ids(seqs: Sequence[Sequence[int]], max_seq_len: int) -> torch.Tensor:
    """[n, max_seq_len] int64, short sequences padded with -1, long ones truncated (`base_dataset.py:16-28`)."""
    out = torch.full((len(seqs), max_seq_len), -1, dtype=torch.int64)
    for i, s in enumerate(seqs):
        t = torch.as_tensor(list(s[:max_seq_len]), dtype=torch.int64)
        out[i, :t.numel()] = t
    return out


def build_text_tensors(ids: Dict[str, Sequence[Sequence[int]]], prefix_index: Dict[str, int], video_start: Dict[str, int],
                       max_seq_len: int, max_feats: int) -> Dict[str, Dict]:
    """Everything `BaseDataset._get_text_token` derives from the tokenised prompts (`base_dataset.py:49-173`):

      text_id      padded ids, padding and the -2 video placeholders -> 0 (`:99-104`)
      label        vqa / vaq: ids from prefix_index on, 0 elsewhere (ignore_index 0, `:65-77`);
                   qav: 0..F-1 at [prefix, prefix+F) clipped to the sequence, -1 elsewhere (`:80-91`)
      label_mask   1.0 where vqa / vaq labels are real tokens; qav: 1.0 at prefix only (`:93-95`)
      video_index  arange(prefix, prefix+F) per task (`:118-120`); video_start: vqa / vaq as given, qav = its prefix
    `ids[task]` is the list of option sequences of one sample (one for training)."""
    out = {k: {} for k in ("text_id", "label", "video_start", "video_index", "label_mask", "prefix_index")}
    F = max_feats
    for task in _TASKS:
        padded = pad_text_ids(ids[task], max_seq_len)
        p = int(prefix_index[task])
        if task == "qav":
            label = torch.full_like(padded, -1)
            n = max(min(max_seq_len - p, F), 0)
            label[:, p:p + n] = torch.arange(n)
            mask = torch.zeros_like(padded, dtype=torch.float32)
            mask[:, p] = 1.0
        else:
            label = padded.clone()
            label[:, :p] = -1
            mask = (label >= 0).float()
            label[label < 0] = 0
        out["text_id"][task] = padded.clamp(min=0)
        out["label"][task] = label
        out["label_mask"][task] = mask
        out["video_index"][task] = torch.arange(p, p + F)
        out["prefix_index"][task] = p
        out["video_start"][task] = p if task == "qav" else int(video_start[task])
    return out


def encode_sample(tokenizer, text: Dict[str, str], answer: int, answer_mapping: Dict[int, str], split: str, max_seq_len: int,
                  max_feats: int, options: Optional[List[str]] = None) -> Dict[str, Dict]:
    """Prompt building + tensors for one sample = `BaseDataset._get_text_token` (`base_dataset.py:30-173`)."""
    kw = dict(text=text, max_feats=max_feats, split=split, answer_mapping=answer_mapping, answer=answer, options=options)
    vqa, vqa_p, vqa_vs = tokenizer.encode_vqa(**kw)
    vaq, vaq_p, vaq_vs = tokenizer.encode_vaq(**kw)
    qav, qav_p = tokenizer.encode_qav(**kw)
    return build_text_tensors({"vqa": vqa, "vaq": vaq, "qav": qav}, {"vqa": vqa_p, "vaq": vaq_p, "qav": qav_p},
                              {"vqa": vqa_vs, "vaq": vaq_vs}, max_seq_len, max_feats)


def pad_text_ids_sub(seqs: Sequence[Sequence[int]], prefix_index: int, prefix_i: int, prefix_main: int, task: str, max_seq_len: int,
                     max_feats: int, q_token_id: int, sub: bool = True):
    """`TVQA._get_padding_id` (`dataloader/tvqa.py:75-108`): like `pad_text_ids`, but a sequence that overflows max_seq_len keeps
    its head [0, prefix_i), its tail [prefix_main, end) and as much of the DIALOGUE [prefix_i, prefix_main) as still fits, instead
    of losing its tail; prefix_index is then re-derived (vqa: S - 4, qav: S - F - 1, vaq: first 'Question' piece + 2 over all rows
    padded so far). Without `sub`, or when the sample has no dialogue, overflow is a plain truncation. Returns (padded, prefix)."""
    S = max_seq_len
    out = torch.full((len(seqs), S), -1, dtype=torch.int64)
    prefix = prefix_index
    for i, s in enumerate(seqs):
        t = torch.as_tensor(list(s), dtype=torch.int64)
        if t.numel() <= S:
            out[i, :t.numel()] = t
            prefix = prefix_index
        elif sub and prefix_i != prefix_main:
            keep = S - (prefix_i + (t.numel() - prefix_main))                  # dialogue tokens that still fit
            out[i, :prefix_i] = t[:prefix_i]
            out[i, prefix_i:prefix_i + keep] = t[prefix_i:prefix_i + keep]
            out[i, prefix_i + keep:] = t[prefix_main:]
            if task == "vqa":
                prefix = S - 4
            elif task == "vaq":
                prefix = int((out == q_token_id).nonzero(as_tuple=True)[1][0]) + 2
            else:
                prefix = S - max_feats - 1
        else:
            out[i] = t[:S]
            prefix = prefix_index
    return out, prefix


def encode_sample_sub(tokenizer, text: Dict[str, str], answer: int, answer_mapping: Dict[int, str], split: str, max_seq_len: int,
                      max_feats: int, sub: bool = True) -> Dict[str, Dict]:
    """Prompt building + tensors for one `--sub` sample = `TVQA._get_text_token` (`dataloader/tvqa.py:110-160`; `vlep.py:104-154` is
    the same code): dialogue prompts, dialogue-aware overflow handling, then the usual labels / masks / video indices."""
    kw = dict(text=text, max_feats=max_feats, split=split, answer_mapping=answer_mapping, answer=answer)
    vqa, vqa_p, vqa_vs, vqa_i, vqa_m = tokenizer.encode_dvqa(**kw)
    vaq, vaq_p, vaq_vs, vaq_i, vaq_m = tokenizer.encode_dvaq(**kw)
    qav, qav_p, qav_i, qav_m = tokenizer.encode_dqav(max_seq_len=max_seq_len, **kw)
    padded, prefix = {}, {}
    for task, seqs, p, pi, pm in (("vqa", vqa, vqa_p, vqa_i, vqa_m), ("vaq", vaq, vaq_p, vaq_i, vaq_m), ("qav", qav, qav_p, qav_i, qav_m)):
        padded[task], prefix[task] = pad_text_ids_sub(seqs, p, pi, pm, task, max_seq_len, max_feats, tokenizer.q_token_id, sub)
    ids = {t: [row.tolist() for row in padded[t]] for t in _TASKS}                      # already padded to S with -1
    out = build_text_tensors(ids, prefix, {"vqa": vqa_vs, "vaq": vaq_vs}, max_seq_len, max_feats)
    # `tvqa.py:139`: the qav frame labels are NOT clipped to the sequence there (prefix + F <= S holds by construction)
    return out


def batch_collate(batch: List[Dict]) -> Dict:
    """The batch dict of `dataloader/__init__.py:28-90`: per-task tensors stacked over samples, python lists for
    `video_start` / `prefix_index` / ids, optional `video` and `audio` (+ lengths)."""
    out: Dict = {"vid": [b["vid"] for b in batch]}
    for name in ("video", "audio"):
        if name in batch[0]:
            out[name] = torch.stack([b[name] for b in batch])
            out[name + "_len"] = torch.tensor([b[name + "_len"] for b in batch], dtype=torch.long)
    out["text"] = [b["text"] for b in batch]
    for name in ("text_id", "label", "video_index", "label_mask"):
        out[name] = {t: torch.stack([b[name][t] for b in batch]) for t in _TASKS}
    for name in ("video_start", "prefix_index"):
        out[name] = {t: [b[name][t] for b in batch] for t in _TASKS}
    out["qid"] = [b["qid"] for b in batch]
    out["answer"] = torch.tensor([b["answer"] for b in batch])
    out["qtype"] = torch.tensor([b["qtype"] for b in batch])
    return out


class PlannedLoader:
    """Wraps an iterable of collated batches; yields `(data, plan)` with `plan` already on the device.

    A daemon thread builds the plan of batch i+depth (pinned int32 staging, `step.BatchPlan` / `step.OptionPlan`) and
    enqueues its H2D copies on `copy_stream` while the compute stream still runs batch i; the consumer makes the
    compute stream wait on the copy's event, never the host. `inference=True` plans option-scoring batches."""

    def __init__(self, loader: Iterable, model, inference: bool = False, depth: int = 2):
        self.loader, self.model, self.inference, self.depth = loader, model, inference, max(depth, 1)
        self.copy_stream = torch.cuda.Stream(device=model._device) if torch.cuda.is_available() else None

    def __len__(self):
        return len(self.loader)

    def _plan(self, data):
        m = self.model
        if self.copy_stream is None:
            return (m.plan_options(data) if self.inference else m.plan_batch(data)), None
        with torch.cuda.stream(self.copy_stream):
            plan = m.plan_options(data) if self.inference else m.plan_batch(data)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        return plan, ev

    def __iter__(self):
        q: "queue.Queue" = queue.Queue(maxsize=self.depth)
        stop = object()
        cancel = threading.Event()
        # weight packing (Transformer.repack) belongs to the consumer's thread: never let the worker trigger it concurrently
        if hasattr(self.model, "_ensure_packed") and torch.cuda.is_available():
            self.model._ensure_packed()

        def put(item) -> bool:
            """Blocking put that gives up when the consumer has gone away (early break, exception, sys.exit in the loop)."""
            while not cancel.is_set():
                try:
                    q.put(item, timeout=0.1)
                    return True
                except queue.Full:
                    continue
            return False

        def work():
            try:
                for data in self.loader:
                    if cancel.is_set() or not put((data,) + self._plan(data)):
                        return
            except BaseException as e:            # surface loader / planning errors in the consumer
                put(e)
            put(stop)

        worker = threading.Thread(target=work, daemon=True)
        worker.start()
        try:
            while True:
                item = q.get()
                if item is stop:
                    return
                if isinstance(item, BaseException):
                    raise item
                data, plan, ev = item
                if ev is not None:
                    cur = torch.cuda.current_stream()
                    cur.wait_event(ev)
                    plan.record_stream(cur)
                yield data, plan
        finally:                                  # also runs when the consumer abandons the generator (GeneratorExit)
            cancel.set()
            while True:                           # release the pinned slots / device plans the worker had queued
                try:
                    q.get_nowait()
                except queue.Empty:
                    break
            worker.join(timeout=5.0)
