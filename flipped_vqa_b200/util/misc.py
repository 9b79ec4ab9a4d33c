"""The parts of `/root/reference/util/misc.py` that sit on the training step's path, same call
signatures: `NativeScalerWithGradNormCount` (`:253-279`), `get_grad_norm_` (`:282-294`),
`init_distributed_mode` (`:220-250`), trainable-only checkpoints (`:297-336`), plus a compact metric
logger with the `global_avg` contract `engine.py:54-56` returns."""
from __future__ import annotations

import datetime
import os
import time
from collections import defaultdict, deque
from pathlib import Path

import torch
import torch.distributed as dist


class SmoothedValue:
    def __init__(self, window_size=20, fmt=None):
        self.deque = deque(maxlen=window_size)
        self.total, self.count = 0.0, 0
        self.fmt = fmt or "{median:.4f} ({global_avg:.4f})"

    def update(self, value, n=1):
        self.deque.append(value)
        self.count += n
        self.total += value * n

    def synchronize_between_processes(self):
        if not is_dist_avail_and_initialized():
            return
        dev = "cuda" if torch.cuda.is_available() and dist.get_backend() == "nccl" else "cpu"
        t = torch.tensor([self.count, self.total], dtype=torch.float64, device=dev)
        dist.barrier()
        dist.all_reduce(t)
        self.count, self.total = int(t[0].item()), float(t[1].item())

    @property
    def median(self):
        d = sorted(self.deque)
        return d[len(d) // 2] if d else 0.0

    @property
    def avg(self):
        return sum(self.deque) / max(len(self.deque), 1)

    @property
    def global_avg(self):
        return self.total / self.count if self.count else 0.0      # `util/misc.py:83-84` (a fractional weight is possible, see log_qtype)

    @property
    def value(self):
        return self.deque[-1] if self.deque else 0.0

    def __str__(self):
        return self.fmt.format(median=self.median, avg=self.avg, global_avg=self.global_avg, value=self.value)


class MetricLogger:
    def __init__(self, delimiter="\t"):
        self.meters = defaultdict(SmoothedValue)
        self.delimiter = delimiter

    def update(self, count=1, **metrics):
        """`util/misc.py:111-118`: the weight is the positional/keyword `count`; a keyword `n=` (as `engine.py:133-135` and the
        per-question-type updates pass it) is therefore a METER called "n", exactly as in the reference."""
        for k, v in metrics.items():
            if v is None:
                continue
            if isinstance(v, torch.Tensor):
                v = v.item()
            self.meters[k].update(float(v), n=count)

    def add_meter(self, name, meter):
        self.meters[name] = meter

    def synchronize_between_processes(self):
        for m in self.meters.values():
            m.synchronize_between_processes()

    def __str__(self):
        return self.delimiter.join(f"{k}: {m}" for k, m in self.meters.items())

    def log_every(self, iterable, print_freq, header=None):
        header = header or ""
        start = time.time()
        n = len(iterable) if hasattr(iterable, "__len__") else -1
        for i, obj in enumerate(iterable):
            yield obj
            if print_freq and (i % print_freq == 0 or i == n - 1):
                print(f"{header} [{i}/{n}] {self}  elapsed {datetime.timedelta(seconds=int(time.time() - start))}")


def is_dist_avail_and_initialized():
    return dist.is_available() and dist.is_initialized()


def get_world_size():
    return dist.get_world_size() if is_dist_avail_and_initialized() else 1


def get_rank():
    return dist.get_rank() if is_dist_avail_and_initialized() else 0


def is_main_process():
    return get_rank() == 0


def save_on_master(*args, **kwargs):
    if is_main_process():
        torch.save(*args, **kwargs)


def save_result(result, result_dir, filename, is_json=True, is_list=True):
    """`util/misc.py:570-610`: every rank writes `<filename>_rank<r>.{json,pth}`; after a barrier the main process merges all
    ranks into `<filename>.{json,pth}` (the reference writes the merged JSON over the LAST rank's file and leaves the final path
    empty - `:603` - here it goes to the path that is returned)."""
    import json
    ext = "json" if is_json else "pth"
    part = os.path.join(result_dir, f"{filename}_rank{get_rank()}.{ext}")
    final = os.path.join(result_dir, f"{filename}.{ext}")
    if is_json:
        with open(part, "w") as f:
            json.dump(result, f, default=lambda o: o.tolist() if hasattr(o, "tolist") else str(o))
    else:
        torch.save(result, part)
    if is_dist_avail_and_initialized():
        dist.barrier()
    if is_main_process():
        merged = [] if is_list else {}
        for r in range(get_world_size()):
            path = os.path.join(result_dir, f"{filename}_rank{r}.{ext}")
            res = json.load(open(path)) if is_json else torch.load(path, weights_only=False)
            if is_list:
                merged += res
            else:
                merged.update(res)
        if is_json:
            with open(final, "w") as f:
                json.dump(merged, f)
        else:
            torch.save(merged, final)
        print("result file saved to %s" % final)
    if is_dist_avail_and_initialized():
        dist.barrier()
    return final


def init_distributed_mode(args):
    """One process per GPU from torchrun's env (`util/misc.py:230-233`); NCCL over NVLink 5."""
    if "RANK" in os.environ and "WORLD_SIZE" in os.environ:
        args.rank = int(os.environ["RANK"])
        args.world_size = int(os.environ["WORLD_SIZE"])
        args.gpu = int(os.environ.get("LOCAL_RANK", 0))
    else:
        print("Not using distributed mode")
        args.distributed = False
        return
    args.distributed = True
    torch.cuda.set_device(args.gpu)
    args.dist_backend = "nccl"
    dist.init_process_group(backend=args.dist_backend, init_method=getattr(args, "dist_url", "env://"),
                            world_size=args.world_size, rank=args.rank)


class NativeScalerWithGradNormCount:
    """Same call path as `util/misc.py:253-279`. bf16 needs no loss scaling, so the GradScaler is
    disabled by default (scale = 1, no inf-check host sync); `enabled=True` keeps the fp16-style
    scaling path alive — the step's backward multiplies by the incoming scaled gradient."""
    state_dict_key = "amp_scaler"

    def __init__(self, enabled: bool = False):
        self._scaler = torch.amp.GradScaler("cuda", enabled=enabled)

    accepts_before_step = True     # engine.train_one_epoch: this scaler can run the finite-loss check between backward and the update

    def __call__(self, loss, optimizer, clip_grad=None, parameters=None, create_graph=False, update_grad=True, before_step=None):
        """`before_step` (optional callable) runs after backward has been ENQUEUED and before anything touches the optimizer:
        the place for a host-side read of the loss (it then waits for the forward pass while the GPU is already in backward)."""
        self._scaler.scale(loss).backward(create_graph=create_graph)
        if before_step is not None:
            before_step()
        if update_grad:
            self._scaler.unscale_(optimizer)
            if clip_grad is not None:
                assert parameters is not None
                norm = torch.nn.utils.clip_grad_norm_(parameters, clip_grad)
            else:
                norm = get_grad_norm_(parameters)
            self._scaler.step(optimizer)
            self._scaler.update()
        else:
            norm = None
        return norm

    def state_dict(self):
        return self._scaler.state_dict()

    def load_state_dict(self, state_dict):
        self._scaler.load_state_dict(state_dict)


def get_grad_norm_(parameters, norm_type: float = 2.0) -> torch.Tensor:
    """L2 norm over parameters that have a gradient (`util/misc.py:282-294`), as one foreach op; stays on
    the device (no host sync)."""
    if isinstance(parameters, torch.Tensor):
        parameters = [parameters]
    grads = [p.grad.detach() for p in parameters if p.grad is not None]
    if not grads:
        return torch.tensor(0.0)
    if norm_type == float("inf"):
        return torch.stack([g.abs().max() for g in grads]).max()
    return torch.linalg.vector_norm(torch.stack(torch._foreach_norm(grads, norm_type)), norm_type)


def trainable_state(model_without_ddp):
    """`util/misc.py:303-306`: only parameters whose name matches the trainable rule are checkpointed."""
    return {n: p for n, p in model_without_ddp.named_parameters()
            if any(s in n for s in ("gate", "adapter", "temporal_emb", "visual_proj"))}


def save_model(args, epoch, model, model_without_ddp, optimizer, loss_scaler, name):
    """Same file layout as `util/misc.py:297-317`: `<output_dir>/<name>.pth` (the reference's driver passes
    name='checkpoint_best', `train.py:141`) holding model (trainables only) / optimizer / epoch / scaler / args."""
    output_dir = Path(args.output_dir)
    to_save = {"model": {k: v.detach().clone() for k, v in trainable_state(model_without_ddp).items()},
               "optimizer": optimizer.state_dict(), "epoch": epoch,
               "scaler": loss_scaler.state_dict() if loss_scaler is not None else None, "args": args}
    save_on_master(to_save, output_dir / f"{name}.pth")


def load_model(args, model_without_ddp, optimizer, loss_scaler):
    """`util/misc.py:323-336`: resume trainables (strict=False), optimizer, scaler, start_epoch."""
    if not getattr(args, "resume", ""):
        return
    checkpoint = torch.load(args.resume, map_location="cpu", weights_only=False)
    model_without_ddp.load_state_dict(checkpoint["model"], strict=False)
    print(f"Resume checkpoint {args.resume}")
    if "optimizer" in checkpoint and "epoch" in checkpoint and not getattr(args, "eval", False):   # `util/misc.py:331`
        optimizer.load_state_dict(checkpoint["optimizer"])
        args.start_epoch = checkpoint["epoch"] + 1
        if loss_scaler is not None and checkpoint.get("scaler") is not None:
            loss_scaler.load_state_dict(checkpoint["scaler"])


# ------------------------------------------------------------------------------------------------
# per-question-type accuracy meters of the validation loop (`util/misc.py:361-532`)
# ------------------------------------------------------------------------------------------------
_VALOR_TYPES = ("count", "temporal", "desc", "action", "loc", "rel_pos")
_MUSIC_TYPES = ("Temporal", "Existential", "Comparative", "Location", "Counting")


def get_qtype_mapping(dataset_name: str) -> dict:
    """`util/misc.py:365-412`: question-type name -> id (0 is reserved for the total)."""
    if dataset_name == "nextqa":
        return {k: i + 1 for i, k in enumerate(("CH", "CW", "TN", "TC", "TP", "DL", "DC", "DO"))}
    if dataset_name == "star":
        return {k: i + 1 for i, k in enumerate(("In", "Seq", "Pre", "Feas"))}
    if dataset_name == "valor32k":
        m = {f"{t}_{mod}": 3 * i + j + 1 for i, t in enumerate(_VALOR_TYPES) for j, mod in enumerate(("visual", "audio", "both"))}
        m.update(audio_both=19, audio_visual=20)
        return m
    if dataset_name == "musicavqa":
        return {f"{mod}_{t}": 5 * i + j + 1 for i, mod in enumerate(("Audio", "Visual", "Audio-Visual")) for j, t in enumerate(_MUSIC_TYPES)}
    return {}


def _qtype_groups(dataset_name: str):
    """(meter name, count-meter name, question-type ids) in the reference's update order (`util/misc.py:443-524`)."""
    if dataset_name == "valor32k":
        g = [("audio", [2, 5, 8, 11, 14, 17]), ("visual", [1, 4, 7, 10, 13, 16, 20]), ("both", [3, 6, 9, 12, 15, 18, 19])]
        g += [(t, [3 * i + 1, 3 * i + 2, 3 * i + 3]) for i, t in enumerate(_VALOR_TYPES)] + [("audio_second", [19, 20])]
        return g
    if dataset_name == "musicavqa":
        g = [("audio", [1, 2, 3, 4, 5]), ("visual", [6, 7, 8, 9, 10]), ("audio_visual", [11, 12, 13, 14, 15])]
        return g + [(t.lower(), [j + 1, j + 6, j + 11]) for j, t in enumerate(_MUSIC_TYPES)]
    return []


def log_qtype(data, eval, metric_logger: MetricLogger, args):
    """`util/misc.py:526-532` (+ `:414-524`): per-batch hit / count per question type -> the reference's meters (nextqa: C, T,
    D, Total; star: In, Seq, Pre, Feas, Total; valor32k / musicavqa: one score meter and one `n_*` meter per group)."""
    eps = 1e-10
    dataset = getattr(args, "dataset", None)
    qtype2id = get_qtype_mapping(dataset)
    if not qtype2id:
        return
    freq = {i: [0.0, 0.0] for i in qtype2id.values()}
    freq[0] = [0.0, 0.0]
    hits = [float(v) for v in torch.as_tensor(eval).flatten().tolist()]
    qts = [int(q) for q in torch.as_tensor(data["qtype"]).flatten().tolist()]
    for qt, v in zip(qts, hits):
        for key in (qt, 0):
            freq[key][0] += v
            freq[key][1] += 1
    tot = lambda ids: (sum(freq[i][0] for i in ids), sum(freq[i][1] for i in ids))
    ratio = lambda f: f[0] / f[1] if f[1] != 0 else 0.0                       # getCount, `:361-363`
    if dataset == "nextqa":
        for name, ids in (("C", (1, 2)), ("T", (3, 4, 5)), ("D", (6, 7, 8))):
            s, c = tot(ids)
            metric_logger.update(n=c + eps, **{name: s / (c + eps)})
        metric_logger.update(n=freq[0][1] + eps, Total=ratio(freq[0]))
    elif dataset == "star":
        for i, name in enumerate(("In", "Seq", "Pre", "Feas")):
            metric_logger.update(n=freq[i + 1][1] + eps, **{name: ratio(freq[i + 1])})
        metric_logger.update(n=freq[0][1] + eps, Total=ratio(freq[0]))
    else:
        # one update(**...) call as in `util/misc.py:479-490`: for valor32k the score of the 'count' group is passed as `count=`, i.e.
        # it binds to update()'s WEIGHT parameter (there is no 'count' meter and that batch's meters are weighted by it) - kept
        upd = {}
        for name, ids in _qtype_groups(dataset):
            s, c = tot(ids)
            upd["n_" + name] = c + eps
            upd[name] = s / (c + eps)
        metric_logger.update(**upd)
