"""Data-parallel training across the GPUs of one box (replaces `train.py:115-117` DDP + the implicit
reducer all-reduce).

The batch shards by sample, every rank holds the frozen base, and the ONLY exchange is the all-reduce
(mean) of the trainable gradients: 4 499 456 fp32 = 18 MB for 7B. `GradSync` issues it from inside
the hand-written backward, on NCCL's stream, overlapped with the remaining layers:
  * adapter-prompt gradient rows become final layer by layer (last layer first) -> reduced in chunks
    of `chunk_layers` layers while backward continues;
  * gates + visual_proj + temporal_emb are one contiguous tail of the flat buffer, final only after
    layer 0 -> one message at the end.
The reference's DDP reduces on EVERY micro-step (no `no_sync()` under accum_iter, `engine.py:37-41`). Here the reduce can be
skipped on non-boundary micro-steps (`DataParallel.require_backward_grad_sync = False`, the attribute torch's DDP `no_sync()`
toggles; `engine.train_one_epoch` sets it from `update_grad`): the local gradient of a skipped micro-step is remembered, and
the boundary micro-step reduces (remembered + own) and returns mean(remembered + own) - remembered, so that `.grad`
(= remembered + returned, accumulated by autograd) ends up as the rank mean of the accumulated gradient - the same value as
averaging every micro-step (mean is linear), with 1/accum_iter of the messages.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist


class GradSync:
    def __init__(self, grad_buffers, n_layers_run: int, adapter_len: int, dim: int, group=None, chunk_layers: int = 8, chunk_group=None):
        self.gb = grad_buffers
        self.L, self.A, self.d = n_layers_run, adapter_len, dim
        self.group = group
        # communicator of the EARLY (overlapped) chunk messages; None = the same as the late message's. DataParallel passes a second NCCL
        # communicator limited to a couple of CTAs: an overlapped all-reduce only has to finish before the end of backward, and every
        # SM it holds is one the persistent GEMMs (sized for all 148) cannot use
        self.chunk_group = chunk_group if chunk_group is not None else group
        self.world = dist.get_world_size(group)
        self.chunk = max(1, chunk_layers)
        self.works: List = []
        self.messages = 0
        self.enabled = True            # False: this micro-step's gradients stay local (accumulation, see module docstring)
        self.acc = None                # sum of the local flat gradients of the skipped micro-steps since the last reduce
        self.timing = None             # list -> (event before, event after) the stream-level wait on NCCL in finish() per step

    def _reduce(self, t: torch.Tensor, group=None):
        self.works.append(dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group if group is not None else self.group, async_op=True))
        self.messages += 1

    def warm_up(self, rounds: int = 3):
        """Set-up, not a step: run the step's exact message pattern (chunk sizes, tail) on a scratch buffer so that NCCL's
        lazily built channels / proxy connections exist before the first real backward (the first all-reduces of a size
        class otherwise cost milliseconds on one rank and every other rank waits for it)."""
        if self.world == 1:
            return
        scratch = torch.zeros_like(self.gb.flat)
        for _ in range(rounds):
            works = []
            for l in range(self.L - 1, -1, -1):
                if l % self.chunk == 0:
                    hi = min(l + self.chunk, self.L)
                    works.append(dist.all_reduce(scratch[l * self.A * self.d: hi * self.A * self.d], group=self.chunk_group, async_op=True))
            works.append(dist.all_reduce(scratch[self.gb.late_offset:], group=self.group, async_op=True))
            for w in works:
                w.wait()
        torch.cuda.synchronize()

    def layer_done(self, l: int):
        """Called by the backward pass right after layer l's adapter gradient rows were written."""
        if self.world == 1 or not self.enabled:
            return
        if l % self.chunk == 0:                                   # layers [l, min(l+chunk, L)) are final
            hi = min(l + self.chunk, self.L)
            lo_, hi_ = l * self.A * self.d, hi * self.A * self.d
            if self.acc is not None:
                self.gb.flat[lo_:hi_].add_(self.acc[lo_:hi_])
            self._reduce(self.gb.flat[lo_:hi_], self.chunk_group)

    def finish(self):
        """Late message (gates, visual_proj, temporal_emb), wait for everything, turn the sum into a mean."""
        if self.world == 1:
            return
        if not self.enabled:                                      # accumulate locally, no message
            if self.acc is None:
                self.acc = self.gb.flat.clone()
            else:
                self.acc.add_(self.gb.flat)
            return
        late = self.gb.late_offset
        if self.acc is not None:
            self.gb.flat[late:].add_(self.acc[late:])
        self._reduce(self.gb.flat[late:])
        if self.timing is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        for w in self.works:
            w.wait()                                              # stream-level wait for NCCL, not a host sync
        if self.timing is not None:
            e1.record()
            self.timing.append((e0, e1))
        self.works = []
        self.gb.flat.mul_(1.0 / self.world)
        if self.acc is not None:                                  # .grad already holds the skipped micro-steps' local gradients
            self.gb.flat.sub_(self.acc)
            self.acc = None


class DataParallel(torch.nn.Module):
    """`model = DataParallel(model)`; exposes `.module` like torch's DDP so `train.py:117`,
    `util/misc.py:297-317` keep working. Broadcasts the trainable parameters from rank 0 once (C2)."""

    def __init__(self, module, group=None, chunk_layers: int = 8, broadcast: bool = True, chunk_ctas: int = 0):
        super().__init__()
        self.module = module
        self.group = group
        self.chunk_layers = chunk_layers
        # optional second NCCL communicator for the overlapped chunk messages, capped at `chunk_ctas` CTAs (0 = default: use `group` for
        # everything). Measured at 2 GPUs (profiles/r2_ab_chunk_ctas.txt): caps of 1 / 2 CTAs do not reduce the slow-down of the overlapped GEMMs.
        self.chunk_group = None
        if chunk_ctas > 0 and dist.is_initialized() and dist.get_world_size(group) > 1 and dist.get_backend(group) == "nccl":
            try:
                opts = dist.ProcessGroupNCCL.Options()
                opts.config.max_ctas = int(chunk_ctas)
                opts.config.min_ctas = 1
                self.chunk_group = dist.new_group(ranks=(dist.get_process_group_ranks(group) if group is not None else None), pg_options=opts)
            except Exception:                                 # older torch / NCCL without communicator config: one communicator
                self.chunk_group = None
        self.require_backward_grad_sync = True                  # same attribute as torch DDP (`no_sync()` clears it)
        if broadcast and dist.is_initialized() and dist.get_world_size(group) > 1:
            for p in module.parameters():
                if p.requires_grad:
                    dist.broadcast(p.data, src=0, group=group)

    def _attach(self):
        m = self.module
        m._ensure_packed()
        if m.grad_sync is None or m.grad_sync.gb is not m._grad_buffers:
            m.grad_sync = GradSync(m._grad_buffers, len(m.run_layers()), m.adapter_len, m.params.dim, self.group, self.chunk_layers,
                                   chunk_group=self.chunk_group)
            m._engine.adapter_grad_chunk = self.chunk_layers    # adapter gradient rows become final in the chunks GradSync reduces
            if m._grad_buffers.flat.is_cuda:
                m.grad_sync.warm_up()
        m.grad_sync.enabled = bool(self.require_backward_grad_sync)

    def no_sync(self):
        """Context manager with torch DDP's semantics: backward passes inside it keep their gradients local."""
        import contextlib

        @contextlib.contextmanager
        def ctx():
            prev, self.require_backward_grad_sync = self.require_backward_grad_sync, False
            try:
                yield
            finally:
                self.require_backward_grad_sync = prev
        return ctx()

    def forward(self, data, inference: bool = False):
        self._attach()
        return self.module(data, inference=inference)

    def forward_plan(self, plan):
        self._attach()
        return self.module.forward_plan(plan)

    def plan_batch(self, data, inference: bool = False):
        return self.module.plan_batch(data, inference)
