"""CPU, world_size 2, gloo: the data-parallel gradient exchange (`flipped_vqa_b200/dp.py`, replacing the
DDP reducer of `train.py:115-117`). Each rank fills its flat gradient buffer the way the hand-written
backward does (adapter rows layer by layer, last layer first; then the late tail), calling
`GradSync.layer_done` / `finish`; the result must equal the mean over ranks, every element exactly once."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

L, A, D, H, VD, F = 5, 10, 64, 4, 24, 10


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank_grads(rank):
    from flipped_vqa_b200.step import GradBuffers
    gb = GradBuffers(L, A, D, H, VD, F, "cpu")
    g = torch.Generator().manual_seed(100 + rank)
    return gb, torch.randn(gb.flat.numel(), generator=g)


def _fill(gb, sync, vals):
    """One backward pass as the hand-written step drives GradSync; returns what autograd would add to `.grad`."""
    row = A * D
    for l in range(L - 1, -1, -1):                           # backward order
        gb.flat[l * row:(l + 1) * row] = vals[l * row:(l + 1) * row]
        sync.layer_done(l)
    gb.flat[gb.late_offset:] = vals[gb.late_offset:]          # gates, visual_proj, temporal_emb: final at the end
    sync.finish()
    return gb.flat.clone()


def _accum_worker(rank, world, port, chunk, q):
    """accum_iter = 3: two micro-steps with require_backward_grad_sync False, then the boundary micro-step."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from flipped_vqa_b200.dp import GradSync
        gb, _ = _rank_grads(rank)
        sync = GradSync(gb, L, A, D, chunk_layers=chunk)
        grad = torch.zeros_like(gb.flat)                      # the parameters' .grad, accumulated by autograd
        for micro in range(3):
            g = torch.Generator().manual_seed(1000 * micro + rank)
            sync.enabled = micro == 2
            grad += _fill(gb, sync, torch.randn(gb.flat.numel(), generator=g))
        q.put((rank, grad, sync.messages))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("chunk", [2, 8])
def test_grad_sync_skips_non_boundary_micro_steps(chunk):
    """with accum_iter > 1 only the boundary micro-step talks; .grad ends as the rank mean of the accumulated
    gradient - what the reference's every-micro-step DDP reduce (`engine.py:37-41`) produces, with 1/accum_iter of the messages."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_accum_worker, args=(r, world, port, chunk, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n = got[0][1].numel()
    expect = sum(torch.randn(n, generator=torch.Generator().manual_seed(1000 * m + r)) for m in range(3) for r in range(world)) / world
    for rank, grad, messages in got:
        assert torch.allclose(grad, expect, atol=1e-5), f"rank {rank}"
        assert messages == -(-L // chunk) + 1                 # ONE round of messages for three micro-steps


def _worker(rank, world, port, chunk, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from flipped_vqa_b200.dp import GradSync
        gb, vals = _rank_grads(rank)
        sync = GradSync(gb, L, A, D, chunk_layers=chunk)
        row = A * D
        for l in range(L - 1, -1, -1):                       # backward order
            gb.flat[l * row:(l + 1) * row] = vals[l * row:(l + 1) * row]
            sync.layer_done(l)
        gb.flat[gb.late_offset:] = vals[gb.late_offset:]      # gates, visual_proj, temporal_emb: final at the end
        sync.finish()
        q.put((rank, gb.flat.clone(), sync.messages))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("chunk", [1, 2, 8])
def test_grad_sync_mean_over_two_ranks(chunk):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, chunk, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = sum(_rank_grads(r)[1] for r in range(world)) / world
    for rank, flat, messages in got:
        assert torch.allclose(flat, expect, atol=1e-6), f"rank {rank}: all-reduced gradients differ from the mean"
        assert messages == -(-L // chunk) + 1                 # adapter chunks + one late message


def test_grad_sync_is_noop_for_one_rank():
    port = _free_port()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK="0", WORLD_SIZE="1")
    dist.init_process_group("gloo", rank=0, world_size=1)
    try:
        from flipped_vqa_b200.dp import GradSync
        gb, vals = _rank_grads(0)
        gb.flat.copy_(vals)
        sync = GradSync(gb, L, A, D)
        for l in range(L - 1, -1, -1):
            sync.layer_done(l)
        sync.finish()
        assert torch.equal(gb.flat, vals) and sync.messages == 0
    finally:
        dist.destroy_process_group()
