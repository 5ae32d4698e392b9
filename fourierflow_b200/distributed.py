"""Batch sharding of the rollout over the GPUs of one box.

Samples are independent through the whole forward and rollout, parameters are replicated, and nothing is
exchanged between layers or steps.  The single collective is the all-gather of the per-sample relative-L2
values for the ``LpLoss`` mean (modules/loss.py:41-42) — n_steps x B_local floats per rank, once per
rollout.  One process per GPU; ``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests) is plumbing.
"""
from __future__ import annotations

import os
from typing import Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: str = "nccl") -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment; initialises the process group."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # keep stdout for the caller's own output
        kwargs = {}
        if backend == "nccl" and torch.cuda.is_available():
            torch.cuda.set_device(local)
            kwargs["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kwargs)
    return rank, world, local


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced slice [lo, hi) of ``n`` samples for ``rank`` (first n % world ranks get one more)."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    lo, hi = shard_bounds(t.shape[0], rank, world)
    return t[lo:hi]


def gather_sample_losses(local: torch.Tensor, n_global: int) -> torch.Tensor:
    """all-gather ``local[n_steps, B_local]`` → ``[n_steps, n_global]`` (ragged shards padded, then cut)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    n_steps = local.shape[0]
    width = -(-n_global // world)
    padded = local.new_zeros(n_steps, width)
    padded[:, :local.shape[1]] = local
    flat = local.new_empty(world * n_steps, width)          # concatenated form: accepted by NCCL and gloo
    dist.all_gather_into_tensor(flat, padded.contiguous())
    out = flat.view(world, n_steps, width)
    pieces = []
    for r in range(world):
        lo, hi = shard_bounds(n_global, r, world)
        pieces.append(out[r, :, :hi - lo])
    return torch.cat(pieces, dim=1)


class LossGatherer:
    """The loss all-gather off the compute stream: ``submit(local[n_steps, B_local])`` copies the per-sample losses
    into a preallocated (padded) send buffer and enqueues ``all_gather_into_tensor`` on a side stream that waits only
    for the kernel that produced ``local``; ``result()`` joins the side stream and returns ``[n_steps, n_global]``.
    No allocation, slicing or concatenation per call — the collective is a few hundred bytes and latency-bound, so
    what matters is that it neither blocks the next forward nor is rebuilt around it."""

    def __init__(self, n_steps: int, b_local: int, n_global: int, device: torch.device):
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.n_steps, self.n_global = n_steps, n_global
        self.width = -(-n_global // self.world)
        self.send = torch.zeros(n_steps, self.width, device=device)
        self.recv = torch.empty(self.world * n_steps, self.width, device=device)
        self.cuda = device.type == "cuda"
        self.stream = torch.cuda.Stream(device) if self.cuda else None
        self.pending = False

    def submit(self, local: torch.Tensor) -> None:
        if self.world == 1:
            self.send[:, :local.shape[1]].copy_(local)
            self.pending = True
            return
        if self.cuda:
            self.stream.wait_stream(torch.cuda.current_stream(local.device))
            with torch.cuda.stream(self.stream):
                self.send[:, :local.shape[1]].copy_(local, non_blocking=True)
                dist.all_gather_into_tensor(self.recv, self.send)
            local.record_stream(self.stream)
        else:
            self.send[:, :local.shape[1]].copy_(local)
            dist.all_gather_into_tensor(self.recv, self.send)
        self.pending = True

    def result(self) -> torch.Tensor:
        if self.cuda and self.world > 1:
            torch.cuda.current_stream(self.send.device).wait_stream(self.stream)
        if self.world == 1:
            return self.send[:, :self.n_global]
        out = self.recv.view(self.world, self.n_steps, self.width)
        if self.n_global % self.world == 0:
            return out.permute(1, 0, 2).reshape(self.n_steps, self.n_global)
        pieces = []
        for r in range(self.world):
            lo, hi = shard_bounds(self.n_global, r, self.world)
            pieces.append(out[r, :, :hi - lo])
        return torch.cat(pieces, dim=1)


def rollout_loss(local: torch.Tensor, n_global: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """(loss, step_losses) over the GLOBAL batch from per-rank per-sample losses: mean over samples per step,
    summed over steps (routines/grid_2d_markov.py:313-315)."""
    allv = gather_sample_losses(local, n_global)
    step = allv.mean(dim=1)
    return step.sum(), step


def allreduce_gradients(params, world_size: int = None, group=None) -> int:
    """Data-parallel gradient averaging for batch shards — what the reference's dormant DDP plugin would do
    (commands/train.py:83-84): ONE all-reduce per step over a flat bucket of every ``.grad`` (2 M floats for the C2
    model: latency-sized, so one bucket), then mean = sum / world.  Every rank must hold gradients for the same
    parameters in the same order.  Returns the number of elements reduced (0 when there is nothing to exchange)."""
    if not dist.is_initialized():
        return 0
    world = world_size or dist.get_world_size(group)
    grads = [p.grad for p in params if p.grad is not None]
    if world == 1 or not grads:
        return 0
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.div_(world)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
    return off
