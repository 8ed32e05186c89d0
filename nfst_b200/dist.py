"""Multi-GPU plumbing: lattices are independent, so a batch is sharded across ranks by arc
count and the only collective of a training step is ONE all-reduce of {sum logZ, dtheta}
(what Lightning DDP does implicitly for the reference, src/trainer/tr_trainer.py:80-86).

One process per GPU, ``torch.distributed`` (NCCL on GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

import heapq
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_by_arcs(arc_counts: Sequence[int], world_size: int) -> List[List[int]]:
    """Longest-processing-time greedy partition of lattice indices into ``world_size`` bins
    with near-equal total arcs.  Deterministic; every rank computes the same answer."""
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    order = sorted(range(len(arc_counts)), key=lambda i: (-int(arc_counts[i]), i))
    heap = [(0, r) for r in range(world_size)]
    bins: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        load, r = heapq.heappop(heap)
        bins[r].append(i)
        heapq.heappush(heap, (load + int(arc_counts[i]), r))
    for b in bins:
        b.sort()
    return bins


def my_shard(arc_counts: Sequence[int]) -> List[int]:
    """Indices of the lattices this rank owns (all of them without a process group)."""
    if not (dist.is_available() and dist.is_initialized()):
        return list(range(len(arc_counts)))
    return shard_by_arcs(arc_counts, dist.get_world_size())[dist.get_rank()]


class PendingReduce:
    """An all-reduce of {loss, dtheta} in flight (``async_op=True``): ``wait()`` makes the current stream wait for
    it and returns ``(loss, dtheta)``."""

    def __init__(self, flat: torch.Tensor, work, loss_shape, dtheta_shape):
        self.flat, self.work, self.loss_shape, self.dtheta_shape = flat, work, loss_shape, dtheta_shape

    def wait(self) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        if self.work is not None:
            self.work.wait()
            self.work = None
        loss = self.flat[0].reshape(self.loss_shape)
        return loss, (self.flat[1:].reshape(self.dtheta_shape) if self.dtheta_shape is not None else None)


def all_reduce_loss_and_grad(loss_sum: torch.Tensor, dtheta: Optional[torch.Tensor] = None, *, async_op: bool = False):
    """Sum ``loss_sum`` (scalar) and ``dtheta`` ([V]) over ranks with a single all-reduce of
    one flat buffer, enqueued right behind the fused backward.  Returns ``(loss, dtheta)``; with
    ``async_op=True`` a ``PendingReduce`` instead (the all-reduce runs on the process group's own stream, the
    caller's stream goes on with the next step and calls ``wait()`` when it needs the sums)."""
    solo = not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1
    parts = [loss_sum.reshape(1).to(torch.float32)]
    if dtheta is not None:
        parts.append(dtheta.reshape(-1).to(torch.float32))
    if solo and not async_op:
        return loss_sum, dtheta
    flat = torch.cat(parts)
    work = None if solo else dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=async_op)
    pending = PendingReduce(flat, work if async_op else None, loss_sum.shape, dtheta.shape if dtheta is not None else None)
    return pending if async_op else pending.wait()


def gather_ragged(local: torch.Tensor, indices: Sequence[int], total: int) -> torch.Tensor:
    """Reassemble a per-lattice result ([n_local]) into batch order ([total]) on every rank
    (decode / Viterbi scores; small)."""
    out = torch.zeros(total, dtype=local.dtype, device=local.device)
    if len(indices):
        out[torch.as_tensor(list(indices), device=local.device)] = local
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(out, op=dist.ReduceOp.SUM)
    return out
