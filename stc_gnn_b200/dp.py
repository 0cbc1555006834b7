"""Batch data-parallel plumbing for the cell stack: one process per GPU, contiguous batch
shards, no collective on the data path, ONE all-reduce of a flat fp32 gradient bucket per step.

The reference has no distributed code at all (single process, `framework/Main.py:11-12`); the cell is
independent per batch element (`B` is a pure batch index in every einsum, `framework/STC_GNN.py:37-42`), so the
only exchange is the sum of the parameter gradients -- plus `dGs [N,N]` / `dGc [C,C]` when the supports are
learned.  At SF sizes the bucket is ~32 K floats (latency-bound): a single call after the last backward cell.

Parity definition: the supports `Gs, Gc` are *inputs* of this path.  The reference's `MGP_Gen` sums its score
matrices over the batch before relu/softmax (`STC_GNN.py:231-232,239-240`), so a sharded batch changes the
supports unless those pre-relu scores are all-reduced too; that generator is outside the hot path (SURVEY §8e
caveat) -- here every rank is handed the same supports and DP gradients equal the single-process gradients of
the concatenated batch (tests/test_multi_cpu.py).

Works with any `torch.distributed` backend: NCCL over NVLink on the GPU box, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [start, stop) of `total` items for `rank`; the first `total % world` ranks get one more."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_batch(x: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """This rank's contiguous slice of a global batch (a view; dim 0 is the batch)."""
    s, e = shard_bounds(x.shape[0], rank, world)
    return x[s:e]


class GradBucket:
    """One flat fp32 buffer holding the gradients of a fixed list of tensors (parameters and, optionally,
    learned supports).  pack -> all_reduce -> unpack; absent gradients (`.grad is None`) count as zeros so that
    every rank contributes the same layout."""

    def __init__(self, tensors: Sequence[torch.Tensor]):
        self.tensors: List[torch.Tensor] = list(tensors)
        if not self.tensors:
            raise ValueError("GradBucket needs at least one tensor")
        dev = self.tensors[0].device
        for t in self.tensors:
            if t.dtype != torch.float32 or t.device != dev:
                raise RuntimeError("GradBucket tensors must be float32 and live on one device")
        self.sizes = [t.numel() for t in self.tensors]
        self.offsets = [0]
        for n in self.sizes:
            self.offsets.append(self.offsets[-1] + n)
        self.flat = torch.zeros(self.offsets[-1], dtype=torch.float32, device=dev)

    @property
    def numel(self) -> int:
        return self.offsets[-1]

    def pack(self) -> torch.Tensor:
        for t, o, n in zip(self.tensors, self.offsets, self.sizes):
            dst = self.flat[o:o + n]
            if t.grad is None:
                dst.zero_()
            else:
                dst.copy_(t.grad.reshape(-1))
        return self.flat

    def unpack(self) -> None:
        for t, o, n in zip(self.tensors, self.offsets, self.sizes):
            g = self.flat[o:o + n].view(t.shape)
            if t.grad is None:
                t.grad = g.clone()
            else:
                t.grad.copy_(g)

    def allreduce(self, group: Optional[dist.ProcessGroup] = None, average: bool = False) -> torch.Tensor:
        """Sum (or mean) the bucket over the group; returns the flat buffer. No-op outside a process group."""
        self.pack()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            if average:
                self.flat.div_(dist.get_world_size(group))
        self.unpack()
        return self.flat


def gradient_tensors(module: torch.nn.Module, extra: Iterable[torch.Tensor] = ()) -> List[torch.Tensor]:
    """Parameters of `module` that require grad (registration order = the reference's Adam order) + extras."""
    return [p for p in module.parameters() if p.requires_grad] + [t for t in extra if t is not None]


def allreduce_gradients(module: torch.nn.Module, extra: Iterable[torch.Tensor] = (),
                        group: Optional[dist.ProcessGroup] = None, average: bool = False,
                        bucket: Optional[GradBucket] = None) -> GradBucket:
    """All-reduce every gradient of `module` (+ `extra` leaves such as learned Gs, Gc) in one flat bucket.
    Pass the returned bucket back in on later steps to reuse its buffer."""
    if bucket is None:
        bucket = GradBucket(gradient_tensors(module, extra))
    bucket.allreduce(group=group, average=average)
    return bucket


class _SumOverRanks(torch.autograd.Function):
    """y = sum over ranks of x (every rank receives y).  With per-rank losses L_q that all depend on y, the gradient
    of the total loss w.r.t. this rank's x is sum_q dL_q/dy: the backward is the same all-reduce of the incoming gradient."""

    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        y = x.detach().clone()
        dist.all_reduce(y, op=dist.ReduceOp.SUM, group=group)
        return y

    @staticmethod
    def backward(ctx, dy):
        g = dy.contiguous().clone()
        dist.all_reduce(g, op=dist.ReduceOp.SUM, group=ctx.group)
        return g, None


def sum_over_ranks(x: torch.Tensor, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """Differentiable cross-rank sum for quantities the reference sums over the WHOLE batch before a non-linearity.

    The case on this path's doorstep: `MGP_Gen` builds its score matrices `Ps [N,N]`, `Pc [C,C]` with einsums that sum
    over b and t (`framework/STC_GNN.py:231, 239`) and only then applies relu / softmax, so under batch data-parallelism
    each rank would generate different supports from its shard.  Wrapping those two einsum results in `sum_over_ranks`
    (40 KB + 100 B at SF sizes) makes `Gs, Gc` -- and, through this function's backward, the generator's gradients --
    equal to the single-process global-batch values; the cells downstream need nothing else (their `dGs`, `dGc` stay
    per-rank partial sums of per-rank losses, which is exactly what the backward all-reduce here adds up).
    Identity outside a process group or at world size 1."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return x
    return _SumOverRanks.apply(x, group)
