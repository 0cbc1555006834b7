"""Batch data-parallel form of the reference's graph-pair generator (SURVEY.md §8e caveat, row f3).

`MGP_Gen.forward` (/root/reference/framework/STC_GNN.py:227-243) builds two score matrices by contracting
`tanh(alpha X Wu)` with `tanh(alpha X Wv)` over the batch AND time axes (:231, :239) and only then applies relu /
softmax / `MixedFusion`.  Under batch data-parallelism every rank would therefore generate different supports from its
shard.  `install(dp_group=...)` rebinds `MGP_Gen.forward` to `mgp_forward_dp` below:

  * forward: the per-shard scores are summed over the ranks *before* the non-linearity (40 KB + 100 B at SF sizes), so
    every rank generates the single-process global-batch `Gs`, `Gc`;
  * backward: the gradients arriving at `Gs [N,N]` / `Gc [C,C]` from this rank's cells are all-reduced right there
    (`reduce_grad`), i.e. BEFORE they flow into `MixedFusion`.  Every rank then back-propagates the same, global
    dL/dGs through its replica of the fusion layers, so the 2 x `Linear(N^2, N^2)` gradients (800 MB at N = 100,
    `STC_GNN.py:250-251`) come out identical and complete on every rank and are never communicated.  Only the
    shard-dependent parameters -- the cells, `params_S/params_C` (their gradients are partial sums over the shard's
    (b,t) rows) and `out_proj` -- go through the flat gradient bucket (`dp_bucket_parameters`).

The reference file is not edited; the module's own parameters (`params_S/params_C['Wu','Wv']`, `aggreg_S/aggreg_C`)
are used as they are, so checkpoints stay interchangeable.

Pure host-side torch on whatever device the module lives on (this is not the hot path: one call per forward).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist



def _active(group) -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


class _SumForward(torch.autograd.Function):
    """y = sum over ranks of x; the incoming gradient is already the global dL/dy on every rank (see `_ReduceGrad`),
    and dy/dx_rank = I, so backward is the identity -- times `grad_scale`: what flows on from here are the shard-
    dependent generator parameters (`params_S/C`), whose full gradient is the SUM of the per-rank pieces; when the
    job averages its gradient bucket (per-rank losses are shard means), the pieces are pre-multiplied by the world size
    so that the average is that sum."""

    @staticmethod
    def forward(ctx, x, group, grad_scale):
        ctx.grad_scale = grad_scale
        y = x.detach().clone()
        dist.all_reduce(y, op=dist.ReduceOp.SUM, group=group)
        return y

    @staticmethod
    def backward(ctx, dy):
        return (dy * ctx.grad_scale if ctx.grad_scale != 1.0 else dy), None, None


class _ReduceGrad(torch.autograd.Function):
    """Identity whose backward sums (or averages) the gradient over the ranks."""

    @staticmethod
    def forward(ctx, x, group, average):
        ctx.group, ctx.average = group, average
        return x.view_as(x)

    @staticmethod
    def backward(ctx, dy):
        g = dy.contiguous().clone()
        dist.all_reduce(g, op=dist.ReduceOp.SUM, group=ctx.group)
        if ctx.average:
            g.div_(dist.get_world_size(ctx.group))
        return g, None, None


def sum_forward(x, group=None, average=False):
    if not _active(group):
        return x
    return _SumForward.apply(x, group, float(dist.get_world_size(group)) if average else 1.0)


def reduce_grad(x, group=None, average=False):
    return _ReduceGrad.apply(x, group, average) if _active(group) else x


def pair_scores(X: torch.Tensor, Wu: torch.Tensor, Wv: torch.Tensor, alpha: float) -> torch.Tensor:
    """M - M^T with M[i,j] = sum_{b,t,h} tanh(alpha X Wu)[b,t,i,h] * tanh(alpha X Wv)[b,t,j,h]   (STC_GNN.py:229-231).

    X [B,T,I,J], Wu/Wv [J,h] -> [I,I].  The reference's second einsum ('btmh,btnh->mn' of V,U) is the transpose of the
    first, so one contraction over the flattened (b,t) axis serves both."""
    U = torch.tanh(alpha * (X @ Wu)).flatten(0, 1)
    V = torch.tanh(alpha * (X @ Wv)).flatten(0, 1)
    M = torch.einsum("rih,rjh->ij", U, V)
    return M - M.t()


def mgp_forward_dp(self, X_seq: torch.Tensor, As: torch.Tensor, Ac: torch.Tensor,
                   group: Optional[dist.ProcessGroup] = None, average: bool = False):
    """Drop-in body for `MGP_Gen.forward` with the batch-coupled sums made global over `group`.
    `average`: the per-rank losses are means over equal shards, the job's loss is their mean, and the gradient bucket of
    `dp_bucket_parameters` is AVERAGED over the ranks (`GradBucket.allreduce(average=True)`); without it per-rank losses
    add up and the bucket is summed."""
    Ss = sum_forward(pair_scores(X_seq, self.params_S["Wu"], self.params_S["Wv"], self.alpha), group, average)
    Gs = self.aggreg_S(As, torch.softmax(torch.relu(Ss), dim=-1))
    Xc = X_seq.transpose(2, 3)
    Sc = sum_forward(pair_scores(Xc, self.params_C["Wu"], self.params_C["Wv"], self.alpha), group, average)
    Gc = self.aggreg_C(Ac, torch.softmax(torch.relu(Sc), dim=-1))
    return reduce_grad(Gs, group, average), reduce_grad(Gc, group, average)


def dp_bucket_parameters(model: torch.nn.Module):
    """Parameters whose gradients are partial sums over this rank's batch shard, i.e. everything except the fusion
    layers of the generator (`mix_graph_pair.aggreg_S/aggreg_C`), whose gradients `mgp_forward_dp` already makes global.
    Hand the list to `dp.GradBucket`."""
    return [p for n, p in model.named_parameters() if p.requires_grad and ".aggreg_" not in "." + n]


def patch_generator(stc_gnn_module, group: Optional[dist.ProcessGroup] = None, average: bool = False):
    """Rebind `MGP_Gen.forward` in the (already imported) reference module; returns the previous forward."""
    cls = stc_gnn_module.MGP_Gen
    prev = cls.forward
    if not hasattr(cls, "_reference_forward"):
        cls._reference_forward = prev

    def forward(self, X_seq, As, Ac):
        return mgp_forward_dp(self, X_seq, As, Ac, group, average)

    cls.forward = forward
    return prev


def unpatch_generator(stc_gnn_module):
    cls = stc_gnn_module.MGP_Gen
    ref = getattr(cls, "_reference_forward", None)
    if ref is not None:
        cls.forward = ref
