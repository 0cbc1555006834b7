"""Spatial row-partition of a constant CSR support with per-hop halo exchange (BASELINE config 4).

The spatial mode product `Y[b,m,:] = sum_n Gs[n,m] X[b,n,:]` (`framework/STC_GNN.py:37`) is independent per
output node given the features of that node's in-neighbours.  Nodes are split in `world` contiguous blocks
(callers order nodes so that blocks are spatially compact, e.g. Morton / strip order); rank r owns block r of
every [B, N, C, L] tensor.  One Chebyshev hop = exchange the boundary-node slabs (`halo` rows of width B*C*L)
with the ranks that own them, then apply the local operator whose columns are renumbered into
`[local nodes | halo nodes]`.  The categorical mix and the gate contraction are node-local and never communicate.
The adjoint (backward: `dX[b,n,:] = sum_m Gs[n,m] dY[b,m,:]`) is the same scheme on the un-transposed graph; both
directions use one halo set (the union of what either needs) so that forward and backward share buffer layouts.
Parameter gradients are partial sums over a rank's nodes and are all-reduced once per cell backward.

Everything here is host logic + `torch.distributed` plumbing (NCCL all-to-all over NVSwitch on the GPU box, gloo in
the CPU tests); the arithmetic is `stc_support_apply` of libstc_b200.so.  There is no CPU arithmetic path: on
non-CUDA tensors `apply` requires the caller to inject `apply_fn` (the tests inject a CPU checker).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional

import torch
import torch.distributed as dist

from .dp import shard_bounds


@dataclass
class HaloPlan:
    """Exchange + local operator of ONE direction for ONE rank.

    Operator rows are this rank's output nodes (local numbering 0..nloc-1); its columns index the extended input
    `[local nodes (nloc) | halo nodes (nhalo)]`, halo nodes ordered by owner rank, then by global id.
    """
    rank: int
    world: int
    start: int                      # first global node of this rank's block
    nloc: int
    nhalo: int
    send_idx: List[torch.Tensor]    # per peer: LOCAL indices of the rows this rank sends (sorted by global id)
    recv_counts: List[int]          # per peer: rows received
    halo_global: torch.Tensor       # [nhalo] global ids of the halo nodes, in extended order
    op_row: torch.Tensor            # COO of the local operator: output node (local)
    op_col: torch.Tensor            #                            input node (extended numbering)
    op_val: torch.Tensor

    @property
    def next(self) -> int:
        return self.nloc + self.nhalo

    @property
    def send_counts(self) -> List[int]:
        return [int(s.numel()) for s in self.send_idx]


def _csr_to_coo(rowptr: torch.Tensor, col: torch.Tensor):
    rowptr = rowptr.long().cpu()
    counts = rowptr[1:] - rowptr[:-1]
    row = torch.repeat_interleave(torch.arange(rowptr.numel() - 1), counts)
    return row, col.long().cpu()


def build_plan(out_idx: torch.Tensor, in_idx: torch.Tensor, vals: torch.Tensor, num_nodes: int, rank: int,
               world: int, symmetric_halo: bool = False) -> HaloPlan:
    """Plan for the operator `Y[out] += val * X[in]` given as global COO triplets (host tensors).

    Deterministic and communication-free: every rank derives what each peer needs from the replicated graph.
    `symmetric_halo`: the halo set is the one of the operator AND of its transpose (a superset of what this operator
    reads), so that the plans of both directions share one extended node set -- the backward of the partitioned cell
    runs its adjoint hops on buffers laid out for the forward.  For a structurally symmetric graph nothing is added.
    """
    out_idx, in_idx, vals = out_idx.long().cpu(), in_idx.long().cpu(), vals.float().cpu()
    pair_out, pair_in = (torch.cat([out_idx, in_idx]), torch.cat([in_idx, out_idx])) if symmetric_halo else (out_idx, in_idx)
    bounds = [shard_bounds(num_nodes, p, world) for p in range(world)]
    starts = torch.tensor([b[0] for b in bounds] + [num_nodes])
    owner_of = lambda idx: torch.bucketize(idx, starts[1:], right=True)   # block index of each global node
    out_owner = owner_of(out_idx)
    pair_out_owner, pair_in_owner = owner_of(pair_out), owner_of(pair_in)

    def halo_of(p: int) -> torch.Tensor:   # sorted global ids rank p needs from other ranks
        need = pair_in[(pair_out_owner == p) & (pair_in_owner != p)]
        return torch.unique(need)          # sorted ascending => grouped by owner (blocks are contiguous)

    s, e = bounds[rank]
    my_halo = halo_of(rank)
    my_halo_owner = owner_of(my_halo)
    recv_counts = [int((my_halo_owner == p).sum()) for p in range(world)]
    send_idx = []
    for p in range(world):
        if p == rank:
            send_idx.append(torch.empty(0, dtype=torch.long))
            continue
        hp = halo_of(p)
        send_idx.append(hp[(hp >= s) & (hp < e)] - s)
    mine = out_owner == rank
    r_loc = out_idx[mine] - s
    c_glob = in_idx[mine]
    c_local = (c_glob >= s) & (c_glob < e)
    pos = torch.searchsorted(my_halo, c_glob.clamp(min=0))      # extended slot of halo columns
    c_ext = torch.where(c_local, c_glob - s, (e - s) + pos)
    return HaloPlan(rank=rank, world=world, start=s, nloc=e - s, nhalo=int(my_halo.numel()), send_idx=send_idx,
                    recv_counts=recv_counts, halo_global=my_halo, op_row=r_loc, op_col=c_ext, op_val=vals[mine])


def exchange(plan: HaloPlan, X_local: torch.Tensor, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """[B, nloc, W] -> [B, nloc + nhalo, W]: append the halo rows received from their owners.

    One all-to-all of packed rows (`[rows, B*W]`, row-major so that split sizes count rows).  With a single rank
    (or no halo anywhere) it degenerates to a copy."""
    B, nloc, W = X_local.shape
    assert nloc == plan.nloc, (nloc, plan.nloc)
    send_rows = [X_local[:, idx.to(X_local.device)].permute(1, 0, 2).reshape(-1, B * W) for idx in plan.send_idx]
    send = torch.cat(send_rows, dim=0).contiguous() if send_rows else X_local.new_empty(0, B * W)
    recv = X_local.new_empty(plan.nhalo, B * W)
    if plan.world > 1:
        dist.all_to_all_single(recv, send, output_split_sizes=plan.recv_counts, input_split_sizes=plan.send_counts,
                               group=group)
    halo = recv.view(plan.nhalo, B, W).permute(1, 0, 2)
    return torch.cat([X_local, halo], dim=1).contiguous()


def exchange_into(plan: HaloPlan, X_ext: torch.Tensor, group: Optional[dist.ProcessGroup] = None) -> None:
    """In-place form of `exchange` on an extended tensor [B, nloc + nhalo, W]: only the boundary rows are packed and
    only the halo rows are written (the local block is never copied)."""
    B, next_, W = X_ext.shape
    assert next_ == plan.next, (next_, plan.next)
    if plan.world == 1:     # a property every rank agrees on: with peers, every rank enters the collective below
        return
    send_rows = [X_ext[:, idx.to(X_ext.device)].permute(1, 0, 2).reshape(-1, B * W) for idx in plan.send_idx]
    send = torch.cat(send_rows, dim=0).contiguous() if send_rows else X_ext.new_empty(0, B * W)
    recv = X_ext.new_empty(plan.nhalo, B * W)
    if plan.world > 1:
        dist.all_to_all_single(recv, send, output_split_sizes=plan.recv_counts, input_split_sizes=plan.send_counts,
                               group=group)
    X_ext[:, plan.nloc:] = recv.view(plan.nhalo, B, W).permute(1, 0, 2)


class PartitionedSupport:
    """A constant CSR spatial support `Gs` [N, N] row-partitioned over `world` ranks.

    `rowptr, col, vals` is the GLOBAL `Gs` in CSR on the host (replicated; N = 65,536, nnz ~ 0.5 M is 6 MB).
    `fwd` is the plan of the forward mode product (operator Gs^T: output m, input n for every Gs[n,m] != 0),
    `bwd` the plan of its adjoint (operator Gs).
    """

    def __init__(self, rowptr: torch.Tensor, col: torch.Tensor, vals: torch.Tensor, num_nodes: int, rank: int,
                 world: int, group: Optional[dist.ProcessGroup] = None):
        n_idx, m_idx = _csr_to_coo(rowptr, col)      # Gs[n, m]
        self.N, self.rank, self.world, self.group = int(num_nodes), rank, world, group
        # one extended node set [local | halo] for both directions (see build_plan)
        self.fwd = build_plan(m_idx, n_idx, vals, num_nodes, rank, world, symmetric_halo=True)
        self.bwd = build_plan(n_idx, m_idx, vals, num_nodes, rank, world, symmetric_halo=True)
        assert self.fwd.nhalo == self.bwd.nhalo and torch.equal(self.fwd.halo_global, self.bwd.halo_global)
        self.start, self.nloc = self.fwd.start, self.fwd.nloc
        # STC_Cell / RecurrentStack accept this object as `Gs` (blocks in, blocks out).  True: every cell backward
        # all-reduces its parameter gradients; False: the caller reduces once per step (dp.allreduce_gradients).
        self.reduce_per_cell = True
        self._dev = {}

    @classmethod
    def from_dense(cls, G: torch.Tensor, rank: int, world: int, group=None) -> "PartitionedSupport":
        s = G.detach().cpu().to_sparse_csr()
        return cls(s.crow_indices(), s.col_indices(), s.values(), G.shape[0], rank, world, group)

    def local_slice(self, X: torch.Tensor, node_dim: int = 1) -> torch.Tensor:
        """This rank's node block of a global tensor."""
        return X.narrow(node_dim, self.start, self.nloc)

    # -- device form of the local operators (square, padded with empty rows so the square C-ABI kernel applies) --
    def device_support(self, direction: str, device):
        """CsrSupport G_loc with `support_apply(G_loc, X_ext, transpose=(direction == 'fwd'))[:, :nloc]` == operator."""
        key = (direction, str(device))
        if key not in self._dev:
            from .support import CsrSupport
            plan = self.fwd if direction == "fwd" else self.bwd
            n = plan.next
            # fwd: operator A_f[m_loc, n_ext] is applied as G_loc^T with G_loc[n_ext, m_loc]; bwd: G_loc = A_b
            r, c = (plan.op_col, plan.op_row) if direction == "fwd" else (plan.op_row, plan.op_col)
            g = torch.sparse_coo_tensor(torch.stack([r, c]), plan.op_val, size=(n, n)).coalesce().to_sparse_csr()
            self._dev[key] = CsrSupport(g.crow_indices().to(device), g.col_indices().to(device),
                                        g.values().to(device), n)
        return self._dev[key]

    def apply(self, X_local: torch.Tensor, direction: str = "fwd", alpha: float = 1.0, beta: float = 0.0,
              Z_local: Optional[torch.Tensor] = None,
              apply_fn: Optional[Callable[[HaloPlan, torch.Tensor], torch.Tensor]] = None) -> torch.Tensor:
        """One hop on this rank's block: `alpha * A X + beta * Z`, A = Gs^T ('fwd') or Gs ('bwd').

        X_local, Z_local: [B, nloc, ...feature axes]; returns the same shape.  `apply_fn(plan, X_ext) -> [B, nloc, W]`
        replaces the CUDA kernel (tests inject a CPU checker; there is no built-in CPU arithmetic)."""
        plan = self.fwd if direction == "fwd" else self.bwd
        shape = X_local.shape
        B = shape[0]
        X3 = X_local.reshape(B, plan.nloc, -1)
        X_ext = exchange(plan, X3, self.group)
        if apply_fn is not None:
            Y = apply_fn(plan, X_ext)
            Y = alpha * Y + (beta * Z_local.reshape(B, plan.nloc, -1) if beta != 0.0 else 0.0)
            return Y.reshape(shape)
        if not X_local.is_cuda:
            raise RuntimeError("PartitionedSupport.apply needs CUDA tensors (there is no CPU path)")
        from .support import support_apply
        Z_ext = None
        if beta != 0.0:
            Z3 = Z_local.reshape(B, plan.nloc, -1)
            Z_ext = torch.cat([Z3, Z3.new_zeros(B, plan.nhalo, Z3.shape[-1])], dim=1)
        Y_ext = support_apply(self.device_support(direction, X_local.device), X_ext, transpose=(direction == "fwd"),
                              alpha=alpha, beta=beta, Z=Z_ext)
        return Y_ext[:, :plan.nloc].reshape(shape)

    # -- device-side exchange state: packed index lists, interior / boundary row lists, persistent buffers --
    def _device_state(self, direction: str, device):
        key = ("xchg", direction, str(device))
        st = self._dev.get(key)
        if st is None:
            plan = self.fwd if direction == "fwd" else self.bwd
            send = torch.cat(plan.send_idx) if plan.send_idx else torch.empty(0, dtype=torch.long)
            has_halo_col = torch.zeros(plan.nloc, dtype=torch.bool)
            has_halo_col[plan.op_row[plan.op_col >= plan.nloc]] = True
            interior = torch.nonzero(~has_halo_col).flatten()
            # boundary pass: local rows that read a halo column + the halo rows themselves (empty operator rows: they
            # come out as beta * Z, i.e. finite -- the node-local stages run over the extended node set)
            boundary = torch.cat([torch.nonzero(has_halo_col).flatten(), torch.arange(plan.nloc, plan.next)])
            st = self._dev[key] = {
                "send_idx": send.to(torch.int32).to(device), "nsend": int(send.numel()),
                "interior": interior.to(torch.int32).to(device), "boundary": boundary.to(torch.int32).to(device),
                "buffers": {}, "side": torch.cuda.Stream(device=device), "ready": torch.cuda.Event(), "done": torch.cuda.Event(),
            }
        return st

    def _buffers(self, st, plan, slab: int, device):
        buf = st["buffers"].get(slab)
        if buf is None:
            if len(st["buffers"]) > 4:
                st["buffers"].clear()
            buf = st["buffers"][slab] = (torch.empty(st["nsend"], slab, dtype=torch.float32, device=device),
                                         torch.empty(plan.nhalo, slab, dtype=torch.float32, device=device))
        return buf

    def hop_ext(self, X_ext: torch.Tensor, out_ext: torch.Tensor, Z_ext: Optional[torch.Tensor] = None,
                alpha: float = 1.0, beta: float = 0.0, direction: str = "fwd") -> None:
        """One hop on EXTENDED tensors [B, nloc + nhalo, ...] (CUDA): `out_ext = alpha * A X_ext + beta * Z_ext` with the
        local operator (square over the extended index set; the halo nodes have empty operator rows).

        The halo rows of `X_ext` are refreshed in place from their owners ON A SIDE STREAM -- `stc_halo_pack` of the
        boundary rows into a persistent send buffer, one NCCL all-to-all of packed rows, `stc_halo_unpack` into the halo
        rows -- while the main stream already computes the INTERIOR output rows (those whose neighbours are all local,
        `stc_support_apply_rows`); the main stream then waits for the exchange and computes the boundary rows.  No
        tensor is copied, sliced or allocated per hop."""
        from . import _lib
        from .support import support_apply
        plan = self.fwd if direction == "fwd" else self.bwd
        B = X_ext.shape[0]
        X3 = X_ext.view(B, plan.next, -1)
        W = X3.shape[-1]
        out3 = out_ext.view(B, plan.next, -1)
        Z3 = None if Z_ext is None else Z_ext.view(B, plan.next, -1)
        sup = self.device_support(direction, X_ext.device)
        tr = direction == "fwd"
        if plan.world == 1:
            support_apply(sup, X3, transpose=tr, alpha=alpha, beta=beta, Z=Z3, out=out3)
            return
        lib = _lib.load()
        st = self._device_state(direction, X_ext.device)
        send, recv = self._buffers(st, plan, B * W, X_ext.device)
        main, side = torch.cuda.current_stream(), st["side"]
        st["ready"].record(main)                     # X_ext's local rows are final once the main stream gets here
        with torch.cuda.stream(side):
            side.wait_event(st["ready"])
            _lib.check(lib.stc_halo_pack(X3.data_ptr(), plan.next * W, W, B, st["send_idx"].data_ptr(), st["nsend"],
                                         send.data_ptr(), side.cuda_stream), "stc_halo_pack")
            dist.all_to_all_single(recv, send, output_split_sizes=plan.recv_counts, input_split_sizes=plan.send_counts,
                                   group=self.group)
            _lib.check(lib.stc_halo_unpack(recv.data_ptr(), W, B, plan.nloc, plan.nhalo, X3.data_ptr(), plan.next * W,
                                           side.cuda_stream), "stc_halo_unpack")
            st["done"].record(side)
        if st["interior"].numel():
            support_apply(sup, X3, transpose=tr, alpha=alpha, beta=beta, Z=Z3, out=out3, rows=st["interior"])
        main.wait_event(st["done"])
        support_apply(sup, X3, transpose=tr, alpha=alpha, beta=beta, Z=Z3, out=out3, rows=st["boundary"])

    def spatial_terms(self, X_local: torch.Tensor, Ks: int, apply_fn=None) -> List[torch.Tensor]:
        """Feature-side Chebyshev terms Y_0..Y_{Ks-1} of this rank's block (Y_1 = Gs^T Y_0, Y_k = 2 Gs^T Y_{k-1} - Y_{k-2};
        `framework/STC_GNN.py:24-29` applied on the feature side): Ks-1 hops, one halo exchange each."""
        terms = [X_local]
        for k in range(1, Ks):
            if k == 1:
                terms.append(self.apply(terms[0], "fwd", apply_fn=apply_fn))
            else:
                terms.append(self.apply(terms[k - 1], "fwd", alpha=2.0, beta=-1.0, Z_local=terms[k - 2], apply_fn=apply_fn))
        return terms


def adjoint_chain_ext(ps: PartitionedSupport, ybar: List[torch.Tensor], hop: Optional[Callable] = None) -> torch.Tensor:
    """Reverse of the feature-side recurrence (Y_1 = Gs^T Y_0, Y_k = 2 Gs^T Y_{k-1} - Y_{k-2}) on EXTENDED adjoints.

    `ybar[k]` [B, nloc + nhalo, ...] holds dL/dY_k of this rank's nodes in its local rows (halo rows: anything).  In
    place, for k = Ks-1 .. 1:  ybar[k-1] += (2 if k >= 2 else 1) * Gs ybar[k]  (one halo exchange of ybar[k], adjoint
    plan) and ybar[k-2] -= ybar[k] (local rows only).  Returns ybar[0], whose local rows are dL/dY_0.
    `hop(X_ext, out_ext, Z_ext, alpha, beta, direction)` replaces `ps.hop_ext` (the CPU tests inject a checker)."""
    hop = hop or (lambda X, out, Z, alpha, beta, direction: ps.hop_ext(X, out, Z_ext=Z, alpha=alpha, beta=beta,
                                                                       direction=direction))
    n = ps.nloc
    for k in range(len(ybar) - 1, 0, -1):
        hop(ybar[k], ybar[k - 1], ybar[k - 1], 2.0 if k >= 2 else 1.0, 1.0, "bwd")
        if k >= 2:
            ybar[k - 2][:, :n].sub_(ybar[k][:, :n])
    return ybar[0]


class _PartitionedCell(torch.autograd.Function):
    """One STC cell on this rank's node block; forward and backward are the C-ABI stages of include/stc_b200.h with
    the spatial hops (halo exchange + `stc_support_apply`) in between."""

    @staticmethod
    def forward(ctx, Gc, Xt, Ht_1, Wg, bg, Wc, bc, cfg):
        from . import _lib
        from .cell import _activation_code, _ptr
        ps, Ks, Kc, activation, reduce_params = cfg
        lib = _lib.load()
        B, n, C, Din = Xt.shape
        h = Ht_1.shape[-1]
        # Everything lives in the EXTENDED node set [local | halo]: the node-local stages simply run over
        # nloc + nhalo nodes (the halo rows, ~1 % at k = 8, are recomputed garbage that the next exchange overwrites),
        # so the spatial terms are written straight into the `saved` regions the stages read -- no slicing, no compaction.
        ne = ps.fwd.next
        Gc = Gc.detach().contiguous()
        Wg, Wc = Wg.detach().contiguous(), Wc.detach().contiguous()
        bg = bg.detach().contiguous() if bg is not None else None
        bc = bc.detach().contiguous() if bc is not None else None
        Xe = Xt.new_zeros(B, ne, C, Din)
        Xe[:, :n] = Xt.detach()
        He = Ht_1.new_zeros(B, ne, C, h)
        He[:, :n] = Ht_1.detach()
        dims = _lib.StcDims(B, ne, C, Din, h, Ks, Kc, _activation_code(activation), 1 if bg is not None else 0)
        lay = _lib.saved_layout(dims)
        saved = torch.empty(lib.stc_cell_saved_bytes(dims) // 4, dtype=torch.float32, device=Xt.device)
        Hn = torch.empty_like(He)
        R = B * ne * C
        stream = torch.cuda.current_stream().cuda_stream

        def region(name, k, width):
            o = lay[name] + k * R * width
            return saved[o:o + R * width].view(B, ne, C, width)

        def stage(which):
            _lib.check(lib.stc_cell_fwd_stage(dims, which, Gc.data_ptr(), Xe.data_ptr(), ne * C * Din, He.data_ptr(),
                                              Wg.data_ptr(), _ptr(bg), Wc.data_ptr(), _ptr(bc), Hn.data_ptr(),
                                              saved.data_ptr(), saved.numel() * 4, stream), "stc_cell_fwd_stage")
            _lib.note_launches()

        def chain(term0, name, width, first):
            """Y_1 = Gs^T Y_0, Y_k = 2 Gs^T Y_{k-1} - Y_{k-2} into regions `name`[first ...]."""
            prev2, prev = None, term0
            for k in range(1, Ks):
                out = region(name, first + k - 1, width)
                if k == 1:
                    ps.hop_ext(prev, out)
                else:
                    ps.hop_ext(prev, out, Z_ext=prev2, alpha=2.0, beta=-1.0)
                prev2, prev = prev, out

        chain(Xe, "Yx", Din, 0)
        chain(He, "Yh", h, 0)
        stage(_lib.STAGE_GATES)
        chain(region("Yr", 0, h), "Yr", h, 1)
        stage(_lib.STAGE_CANDI)
        ctx.cfg, ctx.dims = cfg, dims
        ctx.has_bias = bg is not None
        ctx.keep = (Gc, Xe, He, Wg, Wc, saved)
        return Hn[:, :n].contiguous()

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dHn):
        from . import _lib
        ps, Ks, Kc, activation, reduce_params = ctx.cfg
        Gc, Xe, He, Wg, Wc, saved = ctx.keep
        dims = ctx.dims
        lib = _lib.load()
        B, ne, C, Din = Xe.shape
        h = He.shape[-1]
        n = ps.nloc
        dev = Xe.device
        stream = torch.cuda.current_stream().cuda_stream
        dHe = He.new_zeros(B, ne, C, h)          # halo rows carry no output gradient: they stay out of every dW
        dHe[:, :n] = dHn
        scratch = torch.empty(lib.stc_cell_bwd_scratch_bytes(dims) // 4, dtype=torch.float32, device=dev)
        lay = _lib.scratch_layout(dims)
        R = B * ne * C
        dXe = torch.empty_like(Xe)
        dHp = torch.empty_like(He)
        P, L = Ks * Kc, Din + h
        # one flat buffer for every replicated-parameter gradient: the all-reduce below is a single call
        sizes = [P * L * 2 * h, 2 * h if ctx.has_bias else 0, P * L * h, h if ctx.has_bias else 0, C * C]
        flat = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)
        dWg, dbg, dWc, dbc, dGc = flat.split(sizes)
        pz = lambda t: t.data_ptr() if t.numel() else None

        def region(name, k, width):
            o = lay[name] + k * R * width
            return scratch[o:o + R * width].view(B, ne, C, width)

        def stage(which):
            _lib.check(lib.stc_cell_bwd_stage(dims, which, Gc.data_ptr(), Xe.data_ptr(), ne * C * Din, He.data_ptr(),
                                              Wg.data_ptr(), Wc.data_ptr(), dHe.data_ptr(), dXe.data_ptr(), dHp.data_ptr(),
                                              dWg.data_ptr(), pz(dbg), dWc.data_ptr(), pz(dbc), dGc.data_ptr(), 1,
                                              saved.data_ptr(), saved.numel() * 4, scratch.data_ptr(),
                                              scratch.numel() * 4, stream), "stc_cell_bwd_stage")
            _lib.note_launches()

        stage(_lib.STAGE_CANDI)
        dYr = [region("dYr", k, h) for k in range(Ks)]
        adjoint_chain_ext(ps, dYr)
        dYr[0][:, n:].zero_()                    # d(r*H) of the halo rows belongs to their owners
        stage(_lib.STAGE_GATES)
        adjoint_chain_ext(ps, [dXe] + [region("dYx", k, Din) for k in range(Ks - 1)])
        adjoint_chain_ext(ps, [dHp] + [region("dYh", k, h) for k in range(Ks - 1)])
        if reduce_params and ps.world > 1:
            dist.all_reduce(flat, group=ps.group)
        need = ctx.needs_input_grad
        return (dGc.view(C, C) if need[0] else None, dXe[:, :n].contiguous() if need[1] else None,
                dHp[:, :n].contiguous() if need[2] else None, dWg.view(P * L, 2 * h) if need[3] else None,
                dbg if (need[4] and ctx.has_bias) else None, dWc.view(P * L, h) if need[5] else None,
                dbc if (need[6] and ctx.has_bias) else None, None)


def partitioned_cell_forward(ps: PartitionedSupport, Gc: torch.Tensor, Xt: torch.Tensor, Ht_1: torch.Tensor,
                             Wg: torch.Tensor, bg: Optional[torch.Tensor], Wc: torch.Tensor,
                             bc: Optional[torch.Tensor], Ks: int, Kc: int, activation=None,
                             reduce_params: bool = True) -> torch.Tensor:
    """`STC_Cell.forward` (`framework/STC_GNN.py:65-79`) on this rank's node block of a row-partitioned graph.

    Xt [B, nloc, C, Din], Ht_1 [B, nloc, C, h] are the local blocks; returns the local block of H'.  The spatial
    hops run here (halo exchange + `stc_support_apply`), the node-local stages are `stc_cell_fwd_stage` /
    `stc_cell_bwd_stage` (include/stc_b200.h).  Differentiable w.r.t. Gc, Xt, Ht_1 and the parameters (the support is a
    constant CSR graph: no dGs).  The parameters and Gc are replicated over the ranks, so their gradients are partial
    sums over this rank's nodes: with `reduce_params` (default) the backward all-reduces them in ONE flat bucket and
    every rank receives the full gradient (all ranks must then run the backward); pass False to reduce later, e.g.
    once per step through `dp.GradBucket` after a time loop.
    """
    for name, t in (("Gc", Gc), ("Xt", Xt), ("Ht_1", Ht_1), ("gates.W", Wg), ("candi.W", Wc)):
        if not t.is_cuda or t.dtype != torch.float32:
            raise RuntimeError(f"partitioned cell: {name} must be a float32 CUDA tensor (there is no CPU path)")
    B, n, C, Din = Xt.shape
    h = Ht_1.shape[-1]
    if n != ps.nloc or Ht_1.shape != (B, n, C, h):
        raise RuntimeError(f"local block has {n} nodes, the partition owns {ps.nloc}; Ht_1 {tuple(Ht_1.shape)}")
    if (bg is None) != (bc is None):
        raise RuntimeError("partitioned cell: gates.b and candi.b must both be given or both be None")
    if B == 0:   # empty batch (the same on every rank: the batch is replicated, the nodes are split): nothing to exchange
        return Ht_1.new_zeros(0, n, C, h)
    return _PartitionedCell.apply(Gc, Xt, Ht_1, Wg, bg, Wc, bc, (ps, Ks, Kc, activation, reduce_params))
