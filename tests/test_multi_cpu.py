"""World-size-2 `gloo` tests of the multi-GPU host logic (no GPU needed):

* dp.py    -- flat-bucket gradient all-reduce: DP gradients == single-process gradients of the concatenated batch
* halo.py  -- partition plans + halo exchange: partitioned spatial terms / cell == the unpartitioned oracle

The arithmetic in these tests is the oracle's (tests may use it); the product arithmetic is CUDA-only.
"""
import os
import socket
import traceback

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import stc_oracle as O
from stc_gnn_b200 import dp, halo
from tests.helpers import random_case

WORLD = 2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, fn, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        torch.set_num_threads(1)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        fn(rank, world)
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, None))
    except Exception:  # pragma: no cover - reported by the parent
        q.put((rank, traceback.format_exc()))


def run_ranks(fn, world=WORLD):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, fn, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    errs = [f"rank {r}:\n{e}" for r, e in results if e]
    assert not errs, "\n".join(errs)


# ------------------------------------------------------------------------------------------------
# helpers (module level: spawn pickles the worker functions)
# ------------------------------------------------------------------------------------------------
def _sparse_graph(N, seed, keep=0.08):
    g = torch.Generator().manual_seed(seed)
    G = torch.rand(N, N, generator=g) * (torch.rand(N, N, generator=g) < keep)
    idx = torch.arange(N)
    G[idx, (idx + 1) % N] = 0.5          # every node has an out- and an in-neighbour, some across the block boundary
    G[(idx + 7) % N, idx] += 0.25
    return (G / G.sum(0, keepdim=True).clamp(min=1e-3)).double()


def _oracle_apply(plan, X_ext):
    A = torch.sparse_coo_tensor(torch.stack([plan.op_row, plan.op_col]), plan.op_val.to(X_ext.dtype),
                                size=(plan.nloc, plan.next)).coalesce()
    B, n, W = X_ext.shape
    return torch.sparse.mm(A, X_ext.permute(1, 0, 2).reshape(n, B * W)).view(plan.nloc, B, W).permute(1, 0, 2)


def _dp_case(rank, world):
    cfg = dict(B=6, N=12, C=3, Din=2, h=4, Ks=3, Kc=2)
    t = random_case(cfg["B"], cfg["N"], cfg["C"], cfg["Din"], cfg["h"], cfg["Ks"], cfg["Kc"], seed=3)
    names = ["Gs", "Gc", "Wg", "bg", "Wc", "bc"]

    def grads(Xt, H, dHn):
        leaves = {k: t[k].clone().requires_grad_(True) for k in names}
        Hn = O.stc_cell(leaves["Gs"], leaves["Gc"], Xt, H, leaves["Wg"], leaves["bg"], leaves["Wc"], leaves["bc"],
                        cfg["Ks"], cfg["Kc"])
        Hn.backward(dHn)
        return leaves

    full = grads(t["Xt"], t["H"], t["dHn"])
    s, e = dp.shard_bounds(cfg["B"], rank, world)
    assert (s, e) == (rank * 3, rank * 3 + 3)
    mine = grads(dp.shard_batch(t["Xt"], rank, world), dp.shard_batch(t["H"], rank, world),
                 dp.shard_batch(t["dHn"], rank, world))
    leaves32 = []
    for k in names:                                 # the bucket is fp32 (what the GPU path reduces)
        p = mine[k].detach().float().requires_grad_(True)
        p.grad = mine[k].grad.float()
        leaves32.append(p)
    leaves32[3].grad = None                         # an absent gradient must count as zeros on this rank only
    if rank == 0:
        leaves32[3].grad = mine["bg"].grad.float() + (grads(dp.shard_batch(t["Xt"], 1, world),
                                                            dp.shard_batch(t["H"], 1, world),
                                                            dp.shard_batch(t["dHn"], 1, world))["bg"].grad.float())
    bucket = dp.GradBucket(leaves32)
    assert bucket.numel == sum(p.numel() for p in leaves32)
    bucket.allreduce()
    for k, p in zip(names, leaves32):
        O.assert_close(p.grad, full[k].grad, f"DP-reduced d{k} (rank {rank})", rtol=1e-4, atol_scale=1e-5)


def _global_supports_case(rank, world):
    """A batch-summed score followed by a non-linearity (the shape of MGP_Gen, STC_GNN.py:231-232): with sum_over_ranks
    on the pre-relu sum, sharded-batch values and gradients equal the single-process global-batch ones."""
    g = torch.Generator().manual_seed(4)
    X = torch.randn(6, 3, 7, 4, generator=g, dtype=torch.float64)              # [B, T, N, hdim]
    Wu = torch.randn(4, 5, generator=g, dtype=torch.float64)
    dG = torch.randn(7, 7, generator=g, dtype=torch.float64)                     # per-sample loss weight (shared)

    def supports(Xs, W, reduce):
        U = torch.tanh(Xs @ W)
        P = torch.einsum("btnh,btmh->nm", U, U.flip(-1))
        if reduce:
            P = dp.sum_over_ranks(P)
        return torch.softmax(torch.relu(P), dim=-1)

    def loss(G, Xs):                                                             # per-rank loss: depends on G and own shard
        return (G * dG).sum() * Xs.square().mean()

    W_full = Wu.clone().requires_grad_(True)
    G_full = supports(X, W_full, False)
    total = sum(loss(G_full, dp.shard_batch(X, r, world)) for r in range(world))
    total.backward()
    W_r = Wu.clone().requires_grad_(True)
    Xs = dp.shard_batch(X, rank, world)
    G_r = supports(Xs, W_r, True)
    O.assert_close(G_r, G_full, f"global-batch support (rank {rank})", 1e-12, 1e-12)
    loss(G_r, Xs).backward()
    gw = W_r.grad.clone()
    dist.all_reduce(gw)                                                          # the usual DP gradient sum
    O.assert_close(gw, W_full.grad, f"generator gradient (rank {rank})", 1e-10, 1e-12)
    assert dp.sum_over_ranks(X, group=None) is not None


def _halo_case(rank, world):
    N, B, C, L, Ks = 37, 3, 2, 5, 4                 # odd N: uneven blocks
    G = _sparse_graph(N, seed=11)
    ps = halo.PartitionedSupport.from_dense(G, rank, world)
    assert ps.nloc == (19 if rank == 0 else 18) and ps.fwd.nhalo > 0 and ps.bwd.nhalo > 0
    g = torch.Generator().manual_seed(5)
    X = torch.randn(B, N, C, L, generator=g, dtype=torch.float64)
    # forward terms: Ks-1 hops with a halo exchange each
    want = O.spatial_terms(X, G, Ks)
    got = ps.spatial_terms(ps.local_slice(X), Ks, apply_fn=_oracle_apply)
    for k in range(Ks):
        O.assert_close(got[k], ps.local_slice(want[k]), f"partitioned spatial term {k} (rank {rank})", 1e-9, 1e-10)
    # adjoint hop on the un-transposed graph
    dY = torch.randn(B, N, C, L, generator=g, dtype=torch.float64)
    want_b = torch.einsum("nm,bmcl->bncl", G, dY)
    got_b = ps.apply(ps.local_slice(dY), "bwd", apply_fn=_oracle_apply)
    O.assert_close(got_b, ps.local_slice(want_b), f"partitioned adjoint hop (rank {rank})", 1e-9, 1e-10)
    # exchange bookkeeping: what I receive is exactly the owner's rows
    ext = halo.exchange(ps.fwd, ps.local_slice(X).reshape(B, ps.nloc, -1))
    torch.testing.assert_close(ext[:, ps.nloc:], X.reshape(B, N, -1)[:, ps.fwd.halo_global])


def _oracle_hop(ps):
    """CPU stand-in for PartitionedSupport.hop_ext: the real in-place halo refresh, the checker's arithmetic."""
    def hop(X_ext, out_ext, Z_ext, alpha, beta, direction):
        plan = ps.fwd if direction == "fwd" else ps.bwd
        B = X_ext.shape[0]
        X3 = X_ext.view(B, plan.next, -1)
        halo.exchange_into(plan, X3, ps.group)
        Y = alpha * _oracle_apply(plan, X3)
        if beta != 0.0:
            Y = Y + beta * Z_ext.view(B, plan.next, -1)[:, :plan.nloc]
        out_ext.view(B, plan.next, -1)[:, :plan.nloc] = Y
    return hop


def _halo_adjoint_case(rank, world):
    """Adjoint of the Ks-term spatial recurrence on a row partition == autograd of the unpartitioned recurrence;
    both directions share one extended node set even though the graph is not symmetric."""
    N, B, C, L, Ks = 37, 2, 3, 4, 4
    G = _sparse_graph(N, seed=21)
    assert not torch.equal(G != 0, (G != 0).t())
    ps = halo.PartitionedSupport.from_dense(G, rank, world)
    assert ps.fwd.nhalo == ps.bwd.nhalo and torch.equal(ps.fwd.halo_global, ps.bwd.halo_global)
    assert all(torch.equal(a, b) for a, b in zip(ps.fwd.send_idx, ps.bwd.send_idx))
    g = torch.Generator().manual_seed(6)
    X = torch.randn(B, N, C, L, generator=g, dtype=torch.float64).requires_grad_(True)
    dY = [torch.randn(B, N, C, L, generator=g, dtype=torch.float64) for _ in range(Ks)]
    terms = O.spatial_terms(X, G, Ks)
    sum((t * d).sum() for t, d in zip(terms, dY)).backward()
    ne = ps.fwd.next
    ybar = []
    for d in dY:
        e = torch.full((B, ne, C, L), 7.0, dtype=torch.float64)      # halo rows: junk the chain must never read
        e[:, :ps.nloc] = ps.local_slice(d)
        ybar.append(e)
    got = halo.adjoint_chain_ext(ps, ybar, hop=_oracle_hop(ps))
    O.assert_close(got[:, :ps.nloc], ps.local_slice(X.grad), f"partitioned adjoint chain (rank {rank})", 1e-9, 1e-10)
    # the forward recurrence through the same in-place hop (what the CUDA path does with hop_ext)
    hop = _oracle_hop(ps)
    ext = [torch.zeros(B, ne, C, L, dtype=torch.float64) for _ in range(Ks)]
    ext[0][:, :ps.nloc] = ps.local_slice(X.detach())
    for k in range(1, Ks):
        if k == 1:
            hop(ext[0], ext[1], None, 1.0, 0.0, "fwd")
        else:
            hop(ext[k - 1], ext[k], ext[k - 2], 2.0, -1.0, "fwd")
    for k in range(Ks):
        O.assert_close(ext[k][:, :ps.nloc], ps.local_slice(terms[k].detach()), f"in-place forward term {k}", 1e-9, 1e-10)


def _halo_cell_case(rank, world):
    """Whole cell on a row partition: spatial terms via halo hops, everything else node-local."""
    cfg = dict(B=2, N=26, C=3, Din=2, h=4, Ks=3, Kc=2)
    t = random_case(cfg["B"], cfg["N"], cfg["C"], cfg["Din"], cfg["h"], cfg["Ks"], cfg["Kc"], seed=9, sparse_frac=0.8)
    ps = halo.PartitionedSupport.from_dense(t["Gs"], rank, world)
    want = O.stc_cell(t["Gs"], t["Gc"], t["Xt"], t["H"], t["Wg"], t["bg"], t["Wc"], t["bc"], cfg["Ks"], cfg["Kc"])
    got = O.stc_cell(None, t["Gc"], ps.local_slice(t["Xt"]), ps.local_slice(t["H"]), t["Wg"], t["bg"], t["Wc"], t["bc"],
                     cfg["Ks"], cfg["Kc"], spatial_terms_fn=lambda X: ps.spatial_terms(X, cfg["Ks"], apply_fn=_oracle_apply))
    O.assert_close(got, ps.local_slice(want), f"partitioned cell (rank {rank})", 1e-9, 1e-10)


# ------------------------------------------------------------------------------------------------
def test_shard_bounds_cover_the_batch():
    for total in (0, 1, 7, 32, 4096):
        for world in (1, 2, 3, 8):
            spans = [dp.shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(e - s for s, e in spans) - min(e - s for s, e in spans) <= 1
    with pytest.raises(ValueError):
        dp.shard_bounds(4, 2, 2)


def test_plan_is_consistent_across_ranks():
    """What rank p sends to q is exactly what q expects from p (no communication needed to agree)."""
    G = _sparse_graph(41, seed=2)
    for world in (2, 3, 4):
        plans = [halo.PartitionedSupport.from_dense(G, r, world) for r in range(world)]
        for direction in ("fwd", "bwd"):
            for p in range(world):
                for q in range(world):
                    a = getattr(plans[p], direction)
                    b = getattr(plans[q], direction)
                    assert a.send_counts[q] == b.recv_counts[p]
                    sent_global = a.send_idx[q] + a.start
                    off = sum(b.recv_counts[:p])
                    assert torch.equal(sent_global, b.halo_global[off:off + b.recv_counts[p]])
        assert sum(pl.nloc for pl in plans) == 41


def test_single_rank_plan_has_no_halo():
    G = _sparse_graph(16, seed=4)
    ps = halo.PartitionedSupport.from_dense(G, 0, 1)
    assert ps.nloc == 16 and ps.fwd.nhalo == 0 and ps.bwd.nhalo == 0
    X = torch.randn(2, 16, 3, dtype=torch.float64)
    got = ps.apply(X, "fwd", apply_fn=_oracle_apply)
    O.assert_close(got, torch.einsum("nm,bnw->bmw", G, X), "1-rank partition", 1e-9, 1e-10)
    with pytest.raises(RuntimeError):
        ps.apply(X.float(), "fwd")          # CPU tensors without an injected apply_fn: no CPU path


def test_dp_bucket_allreduce_matches_concatenated_batch():
    run_ranks(_dp_case)


def test_halo_partitioned_terms_match_unpartitioned():
    run_ranks(_halo_case)


def test_halo_partitioned_cell_matches_unpartitioned():
    run_ranks(_halo_cell_case)


def test_halo_partitioned_adjoint_chain_matches_autograd():
    run_ranks(_halo_adjoint_case)


def test_sum_over_ranks_makes_batch_coupled_supports_global():
    run_ranks(_global_supports_case)
    x = torch.ones(3, requires_grad=True)
    assert dp.sum_over_ranks(x) is x            # identity outside a process group


# ------------------------------------------------------------------------------------------------
# install(dp_group=...): the reference's MGP_Gen under batch data-parallelism (stc_gnn_b200/mgp.py)
# ------------------------------------------------------------------------------------------------
def _mgp_setup():
    from tests.helpers import import_reference
    ref, _ = import_reference()
    torch.manual_seed(5)
    gen = ref.MGP_Gen(num_nodes=12, num_categories=4, hidden_dim=6).double()
    g = torch.Generator().manual_seed(8)
    X = (torch.rand(6, 5, 12, 4, generator=g) < 0.3).double()                    # [B, T, N, C] incident flags
    As = (torch.rand(12, 12, generator=g) < 0.3).double()
    Ac = torch.rand(4, 4, generator=g, dtype=torch.float64) * 0.3
    wS = torch.randn(6, 12, 12, generator=g, dtype=torch.float64)                # per-sample loss weights
    wC = torch.randn(6, 4, 4, generator=g, dtype=torch.float64)
    return ref, gen, X, As, Ac, wS, wC


def _mgp_dp_case(rank, world):
    """Sharded batch + patched generator: supports and EVERY generator gradient equal the single-process global-batch
    ones; the fusion-layer gradients come out complete on every rank without being communicated.  Both conventions:
    per-rank losses add up (bucket summed) and per-rank losses are shard means whose mean is the job's loss (bucket
    averaged)."""
    from stc_gnn_b200 import mgp
    ref, gen, X, As, Ac, wS, wC = _mgp_setup()
    loss = lambda Gs, Gc, a, b: (Gs * a.sum(0)).sum() + (Gc * b.sum(0)).sum()     # sum of per-sample losses
    for average in (False, True):
        for p in gen.parameters():
            p.grad = None
        Gs_f, Gc_f = gen(X, As, Ac)                                              # stock forward, whole batch
        (loss(Gs_f, Gc_f, wS, wC) / (world if average else 1)).backward()
        full = {n: p.grad.clone() for n, p in gen.named_parameters()}
        for p in gen.parameters():
            p.grad = None
        mgp.patch_generator(ref, group=None, average=average)
        try:
            sl = slice(*dp.shard_bounds(X.shape[0], rank, world))
            Gs_r, Gc_r = gen(X[sl], As, Ac)
            O.assert_close(Gs_r, Gs_f, f"Gs from a shard (rank {rank})", 1e-10, 1e-12)
            O.assert_close(Gc_r, Gc_f, f"Gc from a shard (rank {rank})", 1e-10, 1e-12)
            loss(Gs_r, Gc_r, wS[sl], wC[sl]).backward()
        finally:
            mgp.unpatch_generator(ref)
        shard_params = mgp.dp_bucket_parameters(gen)
        assert len(shard_params) == 4 and all(".aggreg_" not in n for n, p in gen.named_parameters()
                                              if any(p is q for q in shard_params))
        for p in shard_params:                                                    # the usual DP gradient sum / mean
            dist.all_reduce(p.grad)
            if average:
                p.grad.div_(world)
        for n, p in gen.named_parameters():
            O.assert_close(p.grad, full[n], f"generator d{n} (rank {rank}, average={average})", 1e-8, 1e-10)


def test_generator_restatement_matches_reference_single_process():
    from tests.helpers import import_reference
    from stc_gnn_b200 import mgp
    if import_reference()[0] is None:
        pytest.skip("needs the live reference")
    ref, gen, X, As, Ac, _, _ = _mgp_setup()
    with torch.no_grad():
        want = gen(X, As, Ac)
        got = mgp.mgp_forward_dp(gen, X, As, Ac)
    O.assert_close(got[0], want[0], "Gs", 1e-12, 1e-13)
    O.assert_close(got[1], want[1], "Gc", 1e-12, 1e-13)


def test_generator_under_batch_dp_equals_global_batch():
    from tests.helpers import import_reference
    if import_reference()[0] is None:
        pytest.skip("needs the live reference")
    run_ranks(_mgp_dp_case)
