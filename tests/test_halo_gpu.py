"""GPU tests of the row-partitioned path: staged cell forward AND backward (C ABI) + halo hops through stc_support_apply.

World size 1 runs on any GPU box; the 2-rank NCCL test needs two visible GPUs (`gpurun --gpus 2`) and is skipped otherwise.
"""
import os
import socket
import traceback

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import stc_oracle as O
from tests.helpers import random_case

pytestmark = pytest.mark.gpu


def _cell_block(rank, world, device, cfg, seed):
    """Partitioned cell on this rank's node block == the block of the unpartitioned fp64 oracle: H' and every gradient
    (activations: this rank's rows; replicated parameters and Gc: the all-reduced full gradient)."""
    import stc_gnn_b200 as S
    t = random_case(cfg["B"], cfg["N"], cfg["C"], cfg["Din"], cfg["h"], cfg["Ks"], cfg["Kc"], seed=seed, sparse_frac=0.85)
    leaves = {k: t[k].clone().requires_grad_(True) for k in ("Gc", "Xt", "H", "Wg", "bg", "Wc", "bc")}
    want = O.stc_cell(t["Gs"], leaves["Gc"], leaves["Xt"], leaves["H"], leaves["Wg"], leaves["bg"], leaves["Wc"],
                      leaves["bc"], cfg["Ks"], cfg["Kc"])
    want.backward(t["dHn"])
    ps = S.halo.PartitionedSupport.from_dense(t["Gs"], rank, world)
    f = lambda x: x.float().to(device)
    with torch.no_grad():
        got = S.halo.partitioned_cell_forward(ps, f(t["Gc"]), f(ps.local_slice(t["Xt"])), f(ps.local_slice(t["H"])),
                                              f(t["Wg"]), f(t["bg"]), f(t["Wc"]), f(t["bc"]), cfg["Ks"], cfg["Kc"])
        # one adjoint hop as well (backward direction of the exchange)
        dY = torch.randn(cfg["B"], cfg["N"], cfg["C"], cfg["h"], generator=torch.Generator().manual_seed(seed + 1)).double()
        got_b = ps.apply(f(ps.local_slice(dY)), "bwd")
    torch.cuda.synchronize()
    O.assert_close(got.cpu(), ps.local_slice(want.detach()), f"partitioned cell forward (rank {rank}/{world})")
    O.assert_close(got_b.cpu(), ps.local_slice(torch.einsum("nm,bmcl->bncl", t["Gs"], dY)), f"adjoint hop (rank {rank})")
    # gradients through the partitioned path
    g = {k: f(ps.local_slice(t[k]) if k in ("Xt", "H") else t[k]).requires_grad_(True) for k in leaves}
    out = S.halo.partitioned_cell_forward(ps, g["Gc"], g["Xt"], g["H"], g["Wg"], g["bg"], g["Wc"], g["bc"], cfg["Ks"], cfg["Kc"])
    O.assert_close(out.detach().cpu(), ps.local_slice(want.detach()), f"partitioned cell forward, grad mode (rank {rank})")
    out.backward(f(ps.local_slice(t["dHn"])))
    torch.cuda.synchronize()
    for k in leaves:
        ref = leaves[k].grad
        ref = ps.local_slice(ref) if k in ("Xt", "H") else ref
        O.assert_close(g[k].grad.cpu(), ref, f"partitioned d{k} (rank {rank}/{world})")
    return ps


@pytest.mark.parametrize("cfg", [dict(B=2, N=96, C=5, Din=16, h=16, Ks=2, Kc=2), dict(B=3, N=70, C=4, Din=3, h=8, Ks=4, Kc=2),
                                 dict(B=1, N=40, C=8, Din=64, h=64, Ks=3, Kc=2)])
def test_staged_cell_single_rank(cfg):
    ps = _cell_block(0, 1, torch.device("cuda:0"), cfg, seed=21)
    assert ps.fwd.nhalo == 0


def test_recurrent_stack_accepts_a_partitioned_support():
    """STC_Cell / RecurrentStack dispatch on a PartitionedSupport handed in as `Gs`: encoder 2 x T=3 + decoder roll-out,
    output and every gradient == the same stack on the unpartitioned CSR support."""
    import stc_gnn_b200 as S
    N, C, Din, h, Ks, Kc, T, B = 48, 4, 1, 16, 3, 2, 3, 2
    t = random_case(B, N, C, Din, h, Ks, Kc, seed=5, sparse_frac=0.85)
    dev = torch.device("cuda:0")
    csr = S.CsrSupport.from_torch_sparse(t["Gs"].float().to_sparse_csr().to(dev))
    ps = S.halo.PartitionedSupport.from_dense(t["Gs"], 0, 1)
    ps.reduce_per_cell = False
    torch.manual_seed(3)
    stack = S.RecurrentStack(N, C, Ks, Kc, Din, h, 2, 2).to(dev)
    g = torch.Generator().manual_seed(8)
    X = torch.randn(B, T, N, C, Din, generator=g).to(dev)
    dOut = torch.randn(B, 2, N, C, h, generator=g).to(dev)
    res = []
    for Gs in (csr, ps):
        Gc = t["Gc"].float().to(dev).requires_grad_(True)
        Xl = X.clone().requires_grad_(True)
        stack.zero_grad(set_to_none=True)
        out = stack(Gs, Gc, Xl)
        out.backward(dOut)
        res.append([out.detach(), Gc.grad, Xl.grad] + [p.grad.clone() for p in stack.parameters()])
    for i, (a, b) in enumerate(zip(*res)):
        O.assert_close(b.cpu(), a.cpu(), f"stack tensor {i}: partitioned vs CSR support", rtol=1e-4, atol_scale=1e-5)


def test_staged_backward_error_paths():
    import stc_gnn_b200 as S
    t = random_case(2, 16, 2, 2, 8, 2, 2, seed=1)
    ps = S.halo.PartitionedSupport.from_dense(t["Gs"], 0, 1)
    f = lambda x: x.float().cuda()
    with pytest.raises(RuntimeError, match="both be given"):
        S.halo.partitioned_cell_forward(ps, f(t["Gc"]), f(t["Xt"]), f(t["H"]), f(t["Wg"]), f(t["bg"]), f(t["Wc"]), None, 2, 2)
    with pytest.raises(RuntimeError, match="partition owns"):
        S.halo.partitioned_cell_forward(ps, f(t["Gc"]), f(t["Xt"])[:, :8], f(t["H"])[:, :8], f(t["Wg"]), f(t["bg"]),
                                        f(t["Wc"]), f(t["bc"]), 2, 2)
    with pytest.raises(RuntimeError, match="no CPU path"):
        S.halo.partitioned_cell_forward(ps, t["Gc"].float(), f(t["Xt"]), f(t["H"]), f(t["Wg"]), f(t["bg"]), f(t["Wc"]),
                                        f(t["bc"]), 2, 2)


def _nccl_worker(rank, world, port, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        for cfg in (dict(B=2, N=101, C=5, Din=16, h=16, Ks=3, Kc=2), dict(B=1, N=64, C=8, Din=1, h=8, Ks=4, Kc=2)):
            ps = _cell_block(rank, world, dev, cfg, seed=33)
            assert ps.fwd.nhalo > 0
        # DP bucket over NCCL: sum of rank-dependent gradients
        import stc_gnn_b200 as S
        p = torch.nn.Parameter(torch.ones(1000, device=dev))
        p.grad = torch.full_like(p, float(rank + 1))
        S.dp.GradBucket([p]).allreduce()
        assert torch.allclose(p.grad, torch.full_like(p, float(sum(range(1, world + 1)))))
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, None))
    except Exception:  # pragma: no cover
        q.put((rank, traceback.format_exc()))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_halo_cell_two_ranks_nccl():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    errs = [f"rank {r}:\n{e}" for r, e in results if e]
    assert not errs, "\n".join(errs)


def test_row_subset_apply_and_halo_pack_unpack_kernels():
    """The two device pieces of the overlapped hop, on one GPU: `stc_support_apply_rows` computes exactly the listed
    output nodes (the rest of `out` keeps its contents), `stc_halo_pack` / `stc_halo_unpack` move packed row slabs."""
    import stc_gnn_b200 as S
    from stc_gnn_b200 import _lib
    from stc_gnn_b200.support import support_apply
    lib = _lib.load()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(17)
    N, B, W = 90, 3, 40
    G = (torch.rand(N, N, generator=g) * (torch.rand(N, N, generator=g) > 0.8)).float().to(dev)
    csr = S.CsrSupport.from_dense(G)
    X = torch.randn(B, N, W, generator=g).to(dev)
    Z = torch.randn(B, N, W, generator=g).to(dev)
    rows = torch.tensor([0, 3, 4, 17, 50, 89], dtype=torch.int32, device=dev)
    for transpose in (True, False):
        full = support_apply(csr, X, transpose=transpose, alpha=2.0, beta=-1.0, Z=Z)
        out = torch.full((B, N, W), 7.0, device=dev)
        support_apply(csr, X, transpose=transpose, alpha=2.0, beta=-1.0, Z=Z, out=out, rows=rows)
        keep = torch.ones(N, dtype=torch.bool, device=dev)
        keep[rows.long()] = False
        assert torch.equal(out[:, rows.long()], full[:, rows.long()])
        assert torch.equal(out[:, keep], torch.full_like(out[:, keep], 7.0))
        # in-place accumulation form used by the adjoint chain (out aliases Z), split over two row lists
        acc = Z.clone()
        rest = torch.nonzero(keep).flatten().to(torch.int32)
        support_apply(csr, X, transpose=transpose, alpha=2.0, beta=1.0, Z=acc, out=acc, rows=rows)
        support_apply(csr, X, transpose=transpose, alpha=2.0, beta=1.0, Z=acc, out=acc, rows=rest)
        want = support_apply(csr, X, transpose=transpose, alpha=2.0, beta=1.0, Z=Z)
        assert torch.equal(acc, want)
    with pytest.raises(RuntimeError):
        support_apply(G, X, out=torch.empty_like(X), rows=rows)          # dense supports have no row-list form
    for Wq in (40, 7):                                                   # vectorised and scalar widths
        Xe = torch.randn(B, N, Wq, generator=g).to(dev)
        idx = torch.tensor([5, 6, 30, 2], dtype=torch.int32, device=dev)
        send = torch.empty(idx.numel(), B * Wq, device=dev)
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(lib.stc_halo_pack(Xe.data_ptr(), N * Wq, Wq, B, idx.data_ptr(), idx.numel(), send.data_ptr(), st), "pack")
        assert torch.equal(send.view(-1, B, Wq), Xe[:, idx.long()].permute(1, 0, 2))
        before = Xe.clone()
        _lib.check(lib.stc_halo_unpack(send.data_ptr(), Wq, B, N - 4, 4, Xe.data_ptr(), N * Wq, st), "unpack")
        assert torch.equal(Xe[:, N - 4:], before[:, idx.long()]) and torch.equal(Xe[:, :N - 4], before[:, :N - 4])
