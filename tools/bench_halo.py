#!/usr/bin/env python
"""BASELINE configs[3]: one STC cell forward on the N = 65,536 kNN graph (C = 8, F = 64, Ks = 4 = 3 hops), nodes
row-partitioned over the ranks with a halo exchange per hop (stc_gnn_b200/halo.py).  Launch under torchrun:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/bench_halo.py [B] [--train]

`--train`: forward + backward of the cell (gradients w.r.t. Xt, H, Gc and the parameters; the partitioned backward runs
the adjoint hops with their own halo exchanges and all-reduces the parameter gradients in one bucket).

Strong scaling: the global problem is fixed, each rank owns N / world nodes.  Time = max over ranks (CUDA events,
barrier on both sides).  Rank 0 prints one JSON line."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import stc_gnn_b200 as S  # noqa: E402
from stc_gnn_b200 import _lib  # noqa: E402
from stc_gnn_b200.halo import PartitionedSupport, partitioned_cell_forward  # noqa: E402
from stc_gnn_b200.synth import knn_csr  # noqa: E402


def main():
    train = "--train" in sys.argv
    pos = [a for a in sys.argv[1:] if not a.startswith("--")]
    B = int(pos[0]) if pos else 2
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()
    N, C, F, Ks, Kc = 65536, 8, 64, 4, 2
    rp, ci, va = knn_csr(N, 8)
    torch.manual_seed(0)
    cell = S.STC_Cell(N, C, Ks, Kc, F, F).to(dev)
    Gc = (torch.rand(C, C, generator=torch.Generator().manual_seed(1)) / C).to(dev)
    with torch.set_grad_enabled(train):
        if world > 1:
            ps = PartitionedSupport(rp, ci, va, N, rank, world)
            n = ps.nloc
            halo = {"fwd_halo_rows": ps.fwd.nhalo, "local_rows": n}
        else:
            ps = S.CsrSupport(rp.to(dev), ci.to(dev), va.to(dev), N)
            n = N
            halo = {"fwd_halo_rows": 0, "local_rows": n}
        g = torch.Generator().manual_seed(100 + rank)
        X = torch.randn(B, n, C, F, generator=g).to(dev).requires_grad_(train)
        H = (torch.randn(B, n, C, F, generator=g) * 0.5).to(dev).requires_grad_(train)
        dHn = torch.randn(B, n, C, F, generator=g).to(dev)
        Gc.requires_grad_(train)

        def fwd():
            if world > 1:
                return partitioned_cell_forward(ps, Gc, X, H, cell.gates.W, cell.gates.b, cell.candi.W, cell.candi.b, Ks, Kc)
            return cell(Gs=ps, Gc=Gc, Xt=X, Ht_1=H)

        def step():
            out = fwd()
            if train:
                for t in (X, H, Gc, *cell.parameters()):
                    t.grad = None
                out.backward(dHn)

        for _ in range(3):
            step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 10
        e0.record()
        for _ in range(iters):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
    if rank == 0:
        print(json.dumps({"kind": "config4_cell_step_partitioned" if train else "config4_cell_fwd_partitioned", "n_gpus": world, "N": N, "C": C, "F": F, "Ks": Ks, "B": B,
                          "ms": ms, "cell_step_samples_per_s": B / ms * 1e3, "scaling": "strong", **halo}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
