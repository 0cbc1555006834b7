#!/usr/bin/env python
"""BASELINE configs[3]: one STC cell forward on the N = 65,536 kNN graph (C = 8, F = 64, Ks = 4 = 3 hops), nodes
row-partitioned over the ranks with a halo exchange per hop (stc_gnn_b200/halo.py).  Launch under torchrun:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/bench_halo.py [B] [--train]

`--train`: forward + backward of the cell (gradients w.r.t. Xt, H, Gc and the parameters; the partitioned backward runs
the adjoint hops with their own halo exchanges; the parameter gradients are all-reduced ONCE per step in a flat bucket,
`--reduce-per-cell` restores one all-reduce per cell backward).  At world > 1 every rank first checks its block of the
partitioned result (outputs and every gradient) against the unpartitioned cell run on the same GPU -> `parity_ok`.

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


def violations(got, ref, atol_scale):
    got, ref = got.double(), ref.double()
    scale = ref.abs().mean()
    err = (got - ref).abs()
    return int((err > 1e-4 * ref.abs() + atol_scale * scale).sum()), float(err.max() / scale)


def main():
    train = "--train" in sys.argv
    per_cell = "--reduce-per-cell" in sys.argv
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
    params = list(cell.parameters())
    Gc = (torch.rand(C, C, generator=torch.Generator().manual_seed(1)) / C).to(dev)
    # the GLOBAL problem is generated identically on every rank (one seed); a rank keeps its node block
    g = torch.Generator().manual_seed(100)
    Xg = torch.randn(B, N, C, F, generator=g)
    Hg = torch.randn(B, N, C, F, generator=g) * 0.5
    dHg = torch.randn(B, N, C, F, generator=g)
    full = S.CsrSupport(rp.to(dev), ci.to(dev), va.to(dev), N)
    with torch.set_grad_enabled(train):
        if world > 1:
            ps = PartitionedSupport(rp, ci, va, N, rank, world)
            ps.reduce_per_cell = per_cell
            lo, n = ps.start, ps.nloc
            halo = {"fwd_halo_rows": ps.fwd.nhalo, "local_rows": n}
        else:
            ps, lo, n = full, 0, N
            halo = {"fwd_halo_rows": 0, "local_rows": n}
        blk = lambda t: t[:, lo:lo + n].contiguous().to(dev)
        X, H, dHn = blk(Xg).requires_grad_(train), blk(Hg).requires_grad_(train), blk(dHg)
        Gc.requires_grad_(train)
        leaves = params + [Gc]
        bucket = S.dp.GradBucket(leaves) if (world > 1 and train and not per_cell) else None

        def fwd():
            if world > 1:
                return partitioned_cell_forward(ps, Gc, X, H, cell.gates.W, cell.gates.b, cell.candi.W, cell.candi.b, Ks, Kc,
                                                reduce_params=per_cell)
            return cell(Gs=ps, Gc=Gc, Xt=X, Ht_1=H)

        def step():
            out = fwd()
            if train:
                for t in (X, H, *leaves):
                    t.grad = None
                out.backward(dHn)
                if bucket is not None:
                    bucket.allreduce()        # ONE parameter-gradient all-reduce per step
            return out

        # ---- parity: this rank's block of the partitioned result == the unpartitioned result on the same GPU ----
        parity = None
        if world > 1:
            out_p = step().detach().clone()
            got = [out_p] + ([X.grad.clone(), H.grad.clone()] + [t.grad.clone() for t in leaves] if train else [])
            Xf, Hf = Xg.to(dev).requires_grad_(train), Hg.to(dev).requires_grad_(train)
            for t in leaves:
                t.grad = None
            out_f = cell(Gs=full, Gc=Gc, Xt=Xf, Ht_1=Hf)
            want = [out_f.detach()[:, lo:lo + n]]
            if train:
                out_f.backward(dHg.to(dev))
                want += [Xf.grad[:, lo:lo + n], Hf.grad[:, lo:lo + n]] + [t.grad.clone() for t in leaves]
            names = ["H'", "dXt", "dH", "dWg", "dbg", "dWc", "dbc", "dGc"]
            worst, nbad = {}, 0
            for nm, a_, b_ in zip(names, got, want):
                # parameter gradients sum B*N*C = 1 M rows in a different order: the dW floor of tests/test_cell_gpu.py
                nb, w = violations(a_, b_, 3e-5 if nm.startswith("dW") or nm.startswith("db") or nm == "dGc" else 1e-5)
                worst[nm] = w
                nbad += nb
            flag = torch.tensor([float(nbad)], device=dev)
            dist.all_reduce(flag)
            parity = {"parity_ok": float(flag.item()) == 0.0, "worst_err_over_mean_ref_rank0": worst}
            del Xf, Hf, out_f, want, got
            torch.cuda.empty_cache()

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
        rec = {"kind": "config4_cell_step_partitioned" if train else "config4_cell_fwd_partitioned", "n_gpus": world, "N": N,
               "C": C, "F": F, "Ks": Ks, "B": B, "ms": ms, "cell_step_samples_per_s": B / ms * 1e3, "scaling": "strong",
               "param_allreduce": ("per cell backward" if per_cell else "once per step (flat bucket)") if train else None,
               "exchange": "side stream: device pack -> NCCL all-to-all -> device unpack, interior rows computed meanwhile",
               **halo}
        if parity:
            rec.update(parity)
        print(json.dumps(rec), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
