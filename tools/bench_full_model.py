#!/usr/bin/env python
"""Full reference STCGNN with the B200 cell installed, batch data-parallel (BASELINE configs[0] shape and configs[4]):

  python [-m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1] tools/bench_full_model.py \
         [--config sf|longc] [--batch B] [--steps K] [--stock] [--adam]

  sf     STCGNN(100, 5, Ks=Kc=2, 1, 16, 2 layers, horizon 3), T = 9   -- Main.py's defaults (config 1 / 2 shapes)
  longc  STCGNN(100, 64, Ks=Kc=2, 1, 64, 2 layers, horizon 3), T = 48 -- BASELINE config 5: learned dense 64 x 64 Gc

The model is the UNMODIFIED reference (found by the probe: $STC_REF_DIR, /root/reference/framework,
baseline/_ref/framework); `install(dp_group=...)` swaps the cell and makes MGP_Gen batch-data-parallel: the per-step
communication is two tiny score all-reduces forward, dGs/dGc backward and ONE flat bucket of the shard-dependent
gradients -- the 2 x Linear(N^2, N^2) fusion weights (800 MB of gradients at N = 100) are never communicated.
`--stock`: the stock reference cell through PyTorch eager on the same GPU (single rank), for the ratio.
`--adam`: include the reference's optimizer step (Adam lr 2e-3, wd 1e-4, Model_Trainer.py:35) in the timed region.
One step = forward + ComboLoss + backward (+ gradient exchange) of `batch` windows per GPU.  Rank 0 prints one JSON line.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CONFIGS = {"sf": dict(N=100, C=5, h=16, T=9, horizon=3), "longc": dict(N=100, C=64, h=64, T=48, horizon=3)}


def find_reference():
    for cand in (os.environ.get("STC_REF_DIR"), "/root/reference/framework", os.path.join(ROOT, "baseline", "_ref", "framework")):
        if cand and os.path.isfile(os.path.join(cand, "STC_GNN.py")):
            return cand
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="sf", choices=list(CONFIGS))
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--stock", action="store_true")
    ap.add_argument("--adam", action="store_true")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ref_dir = find_reference()
    if ref_dir is None:
        if rank == 0:
            print(json.dumps({"kind": "full_model", "unavailable": "the unmodified reference is not on this box"}))
        return
    sys.dont_write_bytecode = True
    sys.path.insert(0, ref_dir)
    import STC_GNN as ref
    import Model_Trainer as mt
    import stc_gnn_b200 as S
    from stc_gnn_b200 import _lib, mgp
    from stc_gnn_b200.install import install
    from stc_gnn_b200.synth import grid_adjacency

    if not args.stock:
        _lib.load()
        if world > 1:
            install(ref, dp_group=None, dp_average=True)
        else:
            install(ref)
    torch.manual_seed(0)
    model = ref.STCGNN(cfg["N"], cfg["C"], 2, 2, 1, cfg["h"], 2, cfg["horizon"]).to(dev)
    crit = mt.ComboLoss()
    opt = torch.optim.Adam(model.parameters(), lr=2e-3, weight_decay=1e-4) if args.adam else None
    g = torch.Generator().manual_seed(1 + rank)
    B = args.batch
    X = (torch.rand(B, cfg["T"], cfg["N"], cfg["C"], generator=g) < 0.1635).float().to(dev)
    Y = (torch.rand(B, cfg["horizon"], cfg["N"], cfg["C"], generator=g) < 0.1635).float().to(dev)
    As = grid_adjacency(10, 10).to(dev)
    gc = torch.Generator().manual_seed(0)
    Ac = torch.triu(torch.rand(cfg["C"], cfg["C"], generator=gc) * 0.36, 1)
    Ac = (Ac + Ac.t()).to(dev)
    bucket = S.dp.GradBucket(mgp.dp_bucket_parameters(model)) if world > 1 else None

    def step():
        model.zero_grad(set_to_none=True)
        loss = crit(model(X_seq=X, As=As, Ac=Ac), Y)
        loss.backward()
        if bucket is not None:
            bucket.allreduce(average=True)
        if opt is not None:
            opt.step()
        return loss

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    l0 = _lib.LAUNCHES
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank == 0:
        n_params = sum(p.numel() for p in model.parameters())
        print(json.dumps({
            "kind": "full_model_train_step", "config": args.config, "impl": "stock eager" if args.stock else "b200 cell installed",
            "n_gpus": world, "batch_per_gpu": B, "T": cfg["T"], "N": cfg["N"], "C": cfg["C"], "hidden": cfg["h"],
            "ms_per_step": ms, "samples_per_s": B * world / ms * 1e3, "loss": float(loss.item()), "adam_in_step": bool(opt),
            "parameters": n_params, "dp_bucket_floats": bucket.numel if bucket else 0,
            "cell_launches_per_step": (_lib.LAUNCHES - l0) / args.steps,
            "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 2**30, "scaling": "weak"}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
