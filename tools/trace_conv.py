#!/usr/bin/env python
"""Diagnostic: clock64 phase stamps of CTA 0 of the tcgen05 gate-convolution kernels (stc_debug_trace_set).

  python tools/trace_conv.py [B] [Din] > gpurun_out/trace.txt

Prints, per kernel (conv_fwd gates / candidate, conv_bwd_dx gates / candidate), the mean cycle count between
consecutive stamps over CTA 0's first tiles (the first two tiles are dropped: cold weights / pipeline fill).
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stc_gnn_b200 as S  # noqa: E402
from stc_gnn_b200 import _lib  # noqa: E402
from stc_gnn_b200.synth import sf_supports  # noqa: E402

# smem-A kernels (STC_OPT=9) stamp every phase; the A-in-TMEM kernels (default) have no second build / mid-tile wait,
# so their stamps 2 -> 3 coincide
FWD = ["tile start -> row loads issued (smem-A: stage landed)", "-> A operand written + barrier (smem-A: first atom built)",
       "-> (smem-A only: MMA of first atom done, single A buffer)",
       "-> MMAs issued, epilogue operands requested (smem-A: last atom built + MMAs issued)", "-> last MMA done",
       "-> epilogue operands landed", "-> epilogue done", "-> tile sync (next tile start)"]
DX = ["tile start -> elementwise adjoint done", "-> A operand [Ds | Dm] written + barrier (smem-A: first atom built)",
      "-> MMAs issued (smem-A: last atom built + MMAs issued)", "-> dQ partial sums done", "-> last MMA done",
      "-> epilogue done", "-> tile sync", "-> next tile start"]


def report(name, buf, tiles, labels):
    t = buf.view(tiles, 16).cpu().double()
    t = t[(t[:, :8] > 0).all(dim=1)]
    if t.shape[0] < 4:
        print(f"{name}: only {t.shape[0]} traced tiles")
        return
    t = t[2:]
    d = t[:, 1:8] - t[:, :7]
    nxt = t[1:, 0] - t[:-1, 7]
    per_tile = (t[1:, 0] - t[:-1, 0]).mean().item()
    print(f"{name}: {t.shape[0]} tiles, {per_tile:.0f} cycles per tile")
    for i in range(7):
        print(f"   {d[:, i].mean().item():8.0f}  {labels[i]}")
    print(f"   {nxt.mean().item():8.0f}  {labels[7]}")
    if (t[:, 12:14] > 0).all():
        print(f"   A operand (thread 0): values + stores {(t[:, 12] - t[:, 1]).mean().item():.0f}, proxy fence {(t[:, 13] - t[:, 12]).mean().item():.0f}, "
              f"barrier {(t[:, 2] - t[:, 13]).mean().item():.0f}")
    if (t[:, 8:12] > 0).all():
        print(f"   MMA issue (thread 0): first atom {(t[:, 8] - t[:, 2]).mean().item():.0f} cycles after its build barrier, "
              f"commit +{(t[:, 9] - t[:, 8]).mean().item():.0f}; last atom issue done {(t[:, 10] - t[:, 9]).mean().item():.0f} "
              f"after the first commit, commit +{(t[:, 11] - t[:, 10]).mean().item():.0f}")


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    Din = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    dev = torch.device("cuda:0")
    lib = _lib.load()
    tiles = 40
    buf = torch.zeros(2 * tiles * 16, dtype=torch.int64, device=dev)
    Gs, Gc = sf_supports()
    Gs, Gc = Gs.to(dev).requires_grad_(True), Gc.to(dev).requires_grad_(True)
    cell = S.STC_Cell(100, 5, 2, 2, Din, 16).to(dev)
    X = torch.randn(B, 100, 5, Din, device=dev, requires_grad=True)
    H = torch.randn(B, 100, 5, 16, device=dev, requires_grad=True)
    for _ in range(2):
        cell(Gs=Gs, Gc=Gc, Xt=X, Ht_1=H).sum().backward()
    torch.cuda.synchronize()
    print(f"# B={B} Din={Din} STC_OPT={os.environ.get('STC_OPT', 'default')}")
    _lib.check(lib.stc_debug_trace_set(buf.data_ptr(), buf.numel()), "trace_set")
    out = cell(Gs=Gs, Gc=Gc, Xt=X, Ht_1=H)
    torch.cuda.synchronize()
    report("conv_fwd gates", buf[: tiles * 16].clone(), tiles, FWD)
    report("conv_fwd candidate", buf[tiles * 16:].clone(), tiles, FWD)
    buf.zero_()
    out.sum().backward()
    torch.cuda.synchronize()
    report("conv_bwd_dx gates", buf[: tiles * 16].clone(), tiles, DX)
    report("conv_bwd_dx candidate", buf[tiles * 16:].clone(), tiles, DX)
    lib.stc_debug_trace_set(None, 0)


if __name__ == "__main__":
    main()
