#!/usr/bin/env python
"""Side measurements for the BASELINE.json configurations that are not bench.py's headline line:

  support   the spatial-support kernel alone (stc_support_apply) -- "STC graph-conv HBM GB/s vs peak":
            dense SF (N=100), CSR grid N=4096 (config 3), CSR kNN N=65,536 (config 4), forward and adjoint
  config3   one STC cell step forward+backward, N=4096 C=16 F=64 Ks=Kc=2, CSR grid support (per-kernel breakdown)
  config4   one STC cell step forward, N=65,536 C=8 F=64 Ks=4 (3 hops) Kc=2, CSR kNN support
  sweep     bench.py's SF workload over B = 32 ... 4096 (config 2), eager and CUDA-graph replay of the whole step

  python tools/bench_configs.py [support] [config3] [config4] [sweep] > gpurun_out/configs.jsonl

One JSON object per line.  Algorithmic bytes are the compulsory ones (inputs read once + outputs written once +
the support itself); peaks from MEASURED_PEAKS.json when present.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import stc_gnn_b200 as S  # noqa: E402
from stc_gnn_b200 import _lib  # noqa: E402
from stc_gnn_b200.support import support_apply  # noqa: E402
from stc_gnn_b200.synth import grid_csr, knn_csr, sf_supports  # noqa: E402

DEV = torch.device("cuda:0")


def peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def timed(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def emit(**kw):
    print(json.dumps(kw), flush=True)


def kernel_breakdown(fn, iters=3):
    _lib.timing_enable(True)
    _lib.timing_collect()
    for _ in range(iters):
        fn()
    torch.cuda.synchronize()
    _lib.timing_enable(False)
    kinds = _lib.timing_collect()
    return {k: {"ms": v[0] / iters, "launches": v[1] / iters, "alg_GBs": (v[2] / (v[0] * 1e-3) / 1e9) if v[0] else None}
            for k, v in sorted(kinds.items(), key=lambda kv: -kv[1][0])}


def bench_support():
    pk, src = peak()
    cases = []
    Gs, _ = sf_supports()
    cases.append(("dense_sf_N100_W80", Gs.to(DEV), 100, 80, 4096, 4 * 100 * 100))
    r, c, v = grid_csr(64, 64)
    csr = S.CsrSupport(r.to(DEV), c.to(DEV), v.to(DEV), 4096)
    cases.append(("csr_grid_N4096_W1024", csr, 4096, 1024, 16, 8 * csr.nnz + 4 * 4097))
    r, c, v = knn_csr(65536, 8)
    knn = S.CsrSupport(r.to(DEV), c.to(DEV), v.to(DEV), 65536)
    cases.append(("csr_knn_N65536_W512", knn, 65536, 512, 2, 8 * knn.nnz + 4 * 65537))
    cases.append(("csr_knn_N65536_W128", knn, 65536, 128, 8, 8 * knn.nnz + 4 * 65537))
    for name, G, N, W, B, gbytes in cases:
        X = torch.randn(B, N, W, device=DEV)
        Z = torch.randn(B, N, W, device=DEV)
        for label, kw, streams in (("fwd Y=Gs^T X", dict(transpose=True), 2),
                                   ("adj Y=Gs X", dict(transpose=False), 2),
                                   ("cheb Y=2 Gs^T X - Z", dict(transpose=True, alpha=2.0, beta=-1.0, Z=Z), 3)):
            ms = timed(lambda: support_apply(G, X, **kw), 20)
            alg = 4.0 * B * N * W * streams + gbytes
            emit(kind="support", case=name, op=label, B=B, ms=ms, alg_bytes=alg, achieved_GBs=alg / ms / 1e6,
                 peak_GBs=pk, peak_source=src, frac=alg / ms / 1e6 / pk,
                 l2="operands %.0f MB (L2 is 126 MB)" % (4.0 * B * N * W * streams / 1e6))
        del X, Z


def bench_config3():
    pk, src = peak()
    N, C, F, B = 4096, 16, 64, 8
    r, c, v = grid_csr(64, 64)
    csr = S.CsrSupport(r.to(DEV), c.to(DEV), v.to(DEV), N)
    Gc = (torch.rand(C, C) / C).to(DEV)
    cell = S.STC_Cell(N, C, 2, 2, F, F).to(DEV)
    X = torch.randn(B, N, C, F, device=DEV, requires_grad=True)
    H = (torch.randn(B, N, C, F, device=DEV) * 0.5).requires_grad_(True)
    dH = torch.randn(B, N, C, F, device=DEV)

    def step():
        for p in cell.parameters():
            p.grad = None
        X.grad = H.grad = None
        cell(Gs=csr, Gc=Gc, Xt=X, Ht_1=H).backward(dH)

    ms = timed(step, 5, warm=2)
    alg = 4.0 * N * C * (3 * F + 11 * F) * B          # SURVEY 8d ALG_BYTES_TRAIN
    flops = 3 * B * (2.0 * N * C * 4 * (2 * F) * 3 * F)   # gate GEMM fwd + 2x bwd
    emit(kind="config3_cell_step", N=N, C=C, F=F, B=B, Ks=2, Kc=2, support="csr grid 8-neighbour", ms=ms,
         cell_step_samples_per_s=B / ms * 1e3, alg_bytes_train=alg, achieved_GBs=alg / ms / 1e6, peak_GBs=pk,
         frac_hbm=alg / ms / 1e6 / pk, gate_gemm_TFLOPs=flops / ms / 1e9, kernels=kernel_breakdown(step),
         note="gate contraction: wide-state tcgen05 kernels (stc_conv_tc_big.cu), 3xTF32; gate_gemm_TFLOPs counts useful FLOPs once")


def bench_config4():
    pk, src = peak()
    N, C, F, B, Ks = 65536, 8, 64, 2, 4
    r, c, v = knn_csr(N, 8)
    csr = S.CsrSupport(r.to(DEV), c.to(DEV), v.to(DEV), N)
    Gc = (torch.rand(C, C) / C).to(DEV)
    cell = S.STC_Cell(N, C, Ks, 2, F, F).to(DEV)
    X = torch.randn(B, N, C, F, device=DEV)
    H = torch.randn(B, N, C, F, device=DEV) * 0.5

    def step():
        with torch.no_grad():
            cell(Gs=csr, Gc=Gc, Xt=X, Ht_1=H)

    ms = timed(step, 5, warm=2)
    alg = 4.0 * N * C * (F + 2 * F) * B               # SURVEY 8d ALG_BYTES_FWD
    emit(kind="config4_cell_fwd", N=N, C=C, F=F, B=B, Ks=Ks, Kc=2, support="csr kNN k=8 symmetric, nnz=%d" % csr.nnz,
         ms=ms, cell_step_samples_per_s=B / ms * 1e3, alg_bytes_fwd=alg, achieved_GBs=alg / ms / 1e6, peak_GBs=pk,
         frac_hbm=alg / ms / 1e6 / pk, kernels=kernel_breakdown(step))


def bench_config5():
    """LongC-like cell step (BASELINE configs[4] shapes): N = 100 regions, C = 64 categories, F = 64, dense learned supports
    (dGs, dGc required), forward + backward, B = 16."""
    pk, src = peak()
    N, C, F, B = 100, 64, 64, 16
    Gs, _ = sf_supports()
    Gs = Gs.to(DEV).requires_grad_(True)
    Gc = torch.softmax(torch.randn(C, C, generator=torch.Generator().manual_seed(0)), dim=-1).to(DEV).requires_grad_(True)
    cell = S.STC_Cell(N, C, 2, 2, F, F).to(DEV)
    X = torch.randn(B, N, C, F, device=DEV, requires_grad=True)
    H = (torch.randn(B, N, C, F, device=DEV) * 0.5).requires_grad_(True)
    dH = torch.randn(B, N, C, F, device=DEV)

    def step():
        for p in list(cell.parameters()) + [Gs, Gc, X, H]:
            p.grad = None
        cell(Gs=Gs, Gc=Gc, Xt=X, Ht_1=H).backward(dH)

    ms = timed(step, 5, warm=2)
    alg = 4.0 * N * C * (3 * F + 11 * F) * B
    emit(kind="config5_cell_step", N=N, C=C, F=F, B=B, Ks=2, Kc=2, support="dense learned (dGs, dGc)", ms=ms,
         cell_step_samples_per_s=B / ms * 1e3, alg_bytes_train=alg, achieved_GBs=alg / ms / 1e6, peak_GBs=pk,
         kernels=kernel_breakdown(step))


def bench_sweep():
    import subprocess
    for B in (32, 64, 128, 256, 512, 1024, 2048, 4096):
        for graph in (0, 1):
            if graph and B > 1024:
                continue
            cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--batch", str(B), "--steps", "10", "--warmup", "3",
                   "--no-cpu-baseline", "--no-roofline"] + (["--cuda-graph"] if graph else [])
            res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
            line = [l for l in res.stdout.splitlines() if l.startswith("{")]
            if not line:
                emit(kind="sweep", B=B, cuda_graph=bool(graph), error=(res.stderr or res.stdout)[-400:])
                continue
            d = json.loads(line[-1])
            emit(kind="sweep", B=B, cuda_graph=bool(graph), samples_per_s=d["value"], ms_per_step=d["ms_per_step"],
                 e2e_samples_per_s=d["e2e"]["value"], gpu_launches=d["gpu_launches"], sm_mhz=(d.get("clocks") or {}).get("sm_mhz"))


if __name__ == "__main__":
    which = sys.argv[1:] or ["support", "config3", "config4"]
    _lib.load()
    for w in which:
        try:
            {"support": bench_support, "config3": bench_config3, "config4": bench_config4, "config5": bench_config5,
             "sweep": bench_sweep}[w]()
        except Exception as e:  # keep going: each part is an independent measurement
            emit(kind=w, error=repr(e)[:600])
            torch.cuda.synchronize()
