"""Dense support wider than one tile (N > 128): the tiled tcgen05 kernel against the FFMA tile kernel of the same
library (STC_DISABLE_TC_SUPPORT_BIG=1, run in a child process) and against cuBLAS fp32 through torch.einsum (the stock
reference's path for STC_GNN.py:37).  One JSON line per case.  Run on a B200:  python tools/bench_dense_support.py"""
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import stc_gnn_b200 as S  # noqa: E402
from stc_gnn_b200 import _lib  # noqa: E402

CASES = [(1024, 16, 16 * 64), (4096, 8, 16 * 64), (4096, 8, 16 * 128)]   # (N, B, W = C * F)


def time_ms(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    child = os.environ.get("STC_DISABLE_TC_SUPPORT_BIG", "") not in ("", "0")
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(0)
    for N, B, W in CASES:
        G = torch.softmax(torch.randn(N, N, device=dev, generator=g), dim=1)
        X = torch.randn(B, N, W, device=dev, generator=g)
        Y = torch.empty_like(X)
        flop = 2.0 * N * N * B * W
        iters = 5 if N >= 4096 else 20
        ms = time_ms(lambda: S.support_apply(G, X, transpose=True, out=Y), iters)
        row = {"case": f"dense N={N} B={B} W={W}", "kernel": "support_dense (FFMA)" if child else "tc_support_big (3xTF32 tcgen05)",
               "ms": round(ms, 3), "useful_TFLOPs": round(flop / ms * 1e-9, 1)}
        if not child:
            torch.backends.cuda.matmul.allow_tf32 = False
            ms_ref = time_ms(lambda: torch.einsum("bnw,nm->bmw", X, G), iters)
            ref = torch.einsum("bnw,nm->bmw", X.double(), G.double())
            err = (Y.double() - ref).abs().max().item() / ref.abs().mean().item()
            err_ref = (torch.einsum("bnw,nm->bmw", X, G).double() - ref).abs().max().item() / ref.abs().mean().item()
            row.update({"cublas_fp32_einsum_ms": round(ms_ref, 3), "max_err_over_mean_ref": float(f"{err:.2e}"),
                        "cublas_fp32_max_err_over_mean_ref": float(f"{err_ref:.2e}")})
        print(json.dumps(row), flush=True)
        # gradient of the same product w.r.t. the support: dGs[n][m] += sum_{b,j} X[b][n][j] * D[b][m][j]
        D = torch.randn(B, N, W, device=dev, generator=g)
        dG = torch.zeros(N, N, device=dev)
        lib = _lib.load()
        stream = torch.cuda.current_stream().cuda_stream
        ms = time_ms(lambda: _lib.check(lib.stc_support_outer(N, B, W, X.data_ptr(), N * W, D.data_ptr(), 1.0, dG.data_ptr(),
                                                              stream), "stc_support_outer"), iters)
        row = {"case": f"dGs outer N={N} B={B} W={W}", "kernel": "support_outer (FFMA)" if child else "tc_outer, 128x128 blocks",
               "ms": round(ms, 3), "useful_TFLOPs": round(flop / ms * 1e-9, 1)}
        if not child:
            ms_ref = time_ms(lambda: torch.einsum("bnw,bmw->nm", X, D), iters)
            dG.zero_()
            _lib.check(lib.stc_support_outer(N, B, W, X.data_ptr(), N * W, D.data_ptr(), 1.0, dG.data_ptr(), stream), "outer")
            ref = torch.einsum("bnw,bmw->nm", X.double(), D.double())
            err = (dG.double() - ref).abs().max().item() / ref.abs().mean().item()
            row.update({"cublas_fp32_einsum_ms": round(ms_ref, 3), "max_err_over_mean_ref": float(f"{err:.2e}")})
        print(json.dumps(row), flush=True)
    if not child:
        env = dict(os.environ, STC_DISABLE_TC_SUPPORT_BIG="1")
        subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, check=False)


if __name__ == "__main__":
    main()
