"""Error margins of the wide-state parity cases: worst |got - ref| / mean|ref| per tensor and the number of elements
outside the test tolerance (rtol 1e-4 + 1e-5 x mean|ref|), for the config 3 shapes (N = 4096, C = 16, F = 64, CSR).
Run on a B200:  python tests/margins_config3.py [repeats]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import stc_gnn_b200 as S  # noqa: E402
from oracle import stc_oracle as O  # noqa: E402
from tests.helpers import oracle_cell_with_grads  # noqa: E402
from tests.test_cell_gpu import DEV, _csr_case, run_cuda_cell  # noqa: E402


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    N, C, Din, h = 4096, 16, 64, 64
    cfg = dict(B=1, N=N, C=C, Din=Din, h=h, Ks=2, Kc=2, activation=None)
    t, rowptr, col, vals = _csr_case(N, 8, 1, C, Din, h, 2, 2, seed=7)
    csr = S.CsrSupport(rowptr.to(DEV), col.to(DEV), vals.float().to(DEV), N)
    Hn_o, g_o = oracle_cell_with_grads(t, cfg)
    for rep in range(reps):
        Hn, g = run_cuda_cell(t, cfg, Gs_override=csr)
        row = {"rep": rep, "Hn": O.violations(Hn, Hn_o)}
        for k in ("dXt", "dH", "dWg", "dWc", "dbg", "dbc", "dGc"):
            row[k] = O.violations(g[k], g_o[k])
        print(json.dumps({k: (v if k == "rep" else [v[0], float(f"{v[1]:.3e}")]) for k, v in row.items()}))


if __name__ == "__main__":
    main()
