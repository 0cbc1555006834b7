"""Synthetic supports / inputs of the benchmark configurations. Host-side helpers only."""
import torch


def grid_adjacency(rows: int, cols: int) -> torch.Tensor:
    """Binary 8-neighbour adjacency of a rows x cols grid, zero diagonal (the shipped s_adj is the 10x10 case)."""
    idx = torch.arange(rows * cols)
    r, c = idx // cols, idx % cols
    dr = (r[:, None] - r[None, :]).abs()
    dc = (c[:, None] - c[None, :]).abs()
    return ((dr <= 1) & (dc <= 1) & ((dr + dc) > 0)).float()


def sf_supports(seed: int = 0, rows: int = 10, cols: int = 10, C: int = 5):
    """Learned-like dense supports with the structure MixedFusion produces (STC_GNN.py:253-259):
    G = a * A + (1 - a) * P, A the prior (8-neighbour grid / symmetric U(0,0.36) category correlation),
    P a row-softmax of random scores, a = sigmoid(noise).  Dense, positive, denormal-free."""
    g = torch.Generator().manual_seed(seed)
    N = rows * cols
    As = grid_adjacency(rows, cols)
    Ps = torch.softmax(torch.relu(torch.randn(N, N, generator=g) * 3.0), dim=-1)
    a = torch.sigmoid(torch.randn(N, N, generator=g))
    Gs = a * As + (1 - a) * Ps
    Ac = torch.triu(torch.rand(C, C, generator=g) * 0.36, 1)
    Ac = Ac + Ac.t()
    Pc = torch.softmax(torch.relu(torch.randn(C, C, generator=g) * 3.0), dim=-1)
    ac = torch.sigmoid(torch.randn(C, C, generator=g))
    Gc = ac * Ac + (1 - ac) * Pc
    return Gs.contiguous(), Gc.contiguous()


def grid_csr(rows: int, cols: int, scale: float = 1.0 / 8.0):
    """8-neighbour grid adjacency scaled by `scale`, as CSR (rowptr, col, vals) host tensors."""
    A = (grid_adjacency(rows, cols) * scale).to_sparse_csr()
    return A.crow_indices().to(torch.int32), A.col_indices().to(torch.int32), A.values().float()


def knn_csr(N: int, k: int = 8, seed: int = 0):
    """Symmetric k-nearest-neighbour graph of N points ~ U([0,1]^2), values 1/deg(row), nodes in Morton order so that
    contiguous row blocks are spatially compact.  CSR (rowptr, col, vals) host tensors."""
    import numpy as np
    from scipy.spatial import cKDTree

    rng = np.random.default_rng(seed)
    pts = rng.random((N, 2))
    q = (pts * 65535).astype(np.uint64)

    def spread(v):
        v = (v | (v << 16)) & 0x0000FFFF0000FFFF
        v = (v | (v << 8)) & 0x00FF00FF00FF00FF
        v = (v | (v << 4)) & 0x0F0F0F0F0F0F0F0F
        v = (v | (v << 2)) & 0x3333333333333333
        v = (v | (v << 1)) & 0x5555555555555555
        return v

    order = np.argsort(spread(q[:, 0]) | (spread(q[:, 1]) << 1), kind="stable")
    pts = pts[order]
    _, nbr = cKDTree(pts).query(pts, k=k + 1)
    src = np.repeat(np.arange(N), k)
    dst = nbr[:, 1:].reshape(-1)
    a = np.concatenate([src, dst])
    b = np.concatenate([dst, src])
    key = np.unique(a.astype(np.int64) * N + b)
    row, col = key // N, key % N
    deg = np.bincount(row, minlength=N)
    rowptr = np.concatenate([[0], np.cumsum(deg)])
    vals = (1.0 / deg[row]).astype(np.float32)
    return (torch.from_numpy(rowptr.astype(np.int32)), torch.from_numpy(col.astype(np.int32)), torch.from_numpy(vals))
