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
