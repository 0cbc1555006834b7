"""CPU oracle for the STC-GNN hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module.  The product path
(``stc_gnn_b200``) never imports it and has no CPU fallback.

It restates, on CPU tensors (fp64 by default), the algorithm of the reference's
per-timestep spatio-temporal-categorical graph-convolution GRU cell:

* ``BDG_Dif.cheby_poly``  -- /root/reference/framework/STC_GNN.py:24-29
* ``BDG_Dif.forward``     -- /root/reference/framework/STC_GNN.py:31-47
* ``STC_Cell.forward``    -- /root/reference/framework/STC_GNN.py:65-79
* the encoder / decoder roll-out that drives the cell
                          -- /root/reference/framework/STC_GNN.py:107-118, 160-163, 194-204

Parity pin: the reference ships no tests or golden vectors, so the
pin is the reference itself: ``tests/golden/make_golden.py`` imports the real
reference in the build container and commits seeded input/output/gradient
vectors under ``tests/golden/``; ``tests/test_oracle.py`` checks this oracle
against every one of them (and against the live reference when
``/root/reference`` is present).

Two restatements live here on purpose:

* the *lean* one (``bdg_dif`` / ``stc_cell``): feature-side Chebyshev recurrence
  (never forms T_k(Gs)), dense or sparse spatial support, separate Xt / H
  operands.  This is the checker, and the only form that can run N = 65,536.
* the *reference-shaped* one (``bdg_dif_refshape`` / ``stc_cell_refshape``):
  executes the same operator sequence the reference does (matrix-space Chebyshev
  terms, Ks*Kc pairs of mode products, concatenate, one weight contraction), so
  that a CPU timing of it is a fair stand-in ("port") for the reference's CPU
  cost on a box where /root/reference does not exist.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import torch

Tensor = torch.Tensor


# --------------------------------------------------------------------------- #
# tolerance: |a-b| <= rtol*|b| + atol_scale*mean|b|
# --------------------------------------------------------------------------- #
RTOL = 1e-4
ATOL_SCALE = 1e-5


def violations(got: Tensor, ref: Tensor, rtol: float = RTOL, atol_scale: float = ATOL_SCALE) -> Tuple[int, float]:
    """Number of elements outside ``rtol*|ref| + atol_scale*mean|ref|`` and the worst
    error expressed in units of mean|ref|."""
    got = got.detach().double().cpu()
    ref = ref.detach().double().cpu()
    assert got.shape == ref.shape, (got.shape, ref.shape)
    if ref.numel() == 0:
        return 0, 0.0
    scale = ref.abs().mean().item()
    err = (got - ref).abs()
    bad = err > (rtol * ref.abs() + atol_scale * scale)
    worst = (err.max().item() / scale) if scale > 0 else err.max().item()
    return int(bad.sum().item()), worst


def assert_close(got: Tensor, ref: Tensor, what: str = "", rtol: float = RTOL, atol_scale: float = ATOL_SCALE) -> None:
    nbad, worst = violations(got, ref, rtol, atol_scale)
    if nbad:
        raise AssertionError(
            f"{what}: {nbad}/{ref.numel()} elements outside rtol={rtol} + {atol_scale}*mean|ref| "
            f"(worst abs err = {worst:.3e} x mean|ref|)")


# --------------------------------------------------------------------------- #
# supports
# --------------------------------------------------------------------------- #
def _is_sparse(G) -> bool:
    return isinstance(G, Tensor) and G.layout != torch.strided


def support_T_apply(G, X: Tensor) -> Tensor:
    """Y[b,m,...] = sum_n G[n,m] * X[b,n,...]   (the reference's 'bncl,nm->bmcl', STC_GNN.py:37:
    contraction over G's FIRST index, i.e. G^T acts on the node axis)."""
    B, N = X.shape[0], X.shape[1]
    tail = X.shape[2:]
    X2 = X.reshape(B, N, -1)
    if _is_sparse(G):
        Gt = G.to_sparse_coo().t().coalesce()
        # [N, B*F] layout so one sparse product serves the whole batch
        F = X2.shape[-1]
        Xn = X2.permute(1, 0, 2).reshape(N, B * F)
        Yn = torch.sparse.mm(Gt, Xn)
        Y2 = Yn.reshape(N, B, F).permute(1, 0, 2)
    else:
        Y2 = torch.matmul(G.t().unsqueeze(0), X2)
    return Y2.reshape(B, N, *tail)


def spatial_terms(X: Tensor, Gs, Ks: int) -> List[Tensor]:
    """Feature-side Chebyshev recurrence: Y_0 = X, Y_1 = Gs^T X, Y_k = 2 Gs^T Y_{k-1} - Y_{k-2}.
    Equals einsum(X, T_k(Gs)) with T_k from STC_GNN.py:24-29 (no Laplacian rescaling there)."""
    Y = [X]
    if Ks > 1:
        Y.append(support_T_apply(Gs, X))
    for _ in range(2, Ks):
        Y.append(2.0 * support_T_apply(Gs, Y[-1]) - Y[-2])
    return Y


def cheby_matrix_terms(G: Tensor, K: int) -> List[Tensor]:
    """Matrix-space terms [I, G, 2 G T_{k-1} - T_{k-2}, ...]  (STC_GNN.py:24-29). ``K`` is the
    NUMBER of terms (orders 0..K-1)."""
    n = G.shape[0]
    terms = [torch.eye(n, dtype=G.dtype, device=G.device)]
    if K > 1:
        terms.append(G)
    for _ in range(2, K):
        terms.append(2.0 * (G @ terms[-1]) - terms[-2])
    return terms[:K]


def categorical_T_apply(Q: Tensor, Y: Tensor) -> Tensor:
    """Z[b,m,d,l] = sum_c Q[c,d] * Y[b,m,c,l]   ('bmcl,cd->bmdl', STC_GNN.py:38)."""
    return torch.einsum("bmcl,cd->bmdl", Y, Q)


def _activate(x: Tensor, activation: Optional[str]) -> Tensor:
    if activation is None or activation == "none":
        return x
    if activation == "relu":
        return torch.relu(x)
    raise ValueError(f"unsupported activation {activation!r}")


# --------------------------------------------------------------------------- #
# lean restatement
# --------------------------------------------------------------------------- #
def bdg_dif(X: Tensor, Gs, Gc: Tensor, W: Tensor, b: Optional[Tensor], Ks: int, Kc: int,
            activation: Optional[str] = None, spatial_terms_fn=None) -> Tensor:
    """out[b,m,d,:] = sum_{n<Ks, c<Kc}  (T_c(Gc)^T (x) T_n(Gs)^T X)[b,m,d,:] @ W[(n*Kc+c)*L:(n*Kc+c+1)*L]  + b
    (STC_GNN.py:31-47; W row order is (n, c, l) because feat_coll is appended n-outer/c-inner and
    concatenated on the last axis, :35-41).  ``spatial_terms_fn(X) -> [Y_0..Y_{Ks-1}]`` replaces the spatial
    recurrence (the row-partitioned tests inject halo-exchanging hops; Gs is then unused)."""
    B, N, C, L = X.shape
    Hout = W.shape[1]
    assert W.shape[0] == Ks * Kc * L, (W.shape, Ks, Kc, L)
    Wv = W.reshape(Ks, Kc, L, Hout)
    Q = cheby_matrix_terms(Gc, Kc)
    out = None
    terms = spatial_terms_fn(X) if spatial_terms_fn is not None else spatial_terms(X, Gs, Ks)
    for n, Yn in enumerate(terms):
        for c in range(Kc):
            F = Yn if c == 0 else categorical_T_apply(Q[c], Yn)
            term = F @ Wv[n, c]
            out = term if out is None else out + term
    if b is not None:
        out = out + b
    return _activate(out, activation)


def stc_cell(Gs, Gc: Tensor, Xt: Tensor, H: Tensor, Wg: Tensor, bg: Optional[Tensor], Wc: Tensor,
             bc: Optional[Tensor], Ks: int, Kc: int, activation: Optional[str] = None,
             return_gates: bool = False, spatial_terms_fn=None):
    """GRU cell of STC_GNN.py:65-79: [u|r] = sigmoid(conv_g([Xt,H])), c = tanh(conv_c([Xt, r*H])),
    H' = (1-u)*H + u*c.  u is the FIRST h output channels, r the last h (torch.split, :71)."""
    assert Xt.dim() == 4 and H.dim() == 4
    h = H.shape[-1]
    g = bdg_dif(torch.cat([Xt, H], dim=-1), Gs, Gc, Wg, bg, Ks, Kc, activation, spatial_terms_fn)
    u = torch.sigmoid(g[..., :h])
    r = torch.sigmoid(g[..., h:])
    c = torch.tanh(bdg_dif(torch.cat([Xt, r * H], dim=-1), Gs, Gc, Wc, bc, Ks, Kc, activation, spatial_terms_fn))
    Hn = (1.0 - u) * H + u * c
    if return_gates:
        return Hn, u, r, c
    return Hn


# --------------------------------------------------------------------------- #
# reference-shaped restatement (same operator sequence as the reference; for CPU timing)
# --------------------------------------------------------------------------- #
def bdg_dif_refshape(X: Tensor, Gs: Tensor, Gc: Tensor, W: Tensor, b: Optional[Tensor], Ks: int, Kc: int,
                     activation: Optional[str] = None) -> Tensor:
    S = cheby_matrix_terms(Gs, Ks)
    Q = cheby_matrix_terms(Gc, Kc)
    blocks = []
    for n in range(Ks):
        for c in range(Kc):
            blocks.append(torch.einsum("bmcl,cd->bmdl", torch.einsum("bncl,nm->bmcl", X, S[n]), Q[c]))
    out = torch.einsum("bmdk,kh->bmdh", torch.cat(blocks, dim=-1), W)
    if b is not None:
        out = out + b
    return _activate(out, activation)


def stc_cell_refshape(Gs, Gc, Xt, H, Wg, bg, Wc, bc, Ks, Kc, activation=None):
    h = H.shape[-1]
    g = bdg_dif_refshape(torch.cat([Xt, H], dim=-1), Gs, Gc, Wg, bg, Ks, Kc, activation)
    u, r = torch.sigmoid(g[..., :h]), torch.sigmoid(g[..., h:])
    c = torch.tanh(bdg_dif_refshape(torch.cat([Xt, r * H], dim=-1), Gs, Gc, Wc, bc, Ks, Kc, activation))
    return (1.0 - u) * H + u * c


# --------------------------------------------------------------------------- #
# analytic backward -- independent of autograd; used to check the
# formulas the CUDA backward kernels implement, intermediate by intermediate.
# --------------------------------------------------------------------------- #
def support_apply(G, X: Tensor) -> Tensor:
    """Y[b,n',...] = sum_m G[n',m] X[b,m,...]  (adjoint of ``support_T_apply``)."""
    B, N = X.shape[0], X.shape[1]
    tail = X.shape[2:]
    X2 = X.reshape(B, N, -1)
    if _is_sparse(G):
        Gc_ = G.to_sparse_coo().coalesce()
        F = X2.shape[-1]
        Yn = torch.sparse.mm(Gc_, X2.permute(1, 0, 2).reshape(N, B * F))
        Y2 = Yn.reshape(N, B, F).permute(1, 0, 2)
    else:
        Y2 = torch.matmul(G.unsqueeze(0), X2)
    return Y2.reshape(B, N, *tail)


def bdg_dif_backward(X: Tensor, Gs, Gc: Tensor, W: Tensor, dOut: Tensor, Ks: int, Kc: int,
                     need_dGs: bool = True):
    """Backward of the *linear* part of ``bdg_dif`` (dOut is the gradient w.r.t. the pre-activation
    output).  Returns dX, dW, db, dGs (dense only, else None), dGc."""
    B, N, C, L = X.shape
    Hout = W.shape[1]
    Wv = W.reshape(Ks, Kc, L, Hout)
    Q = cheby_matrix_terms(Gc, Kc)
    Y = spatial_terms(X, Gs, Ks)
    db = dOut.sum(dim=(0, 1, 2))
    dW = torch.zeros_like(Wv)
    dQ = [torch.zeros_like(Gc) for _ in range(Kc)]
    dY = []
    for n in range(Ks):
        dYn = torch.zeros_like(X)
        for c in range(Kc):
            F = Y[n] if c == 0 else categorical_T_apply(Q[c], Y[n])
            dW[n, c] = torch.einsum("bmdl,bmdh->lh", F, dOut)
            dF = dOut @ Wv[n, c].t()                               # [B,N,C,L]
            if c == 0:
                dYn = dYn + dF
            else:
                dYn = dYn + torch.einsum("bmdl,cd->bmcl", dF, Q[c])
                dQ[c] = dQ[c] + torch.einsum("bmcl,bmdl->cd", Y[n], dF)
        dY.append(dYn)
    # categorical Chebyshev chain in matrix space (C x C, tiny)
    dGc = torch.zeros_like(Gc)
    for k in range(Kc - 1, 1, -1):
        dGc = dGc + 2.0 * dQ[k] @ Q[k - 1].t()
        dQ[k - 1] = dQ[k - 1] + 2.0 * Gc.t() @ dQ[k]
        dQ[k - 2] = dQ[k - 2] - dQ[k]
    if Kc > 1:
        dGc = dGc + dQ[1]
    # spatial recurrence in reverse (feature side):  Y_k = 2 Gs^T Y_{k-1} - Y_{k-2}
    dGs = None
    dense = not _is_sparse(Gs)
    if need_dGs and dense:
        dGs = torch.zeros_like(Gs)
    ybar = list(dY)
    for k in range(Ks - 1, 1, -1):
        if dGs is not None:
            dGs = dGs + 2.0 * torch.einsum("bncl,bmcl->nm", Y[k - 1], ybar[k])
        ybar[k - 1] = ybar[k - 1] + 2.0 * support_apply(Gs, ybar[k])
        ybar[k - 2] = ybar[k - 2] - ybar[k]
    if Ks > 1:
        if dGs is not None:
            dGs = dGs + torch.einsum("bncl,bmcl->nm", Y[0], ybar[1])
        ybar[0] = ybar[0] + support_apply(Gs, ybar[1])
    return ybar[0], dW.reshape(W.shape), db, dGs, dGc


def _act_grad_mask(post: Tensor, kind: str, activation: Optional[str]) -> Tensor:
    """d(act)/d(pre) recovered from the saved *post-nonlinearity* value.
    relu(pre) > 0  <=>  sigmoid(relu(pre)) > 1/2  <=>  tanh(relu(pre)) > 0."""
    if activation is None or activation == "none":
        return torch.ones_like(post)
    if kind == "sigmoid":
        return (post > 0.5).to(post.dtype)
    return (post > 0.0).to(post.dtype)


def stc_cell_backward(Gs, Gc, Xt, H, Wg, bg, Wc, bc, Ks, Kc, dHn, activation=None, need_dGs=True):
    """Analytic backward of ``stc_cell``. Returns a dict of gradients."""
    h = H.shape[-1]
    Din = Xt.shape[-1]
    _, u, r, c = stc_cell(Gs, Gc, Xt, H, Wg, bg, Wc, bc, Ks, Kc, activation, return_gates=True)
    du = dHn * (c - H)
    dc = dHn * u
    dH = dHn * (1.0 - u)
    dpre_c = dc * (1.0 - c * c) * _act_grad_mask(c, "tanh", activation)
    Xc = torch.cat([Xt, r * H], dim=-1)
    dXc, dWc, dbc, dGs_c, dGc_c = bdg_dif_backward(Xc, Gs, Gc, Wc, dpre_c, Ks, Kc, need_dGs)
    dXt = dXc[..., :Din]
    drH = dXc[..., Din:]
    dr = drH * H
    dH = dH + drH * r
    dpre_g = torch.cat([du * u * (1.0 - u) * _act_grad_mask(u, "sigmoid", activation),
                        dr * r * (1.0 - r) * _act_grad_mask(r, "sigmoid", activation)], dim=-1)
    Xg = torch.cat([Xt, H], dim=-1)
    dXg, dWg, dbg, dGs_g, dGc_g = bdg_dif_backward(Xg, Gs, Gc, Wg, dpre_g, Ks, Kc, need_dGs)
    dXt = dXt + dXg[..., :Din]
    dH = dH + dXg[..., Din:]
    out = dict(dXt=dXt, dH=dH, dWg=dWg, dWc=dWc, dGc=dGc_c + dGc_g,
               dbg=dbg if bg is not None else None, dbc=dbc if bc is not None else None,
               dGs=(dGs_c + dGs_g) if dGs_c is not None else None)
    return out


# --------------------------------------------------------------------------- #
# recurrent roll-out that drives the cell (encoder layers x T, then decoder horizon x layers)
# --------------------------------------------------------------------------- #
class CellParams:
    """Weights of one cell, with the reference's shapes (STC_GNN.py:17-21, 57-58)."""

    def __init__(self, Wg: Tensor, bg: Optional[Tensor], Wc: Tensor, bc: Optional[Tensor]):
        self.Wg, self.bg, self.Wc, self.bc = Wg, bg, Wc, bc

    def tensors(self):
        return [t for t in (self.Wg, self.bg, self.Wc, self.bc) if t is not None]


def xavier_cell_params(Din: int, h: int, Ks: int, Kc: int, gen: torch.Generator, dtype=torch.float64,
                       use_bias: bool = True, bias_scale: float = 0.0) -> CellParams:
    """Xavier-normal W (std = sqrt(2/(fan_in+fan_out)), as nn.init.xavier_normal_ on a 2-D [rows, cols]
    tensor), b = 0 unless ``bias_scale`` is given (tests use non-zero biases to exercise the add)."""
    L = Din + h
    rows = L * Ks * Kc

    def xav(cols):
        std = math.sqrt(2.0 / (rows + cols))
        return torch.randn(rows, cols, generator=gen, dtype=torch.float64).mul_(std).to(dtype)

    Wg = xav(2 * h)
    Wc = xav(h)
    if use_bias:
        bg = (torch.randn(2 * h, generator=gen, dtype=torch.float64) * bias_scale).to(dtype)
        bc = (torch.randn(h, generator=gen, dtype=torch.float64) * bias_scale).to(dtype)
    else:
        bg = bc = None
    return CellParams(Wg, bg, Wc, bc)


def encoder_rollout(cell_fn, Gs, Gc, X_seq: Tensor, enc: Sequence[CellParams], Ks: int, Kc: int,
                    activation=None) -> Tuple[List[Tensor], List[Tensor]]:
    """STC_Encoder.forward (STC_GNN.py:97-123): for each layer, run the cell over t = 0..T-1 starting
    from zeros; the stacked outputs of layer l are the input sequence of layer l+1."""
    B, T = X_seq.shape[0], X_seq.shape[1]
    N, C = X_seq.shape[2], X_seq.shape[3]
    seq = X_seq
    out_seqs, last = [], []
    for p in enc:
        h = p.Wc.shape[1]
        Ht = torch.zeros(B, N, C, h, dtype=X_seq.dtype, device=X_seq.device)
        outs = []
        for t in range(T):
            Ht = cell_fn(Gs, Gc, seq[:, t], Ht, p.Wg, p.bg, p.Wc, p.bc, Ks, Kc, activation)
            outs.append(Ht)
        seq = torch.stack(outs, dim=1)
        out_seqs.append(seq)
        last.append(Ht)
    return out_seqs, last


def decoder_rollout(cell_fn, Gs, Gc, H_last: List[Tensor], dec: Sequence[CellParams], horizon: int,
                    Ks: int, Kc: int, activation=None) -> Tensor:
    """STCGNN.forward decode loop (STC_GNN.py:194-204) over STC_Decoder.forward (:154-166): the decoder's
    first input is the last encoder layer's final state; every step feeds its top-layer state back in and
    carries all layer states forward."""
    states = list(H_last)
    x = states[-1]
    outs = []
    for _ in range(horizon):
        new_states = []
        inp = x
        for l, p in enumerate(dec):
            Hl = cell_fn(Gs, Gc, inp, states[l], p.Wg, p.bg, p.Wc, p.bc, Ks, Kc, activation)
            new_states.append(Hl)
            inp = Hl
        states = new_states
        x = states[-1]
        outs.append(x)
    return torch.stack(outs, dim=1)          # [B, horizon, N, C, h]


def stack_forward(Gs, Gc, X_seq: Tensor, enc: Sequence[CellParams], dec: Sequence[CellParams],
                  horizon: int, Ks: int, Kc: int, activation=None, cell_fn=stc_cell) -> Tensor:
    """Encoder + decoder recurrent stack = STCGNN.forward minus MGP_Gen (:188) and out_proj (:206).
    X_seq: [B,T,N,C,Din]."""
    _, last = encoder_rollout(cell_fn, Gs, Gc, X_seq, enc, Ks, Kc, activation)
    return decoder_rollout(cell_fn, Gs, Gc, last, dec, horizon, Ks, Kc, activation)


# --------------------------------------------------------------------------- #
# synthetic supports used by tests and bench
# --------------------------------------------------------------------------- #
def grid_adjacency(rows: int, cols: int, dtype=torch.float64) -> Tensor:
    """Binary 8-neighbour adjacency of a rows x cols grid, zero diagonal (the shape of the shipped
    ``s_adj``: 10x10 -> 684 non-zeros, row sums 3/5/8)."""
    N = rows * cols
    A = torch.zeros(N, N, dtype=dtype)
    for i in range(rows):
        for j in range(cols):
            for di in (-1, 0, 1):
                for dj in (-1, 0, 1):
                    if di == 0 and dj == 0:
                        continue
                    a, b = i + di, j + dj
                    if 0 <= a < rows and 0 <= b < cols:
                        A[i * cols + j, a * cols + b] = 1.0
    return A


def random_category_support(C: int, gen: torch.Generator, hi: float = 0.36, dtype=torch.float64) -> Tensor:
    """Symmetric, zero-diagonal, U(0, hi) -- the shape of the shipped ``c_cor``."""
    A = torch.rand(C, C, generator=gen, dtype=torch.float64) * hi
    A = torch.triu(A, 1)
    return (A + A.t()).to(dtype)
