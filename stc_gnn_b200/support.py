"""Spatial support handed to the cell: a dense [N,N] tensor (learned Gs, as MGP_Gen produces:
/root/reference/framework/STC_GNN.py:233) or a constant CSR graph (synthetic large-N configs)."""
from __future__ import annotations

import torch

from . import _lib


class CsrSupport:
    """Constant sparse spatial support Gs in CSR, together with Gs^T in CSR.

    The forward mode product contracts over Gs's first index (STC_GNN.py:37), so forward walks rows of
    Gs^T and backward walks rows of Gs; both are built once here.
    """

    def __init__(self, rowptr: torch.Tensor, col: torch.Tensor, vals: torch.Tensor, num_nodes: int):
        if not (rowptr.is_cuda and col.is_cuda and vals.is_cuda):
            raise RuntimeError("CsrSupport tensors must live on a CUDA device (there is no CPU path)")
        self.N = int(num_nodes)
        self.rowptr = rowptr.to(torch.int32).contiguous()
        self.col = col.to(torch.int32).contiguous()
        self.vals = vals.to(torch.float32).contiguous()
        self.nnz = int(self.vals.numel())
        t = torch.sparse_csr_tensor(self.rowptr.long(), self.col.long(), self.vals, size=(self.N, self.N))
        tt = t.to_sparse_coo().t().coalesce().to_sparse_csr()
        self.t_rowptr = tt.crow_indices().to(torch.int32).contiguous()
        self.t_col = tt.col_indices().to(torch.int32).contiguous()
        self.t_vals = tt.values().to(torch.float32).contiguous()
        self.device = self.vals.device

    @classmethod
    def from_dense(cls, G: torch.Tensor) -> "CsrSupport":
        s = G.detach().to_sparse_csr()
        return cls(s.crow_indices(), s.col_indices(), s.values(), G.shape[0])

    @classmethod
    def from_torch_sparse(cls, G: torch.Tensor) -> "CsrSupport":
        s = G.detach().to_sparse_csr()
        return cls(s.crow_indices(), s.col_indices(), s.values(), G.shape[0])

    def to_dense(self) -> torch.Tensor:
        return torch.sparse_csr_tensor(self.rowptr.long(), self.col.long(), self.vals, size=(self.N, self.N)).to_dense()

    def struct(self) -> _lib.StcSupport:
        s = _lib.StcSupport()
        s.kind = _lib.SUPPORT_CSR
        s.nnz = self.nnz
        s.vals, s.rowptr, s.col = self.vals.data_ptr(), self.rowptr.data_ptr(), self.col.data_ptr()
        s.t_vals, s.t_rowptr, s.t_col = self.t_vals.data_ptr(), self.t_rowptr.data_ptr(), self.t_col.data_ptr()
        return s


def dense_struct(G: torch.Tensor) -> _lib.StcSupport:
    s = _lib.StcSupport()
    s.kind = _lib.SUPPORT_DENSE
    s.nnz = G.numel()
    s.vals = G.data_ptr()
    return s


def support_apply(Gs, X: torch.Tensor, transpose: bool = True, alpha: float = 1.0, beta: float = 0.0,
                  Z: torch.Tensor = None, out: torch.Tensor = None, rows: torch.Tensor = None) -> torch.Tensor:
    """Y[b,m,...] = alpha * sum_n A(m,n) X[b,n,...] + beta * Z  with A = Gs^T (transpose) or Gs. CUDA only.
    `out` (contiguous, X's shape) receives the result in place of a fresh tensor.  `rows` (int32 CUDA tensor, CSR
    supports only): compute just those output nodes and leave every other row of `out` untouched."""
    lib = _lib.load()
    if not X.is_cuda or X.dtype != torch.float32:
        raise RuntimeError("support_apply needs a float32 CUDA tensor (there is no CPU path)")
    B, N = X.shape[0], X.shape[1]
    Xc = X.contiguous()
    width = Xc[0, 0].numel() if Xc.numel() else 0
    if out is not None:
        if out.shape != Xc.shape or not out.is_contiguous() or out.dtype != torch.float32 or out.device != Xc.device:
            raise RuntimeError("support_apply: `out` must be a contiguous float32 tensor of X's shape on X's device")
        Y = out
    else:
        Y = torch.empty_like(Xc)
    if isinstance(Gs, CsrSupport):
        st = Gs.struct()
        keep = Gs
    else:
        keep = Gs.detach().contiguous()
        st = dense_struct(keep)
    zc = Z.contiguous() if Z is not None else None
    if rows is not None:
        if rows.dtype != torch.int32 or not rows.is_cuda or not rows.is_contiguous():
            raise RuntimeError("support_apply: `rows` must be a contiguous int32 CUDA tensor")
        if out is None:
            raise RuntimeError("support_apply: `rows` needs `out` (the other rows keep their contents)")
        status = lib.stc_support_apply_rows(st, N, B, width, 1 if transpose else 0, Xc.data_ptr(), N * width,
                                            zc.data_ptr() if zc is not None else None, N * width, Y.data_ptr(),
                                            float(alpha), float(beta), rows.data_ptr(), rows.numel(),
                                            torch.cuda.current_stream().cuda_stream)
        _lib.check(status, "stc_support_apply_rows")
        del keep
        return Y
    status = lib.stc_support_apply(st, N, B, width, 1 if transpose else 0, Xc.data_ptr(), N * width,
                                   zc.data_ptr() if zc is not None else None, N * width, Y.data_ptr(),
                                   float(alpha), float(beta), torch.cuda.current_stream().cuda_stream)
    _lib.check(status, "stc_support_apply")
    del keep
    return Y
