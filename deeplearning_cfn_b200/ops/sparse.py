"""Row-sparse optimizer updates on the dlcfn kernels (csrc/cuda/nn_ops.cu: ksparse_opt, kgroup_adagrad).

Reference O3: SGDMomStdDnsRspDnsKernel / AdamStdDnsRspDnsKernel / _sparse_adagrad_update (optimizer_op.cu:33,128,262)
and contrib group_adagrad (contrib/optimizer_op.cu): only the rows present in the gradient are touched ("lazy update").
"""
from __future__ import annotations

from ctypes import c_float, c_int, c_longlong, c_void_p

import torch

from . import native


def _declare():
    L = native.lib()
    if getattr(L, "_sparse_declared", False):
        return L
    L.dlcfn_sparse_row_update.argtypes = [c_int, c_void_p, c_longlong, c_longlong, c_int, c_void_p, c_int, c_void_p,
                                          c_void_p, c_void_p, c_void_p, c_int] + [c_float] * 8 + [c_void_p]
    L.dlcfn_sparse_row_update.restype = c_int
    L.dlcfn_group_adagrad_update.argtypes = [c_void_p, c_longlong, c_longlong, c_int, c_void_p, c_int, c_void_p, c_void_p,
                                             c_float, c_float, c_float, c_float, c_void_p]
    L.dlcfn_group_adagrad_update.restype = c_int
    L._sparse_declared = True
    return L


def coalesce(row_ids: torch.Tensor, grad_rows: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    """Sorted unique row ids + fp32 gradient rows summed over duplicates."""
    rows, inv = torch.unique(row_ids.reshape(-1).to(torch.int64), return_inverse=True)
    g2 = grad_rows.reshape(row_ids.numel(), -1).float()
    g = torch.zeros((rows.numel(), g2.shape[1]), dtype=torch.float32, device=g2.device).index_add_(0, inv, g2)
    return rows, g


def sparse_row_update(kind: str, rows: torch.Tensor, grad_rows: torch.Tensor, master: torch.Tensor,
                      s1: torch.Tensor | None, s2: torch.Tensor | None, wout: torch.Tensor | None, *, lr: float,
                      wd: float = 0.0, momentum: float = 0.0, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8,
                      rescale: float = 1.0, clip: float = -1.0) -> None:
    """Apply optimizer `kind` (names of native.OPT_KINDS) to rows `rows` (unique int64) of the fp32 table `master`
    [num_rows, row_len] (and its states); `wout` optionally receives the model-dtype copy of the touched rows."""
    L = _declare()
    num_rows = master.shape[0]
    row_len = master.numel() // max(1, num_rows)
    assert master.dtype == torch.float32 and master.is_contiguous() and rows.dtype == torch.int64
    g = grad_rows.contiguous()
    native.check(L.dlcfn_sparse_row_update(native.OPT_KINDS[kind], rows.data_ptr(), rows.numel(), num_rows, row_len,
                                           g.data_ptr(), native.dtype_code(g.dtype), master.data_ptr(), native._ptr(s1),
                                           native._ptr(s2), native._ptr(wout),
                                           native.dtype_code(wout.dtype) if wout is not None else 0, lr, wd, momentum,
                                           beta1, beta2, eps, rescale, clip, native.stream_ptr()), "sparse_row_update")
    native._count()


def group_adagrad_update(rows: torch.Tensor | None, grad_rows: torch.Tensor, weight: torch.Tensor, hist: torch.Tensor, *,
                         lr: float, eps: float = 1e-5, rescale: float = 1.0, clip: float = -1.0) -> None:
    """hist[row] += mean(g_row^2); weight[row] -= lr * g_row / sqrt(hist[row] + eps).  rows=None: dense (all rows)."""
    L = _declare()
    num_rows = weight.shape[0]
    row_len = weight.numel() // max(1, num_rows)
    assert weight.dtype == torch.float32 and weight.is_contiguous() and hist.dtype == torch.float32
    g = grad_rows.contiguous()
    n = rows.numel() if rows is not None else num_rows
    native.check(L.dlcfn_group_adagrad_update(native._ptr(rows), n, num_rows, row_len, g.data_ptr(),
                                              native.dtype_code(g.dtype), weight.data_ptr(), hist.data_ptr(), lr, eps,
                                              rescale, clip, native.stream_ptr()), "group_adagrad_update")
    native._count()
