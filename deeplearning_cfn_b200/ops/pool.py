"""NHWC bf16 max-pooling on the dlcfn kernels (csrc/cuda/pool.cu); anything else goes through torch."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import native


def _declare():
    L = native.lib()
    if getattr(L, "_pool_declared", False):
        return L
    from ctypes import c_int, c_void_p
    L.dlcfn_maxpool_nhwc_fwd.argtypes = [c_void_p, c_void_p, c_void_p] + [c_int] * 7 + [c_void_p]
    L.dlcfn_maxpool_nhwc_fwd.restype = c_int
    L.dlcfn_maxpool_nhwc_bwd.argtypes = [c_void_p, c_void_p, c_void_p] + [c_int] * 7 + [c_void_p]
    L.dlcfn_maxpool_nhwc_bwd.restype = c_int
    L._pool_declared = True
    return L


class _MaxPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, k, s, p):
        L = _declare()
        if not x.is_contiguous(memory_format=torch.channels_last):
            x = x.contiguous(memory_format=torch.channels_last)
        N, C, H, W = x.shape
        OH, OW = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
        y = torch.empty((N, C, OH, OW), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
        idx = torch.empty(N * OH * OW * C, dtype=torch.uint8, device=x.device)
        native.check(L.dlcfn_maxpool_nhwc_fwd(x.data_ptr(), y.data_ptr(), idx.data_ptr(), N, H, W, C, k, s, p,
                                              native.stream_ptr()), "maxpool_fwd")
        native._count()
        ctx.save_for_backward(idx)
        ctx.geom = (N, C, H, W, k, s, p)
        return y

    @staticmethod
    def backward(ctx, dy):
        L = _declare()
        (idx,) = ctx.saved_tensors
        N, C, H, W, k, s, p = ctx.geom
        if not dy.is_contiguous(memory_format=torch.channels_last):
            dy = dy.contiguous(memory_format=torch.channels_last)
        dx = torch.empty((N, C, H, W), dtype=dy.dtype, device=dy.device, memory_format=torch.channels_last)
        native.check(L.dlcfn_maxpool_nhwc_bwd(dy.data_ptr(), idx.data_ptr(), dx.data_ptr(), N, H, W, C, k, s, p,
                                              native.stream_ptr()), "maxpool_bwd")
        native._count()
        return dx, None, None, None


def max_pool2d(x: torch.Tensor, kernel_size: int, stride: int | None = None, padding: int = 0) -> torch.Tensor:
    stride = stride or kernel_size
    if x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 4 and x.shape[1] % 8 == 0 and kernel_size <= 15:
        return _MaxPoolFn.apply(x, kernel_size, stride, padding)
    return F.max_pool2d(x, kernel_size, stride, padding)


class MaxPool2d(nn.Module):
    def __init__(self, kernel_size: int, stride: int | None = None, padding: int = 0):
        super().__init__()
        self.kernel_size, self.stride, self.padding = kernel_size, stride or kernel_size, padding

    def forward(self, x):
        return max_pool2d(x, self.kernel_size, self.stride, self.padding)

    def extra_repr(self):
        return f"kernel_size={self.kernel_size}, stride={self.stride}, padding={self.padding}, kernel=dlcfn-nhwc"
