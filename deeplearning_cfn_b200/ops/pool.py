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
    L.dlcfn_stem_s2d_fwd.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]
    L.dlcfn_stem_s2d_fwd.restype = c_int
    L.dlcfn_stem_s2d_bwd.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]
    L.dlcfn_stem_s2d_bwd.restype = c_int
    from ctypes import c_longlong
    L.dlcfn_pad3to8_nhwc.argtypes = [c_void_p, c_void_p, c_longlong, c_void_p]
    L.dlcfn_pad3to8_nhwc.restype = c_int
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


class _StemS2DFn(torch.autograd.Function):
    """[N,8,H,W] channels_last (3 real channels) -> space-to-depth [N,16,(H+6)/2,(W+6)/2] channels_last with the 7x7/pad-3
    border folded in (csrc/cuda/pool.cu:kstem_s2d_*)."""

    @staticmethod
    def forward(ctx, x8):
        L = _declare()
        if not x8.is_contiguous(memory_format=torch.channels_last):
            x8 = x8.contiguous(memory_format=torch.channels_last)
        N, C, H, W = x8.shape
        y = torch.empty((N, 16, (H + 6) // 2, (W + 6) // 2), dtype=x8.dtype, device=x8.device, memory_format=torch.channels_last)
        native.check(L.dlcfn_stem_s2d_fwd(x8.data_ptr(), y.data_ptr(), N, H, W, native.stream_ptr()), "stem_s2d_fwd")
        native._count()
        ctx.shape = (N, C, H, W)
        return y

    @staticmethod
    def backward(ctx, dy):
        L = _declare()
        N, C, H, W = ctx.shape
        if not dy.is_contiguous(memory_format=torch.channels_last):
            dy = dy.contiguous(memory_format=torch.channels_last)
        dx = torch.empty((N, 8, H, W), dtype=dy.dtype, device=dy.device, memory_format=torch.channels_last)
        native.check(L.dlcfn_stem_s2d_bwd(dy.data_ptr(), dx.data_ptr(), N, H, W, native.stream_ptr()), "stem_s2d_bwd")
        native._count()
        return dx


def stem_space_to_depth(x8: torch.Tensor) -> torch.Tensor:
    return _StemS2DFn.apply(x8)


def pad3to8(x: torch.Tensor) -> torch.Tensor:
    """[N,3,H,W] channels_last bf16 (no grad) -> [N,8,H,W] channels_last with zero channels 3..7, one kernel."""
    if x.requires_grad or not (x.is_cuda and x.dtype == torch.bfloat16 and x.shape[1] == 3 and
                                x.is_contiguous(memory_format=torch.channels_last)):
        return F.pad(x, (0, 0, 0, 0, 0, 5)).contiguous(memory_format=torch.channels_last)
    L = _declare()
    N, _, H, W = x.shape
    y = torch.empty((N, 8, H, W), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
    native.check(L.dlcfn_pad3to8_nhwc(x.data_ptr(), y.data_ptr(), N * H * W, native.stream_ptr()), "pad3to8")
    native._count()
    return y


class _GlobalAvgPoolFn(torch.autograd.Function):
    """[N,C,H,W] -> [N,C] mean over H,W.  The point is the backward: torch's adaptive_avg_pool2d hands back an NCHW
    gradient, which the NHWC BatchNorm backward then has to transpose; here the broadcast gradient is written directly
    in channels_last."""

    @staticmethod
    def forward(ctx, x):
        ctx.shape = x.shape
        return x.mean(dim=(2, 3))

    @staticmethod
    def backward(ctx, dy):
        N, C, H, W = ctx.shape
        g = (dy.float() / float(H * W)).to(dy.dtype)
        return g.view(N, 1, 1, C).expand(N, H, W, C).contiguous().permute(0, 3, 1, 2)


def global_avg_pool(x: torch.Tensor) -> torch.Tensor:
    """Global average pooling to [N, C] with a channels_last gradient."""
    return _GlobalAvgPoolFn.apply(x)
