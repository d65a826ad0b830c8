"""Fused BatchNorm(+ReLU) on the hand-written NHWC bf16 kernels (csrc/cuda/batchnorm.cu).

Reference M3/M4: BatchNorm (cuDNN or one-block-per-channel hand kernels, nn/batch_norm.cu:208-572) and a
separate ReLU Activation.  Profiling the ResNet-50 step on B200 showed the framework's channels-last
batch-norm + ReLU kernels taking >55 % of the step while moving data at a small fraction of HBM
bandwidth; these kernels do the same math in 3 (forward) + 5 (backward) streaming passes.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import native

_MAX_C = 2048
_DEBUG_LAYOUT = __import__("os").environ.get("DLCFN_DEBUG_LAYOUT", "0") == "1"


def _declare():
    L = native.lib()
    if getattr(L, "_bn_declared", False):
        return L
    import ctypes
    from ctypes import c_float, c_int, c_longlong, c_size_t, c_void_p
    L.dlcfn_bn_workspace_floats.argtypes = [c_longlong, c_int]
    L.dlcfn_bn_workspace_floats.restype = c_size_t
    L.dlcfn_bn_forward.argtypes = [c_void_p, c_void_p, c_longlong, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_float, c_float, c_int, c_void_p, c_void_p]
    L.dlcfn_bn_forward.restype = c_int
    L.dlcfn_bn_inference.argtypes = [c_void_p, c_void_p, c_longlong, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_float, c_int, c_void_p, c_void_p]
    L.dlcfn_bn_inference.restype = c_int
    L.dlcfn_bn_backward.argtypes = [c_void_p, c_void_p, c_void_p, c_longlong, c_int, c_void_p, c_void_p, c_void_p,
                                    c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]
    L.dlcfn_bn_backward.restype = c_int
    L.dlcfn_bn_forward_ex.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_longlong, c_int, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_float, c_int, c_void_p, c_void_p]
    L.dlcfn_bn_forward_ex.restype = c_int
    L.dlcfn_bn_backward_ex.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_longlong, c_int, c_void_p, c_void_p,
                                       c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]
    L.dlcfn_bn_backward_ex.restype = c_int
    L.dlcfn_bn_forward_pre.argtypes = [c_void_p, c_void_p, c_longlong, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_void_p, c_float, c_float, c_int, c_void_p, c_int, c_void_p, c_void_p]
    L.dlcfn_bn_forward_pre.restype = c_int
    L._bn_declared = True
    return L


def fused_supported(x: torch.Tensor, C: int) -> bool:
    return x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 4 and C % 8 == 0 and 8 <= C <= _MAX_C


layout_copies: list = []      # (shape, strides) of tensors that had to be re-laid-out (DLCFN_DEBUG_LAYOUT=1)


def _nhwc(x: torch.Tensor) -> torch.Tensor:
    if x.is_contiguous(memory_format=torch.channels_last):
        return x
    if _DEBUG_LAYOUT:
        layout_copies.append((tuple(x.shape), tuple(x.stride())))
    return x.contiguous(memory_format=torch.channels_last)


class _BNReLUFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, momentum, eps, relu, stats=None):
        """`stats` (optional): partial sums [rows, 2, C] of x written by the epilogue of the convolution that produced x
        (ops/conv.py, fused statistics): the statistics pass over x is skipped."""
        L = _declare()
        x = _nhwc(x)
        N, C, H, W = x.shape
        M = N * H * W
        y = torch.empty_like(x, memory_format=torch.channels_last)
        mean = torch.empty(C, dtype=torch.float32, device=x.device)
        invstd = torch.empty(C, dtype=torch.float32, device=x.device)
        ws = torch.empty(L.dlcfn_bn_workspace_floats(M, C), dtype=torch.float32, device=x.device)
        if stats is not None:
            assert stats.dtype == torch.float32 and stats.dim() == 3 and stats.shape[1] == 2 and stats.shape[2] == C
            native.check(L.dlcfn_bn_forward_pre(x.data_ptr(), y.data_ptr(), M, C, native._ptr(gamma), native._ptr(beta),
                                                mean.data_ptr(), invstd.data_ptr(), native._ptr(running_mean),
                                                native._ptr(running_var), momentum, eps, int(relu), stats.data_ptr(),
                                                stats.shape[0], ws.data_ptr(), native.stream_ptr()), "bn_forward_pre")
            native._count(2)
        else:
            native.check(L.dlcfn_bn_forward(x.data_ptr(), y.data_ptr(), M, C, native._ptr(gamma), native._ptr(beta),
                                            mean.data_ptr(), invstd.data_ptr(), native._ptr(running_mean),
                                            native._ptr(running_var), momentum, eps, int(relu), ws.data_ptr(),
                                            native.stream_ptr()), "bn_forward")
            native._count(3)
        ctx.save_for_backward(x, gamma, beta, mean, invstd)
        ctx.relu = relu
        return y

    @staticmethod
    def backward(ctx, dy):
        L = _declare()
        x, gamma, beta, mean, invstd = ctx.saved_tensors
        dy = _nhwc(dy)
        N, C, H, W = x.shape
        M = N * H * W
        dx = torch.empty_like(x, memory_format=torch.channels_last)
        dgamma = torch.empty(C, dtype=torch.float32, device=x.device) if gamma is not None else None
        dbeta = torch.empty(C, dtype=torch.float32, device=x.device) if beta is not None else None
        ws = torch.empty(L.dlcfn_bn_workspace_floats(M, C), dtype=torch.float32, device=x.device)
        native.check(L.dlcfn_bn_backward(x.data_ptr(), dy.data_ptr(), dx.data_ptr(), M, C, native._ptr(gamma),
                                         native._ptr(beta), mean.data_ptr(), invstd.data_ptr(), native._ptr(dgamma),
                                         native._ptr(dbeta), 0, int(ctx.relu), ws.data_ptr(), native.stream_ptr()),
                     "bn_backward")
        native._count(3)
        return dx, dgamma, dbeta, None, None, None, None, None, None


class _AddBNReLUFn(torch.autograd.Function):
    """(z, y) = (a + b, relu?(bn(a + b))) with the add fused into the statistics pass; backward fuses the gradient
    that reaches z through the residual connection into the BN backward apply pass."""

    @staticmethod
    def forward(ctx, a, b, gamma, beta, running_mean, running_var, momentum, eps, relu):
        L = _declare()
        a, b = _nhwc(a), _nhwc(b)
        N, C, H, W = a.shape
        M = N * H * W
        z = torch.empty_like(a, memory_format=torch.channels_last)
        y = torch.empty_like(a, memory_format=torch.channels_last)
        mean = torch.empty(C, dtype=torch.float32, device=a.device)
        invstd = torch.empty(C, dtype=torch.float32, device=a.device)
        ws = torch.empty(L.dlcfn_bn_workspace_floats(M, C), dtype=torch.float32, device=a.device)
        native.check(L.dlcfn_bn_forward_ex(a.data_ptr(), b.data_ptr(), z.data_ptr(), y.data_ptr(), M, C,
                                           native._ptr(gamma), native._ptr(beta), mean.data_ptr(), invstd.data_ptr(),
                                           native._ptr(running_mean), native._ptr(running_var), momentum, eps, int(relu),
                                           ws.data_ptr(), native.stream_ptr()), "bn_forward_ex")
        native._count(3)
        ctx.save_for_backward(z, gamma, beta, mean, invstd)
        ctx.relu = relu
        # z is often unused downstream (first unit of a stage: the shortcut conv takes y): without this autograd would
        # hand backward a materialised NCHW zeros tensor for dz, which then costs a fill, a transposing layout copy and
        # an extra read in the apply kernel - 1.5 ms per ResNet-50 step at batch 256 (tools/copy_diag2.py)
        ctx.set_materialize_grads(False)
        return z, y

    @staticmethod
    def backward(ctx, dz, dy):
        L = _declare()
        z, gamma, beta, mean, invstd = ctx.saved_tensors
        N, C, H, W = z.shape
        M = N * H * W
        if dy is None:           # y unused: only the residual gradient flows
            return dz, dz, None, None, None, None, None, None, None
        dy = _nhwc(dy)
        extra = _nhwc(dz) if dz is not None else None
        dx = torch.empty_like(z, memory_format=torch.channels_last)
        dgamma = torch.empty(C, dtype=torch.float32, device=z.device) if gamma is not None else None
        dbeta = torch.empty(C, dtype=torch.float32, device=z.device) if beta is not None else None
        ws = torch.empty(L.dlcfn_bn_workspace_floats(M, C), dtype=torch.float32, device=z.device)
        native.check(L.dlcfn_bn_backward_ex(z.data_ptr(), dy.data_ptr(), native._ptr(extra), dx.data_ptr(), M, C,
                                            native._ptr(gamma), native._ptr(beta), mean.data_ptr(), invstd.data_ptr(),
                                            native._ptr(dgamma), native._ptr(dbeta), 0, int(ctx.relu), ws.data_ptr(),
                                            native.stream_ptr()), "bn_backward_ex")
        native._count(3)
        return dx, dx, dgamma, dbeta, None, None, None, None, None


def add_batch_norm_relu(a, b, gamma, beta, running_mean, running_var, training: bool, momentum: float, eps: float,
                        relu: bool = True):
    """Returns (z, y) with z = a + b and y = relu?(batch_norm(z))."""
    C = a.shape[1]
    if training and fused_supported(a, C) and b.dtype == a.dtype and b.shape == a.shape and \
            (gamma is None or gamma.dtype == torch.float32):
        return _AddBNReLUFn.apply(a, b, gamma, beta, running_mean, running_var, momentum, eps, relu)
    z = a + b
    return z, batch_norm_relu(z, gamma, beta, running_mean, running_var, training, momentum, eps, relu)


def batch_norm_relu(x, gamma, beta, running_mean, running_var, training: bool, momentum: float, eps: float,
                    relu: bool = True, stats=None):
    C = x.shape[1]
    if fused_supported(x, C) and (gamma is None or gamma.dtype == torch.float32):
        if training:
            return _BNReLUFn.apply(x, gamma, beta, running_mean, running_var, momentum, eps, relu, stats)
        L = _declare()
        x = _nhwc(x)
        N, _, H, W = x.shape
        y = torch.empty_like(x, memory_format=torch.channels_last)
        ws = torch.empty(L.dlcfn_bn_workspace_floats(N * H * W, C), dtype=torch.float32, device=x.device)
        native.check(L.dlcfn_bn_inference(x.data_ptr(), y.data_ptr(), N * H * W, C, native._ptr(gamma),
                                          native._ptr(beta), running_mean.data_ptr(), running_var.data_ptr(), eps,
                                          int(relu), ws.data_ptr(), native.stream_ptr()), "bn_inference")
        native._count(2)
        return y
    if x.is_cuda and not native.lib_path().exists():
        raise native.NativeUnavailable("libdlcfn_cuda.so missing")
    y = F.batch_norm(x, running_mean, running_var, gamma, beta, training, momentum, eps)
    return F.relu(y) if relu else y


class BatchNormReLU2d(nn.BatchNorm2d):
    """nn.BatchNorm2d (same parameters / buffers / state_dict keys) with an optional fused ReLU."""

    def __init__(self, num_features: int, eps: float = 1e-5, momentum: float = 0.1, relu: bool = True):
        super().__init__(num_features, eps=eps, momentum=momentum)
        self.relu = relu

    def forward(self, x, stats=None):
        """`x` may be the (y, stats) pair of a convolution called with `stats=True` (ops/conv.py): the statistics of y were
        accumulated in that convolution's epilogue."""
        if isinstance(x, tuple):
            x, stats = x
        return batch_norm_relu(x, self.weight, self.bias, self.running_mean, self.running_var, self.training,
                               self.momentum, self.eps, self.relu, stats)

    def forward_add(self, a, b):
        """(a + b, bn_relu(a + b)) - the residual add fused into this layer's statistics pass."""
        return add_batch_norm_relu(a, b, self.weight, self.bias, self.running_mean, self.running_var, self.training,
                                   self.momentum, self.eps, self.relu)

    def extra_repr(self):
        return super().extra_repr() + f", relu={self.relu}, kernel=dlcfn-nhwc"
