"""Convolutions on the hand-written tcgen05 GEMM (csrc/cuda/gemm_tcgen05.cu): 1x1 as a plain GEMM (validated, below) and
KxK as an implicit GEMM through im2col-mode TMA (`ConvKxK`, EXPERIMENTAL, at the end of this file).

In NHWC a 1x1 convolution IS a GEMM: the activation [N,H,W,Cin] is the row-major matrix [N*H*W, Cin] as it sits in
memory, the filter [Cout,Cin,1,1] is [Cout,Cin] K-major, and the output matrix [N*H*W, Cout] is the NHWC output.
No im2col, no layout change, no workspace:
    fprop  Y[M,Cout]  = X[M,Cin]  * W[Cout,Cin]^T
    dgrad  dX[M,Cin]  = dY[M,Cout] * W[Cout,Cin]       (W read MN-major by TMA)
    wgrad  dW[Cout,Cin] = dY[M,Cout]^T * X[M,Cin]      (both operands MN-major)
A stride-s 1x1 convolution (ResNet projection shortcuts) first gathers the s-strided pixels with one NHWC kernel
(ops.nn.subsample), then runs the same three GEMMs on the smaller matrix.

Reference: M1 convolution via cuDNN / im2col+cuBLAS (MX/src/operator/nn/convolution.cu:90-160,
nn/cudnn/cudnn_convolution-inl.h:638-793, nn/im2col.cuh:79-132).  2/3 of ResNet-50's convolution layers are 1x1.

The backend is selectable because cuDNN's sm_100 kernels are the yardstick: `set_conv1x1_backend(...)` or
DLCFN_CONV1X1 = `tcgen05` (every 1x1 convolution on the dlcfn GEMM), `cudnn` (none; the measured default: 25.36 ms
per ResNet-50 step vs 25.61 `auto` vs 26.46 `tcgen05`), or `auto` (the strided projection shortcuts only: gather + GEMM +
scatter); tools/conv1x1_probe.py measures both per layer shape.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import native
from .linear import _LinearFn
from .nn import subsample

_BACKEND = os.environ.get("DLCFN_CONV1X1", "cudnn").lower()


def set_conv1x1_backend(name: str) -> None:
    global _BACKEND
    if name not in ("tcgen05", "cudnn", "auto"):
        raise ValueError("backend must be 'tcgen05', 'cudnn' or 'auto'")
    _BACKEND = name


def conv1x1_backend() -> str:
    return _BACKEND


def _eligible(x: torch.Tensor, weight: torch.Tensor) -> bool:
    return (x.is_cuda and x.dtype == torch.bfloat16 and weight.dtype == torch.bfloat16 and x.dim() == 4 and
            weight.shape[2] == 1 and weight.shape[3] == 1 and weight.shape[0] % 8 == 0 and weight.shape[1] % 8 == 0)


def conv1x1(x: torch.Tensor, weight: torch.Tensor, stride: int = 1) -> torch.Tensor:
    """x [N,Cin,H,W] (channels_last), weight [Cout,Cin,1,1] -> [N,Cout,ceil(H/s),ceil(W/s)] channels_last."""
    if not _eligible(x, weight):
        if x.is_cuda and not native.lib_path().exists():
            raise native.NativeUnavailable("CUDA conv1x1 requested but libdlcfn_cuda.so is missing")
        return F.conv2d(x, weight.to(x.dtype), None, stride)
    if not x.is_contiguous(memory_format=torch.channels_last):
        x = x.contiguous(memory_format=torch.channels_last)
    x = subsample(x, stride)
    N, C, H, W = x.shape
    O = weight.shape[0]
    x2 = x.permute(0, 2, 3, 1).reshape(N * H * W, C)          # a view: NHWC memory is already this matrix
    y2 = _LinearFn.apply(x2, weight.reshape(O, C), None)       # [N*H*W, O]
    return y2.view(N, H, W, O).permute(0, 3, 1, 2)            # channels_last view of the GEMM output


class Conv1x1(nn.Conv2d):
    """nn.Conv2d(kernel 1, no padding, no bias) whose CUDA/bf16 path can run on the dlcfn tcgen05 GEMM.  Parameter name
    and shape ([Cout,Cin,1,1]) are those of nn.Conv2d, so checkpoints and the flat parameter buffers do not change."""

    def __init__(self, in_channels: int, out_channels: int, stride: int = 1, bias: bool = False):
        if bias:
            raise ValueError("Conv1x1 has no bias (every 1x1 convolution of the reference symbols is followed by BatchNorm)")
        super().__init__(in_channels, out_channels, 1, stride, 0, bias=False)

    def forward(self, x):
        use = _BACKEND == "tcgen05" or (_BACKEND == "auto" and self.stride[0] > 1)
        if use and _eligible(x, self.weight):
            return conv1x1(x, self.weight, self.stride[0])
        return super().forward(x)

    def extra_repr(self):
        return super().extra_repr() + f", backend={_BACKEND}"


# ------------------------------------------------------------------------------ KxK implicit GEMM (EXPERIMENTAL)
# Written at the end of round 1 with no GPU left to run it on: compiled (SASS shows UTMALDG.4D.IM2COL), off by default,
# validated by tests/test_gpu_experimental.py under DLCFN_EXPERIMENTAL=1.  `DLCFN_CONVKXK=tcgen05` / set_convkxk_backend
# switch ConvKxK modules over; cuDNN stays the default and the yardstick.
#
# The activation tensor never goes through an im2col buffer: the producer warp of the GEMM kernel asks the TMA unit for
# "128 consecutive output pixels, filter tap (r, s), channels c0..c0+63" and receives the [128 x 64] operand tile with the
# halo already zero-filled (cuTensorMapEncodeIm2col).  fprop, the stride-1 dgrad (same kernel on dY with the taps reversed,
# weights read MN-major) and wgrad (split-K over the N*P*Q pixels) all run on it; a strided dgrad falls back to cuDNN.
_KXK_BACKEND = os.environ.get("DLCFN_CONVKXK", "cudnn").lower()


def set_convkxk_backend(name: str) -> None:
    global _KXK_BACKEND
    if name not in ("tcgen05", "cudnn"):
        raise ValueError("backend must be 'tcgen05' or 'cudnn'")
    _KXK_BACKEND = name


def convkxk_backend() -> str:
    return _KXK_BACKEND


def _kxk_eligible(x: torch.Tensor, weight: torch.Tensor, stride, padding, dilation, groups) -> bool:
    if not (x.is_cuda and x.dtype == torch.bfloat16 and weight.dtype == torch.bfloat16 and x.dim() == 4):
        return False
    K, C, R, S = weight.shape
    return (groups == 1 and tuple(dilation) == (1, 1) and stride[0] == stride[1] and padding[0] == padding[1] and
            C % 64 == 0 and K % 8 == 0 and R * S <= 49 and x.shape[1] == C)


class _ConvKxKFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, stride, pad):
        if not x.is_contiguous(memory_format=torch.channels_last):
            x = x.contiguous(memory_format=torch.channels_last)
        w = weight if weight.is_contiguous(memory_format=torch.channels_last) else weight.contiguous(memory_format=torch.channels_last)
        ctx.save_for_backward(x, w)
        ctx.stride, ctx.pad = stride, pad
        return native.conv_fprop(x, w, stride, pad)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        if not dy.is_contiguous(memory_format=torch.channels_last):
            dy = dy.contiguous(memory_format=torch.channels_last)
        dx = dw = None
        if ctx.needs_input_grad[0]:
            if ctx.stride == 1 and w.shape[0] % 64 == 0:
                dx = native.conv_dgrad(dy, w, ctx.pad, (x.shape[2], x.shape[3]))
            else:       # strided dgrad (zero-interleaved gradient) or Cout not a multiple of 64: cuDNN for now
                dx = torch.ops.aten.convolution_backward(dy, x, w, None, [ctx.stride] * 2, [ctx.pad] * 2, [1, 1], False, [0, 0],
                                                         1, [True, False, False])[0]
        if ctx.needs_input_grad[1]:
            dw = native.conv_wgrad(x, dy, (w.shape[2], w.shape[3]), ctx.stride, ctx.pad)
        return dx, dw, None, None


def conv_kxk(x: torch.Tensor, weight: torch.Tensor, stride: int = 1, padding: int = 0) -> torch.Tensor:
    """x [N,C,H,W] channels_last bf16, weight [K,C,R,S] -> [N,K,P,Q] channels_last on the tcgen05 implicit GEMM."""
    return _ConvKxKFn.apply(x, weight, int(stride), int(padding))


class ConvKxK(nn.Conv2d):
    """nn.Conv2d (no bias) whose CUDA / bf16 / NHWC path can run on the dlcfn implicit-GEMM kernel; same parameter names
    and shapes as nn.Conv2d, so checkpoints and flat parameter buffers do not change with the backend."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size: int, stride: int = 1, padding: int = 0, bias: bool = False):
        if bias:
            raise ValueError("ConvKxK has no bias (the convolutions of the reference symbols are followed by BatchNorm)")
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, bias=False)

    def forward(self, x):
        if _KXK_BACKEND == "tcgen05" and _kxk_eligible(x, self.weight, self.stride, self.padding, self.dilation, self.groups):
            return conv_kxk(x, self.weight, self.stride[0], self.padding[0])
        return super().forward(x)

    def extra_repr(self):
        return super().extra_repr() + f", backend={_KXK_BACKEND}"
