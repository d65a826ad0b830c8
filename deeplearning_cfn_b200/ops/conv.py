"""1x1 convolutions on the hand-written tcgen05 GEMM (csrc/cuda/gemm_tcgen05.cu).

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
