"""Convolutions on the hand-written tcgen05 kernels (csrc/cuda/gemm_tcgen05.cu).

Every convolution whose channel counts fit the 128-byte TMA rows runs on the repo's own sm_100a code, forward and backward:

* 1x1 / stride 1: in NHWC the activation [N,H,W,Cin] IS the row-major matrix [N*H*W, Cin], the filter [Cout,Cin,1,1] is
  [Cout,Cin] K-major, and the output matrix [N*H*W, Cout] is the NHWC output - three plain GEMMs, no im2col, no layout
  change, no workspace:
      fprop  Y[M,Cout]    = X[M,Cin]  * W[Cout,Cin]^T
      dgrad  dX[M,Cin]    = dY[M,Cout] * W[Cout,Cin]       (W read MN-major by TMA)
      wgrad  dW[Cout,Cin] = dY[M,Cout]^T * X[M,Cin]        (both operands MN-major, split-K)
* everything else (KxK, and strided 1x1 projection shortcuts): implicit GEMM - the producer warp asks the TMA unit for
  "128 consecutive output pixels, filter tap (r, s), channels c0..c0+63" (im2col-mode tensor map) and receives the operand
  tile with the halo zero-filled; no im2col buffer, no gather kernel.  fprop and wgrad take any stride; dgrad of a strided
  convolution is decomposed by output parity (stride^2 implicit GEMMs over the taps that reach each class, stored to the
  strided positions of dX).

Reference: M1 convolution via cuDNN / im2col+cuBLAS (MX/src/operator/nn/convolution.cu:90-160,
nn/cudnn/cudnn_convolution-inl.h:638-793, nn/im2col.cuh:79-132).

The backends stay selectable because cuDNN's sm_100 kernels are the yardstick: `set_conv1x1_backend(...)` /
`set_convkxk_backend(...)` or DLCFN_CONV1X1 / DLCFN_CONVKXK = `tcgen05` (default) | `cudnn`; tools/conv_probe.py measures
both per layer shape and per op.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import native

_BACKEND = os.environ.get("DLCFN_CONV1X1", "tcgen05").lower()


def set_conv1x1_backend(name: str) -> None:
    global _BACKEND
    if name not in ("tcgen05", "cudnn", "auto"):
        raise ValueError("backend must be 'tcgen05', 'cudnn' or 'auto'")
    _BACKEND = name


def conv1x1_backend() -> str:
    return _BACKEND


class ConvOps:
    """The three kernels of one convolution (square stride, symmetric padding, no bias) on the dlcfn tcgen05 path.
    `x` [N,C,H,W] and `weight` [K,C,R,S] are channels_last bf16 CUDA tensors."""

    def __init__(self, x: torch.Tensor, weight: torch.Tensor, stride: int = 1, pad: int = 0):
        self.x, self.w, self.stride, self.pad = x, weight, int(stride), int(pad)
        K, C, R, S = weight.shape
        self.gemm = (R == 1 and S == 1 and self.stride == 1 and self.pad == 0)
        N, _, H, W = x.shape
        self.P, self.Q = (H + 2 * self.pad - R) // self.stride + 1, (W + 2 * self.pad - S) // self.stride + 1

    @staticmethod
    def eligible(x: torch.Tensor, weight: torch.Tensor, stride, padding, dilation=(1, 1), groups: int = 1) -> bool:
        if not (x.is_cuda and x.dtype == torch.bfloat16 and weight.dtype == torch.bfloat16 and x.dim() == 4):
            return False
        K, C, R, S = weight.shape
        if not (groups == 1 and tuple(dilation) == (1, 1) and stride[0] == stride[1] and padding[0] == padding[1] and x.shape[1] == C):
            return False
        if R == 1 and S == 1 and stride[0] == 1 and padding[0] == 0:
            return C % 8 == 0 and K % 8 == 0
        # implicit GEMM: the im2col'd tensors (x for fprop / wgrad, dy for dgrad) need whole 64-channel TMA rows
        return C % 64 == 0 and K % 64 == 0 and R <= 7 and S <= 7

    def _mat(self, t: torch.Tensor) -> torch.Tensor:
        n, c, h, w = t.shape
        return t.permute(0, 2, 3, 1).reshape(n * h * w, c)          # a view: NHWC memory is already this matrix

    def fprop(self, stats: torch.Tensor | None = None) -> torch.Tensor:
        """`stats`: [rows, 2, K] buffer (native.new_stats) that receives the per-channel partial sums / sums of squares of the
        bf16 output, accumulated in the kernel's epilogue (fused BatchNorm statistics)."""
        x, w = self.x, self.w
        if self.gemm:
            N, C, H, W = x.shape
            y = torch.empty((N, w.shape[0], H, W), dtype=torch.bfloat16, device=x.device, memory_format=torch.channels_last)
            native.gemm_bf16(self._mat(x), w.reshape(w.shape[0], C), out=self._mat(y), stats=stats)
            return y
        return native.conv_fprop(x, w, self.stride, self.pad, stats=stats)

    @staticmethod
    def stats_ok(weight: torch.Tensor) -> bool:
        return weight.shape[0] >= 64 and weight.shape[0] % 8 == 0          # the bf16 TMA-store epilogue form

    def dgrad(self, dy: torch.Tensor) -> torch.Tensor:
        x, w = self.x, self.w
        if self.gemm:
            dx = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device, memory_format=torch.channels_last)
            native.gemm_bf16(self._mat(dy), w.reshape(w.shape[0], w.shape[1]), b_mn=True, out=self._mat(dx))
            return dx
        return native.conv_dgrad(dy, w, self.pad, (x.shape[2], x.shape[3]), stride=self.stride)

    def wgrad(self, dy: torch.Tensor) -> torch.Tensor:
        x, w = self.x, self.w
        if self.gemm:
            dw = torch.empty((w.shape[0], w.shape[1]), dtype=torch.bfloat16, device=x.device)
            native.gemm_bf16(self._mat(dy), self._mat(x), a_mn=True, b_mn=True, out=dw)
            return dw.view(w.shape)           # [K,C,1,1]: contiguous and channels_last at once
        return native.conv_wgrad(x, dy, (w.shape[2], w.shape[3]), self.stride, self.pad)


class _ConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, stride, pad, want_stats=False):
        if not x.is_contiguous(memory_format=torch.channels_last):
            x = x.contiguous(memory_format=torch.channels_last)
        w = weight if weight.is_contiguous(memory_format=torch.channels_last) else weight.contiguous(memory_format=torch.channels_last)
        ctx.save_for_backward(x, w)
        ctx.stride, ctx.pad = stride, pad
        stats = native.new_stats(w.shape[0], x.device) if want_stats and ConvOps.stats_ok(w) else None
        y = ConvOps(x, w, stride, pad).fprop(stats)
        if not want_stats:
            return y
        if stats is not None:
            ctx.mark_non_differentiable(stats)
        return y, stats

    @staticmethod
    def backward(ctx, dy, _dstats=None):
        x, w = ctx.saved_tensors
        if not dy.is_contiguous(memory_format=torch.channels_last):
            dy = dy.contiguous(memory_format=torch.channels_last)
        ops = ConvOps(x, w, ctx.stride, ctx.pad)
        dx = ops.dgrad(dy) if ctx.needs_input_grad[0] else None
        dw = ops.wgrad(dy) if ctx.needs_input_grad[1] else None
        return dx, dw, None, None, None


_FUSE_STATS = os.environ.get("DLCFN_CONV_BN_STATS", "1") != "0"


def conv2d(x: torch.Tensor, weight: torch.Tensor, stride: int = 1, padding: int = 0, stats: bool = False):
    """x [N,C,H,W] channels_last bf16, weight [K,C,R,S] -> [N,K,P,Q] channels_last, all three passes on the tcgen05 kernels.
    stats=True -> (y, partial statistics of y or None): hand the pair to the BatchNormReLU2d that follows."""
    if stats and _FUSE_STATS:
        return _ConvFn.apply(x, weight, int(stride), int(padding), True)
    y = _ConvFn.apply(x, weight, int(stride), int(padding), False)
    return (y, None) if stats else y


def _eligible(x: torch.Tensor, weight: torch.Tensor, stride: int = 1) -> bool:
    return weight.shape[2] == 1 and weight.shape[3] == 1 and ConvOps.eligible(x, weight, (stride, stride), (0, 0))


def conv1x1(x: torch.Tensor, weight: torch.Tensor, stride: int = 1) -> torch.Tensor:
    """x [N,Cin,H,W] (channels_last), weight [Cout,Cin,1,1] -> [N,Cout,ceil(H/s),ceil(W/s)] channels_last."""
    if not _eligible(x, weight, stride):
        if x.is_cuda and not native.lib_path().exists():
            raise native.NativeUnavailable("CUDA conv1x1 requested but libdlcfn_cuda.so is missing")
        return F.conv2d(x, weight.to(x.dtype), None, stride)
    return conv2d(x, weight, stride, 0)


class Conv1x1(nn.Conv2d):
    """nn.Conv2d(kernel 1, no padding, no bias) whose CUDA/bf16 path can run on the dlcfn tcgen05 GEMM.  Parameter name
    and shape ([Cout,Cin,1,1]) are those of nn.Conv2d, so checkpoints and the flat parameter buffers do not change."""

    def __init__(self, in_channels: int, out_channels: int, stride: int = 1, bias: bool = False):
        if bias:
            raise ValueError("Conv1x1 has no bias (every 1x1 convolution of the reference symbols is followed by BatchNorm)")
        super().__init__(in_channels, out_channels, 1, stride, 0, bias=False)

    def forward(self, x, stats: bool = False):
        """stats=True returns (y, statistics-or-None) for the BatchNorm that follows (fused into the epilogue)."""
        use = _BACKEND == "tcgen05" or (_BACKEND == "auto" and self.stride[0] > 1)
        if use and _eligible(x, self.weight, self.stride[0]):
            return conv2d(x, self.weight, self.stride[0], 0, stats=stats)
        y = super().forward(x)
        return (y, None) if stats else y

    def extra_repr(self):
        return super().extra_repr() + f", backend={_BACKEND}"


# ------------------------------------------------------------------------------ KxK implicit GEMM
_KXK_BACKEND = os.environ.get("DLCFN_CONVKXK", "tcgen05").lower()


def set_convkxk_backend(name: str) -> None:
    global _KXK_BACKEND
    if name not in ("tcgen05", "cudnn"):
        raise ValueError("backend must be 'tcgen05' or 'cudnn'")
    _KXK_BACKEND = name


def convkxk_backend() -> str:
    return _KXK_BACKEND


def _kxk_eligible(x: torch.Tensor, weight: torch.Tensor, stride, padding, dilation, groups) -> bool:
    return ConvOps.eligible(x, weight, stride, padding, dilation, groups)


def conv_kxk(x: torch.Tensor, weight: torch.Tensor, stride: int = 1, padding: int = 0) -> torch.Tensor:
    """x [N,C,H,W] channels_last bf16, weight [K,C,R,S] -> [N,K,P,Q] channels_last on the tcgen05 implicit GEMM."""
    return conv2d(x, weight, stride, padding)


class ConvKxK(nn.Conv2d):
    """nn.Conv2d (no bias) whose CUDA / bf16 / NHWC path can run on the dlcfn implicit-GEMM kernel; same parameter names
    and shapes as nn.Conv2d, so checkpoints and flat parameter buffers do not change with the backend."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size: int, stride: int = 1, padding: int = 0, bias: bool = False):
        if bias:
            raise ValueError("ConvKxK has no bias (the convolutions of the reference symbols are followed by BatchNorm)")
        super().__init__(in_channels, out_channels, kernel_size, stride, padding, bias=False)

    def forward(self, x, stats: bool = False):
        if _KXK_BACKEND == "tcgen05" and _kxk_eligible(x, self.weight, self.stride, self.padding, self.dilation, self.groups):
            return conv2d(x, self.weight, self.stride[0], self.padding[0], stats=stats)
        y = super().forward(x)
        return (y, None) if stats else y

    def extra_repr(self):
        return super().extra_repr() + f", backend={_KXK_BACKEND}"
