"""Model-compute ops off the GEMM path, on the dlcfn kernels of csrc/cuda/nn_ops.cu.

  softmax_cross_entropy   reference M7 SoftmaxOutput (softmax-inl.h:166-266): loss forward, (p - onehot) backward
  LocalResponseNorm       reference M9 LRN (nn/lrn.cu); torch semantics (alpha / size) or TF's depth_radius form
  DepthwiseConv2d         reference M10 (nn/depthwise_convolution_tf.cuh)
  subsample               NHWC gather in front of a 1x1 / stride-s convolution evaluated as a GEMM (ops/conv.py)

CUDA bf16 (fp32 logits also for the loss) tensors run on the hand-written kernels; anything else (CPU unit tests,
other dtypes) goes through the equivalent torch expression so the modules stay usable everywhere.
"""
from __future__ import annotations

from ctypes import c_float, c_int, c_longlong, c_void_p

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import native


def _declare():
    L = native.lib()
    if getattr(L, "_nn_declared", False):
        return L
    L.dlcfn_softmax_ce_fwd.argtypes = [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_longlong, c_int, c_longlong,
                                       c_longlong, c_void_p]
    L.dlcfn_softmax_ce_bwd.argtypes = [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_longlong, c_int, c_int,
                                       c_longlong, c_longlong, c_void_p]
    L.dlcfn_lrn_fwd.argtypes = [c_void_p, c_void_p, c_void_p, c_longlong, c_int, c_int, c_float, c_float, c_float,
                                c_void_p]
    L.dlcfn_lrn_bwd.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_longlong, c_int, c_int, c_float, c_float,
                                c_float, c_void_p]
    L.dlcfn_depthwise_fwd.argtypes = [c_void_p, c_void_p, c_void_p] + [c_int] * 7 + [c_void_p]
    L.dlcfn_depthwise_bwd_data.argtypes = [c_void_p, c_void_p, c_void_p] + [c_int] * 7 + [c_void_p]
    L.dlcfn_depthwise_filter_blocks.argtypes = [c_int, c_int, c_int]
    L.dlcfn_depthwise_bwd_filter.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p] + [c_int] * 7 + [c_void_p]
    L.dlcfn_subsample_nhwc_fwd.argtypes = [c_void_p, c_void_p] + [c_int] * 5 + [c_void_p]
    L.dlcfn_subsample_nhwc_bwd.argtypes = [c_void_p, c_void_p] + [c_int] * 5 + [c_void_p]
    for f in ("dlcfn_softmax_ce_fwd", "dlcfn_softmax_ce_bwd", "dlcfn_lrn_fwd", "dlcfn_lrn_bwd", "dlcfn_depthwise_fwd",
              "dlcfn_depthwise_bwd_data", "dlcfn_depthwise_filter_blocks", "dlcfn_depthwise_bwd_filter",
              "dlcfn_subsample_nhwc_fwd", "dlcfn_subsample_nhwc_bwd"):
        getattr(L, f).restype = c_int
    L._nn_declared = True
    return L


# ------------------------------------------------------------------------------------------- softmax cross-entropy
class _SoftmaxCEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, ignore_index, num_classes):
        L = _declare()
        x = logits if logits.stride(-1) == 1 else logits.contiguous()
        M, width = x.shape
        C = num_classes
        labels = labels.to(torch.int64).contiguous()
        loss = torch.empty(M, dtype=torch.float32, device=x.device)
        lse = torch.empty(M, dtype=torch.float32, device=x.device)
        native.check(L.dlcfn_softmax_ce_fwd(x.data_ptr(), native.dtype_code(x.dtype), labels.data_ptr(), loss.data_ptr(),
                                            lse.data_ptr(), M, C, x.stride(0), ignore_index, native.stream_ptr()),
                     "softmax_ce_fwd")
        native._count()
        ctx.save_for_backward(x, labels, lse)
        ctx.ignore_index, ctx.C = ignore_index, C
        return loss

    @staticmethod
    def backward(ctx, dloss):
        L = _declare()
        x, labels, lse = ctx.saved_tensors
        M, width = x.shape
        C = ctx.C
        valid = (labels != ctx.ignore_index) & (labels >= 0) & (labels < C)
        row_scale = (dloss.to(torch.float32) * valid).contiguous()
        dx = torch.empty((M, width), dtype=x.dtype, device=x.device)
        native.check(L.dlcfn_softmax_ce_bwd(x.data_ptr(), native.dtype_code(x.dtype), labels.data_ptr(), lse.data_ptr(),
                                            row_scale.data_ptr(), dx.data_ptr(), M, C, width, x.stride(0), width,
                                            native.stream_ptr()), "softmax_ce_bwd")
        native._count()
        return dx, None, None, None


def softmax_cross_entropy(logits: torch.Tensor, labels: torch.Tensor, ignore_index: int = -100,
                          reduction: str = "mean", num_classes: int | None = None) -> torch.Tensor:
    """Cross-entropy of `logits` [..., C] against integer `labels` [...] with fp32 accumulation.

    On CUDA (fp32 or bf16 logits) forward is one pass over the logits producing per-row loss and log-sum-exp; backward
    is one more pass writing (softmax - onehot) * scale directly in the logits' dtype - the logits are never
    up-cast to an fp32 copy and no probability tensor is materialised.

    `num_classes` < logits.shape[-1] treats the trailing columns as GEMM padding (a vocabulary rounded up for TMA):
    they are excluded from the softmax and receive exact zero gradient, without slicing (= copying) the logits."""
    width = logits.shape[-1]
    C = width if num_classes is None else num_classes
    x2 = logits.reshape(-1, width)
    y1 = labels.reshape(-1)
    if x2.is_cuda and x2.dtype in (torch.float32, torch.bfloat16):
        rows = _SoftmaxCEFn.apply(x2, y1, ignore_index, C)
    else:
        if x2.is_cuda and not native.lib_path().exists():
            raise native.NativeUnavailable("CUDA softmax_cross_entropy requested but libdlcfn_cuda.so is missing")
        rows = F.cross_entropy(x2[:, :C].float(), y1, ignore_index=ignore_index, reduction="none")
    if reduction == "none":
        return rows.view(labels.shape)
    if reduction == "sum":
        return rows.sum()
    valid = (y1 != ignore_index).sum().clamp(min=1)
    return rows.sum() / valid


class SoftmaxCrossEntropyLoss(nn.Module):
    """gluon.loss.SoftmaxCrossEntropyLoss / symbol SoftmaxOutput equivalent."""

    def __init__(self, ignore_index: int = -100, reduction: str = "mean"):
        super().__init__()
        self.ignore_index, self.reduction = ignore_index, reduction

    def forward(self, logits, labels):
        return softmax_cross_entropy(logits, labels, self.ignore_index, self.reduction)


# ----------------------------------------------------------------------------------------------------------- LRN
class _LRNFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, size, alpha_eff, beta, k):
        L = _declare()
        if not x.is_contiguous(memory_format=torch.channels_last):
            x = x.contiguous(memory_format=torch.channels_last)
        N, C, H, W = x.shape
        y = torch.empty_like(x, memory_format=torch.channels_last)
        need_grad = ctx.needs_input_grad[0]
        scale = torch.empty(N * H * W * C, dtype=torch.float32, device=x.device) if need_grad else None
        native.check(L.dlcfn_lrn_fwd(x.data_ptr(), y.data_ptr(), scale.data_ptr() if scale is not None else None,
                                     N * H * W, C, size, alpha_eff, beta, k, native.stream_ptr()), "lrn_fwd")
        native._count()
        if need_grad:
            ctx.save_for_backward(x, scale)
        ctx.cfg = (size, alpha_eff, beta, k)
        return y

    @staticmethod
    def backward(ctx, dy):
        L = _declare()
        x, scale = ctx.saved_tensors
        size, alpha_eff, beta, k = ctx.cfg
        if not dy.is_contiguous(memory_format=torch.channels_last):
            dy = dy.contiguous(memory_format=torch.channels_last)
        N, C, H, W = x.shape
        dx = torch.empty_like(x, memory_format=torch.channels_last)
        native.check(L.dlcfn_lrn_bwd(x.data_ptr(), dy.data_ptr(), scale.data_ptr(), dx.data_ptr(), N * H * W, C, size,
                                     alpha_eff, beta, k, native.stream_ptr()), "lrn_bwd")
        native._count()
        return dx, None, None, None, None


def local_response_norm(x: torch.Tensor, size: int, alpha: float = 1e-4, beta: float = 0.75, k: float = 1.0,
                        tf_style: bool = False) -> torch.Tensor:
    """Across-channel LRN.  torch / cuDNN semantics by default: scale = k + alpha/size * sum(x^2 over `size` channels);
    `tf_style=True` uses alpha undivided (tf.nn.lrn with depth_radius = (size-1)/2)."""
    alpha_eff = alpha if tf_style else alpha / size
    if x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 4:
        return _LRNFn.apply(x, size, float(alpha_eff), float(beta), float(k))
    if x.is_cuda and not native.lib_path().exists():
        raise native.NativeUnavailable("CUDA LRN requested but libdlcfn_cuda.so is missing")
    return F.local_response_norm(x, size, alpha_eff * size, beta, k)


class LocalResponseNorm(nn.Module):
    def __init__(self, size: int, alpha: float = 1e-4, beta: float = 0.75, k: float = 1.0, tf_style: bool = False):
        super().__init__()
        self.size, self.alpha, self.beta, self.k, self.tf_style = size, alpha, beta, k, tf_style

    def forward(self, x):
        return local_response_norm(x, self.size, self.alpha, self.beta, self.k, self.tf_style)

    def extra_repr(self):
        return f"{self.size}, alpha={self.alpha}, beta={self.beta}, k={self.k}, tf_style={self.tf_style}, kernel=dlcfn-nhwc"


# ----------------------------------------------------------------------------------------------- depthwise conv
class _DepthwiseFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, stride, padding):
        L = _declare()
        if not x.is_contiguous(memory_format=torch.channels_last):
            x = x.contiguous(memory_format=torch.channels_last)
        N, C, H, W = x.shape
        K = weight.shape[-1]
        w_khwc = weight.detach().reshape(C, K, K).permute(1, 2, 0).contiguous().float()      # [K,K,C] fp32
        OH, OW = (H + 2 * padding - K) // stride + 1, (W + 2 * padding - K) // stride + 1
        y = torch.empty((N, C, OH, OW), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
        native.check(L.dlcfn_depthwise_fwd(x.data_ptr(), w_khwc.data_ptr(), y.data_ptr(), N, H, W, C, K, stride, padding,
                                           native.stream_ptr()), "depthwise_fwd")
        native._count()
        ctx.save_for_backward(x, w_khwc)
        ctx.geom = (N, C, H, W, K, stride, padding, OH, OW, weight.dtype)
        return y

    @staticmethod
    def backward(ctx, dy):
        L = _declare()
        x, w_khwc = ctx.saved_tensors
        N, C, H, W, K, S, P, OH, OW, wdtype = ctx.geom
        if not dy.is_contiguous(memory_format=torch.channels_last):
            dy = dy.contiguous(memory_format=torch.channels_last)
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty((N, C, H, W), dtype=dy.dtype, device=dy.device, memory_format=torch.channels_last)
            native.check(L.dlcfn_depthwise_bwd_data(dy.data_ptr(), w_khwc.data_ptr(), dx.data_ptr(), N, H, W, C, K, S, P,
                                                    native.stream_ptr()), "depthwise_bwd_data")
            native._count()
        if ctx.needs_input_grad[1]:
            B = L.dlcfn_depthwise_filter_blocks(N, OH, OW)
            partial = torch.empty(B * K * K * C, dtype=torch.float32, device=dy.device)
            dwk = torch.empty((K, K, C), dtype=torch.float32, device=dy.device)
            native.check(L.dlcfn_depthwise_bwd_filter(x.data_ptr(), dy.data_ptr(), partial.data_ptr(), dwk.data_ptr(), N, H,
                                                      W, C, K, S, P, native.stream_ptr()), "depthwise_bwd_filter")
            native._count(2)
            dw = dwk.permute(2, 0, 1).reshape(C, 1, K, K).to(wdtype)
        return dx, dw, None, None


def depthwise_conv2d(x: torch.Tensor, weight: torch.Tensor, stride: int = 1, padding: int = 0) -> torch.Tensor:
    """weight [C,1,K,K] (channel multiplier 1)."""
    C = x.shape[1]
    if x.is_cuda and x.dtype == torch.bfloat16 and C % 8 == 0 and weight.shape[0] == C and weight.shape[-1] <= 11:
        return _DepthwiseFn.apply(x, weight, stride, padding)
    if x.is_cuda and not native.lib_path().exists():
        raise native.NativeUnavailable("CUDA depthwise conv requested but libdlcfn_cuda.so is missing")
    return F.conv2d(x, weight.to(x.dtype), None, stride, padding, 1, C)


class DepthwiseConv2d(nn.Module):
    def __init__(self, channels: int, kernel_size: int, stride: int = 1, padding: int = 0):
        super().__init__()
        self.channels, self.kernel_size, self.stride, self.padding = channels, kernel_size, stride, padding
        self.weight = nn.Parameter(torch.empty(channels, 1, kernel_size, kernel_size))
        nn.init.kaiming_normal_(self.weight, mode="fan_out", nonlinearity="relu")

    def forward(self, x):
        return depthwise_conv2d(x, self.weight, self.stride, self.padding)


# ---------------------------------------------------------------------------------------------------- subsample
class _SubsampleFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, stride):
        L = _declare()
        if not x.is_contiguous(memory_format=torch.channels_last):
            x = x.contiguous(memory_format=torch.channels_last)
        N, C, H, W = x.shape
        OH, OW = (H - 1) // stride + 1, (W - 1) // stride + 1
        y = torch.empty((N, C, OH, OW), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
        native.check(L.dlcfn_subsample_nhwc_fwd(x.data_ptr(), y.data_ptr(), N, H, W, C, stride, native.stream_ptr()),
                     "subsample_fwd")
        native._count()
        ctx.geom = (N, C, H, W, stride)
        return y

    @staticmethod
    def backward(ctx, dy):
        L = _declare()
        N, C, H, W, stride = ctx.geom
        if not dy.is_contiguous(memory_format=torch.channels_last):
            dy = dy.contiguous(memory_format=torch.channels_last)
        dx = torch.empty((N, C, H, W), dtype=dy.dtype, device=dy.device, memory_format=torch.channels_last)
        native.check(L.dlcfn_subsample_nhwc_bwd(dy.data_ptr(), dx.data_ptr(), N, H, W, C, stride, native.stream_ptr()),
                     "subsample_bwd")
        native._count()
        return dx, None


def subsample(x: torch.Tensor, stride: int) -> torch.Tensor:
    """x[:, :, ::stride, ::stride] as a dense channels_last tensor (one gather kernel; backward = one scatter-with-zeros)."""
    if stride == 1:
        return x
    if x.is_cuda and x.dtype == torch.bfloat16 and x.shape[1] % 8 == 0:
        return _SubsampleFn.apply(x, stride)
    return x[:, :, ::stride, ::stride].contiguous(memory_format=torch.channels_last)
