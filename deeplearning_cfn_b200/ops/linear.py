"""Linear layer on the hand-written tcgen05 GEMM (forward, dgrad and wgrad all on csrc/cuda/gemm_tcgen05.cu).

Reference analogue: FullyConnected -> linalg_gemm -> cublasSgemmEx (reference M2,
fully_connected-inl.h:110-171, linalg_impl.h:189-273).  Here the three GEMMs are
  fwd   Y[M,N]  = X[M,K]  * W[N,K]^T            (both operands K-major)
  dgrad dX[M,K] = dY[M,N] * W[N,K]              (B = W read MN-major: no transpose materialised)
  wgrad dW[N,K] = dY[M,N]^T * X[M,K]            (A and B both MN-major)
On CPU tensors (unit tests without a GPU) the same math runs through torch.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import native


_COLSUM_MIN_ROWS = int(__import__("os").environ.get("DLCFN_COLSUM_MIN_ROWS", "2048"))


class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        x2 = x.reshape(-1, x.shape[-1])
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        ctx.save_for_backward(x2, weight)
        ctx.has_bias = bias is not None
        ctx.x_shape = x.shape
        bias_f32 = None
        if bias is not None:
            bias_f32 = bias if bias.dtype == torch.float32 else bias.float()
        # the result is allocated in its final shape and handed to the GEMM as a 2-D alias: returning `y.view(...)` would
        # make the Function's output a view, which autograd forbids to modify in place (classifier stacks follow their
        # Linear layers with nn.ReLU(inplace=True))
        y = torch.empty((*x.shape[:-1], weight.shape[0]), dtype=torch.bfloat16, device=x.device)
        native.gemm_bf16(x2, weight, bias=bias_f32, out=y.view(-1, weight.shape[0]))
        return y

    @staticmethod
    def backward(ctx, dy):
        x2, weight = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1])
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = native.gemm_bf16(dy2, weight, b_mn=True, out_dtype=torch.bfloat16).view(ctx.x_shape)
        if ctx.needs_input_grad[1]:
            dw = native.gemm_bf16(dy2, x2, a_mn=True, b_mn=True, out_dtype=torch.bfloat16)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            # tall gradients (BERT: 8192 tokens x 1024 .. 30522 columns) go through the streaming column-sum kernel - torch's
            # strided reduce took 2.1 ms per BERT-large step for them; short ones (a classifier's 256 rows) are not worth a launch pair
            ok = (dy2.is_cuda and dy2.shape[0] >= _COLSUM_MIN_ROWS and dy2.shape[1] % 8 == 0 and dy2.stride(0) % 8 == 0
                  and dy2.data_ptr() % 16 == 0)
            db = native.colsum(dy2) if ok else dy2.sum(0, dtype=torch.float32)
        return dx, dw, db


def linear(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor | None = None) -> torch.Tensor:
    """y = x @ weight.T + bias.  CUDA bf16 -> tcgen05 kernels; anything else -> torch (CPU tests)."""
    if x.is_cuda and x.dtype == torch.bfloat16 and weight.dtype == torch.bfloat16:
        ok = weight.shape[1] % 8 == 0 and weight.shape[0] % 8 == 0 and weight.is_contiguous()
        if not ok:
            raise ValueError(f"tcgen05 linear needs in/out features % 8 == 0 and contiguous weight, got {tuple(weight.shape)}")
        y = _LinearFn.apply(x, weight, bias)
        return y
    if x.is_cuda and not native.lib_path().exists():
        raise native.NativeUnavailable("CUDA linear requested but libdlcfn_cuda.so is missing")
    b = bias.to(x.dtype) if bias is not None else None
    return F.linear(x, weight.to(x.dtype), b)


class Linear(nn.Module):
    """Drop-in nn.Linear whose CUDA/bf16 path is the dlcfn tcgen05 GEMM.

    TMA needs 16-byte row strides, so the weight is stored with in/out features rounded up to multiples of 8;
    the padding rows/columns are zero, receive zero gradient, and are sliced off the output."""

    def __init__(self, in_features: int, out_features: int, bias: bool = True):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.in_padded = (in_features + 7) // 8 * 8
        self.out_padded = (out_features + 7) // 8 * 8
        self.weight = nn.Parameter(torch.zeros(self.out_padded, self.in_padded))
        self.bias = nn.Parameter(torch.zeros(self.out_padded)) if bias else None
        with torch.no_grad():
            nn.init.kaiming_uniform_(self.weight[:out_features, :in_features], a=math.sqrt(5))

    def forward(self, x):
        if self.in_padded != self.in_features:
            x = F.pad(x, (0, self.in_padded - self.in_features))
        y = linear(x, self.weight, self.bias)
        return y if self.out_padded == self.out_features else y[..., : self.out_features]

    def extra_repr(self):
        return f"in_features={self.in_features}, out_features={self.out_features}, bias={self.bias is not None}, kernel=tcgen05"
