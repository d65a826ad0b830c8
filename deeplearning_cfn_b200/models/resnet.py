"""ResNet (pre-activation v2) - the architecture of the reference's ImageNet benchmark symbol.

Parity target: MX/example/image-classification/symbols/resnet.py:29-177 (bottleneck units
[3,4,6,3], filters [64,256,512,1024,2048], BatchNorm eps 2e-5 momentum 0.9, `bn_data` on the input,
7x7/2 stem + BN + ReLU + 3x3/2 max-pool for ImageNet, 3x3 stem for CIFAR).  ResNet-50 has 157
learnable tensors / 25.55 M parameters, which this module reproduces.

B200 shape: channels_last activations, bf16 conv/FC weights with fp32 BatchNorm parameters.  Every convolution runs on the
hand-written tcgen05 kernels (ops/conv.py; cuDNN only as the selectable yardstick): the 1x1 layers (2/3 of the network) are
plain GEMMs on the NHWC matrix, the 3x3 layers and the strided projection shortcuts implicit GEMMs through im2col-mode TMA,
the 7x7 stem + `bn_data` one fused function over the space-to-depth image (`_FusedStemFn`); bn2 / bn3 / bn0 take their
statistics from the epilogue of the convolution in front of them; the classifier GEMM (`ops.linear.Linear`) can consume
parameter-server-resident weights directly.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..ops.batchnorm import BatchNormReLU2d
from ..ops.conv import Conv1x1, ConvKxK
from ..ops.linear import Linear
from ..ops.pool import MaxPool2d, global_avg_pool

_FUSED_STEM = __import__("os").environ.get("DLCFN_FUSED_STEM", "1") != "0"
_EPS = 2e-5
_MOM = 0.1  # torch momentum = 1 - mxnet bn_mom(0.9)


def _bn(c: int, affine_gamma: bool = True, relu: bool = True) -> nn.BatchNorm2d:
    """BatchNorm followed by ReLU, fused in one dlcfn kernel pair (every BN of the reference symbol except
    `bn_data` is followed by a ReLU activation)."""
    bn = BatchNormReLU2d(c, eps=_EPS, momentum=_MOM, relu=relu)
    if not affine_gamma:  # fix_gamma=True in the reference: gamma frozen at 1
        bn.weight.requires_grad_(False)
    return bn


class BottleneckV2(nn.Module):
    def __init__(self, cin: int, cout: int, stride: int, dim_match: bool):
        super().__init__()
        mid = cout // 4
        self.bn1 = _bn(cin)
        self.conv1 = Conv1x1(cin, mid)
        self.bn2 = _bn(mid)
        self.conv2 = ConvKxK(mid, mid, 3, stride, 1)
        self.bn3 = _bn(mid)
        self.conv3 = Conv1x1(mid, cout)
        self.sc = None if dim_match else Conv1x1(cin, cout, stride)

    def forward(self, x):
        """`x` is either a tensor or the un-summed pair (branch, shortcut) of the previous unit; the pair's add is fused
        into this unit's first BatchNorm.  Returns this unit's own un-summed pair."""
        if isinstance(x, tuple):
            x, a = self.bn1.forward_add(x[0], x[1])
        else:
            a = self.bn1(x)  # BN + ReLU fused
        # stats=True: the convolution's epilogue accumulates the BatchNorm statistics of its output (ops/conv.py)
        y = self.conv1(a, stats=True)
        y = self.conv2(self.bn2(y), stats=True)
        y = self.conv3(self.bn3(y))
        return y, (x if self.sc is None else self.sc(a))


class BasicV2(nn.Module):
    def __init__(self, cin: int, cout: int, stride: int, dim_match: bool):
        super().__init__()
        self.bn1 = _bn(cin)
        self.conv1 = ConvKxK(cin, cout, 3, stride, 1)
        self.bn2 = _bn(cout)
        self.conv2 = ConvKxK(cout, cout, 3, 1, 1)
        self.sc = None if dim_match else Conv1x1(cin, cout, stride)

    def forward(self, x):
        if isinstance(x, tuple):
            x, a = self.bn1.forward_add(x[0], x[1])
        else:
            a = self.bn1(x)
        y = self.conv2(self.bn2(self.conv1(a, stats=True)))
        return y, (x if self.sc is None else self.sc(a))


_IMAGENET_UNITS = {18: [2, 2, 2, 2], 34: [3, 4, 6, 3], 50: [3, 4, 6, 3], 101: [3, 4, 23, 3],
                   152: [3, 8, 36, 3], 200: [3, 24, 36, 3], 269: [3, 30, 48, 8]}


class DataBatchNorm(BatchNormReLU2d):
    """`bn_data`: BatchNorm over the 3 input channels (gamma fixed, reference `fix_gamma=True`).  On CUDA/bf16 the
    image is zero-padded to 8 channels first so that the dlcfn NHWC kernels (16-byte channel vectors) apply; the
    parameters and running statistics keep their reference shape [3].  Returns the 8-channel tensor (channels 3..7
    are exactly zero) - `StemConv` knows how to consume it."""

    def forward(self, x):
        if not (x.is_cuda and x.dtype == torch.bfloat16 and self.num_features == 3 and x.shape[1] == 3):
            return super().forward(x)
        from ..ops.batchnorm import batch_norm_relu
        from ..ops.pool import pad3to8
        x8 = pad3to8(x)
        g8 = F.pad(self.weight, (0, 5), value=1.0)
        b8 = F.pad(self.bias, (0, 5))
        rm8, rv8 = F.pad(self.running_mean, (0, 5)), F.pad(self.running_var, (0, 5), value=1.0)
        y8 = batch_norm_relu(x8, g8, b8, rm8, rv8, self.training, self.momentum, self.eps, relu=False)
        if self.training:
            with torch.no_grad():
                self.running_mean.copy_(rm8[:3])
                self.running_var.copy_(rv8[:3])
        return y8


class StemConv(nn.Conv2d):
    """The first convolution (Cin = 3).  Tensor-core implicit-GEMM kernels want 16-byte (8 x bf16) channel
    vectors and a deep reduction; with 3 input channels cuDNN pads and re-lays-out the activation on every call and
    then runs a K = 147 GEMM on a legacy kernel.  The parameter keeps the reference shape [Cout, 3, k, k]; on
    CUDA/bf16 the same convolution is evaluated in a tensor-core friendly form:
      * 7x7 / stride 2 / pad 3 (ImageNet stem): space-to-depth by 2.  With the kernel zero-extended to 8x8 taps,
        out[o] = sum_k w[k] xp[2o+k] = sum_{a<4, b<2} w[2a+b] s2d[o+a][b], i.e. a 4x4 stride-1 convolution over the
        (115 x 115 x 12 -> 16 channel) space-to-depth image - identical math, K = 256;
      * anything else: input and weight zero-padded to 8 channels."""

    def forward(self, x):
        if not (x.is_cuda and x.dtype == torch.bfloat16 and self.in_channels == 3):
            return super().forward(x[:, :3] if x.shape[1] != 3 else x)
        w = self.weight
        if x.shape[1] == 8 and not (self.kernel_size == (7, 7) and self.stride == (2, 2)):
            w8 = F.pad(w, (0, 0, 0, 0, 0, 5)).contiguous(memory_format=torch.channels_last)   # input already 8-channel
            return F.conv2d(x, w8, self.bias, self.stride, self.padding)
        if self.kernel_size == (7, 7) and self.stride == (2, 2) and self.padding == (3, 3) and x.shape[-1] % 2 == 0 \
                and x.shape[-2] % 2 == 0:
            from ..ops.pool import stem_space_to_depth
            if x.shape[1] == 3:
                x = F.pad(x, (0, 0, 0, 0, 0, 5))
            s2d = stem_space_to_depth(x)                                            # one fused kernel each way
            O, C = w.shape[0], 3
            w8 = F.pad(w, (0, 1, 0, 1))                                            # [O,3,8,8], last tap row/col zero
            w4 = w8.reshape(O, C, 4, 2, 4, 2).permute(0, 1, 3, 5, 2, 4).reshape(O, C * 4, 4, 4)
            w4 = F.pad(w4, (0, 0, 0, 0, 0, 4)).contiguous(memory_format=torch.channels_last)
            return F.conv2d(s2d, w4, self.bias, 1, 0)
        if x.shape[1] != 3:
            x = x[:, :3]
        x8 = F.pad(x, (0, 0, 0, 0, 0, 5)).contiguous(memory_format=torch.channels_last)
        w8 = F.pad(w, (0, 0, 0, 0, 0, 5)).contiguous(memory_format=torch.channels_last)
        return F.conv2d(x8, w8, self.bias, self.stride, self.padding)


def _stem_w4(w: torch.Tensor) -> torch.Tensor:
    """[O,3,7,7] -> the 4x4 filter over the 16-channel space-to-depth image, [O,16,4,4] channels_last:
    w4[o, c*4 + i*2 + j, a, b] = w[o, c, 2a+i, 2b+j] (zero where 2a+i or 2b+j = 7, channels 12..15 zero)."""
    O, C = w.shape[0], 3
    w8 = F.pad(w, (0, 1, 0, 1))
    w4 = w8.reshape(O, C, 4, 2, 4, 2).permute(0, 1, 3, 5, 2, 4).reshape(O, C * 4, 4, 4)
    return F.pad(w4, (0, 0, 0, 0, 0, 4)).contiguous(memory_format=torch.channels_last)


def _stem_w4_adjoint(dw4: torch.Tensor) -> torch.Tensor:
    """Gradient of `_stem_w4`: [O,16,4,4] -> [O,3,7,7]."""
    O = dw4.shape[0]
    d = dw4[:, :12].reshape(O, 3, 2, 2, 4, 4).permute(0, 1, 4, 2, 5, 3).reshape(O, 3, 8, 8)
    return d[:, :, :7, :7].contiguous()


_TAP_MASKS: dict = {}


def _stem_tap_masks(H: int, W: int, P: int, Q: int, device):
    """A[r, p] = 1 iff output row p's filter row r reads a real image row (2p + r - 3 in [0, H)); B likewise for columns."""
    key = (H, W, P, Q, str(device))
    if key not in _TAP_MASKS:
        def mask(extent, outs):
            m = torch.zeros(7, outs)
            for t in range(7):
                lo, hi = max(0, (3 - t + 1) // 2), min(outs - 1, (extent + 2 - t) // 2)
                m[t, lo:hi + 1] = 1.0
            return m.to(device)
        _TAP_MASKS[key] = (mask(H, P), mask(W, Q))
    return _TAP_MASKS[key]


class _FusedStemFn(torch.autograd.Function):
    """bn_data (BatchNorm over the 3 image channels, gamma fixed) + 7x7 / stride 2 / pad 3 convolution, all on dlcfn kernels:
    pad 3->8 channels, fused BatchNorm, space-to-depth (one kernel each), then the 4x4 / stride-1 convolution over the
    16-channel s2d image as an implicit GEMM whose "pixels" are 4 adjacent columns (ops/native.py:_stem_view).

    Backward needs no input gradient (the image): the filter gradient is one split-K implicit GEMM, and the only thing the
    BatchNorm wants - d beta[c] = sum over real pixels of the convolution's input gradient - is, by linearity,
        sum_{k,r,s} w[k,c,r,s] * T[k,r,s],     T[k,r,s] = sum_n sum_{(p,q): pixel (2p+r-3, 2q+s-3) inside the image} dy[n,k,p,q]
    i.e. rectangle sums of the batch-summed output gradient: one reduction over dy instead of a dgrad and a BatchNorm
    backward over the 224x224 image (reference: MX/src/operator/nn/batch_norm.cu backward with fix_gamma)."""

    @staticmethod
    def forward(ctx, x, w, bn_bias, bn: "DataBatchNorm", want_stats: bool):
        from ..ops import native
        from ..ops.batchnorm import batch_norm_relu
        from ..ops.pool import pad3to8, stem_space_to_depth
        x8 = pad3to8(x)
        g8 = F.pad(bn.weight.detach(), (0, 5), value=1.0)
        b8 = F.pad(bn_bias.detach(), (0, 5))
        rm8, rv8 = F.pad(bn.running_mean, (0, 5)), F.pad(bn.running_var, (0, 5), value=1.0)
        y8 = batch_norm_relu(x8, g8, b8, rm8, rv8, bn.training, bn.momentum, bn.eps, relu=False)
        if bn.training:
            bn.running_mean.copy_(rm8[:3])
            bn.running_var.copy_(rv8[:3])
        s2d = stem_space_to_depth(y8)
        stats = native.new_stats(w.shape[0], x.device) if want_stats else None
        y = native.stem_conv_fprop(s2d, _stem_w4(w.detach()), stats)
        ctx.save_for_backward(s2d, w)
        ctx.hw = (x.shape[2], x.shape[3])
        if stats is not None:
            ctx.mark_non_differentiable(stats)
        return y, stats

    @staticmethod
    def backward(ctx, dy, _dstats=None):
        from ..ops import native
        s2d, w = ctx.saved_tensors
        if not dy.is_contiguous(memory_format=torch.channels_last):
            dy = dy.contiguous(memory_format=torch.channels_last)
        dw = _stem_w4_adjoint(native.stem_conv_wgrad(s2d, dy)).to(w.dtype) if ctx.needs_input_grad[1] else None
        dbeta = None
        if ctx.needs_input_grad[2]:
            H, W = ctx.hw
            D = native.batch_sum(dy)                                        # [K, P, Q] fp32 (strides of dy[0]): one pass over dy
            A, B = _stem_tap_masks(H, W, D.shape[1], D.shape[2], D.device)  # [7, P], [7, Q] 0/1: outputs whose tap hits a real pixel
            T = torch.einsum("rp,kpq,sq->krs", A, D, B)                     # rectangle sums as two small contractions
            dbeta = torch.einsum("kcrs,krs->c", w.float(), T)
        return None, dw, dbeta, None, None


class ResNetV2(nn.Module):
    def __init__(self, num_layers: int = 50, num_classes: int = 1000, image_size: int = 224):
        super().__init__()
        if image_size <= 32:  # CIFAR variant (resnet.py:165-177 first branch)
            if (num_layers - 2) % 9 == 0 and num_layers >= 164:
                units, filters, bottleneck = [(num_layers - 2) // 9] * 3, [16, 64, 128, 256], True
            elif (num_layers - 2) % 6 == 0 and num_layers < 164:
                units, filters, bottleneck = [(num_layers - 2) // 6] * 3, [16, 16, 32, 64], False
            else:
                raise ValueError(f"no CIFAR ResNet with {num_layers} layers")
        else:
            if num_layers not in _IMAGENET_UNITS:
                raise ValueError(f"no ImageNet ResNet with {num_layers} layers")
            units = _IMAGENET_UNITS[num_layers]
            bottleneck = num_layers >= 50
            filters = [64, 256, 512, 1024, 2048] if bottleneck else [64, 64, 128, 256, 512]
        self.bn_data = DataBatchNorm(3, eps=_EPS, momentum=_MOM, relu=False)
        self.bn_data.weight.requires_grad_(False)      # fix_gamma=True
        if image_size <= 32:
            self.stem = nn.Sequential(StemConv(3, filters[0], 3, 1, 1, bias=False))
        else:
            self.stem = nn.Sequential(StemConv(3, filters[0], 7, 2, 3, bias=False), _bn(filters[0]),
                                      MaxPool2d(3, 2, 1))
        unit = BottleneckV2 if bottleneck else BasicV2
        blocks = []
        cin = filters[0]
        for i, n in enumerate(units):
            blocks.append(unit(cin, filters[i + 1], 1 if i == 0 else 2, False))
            cin = filters[i + 1]
            for _ in range(n - 1):
                blocks.append(unit(cin, cin, 1, True))
        self.body = nn.Sequential(*blocks)
        self.bn1 = _bn(cin)
        self.fc1 = Linear(cin, num_classes)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")

    def _fused_stem_ok(self, x) -> bool:
        c = self.stem[0]
        return (len(self.stem) == 3 and x.is_cuda and x.dtype == torch.bfloat16 and x.shape[1] == 3 and not x.requires_grad and
                c.kernel_size == (7, 7) and c.stride == (2, 2) and c.padding == (3, 3) and c.out_channels % 64 == 0 and
                c.weight.dtype == torch.bfloat16 and x.shape[2] % 2 == 0 and x.shape[3] % 2 == 0 and
                x.is_contiguous(memory_format=torch.channels_last) and _FUSED_STEM)

    def forward(self, x):
        if self._fused_stem_ok(x):
            y = _FusedStemFn.apply(x, self.stem[0].weight, self.bn_data.bias, self.bn_data, self.training)
            x = self.stem[2](self.stem[1](y))  # bn0 takes (y, statistics from the convolution epilogue), then max-pool
        else:
            x = self.stem(self.bn_data(x))
        x = self.body(x)                       # (branch, shortcut) of the last unit, still un-summed
        _, x = self.bn1.forward_add(x[0], x[1])
        return self.fc1(global_avg_pool(x))


def resnet(num_layers: int = 50, num_classes: int = 1000, image_size: int = 224) -> ResNetV2:
    return ResNetV2(num_layers, num_classes, image_size)


def to_b200_layout(model: nn.Module, dtype: torch.dtype = torch.bfloat16) -> nn.Module:
    """channels_last + `dtype` weights for conv/linear; BatchNorm parameters and statistics stay fp32
    (cuDNN mixed-dtype batch norm) - the worker half of the reference's multi_precision mode."""
    from ..ops.nn import DepthwiseConv2d
    model = model.to(memory_format=torch.channels_last)
    for m in model.modules():
        if isinstance(m, (nn.Conv2d, nn.Linear, Linear, nn.Embedding, DepthwiseConv2d)):
            m.to(dtype)
        elif isinstance(m, (nn.LayerNorm,)):
            m.to(dtype)
    return model
