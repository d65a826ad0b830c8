"""The rest of the reference's image-classification model families, assembled from the dlcfn operators.

The reference ships one symbol file per family under `MX/example/image-classification/symbols/` (`train_imagenet.py
--network <name>`, `common/fit.py`) and the gluon model zoo behind `image_classification.py --model <name>`
(R/README.md:139).  `models/resnet.py` (v2), `vgg.py`, `inception.py` (v3) and `small.py` cover the ones the first-party
commands name; this file adds the others so that `--network` / `--model` accept what the reference accepts:

    resnet-v1 (symbols/resnet-v1.py: post-activation units, stride on the first 1x1)      ResNetV1
    resnext   (symbols/resnext.py: grouped 3x3, mid = cout/2, num_group 32)                ResNetV1(groups=32)
    mobilenet (symbols/mobilenet.py: depthwise-separable stack, alpha multiplier)          MobileNet
    mobilenetv2 (symbols/mobilenetv2.py: inverted residuals, ReLU6, multiplier)            MobileNetV2
    googlenet (symbols/googlenet.py: conv+ReLU inception modules, no BatchNorm)            GoogLeNet
    inception-bn (symbols/inception-bn.py)                                                 InceptionBN
    inception-v4 (symbols/inception-v4.py)                                                 InceptionV4
    inception-resnet-v2 (symbols/inception-resnet-v2.py)                                   InceptionResNetV2
    squeezenet1.0 / 1.1, densenet121/161/169/201 (gluon model zoo)                         SqueezeNet, DenseNet

B200 shape: every conv -> BatchNorm -> ReLU triple is one convolution followed by ONE fused BatchNorm+ReLU kernel pair
(`ops.batchnorm.BatchNormReLU2d`, NHWC/bf16), ungrouped 1x1 convolutions are `ops.conv.Conv1x1` (a GEMM on the NHWC
matrix), depthwise 3x3 convolutions are the hand-written NHWC kernel (`ops.nn.DepthwiseConv2d`), max-pools the NHWC kernel
of `ops.pool`, classifiers the tcgen05 `ops.linear.Linear`.  Everything falls back to aten on CPU tensors.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ..ops.batchnorm import BatchNormReLU2d
from ..ops.conv import Conv1x1
from ..ops.linear import Linear
from ..ops.nn import DepthwiseConv2d
from ..ops.pool import MaxPool2d, global_avg_pool

_EPS, _MOM = 2e-5, 0.1


def _pair(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


def conv_bn(cin: int, cout: int, k=1, s=1, p=0, groups: int = 1, relu: bool = True, eps: float = _EPS) -> nn.Sequential:
    """conv (no bias) -> fused BatchNorm(+ReLU)."""
    if _pair(k) == (1, 1) and groups == 1 and _pair(p) == (0, 0) and isinstance(s, int):
        conv = Conv1x1(cin, cout, s)
    else:
        conv = nn.Conv2d(cin, cout, k, s, p, groups=groups, bias=False)
    return nn.Sequential(conv, BatchNormReLU2d(cout, eps=eps, momentum=_MOM, relu=relu))


def conv_relu(cin: int, cout: int, k=1, s=1, p=0) -> nn.Sequential:
    return nn.Sequential(nn.Conv2d(cin, cout, k, s, p), nn.ReLU(inplace=True))


def _init(model: nn.Module, xavier: bool = False) -> None:
    """MSRA-style init for the BatchNorm networks, Xavier (magnitude 2, the reference's default initializer of fit.py) for the
    plain conv+ReLU ones, which diverge at lr 0.1 with the former."""
    for m in model.modules():
        if isinstance(m, nn.Conv2d) and xavier:
            nn.init.xavier_uniform_(m.weight, gain=2 ** 0.5)
            if m.bias is not None:
                nn.init.zeros_(m.bias)
        elif isinstance(m, nn.Conv2d):
            nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            if m.bias is not None:
                nn.init.zeros_(m.bias)
        elif isinstance(m, DepthwiseConv2d):
            nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")


class Concat(nn.Module):
    """Parallel branches concatenated over channels (every inception block)."""

    def __init__(self, *branches: nn.Module):
        super().__init__()
        self.branches = nn.ModuleList(branches)

    def forward(self, x):
        return torch.cat([b(x) for b in self.branches], 1)


# ------------------------------------------------------------------------------------------------ ResNet v1 / ResNeXt
class UnitV1(nn.Module):
    def __init__(self, cin: int, cout: int, stride: int, dim_match: bool, bottleneck: bool, groups: int):
        super().__init__()
        if bottleneck and groups > 1:      # ResNeXt: wide grouped 3x3 carries the stride
            mid = cout // 2
            body = [conv_bn(cin, mid, 1), conv_bn(mid, mid, 3, stride, 1, groups), conv_bn(mid, cout, 1, relu=False)]
        elif bottleneck:                   # resnet-v1.py:49-57: the stride sits on the first 1x1
            mid = cout // 4
            body = [conv_bn(cin, mid, 1, stride), conv_bn(mid, mid, 3, 1, 1), conv_bn(mid, cout, 1, relu=False)]
        else:
            body = [conv_bn(cin, cout, 3, stride, 1), conv_bn(cout, cout, 3, 1, 1, relu=False)]
        self.body = nn.Sequential(*body)
        self.sc = None if dim_match else conv_bn(cin, cout, 1, stride, relu=False)

    def forward(self, x):
        return torch.relu_(self.body(x) + (x if self.sc is None else self.sc(x)))


_UNITS = {18: [2, 2, 2, 2], 34: [3, 4, 6, 3], 50: [3, 4, 6, 3], 101: [3, 4, 23, 3], 152: [3, 8, 36, 3],
          200: [3, 24, 36, 3], 269: [3, 30, 48, 8]}


class ResNetV1(nn.Module):
    """Post-activation ResNet; `groups=32` gives ResNeXt-<layers> (32 x 4d at 50 / 101 layers)."""

    def __init__(self, num_layers: int = 50, num_classes: int = 1000, image_size: int = 224, groups: int = 1):
        super().__init__()
        if image_size <= 32:
            if (num_layers - 2) % 9 == 0 and num_layers >= 164:
                units, filters, bottleneck = [(num_layers - 2) // 9] * 3, [16, 64, 128, 256], True
            elif (num_layers - 2) % 6 == 0 and num_layers < 164:
                units, filters, bottleneck = [(num_layers - 2) // 6] * 3, [16, 16, 32, 64], False
            else:
                raise ValueError(f"no CIFAR ResNet with {num_layers} layers")
            stem = [conv_bn(3, filters[0], 3, 1, 1)]
        else:
            if num_layers not in _UNITS:
                raise ValueError(f"no ImageNet ResNet with {num_layers} layers")
            units, bottleneck = _UNITS[num_layers], num_layers >= 50
            filters = [64, 256, 512, 1024, 2048] if bottleneck else [64, 64, 128, 256, 512]
            stem = [conv_bn(3, filters[0], 7, 2, 3), MaxPool2d(3, 2, 1)]
        blocks, cin = [], filters[0]
        for i, n in enumerate(units):
            blocks.append(UnitV1(cin, filters[i + 1], 1 if i == 0 else 2, False, bottleneck, groups))
            cin = filters[i + 1]
            blocks += [UnitV1(cin, cin, 1, True, bottleneck, groups) for _ in range(n - 1)]
        self.features = nn.Sequential(*stem, *blocks)
        self.fc1 = Linear(cin, num_classes)
        _init(self)

    def forward(self, x):
        return self.fc1(global_avg_pool(self.features(x)))


# ------------------------------------------------------------------------------------------------------- MobileNets
def _dw_bn(c: int, stride: int, relu: bool = True) -> nn.Sequential:
    return nn.Sequential(DepthwiseConv2d(c, 3, stride, 1), BatchNormReLU2d(c, eps=_EPS, momentum=_MOM, relu=relu))


class MobileNet(nn.Module):
    """MobileNet v1: 3x3/2 stem then 13 depthwise-separable pairs (mobilenet.py:85-140); `alpha` in {0.25, 0.5, 0.75, 1}."""
    _PAIRS = [(2, 1), (4, 2), (4, 1), (8, 2), (8, 1), (16, 2), (16, 1), (16, 1), (16, 1), (16, 1), (16, 1), (32, 2), (32, 1)]

    def __init__(self, num_classes: int = 1000, alpha: float = 1.0):
        super().__init__()
        base = int(32 * alpha)
        layers, cin = [conv_bn(3, base, 3, 2, 1)], base
        for mult, stride in self._PAIRS:
            layers += [_dw_bn(cin, stride), conv_bn(cin, base * mult, 1)]
            cin = base * mult
        self.features = nn.Sequential(*layers)
        self.fc = Linear(cin, num_classes)
        _init(self)

    def forward(self, x):
        return self.fc(global_avg_pool(self.features(x)))


class _ReLU6BN(nn.Sequential):
    def __init__(self, c: int):
        super().__init__(BatchNormReLU2d(c, eps=_EPS, momentum=_MOM, relu=False), nn.ReLU6(inplace=True))


class InvertedResidual(nn.Module):
    def __init__(self, cin: int, cout: int, stride: int, expand: int):
        super().__init__()
        mid = cin * expand
        layers = []
        if expand != 1:
            layers += [Conv1x1(cin, mid), _ReLU6BN(mid)]
        layers += [DepthwiseConv2d(mid, 3, stride, 1), _ReLU6BN(mid), Conv1x1(mid, cout),
                   BatchNormReLU2d(cout, eps=_EPS, momentum=_MOM, relu=False)]
        self.body = nn.Sequential(*layers)
        self.residual = stride == 1 and cin == cout

    def forward(self, x):
        return x + self.body(x) if self.residual else self.body(x)


class MobileNetV2(nn.Module):
    _CFG = [(1, 16, 1, 1), (6, 24, 2, 2), (6, 32, 3, 2), (6, 64, 4, 2), (6, 96, 3, 1), (6, 160, 3, 2), (6, 320, 1, 1)]

    def __init__(self, num_classes: int = 1000, multiplier: float = 1.0):
        super().__init__()
        def ch(c):          # channel counts stay multiples of 8: the NHWC kernels move 16-byte channel vectors
            return max(8, int(c * multiplier + 4) // 8 * 8)
        cin = ch(32)
        layers = [nn.Conv2d(3, cin, 3, 2, 1, bias=False), _ReLU6BN(cin)]
        for t, c, n, s in self._CFG:
            for i in range(n):
                layers.append(InvertedResidual(cin, ch(c), s if i == 0 else 1, t))
                cin = ch(c)
        last = ch(1280) if multiplier > 1.0 else 1280
        layers += [Conv1x1(cin, last), _ReLU6BN(last)]
        self.features = nn.Sequential(*layers)
        self.fc = Linear(last, num_classes)
        _init(self)

    def forward(self, x):
        return self.fc(global_avg_pool(self.features(x)))


# --------------------------------------------------------------------------------------------------------- GoogLeNet
def _inception_v1(cin, c1, c3r, c3, c5r, c5, proj) -> Concat:
    return Concat(conv_relu(cin, c1, 1),
                  nn.Sequential(conv_relu(cin, c3r, 1), conv_relu(c3r, c3, 3, 1, 1)),
                  nn.Sequential(conv_relu(cin, c5r, 1), conv_relu(c5r, c5, 5, 1, 2)),
                  nn.Sequential(nn.MaxPool2d(3, 1, 1), conv_relu(cin, proj, 1)))


class GoogLeNet(nn.Module):
    _CFG = [(64, 96, 128, 16, 32, 32), (128, 128, 192, 32, 96, 64), "pool", (192, 96, 208, 16, 48, 64),
            (160, 112, 224, 24, 64, 64), (128, 128, 256, 24, 64, 64), (112, 144, 288, 32, 64, 64),
            (256, 160, 320, 32, 128, 128), "pool", (256, 160, 320, 32, 128, 128), (384, 192, 384, 48, 128, 128)]

    def __init__(self, num_classes: int = 1000):
        super().__init__()
        layers = [conv_relu(3, 64, 7, 2, 3), nn.MaxPool2d(3, 2, 1), conv_relu(64, 64, 1), conv_relu(64, 192, 3, 1, 1),
                  nn.MaxPool2d(3, 2, 1)]
        cin = 192
        for c in self._CFG:
            if c == "pool":
                layers.append(nn.MaxPool2d(3, 2, 1))
            else:
                layers.append(_inception_v1(cin, *c))
                cin = c[0] + c[2] + c[4] + c[5]
        self.features = nn.Sequential(*layers)
        self.fc = Linear(cin, num_classes)
        _init(self, xavier=True)

    def forward(self, x):
        return self.fc(global_avg_pool(self.features(x)))


# ------------------------------------------------------------------------------------------------------ Inception-BN
def _incbn_a(cin, c1, c3r, c3, d3r, d3, pool, proj) -> Concat:
    p = nn.AvgPool2d(3, 1, 1) if pool == "avg" else nn.MaxPool2d(3, 1, 1)
    return Concat(conv_bn(cin, c1, 1),
                  nn.Sequential(conv_bn(cin, c3r, 1), conv_bn(c3r, c3, 3, 1, 1)),
                  nn.Sequential(conv_bn(cin, d3r, 1), conv_bn(d3r, d3, 3, 1, 1), conv_bn(d3, d3, 3, 1, 1)),
                  nn.Sequential(p, conv_bn(cin, proj, 1)))


def _incbn_b(cin, c3r, c3, d3r, d3) -> Concat:
    return Concat(nn.Sequential(conv_bn(cin, c3r, 1), conv_bn(c3r, c3, 3, 2, 1)),
                  nn.Sequential(conv_bn(cin, d3r, 1), conv_bn(d3r, d3, 3, 1, 1), conv_bn(d3, d3, 3, 2, 1)),
                  nn.MaxPool2d(3, 2, 1))


class InceptionBN(nn.Module):
    _CFG = [("a", 64, 64, 64, 64, 96, "avg", 32), ("a", 64, 64, 96, 64, 96, "avg", 64), ("b", 128, 160, 64, 96),
            ("a", 224, 64, 96, 96, 128, "avg", 128), ("a", 192, 96, 128, 96, 128, "avg", 128),
            ("a", 160, 128, 160, 128, 160, "avg", 128), ("a", 96, 128, 192, 160, 192, "avg", 128), ("b", 128, 192, 192, 256),
            ("a", 352, 192, 320, 160, 224, "avg", 128), ("a", 352, 192, 320, 192, 224, "max", 128)]

    def __init__(self, num_classes: int = 1000):
        super().__init__()
        layers = [conv_bn(3, 64, 7, 2, 3), MaxPool2d(3, 2, 1), conv_bn(64, 64, 1), conv_bn(64, 192, 3, 1, 1), MaxPool2d(3, 2, 1)]
        cin = 192
        for c in self._CFG:
            if c[0] == "a":
                layers.append(_incbn_a(cin, *c[1:]))
                cin = c[1] + c[3] + c[5] + c[7]
            else:
                layers.append(_incbn_b(cin, *c[1:]))
                cin = cin + c[2] + c[4]
        self.features = nn.Sequential(*layers)
        self.fc = Linear(cin, num_classes)
        _init(self)

    def forward(self, x):
        return self.fc(global_avg_pool(self.features(x)))


# ------------------------------------------------------------------------------------------------------ Inception-v4
class _StemV4(nn.Module):
    def __init__(self):
        super().__init__()
        self.a = nn.Sequential(conv_bn(3, 32, 3, 2), conv_bn(32, 32, 3), conv_bn(32, 64, 3, 1, 1))
        self.b = Concat(nn.MaxPool2d(3, 2), conv_bn(64, 96, 3, 2))
        self.c = Concat(nn.Sequential(conv_bn(160, 64, 1), conv_bn(64, 96, 3)),
                        nn.Sequential(conv_bn(160, 64, 1), conv_bn(64, 64, (1, 7), 1, (0, 3)), conv_bn(64, 64, (7, 1), 1, (3, 0)),
                                      conv_bn(64, 96, 3)))
        self.d = Concat(conv_bn(192, 192, 3, 2), nn.MaxPool2d(3, 2))

    def forward(self, x):
        return self.d(self.c(self.b(self.a(x))))


def _v4_a(cin=384) -> Concat:
    return Concat(nn.Sequential(nn.AvgPool2d(3, 1, 1, count_include_pad=False), conv_bn(cin, 96, 1)), conv_bn(cin, 96, 1),
                  nn.Sequential(conv_bn(cin, 64, 1), conv_bn(64, 96, 3, 1, 1)),
                  nn.Sequential(conv_bn(cin, 64, 1), conv_bn(64, 96, 3, 1, 1), conv_bn(96, 96, 3, 1, 1)))


def _v4_ra(cin=384, k=192, l=224, m=256, n=384) -> Concat:
    return Concat(nn.MaxPool2d(3, 2), conv_bn(cin, n, 3, 2),
                  nn.Sequential(conv_bn(cin, k, 1), conv_bn(k, l, 3, 1, 1), conv_bn(l, m, 3, 2)))


def _v4_b(cin=1024) -> Concat:
    return Concat(nn.Sequential(nn.AvgPool2d(3, 1, 1, count_include_pad=False), conv_bn(cin, 128, 1)), conv_bn(cin, 384, 1),
                  nn.Sequential(conv_bn(cin, 192, 1), conv_bn(192, 224, (1, 7), 1, (0, 3)), conv_bn(224, 256, (7, 1), 1, (3, 0))),
                  nn.Sequential(conv_bn(cin, 192, 1), conv_bn(192, 192, (1, 7), 1, (0, 3)), conv_bn(192, 224, (7, 1), 1, (3, 0)),
                                conv_bn(224, 224, (1, 7), 1, (0, 3)), conv_bn(224, 256, (7, 1), 1, (3, 0))))


def _v4_rb(cin=1024) -> Concat:
    return Concat(nn.MaxPool2d(3, 2), nn.Sequential(conv_bn(cin, 192, 1), conv_bn(192, 192, 3, 2)),
                  nn.Sequential(conv_bn(cin, 256, 1), conv_bn(256, 256, (1, 7), 1, (0, 3)), conv_bn(256, 320, (7, 1), 1, (3, 0)),
                                conv_bn(320, 320, 3, 2)))


class _V4C(nn.Module):
    def __init__(self, cin=1536):
        super().__init__()
        self.pool = nn.Sequential(nn.AvgPool2d(3, 1, 1, count_include_pad=False), conv_bn(cin, 256, 1))
        self.b1 = conv_bn(cin, 256, 1)
        self.b2 = conv_bn(cin, 384, 1)
        self.b2a, self.b2b = conv_bn(384, 256, (1, 3), 1, (0, 1)), conv_bn(384, 256, (3, 1), 1, (1, 0))
        self.b3 = nn.Sequential(conv_bn(cin, 384, 1), conv_bn(384, 448, (1, 3), 1, (0, 1)), conv_bn(448, 512, (3, 1), 1, (1, 0)))
        self.b3a, self.b3b = conv_bn(512, 256, (3, 1), 1, (1, 0)), conv_bn(512, 256, (1, 3), 1, (0, 1))

    def forward(self, x):
        b2, b3 = self.b2(x), self.b3(x)
        return torch.cat([self.pool(x), self.b1(x), self.b2a(b2), self.b2b(b2), self.b3a(b3), self.b3b(b3)], 1)


class InceptionV4(nn.Module):
    """299x299 input: stem (384 ch, 35x35) -> 4 x A -> reduction A (1024, 17x17) -> 7 x B -> reduction B (1536, 8x8) -> 3 x C."""

    def __init__(self, num_classes: int = 1000, dropout: float = 0.2):
        super().__init__()
        self.features = nn.Sequential(_StemV4(), *[_v4_a() for _ in range(4)], _v4_ra(), *[_v4_b() for _ in range(7)],
                                      _v4_rb(), *[_V4C() for _ in range(3)])
        self.drop = nn.Dropout(dropout)
        self.fc = Linear(1536, num_classes)
        _init(self)

    def forward(self, x):
        return self.fc(self.drop(global_avg_pool(self.features(x))))


# ----------------------------------------------------------------------------------------------- Inception-ResNet-v2
class _ResBlock(nn.Module):
    """x + scale * conv1x1(concat(branches)), then ReLU (block35 / block17 / block8)."""

    def __init__(self, cin: int, branches: list, cat: int, scale: float, relu: bool = True):
        super().__init__()
        self.branches = Concat(*branches)
        self.up = nn.Conv2d(cat, cin, 1)
        self.scale, self.relu = scale, relu

    def forward(self, x):
        y = x + self.scale * self.up(self.branches(x))
        return torch.relu_(y) if self.relu else y


def _block35(c=320, scale=0.17):
    return _ResBlock(c, [conv_bn(c, 32, 1), nn.Sequential(conv_bn(c, 32, 1), conv_bn(32, 32, 3, 1, 1)),
                         nn.Sequential(conv_bn(c, 32, 1), conv_bn(32, 48, 3, 1, 1), conv_bn(48, 64, 3, 1, 1))], 128, scale)


def _block17(c=1088, scale=0.1):
    return _ResBlock(c, [conv_bn(c, 192, 1), nn.Sequential(conv_bn(c, 128, 1), conv_bn(128, 160, (1, 7), 1, (0, 3)),
                                                           conv_bn(160, 192, (7, 1), 1, (3, 0)))], 384, scale)


def _block8(c=2080, scale=0.2, relu=True):
    return _ResBlock(c, [conv_bn(c, 192, 1), nn.Sequential(conv_bn(c, 192, 1), conv_bn(192, 224, (1, 3), 1, (0, 1)),
                                                           conv_bn(224, 256, (3, 1), 1, (1, 0)))], 448, scale, relu)


class InceptionResNetV2(nn.Module):
    def __init__(self, num_classes: int = 1000, dropout: float = 0.2):
        super().__init__()
        stem = [conv_bn(3, 32, 3, 2), conv_bn(32, 32, 3), conv_bn(32, 64, 3, 1, 1), nn.MaxPool2d(3, 2), conv_bn(64, 80, 1),
                conv_bn(80, 192, 3), nn.MaxPool2d(3, 2),
                Concat(conv_bn(192, 96, 1), nn.Sequential(conv_bn(192, 48, 1), conv_bn(48, 64, 5, 1, 2)),
                       nn.Sequential(conv_bn(192, 64, 1), conv_bn(64, 96, 3, 1, 1), conv_bn(96, 96, 3, 1, 1)),
                       nn.Sequential(nn.AvgPool2d(3, 1, 1, count_include_pad=False), conv_bn(192, 64, 1)))]          # 320
        red_a = Concat(conv_bn(320, 384, 3, 2), nn.Sequential(conv_bn(320, 256, 1), conv_bn(256, 256, 3, 1, 1), conv_bn(256, 384, 3, 2)),
                       nn.MaxPool2d(3, 2))                                                                            # 1088
        red_b = Concat(nn.Sequential(conv_bn(1088, 256, 1), conv_bn(256, 384, 3, 2)),
                       nn.Sequential(conv_bn(1088, 256, 1), conv_bn(256, 288, 3, 2)),
                       nn.Sequential(conv_bn(1088, 256, 1), conv_bn(256, 288, 3, 1, 1), conv_bn(288, 320, 3, 2)),
                       nn.MaxPool2d(3, 2))                                                                            # 2080
        self.features = nn.Sequential(*stem, *[_block35() for _ in range(10)], red_a, *[_block17() for _ in range(20)], red_b,
                                      *[_block8() for _ in range(9)], _block8(relu=False), conv_bn(2080, 1536, 1))
        self.drop = nn.Dropout(dropout)
        self.fc = Linear(1536, num_classes)
        _init(self)

    def forward(self, x):
        return self.fc(self.drop(global_avg_pool(self.features(x))))


# -------------------------------------------------------------------------------------------- SqueezeNet / DenseNet
def _fire(cin, s, e1, e3) -> nn.Sequential:
    return nn.Sequential(conv_relu(cin, s, 1), Concat(conv_relu(s, e1, 1), conv_relu(s, e3, 3, 1, 1)))


class SqueezeNet(nn.Module):
    def __init__(self, version: str = "1.0", num_classes: int = 1000):
        super().__init__()
        pool = lambda: nn.MaxPool2d(3, 2, ceil_mode=True)     # noqa: E731
        if version == "1.0":
            f = [conv_relu(3, 96, 7, 2), pool(), _fire(96, 16, 64, 64), _fire(128, 16, 64, 64), _fire(128, 32, 128, 128), pool(),
                 _fire(256, 32, 128, 128), _fire(256, 48, 192, 192), _fire(384, 48, 192, 192), _fire(384, 64, 256, 256), pool(),
                 _fire(512, 64, 256, 256)]
        elif version == "1.1":
            f = [conv_relu(3, 64, 3, 2), pool(), _fire(64, 16, 64, 64), _fire(128, 16, 64, 64), pool(), _fire(128, 32, 128, 128),
                 _fire(256, 32, 128, 128), pool(), _fire(256, 48, 192, 192), _fire(384, 48, 192, 192), _fire(384, 64, 256, 256),
                 _fire(512, 64, 256, 256)]
        else:
            raise ValueError("SqueezeNet version must be '1.0' or '1.1'")
        self.features = nn.Sequential(*f, nn.Dropout(0.5))
        self.classifier = conv_relu(512, num_classes, 1)
        _init(self, xavier=True)

    def forward(self, x):
        return torch.flatten(global_avg_pool(self.classifier(self.features(x))), 1)


class _DenseLayer(nn.Module):
    def __init__(self, cin: int, growth: int, bn_size: int):
        super().__init__()
        self.body = nn.Sequential(BatchNormReLU2d(cin, eps=1e-5, momentum=_MOM), Conv1x1(cin, bn_size * growth),
                                  BatchNormReLU2d(bn_size * growth, eps=1e-5, momentum=_MOM),
                                  nn.Conv2d(bn_size * growth, growth, 3, 1, 1, bias=False))

    def forward(self, x):
        return torch.cat([x, self.body(x)], 1)


class DenseNet(nn.Module):
    _SPEC = {121: (64, 32, [6, 12, 24, 16]), 161: (96, 48, [6, 12, 36, 24]), 169: (64, 32, [6, 12, 32, 32]),
             201: (64, 32, [6, 12, 48, 32])}

    def __init__(self, num_layers: int = 121, num_classes: int = 1000, bn_size: int = 4):
        super().__init__()
        init, growth, blocks = self._SPEC[num_layers]
        layers = [conv_bn(3, init, 7, 2, 3, eps=1e-5), MaxPool2d(3, 2, 1)]
        c = init
        for i, n in enumerate(blocks):
            for _ in range(n):
                layers.append(_DenseLayer(c, growth, bn_size))
                c += growth
            if i != len(blocks) - 1:
                layers += [BatchNormReLU2d(c, eps=1e-5, momentum=_MOM), Conv1x1(c, c // 2), nn.AvgPool2d(2, 2)]
                c //= 2
        layers.append(BatchNormReLU2d(c, eps=1e-5, momentum=_MOM))
        self.features = nn.Sequential(*layers)
        self.fc = Linear(c, num_classes)
        _init(self)

    def forward(self, x):
        return self.fc(global_avg_pool(self.features(x)))
