"""Inception-v3 (MX/example/image-classification/symbols/inception-v3.py; 299x299 input) - one of the three
networks of the reference's published scaling table (README.md:303-319)."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..ops.batchnorm import BatchNormReLU2d
from ..ops.linear import Linear


class ConvBN(nn.Sequential):
    def __init__(self, cin, cout, k, stride=1, pad=0):
        super().__init__(nn.Conv2d(cin, cout, k, stride, pad, bias=False), BatchNormReLU2d(cout, eps=1e-3, relu=True))


class A(nn.Module):
    def __init__(self, cin, pool):
        super().__init__()
        self.b1 = ConvBN(cin, 64, 1)
        self.b5 = nn.Sequential(ConvBN(cin, 48, 1), ConvBN(48, 64, 5, pad=2))
        self.b3 = nn.Sequential(ConvBN(cin, 64, 1), ConvBN(64, 96, 3, pad=1), ConvBN(96, 96, 3, pad=1))
        self.bp = ConvBN(cin, pool, 1)

    def forward(self, x):
        return torch.cat([self.b1(x), self.b5(x), self.b3(x), self.bp(F.avg_pool2d(x, 3, 1, 1))], 1)


class B(nn.Module):
    def __init__(self, cin):
        super().__init__()
        self.b3 = ConvBN(cin, 384, 3, 2)
        self.bd = nn.Sequential(ConvBN(cin, 64, 1), ConvBN(64, 96, 3, pad=1), ConvBN(96, 96, 3, 2))

    def forward(self, x):
        return torch.cat([self.b3(x), self.bd(x), F.max_pool2d(x, 3, 2)], 1)


class C(nn.Module):
    def __init__(self, cin, c7):
        super().__init__()
        self.b1 = ConvBN(cin, 192, 1)
        self.b7 = nn.Sequential(ConvBN(cin, c7, 1), ConvBN(c7, c7, (1, 7), pad=(0, 3)), ConvBN(c7, 192, (7, 1), pad=(3, 0)))
        self.bd = nn.Sequential(ConvBN(cin, c7, 1), ConvBN(c7, c7, (7, 1), pad=(3, 0)), ConvBN(c7, c7, (1, 7), pad=(0, 3)),
                                ConvBN(c7, c7, (7, 1), pad=(3, 0)), ConvBN(c7, 192, (1, 7), pad=(0, 3)))
        self.bp = ConvBN(cin, 192, 1)

    def forward(self, x):
        return torch.cat([self.b1(x), self.b7(x), self.bd(x), self.bp(F.avg_pool2d(x, 3, 1, 1))], 1)


class D(nn.Module):
    def __init__(self, cin):
        super().__init__()
        self.b3 = nn.Sequential(ConvBN(cin, 192, 1), ConvBN(192, 320, 3, 2))
        self.b7 = nn.Sequential(ConvBN(cin, 192, 1), ConvBN(192, 192, (1, 7), pad=(0, 3)), ConvBN(192, 192, (7, 1), pad=(3, 0)),
                                ConvBN(192, 192, 3, 2))

    def forward(self, x):
        return torch.cat([self.b3(x), self.b7(x), F.max_pool2d(x, 3, 2)], 1)


class E(nn.Module):
    def __init__(self, cin):
        super().__init__()
        self.b1 = ConvBN(cin, 320, 1)
        self.b3 = ConvBN(cin, 384, 1)
        self.b3a, self.b3b = ConvBN(384, 384, (1, 3), pad=(0, 1)), ConvBN(384, 384, (3, 1), pad=(1, 0))
        self.bd = nn.Sequential(ConvBN(cin, 448, 1), ConvBN(448, 384, 3, pad=1))
        self.bda, self.bdb = ConvBN(384, 384, (1, 3), pad=(0, 1)), ConvBN(384, 384, (3, 1), pad=(1, 0))
        self.bp = ConvBN(cin, 192, 1)

    def forward(self, x):
        b3 = self.b3(x)
        bd = self.bd(x)
        return torch.cat([self.b1(x), self.b3a(b3), self.b3b(b3), self.bda(bd), self.bdb(bd),
                          self.bp(F.avg_pool2d(x, 3, 1, 1))], 1)


class InceptionV3(nn.Module):
    def __init__(self, num_classes: int = 1000):
        super().__init__()
        self.stem = nn.Sequential(ConvBN(3, 32, 3, 2), ConvBN(32, 32, 3), ConvBN(32, 64, 3, pad=1), nn.MaxPool2d(3, 2),
                                  ConvBN(64, 80, 1), ConvBN(80, 192, 3), nn.MaxPool2d(3, 2))
        self.mixed = nn.Sequential(A(192, 32), A(256, 64), A(288, 64), B(288), C(768, 128), C(768, 160), C(768, 160),
                                   C(768, 192), D(768), E(1280), E(2048))
        self.drop = nn.Dropout(0.5)
        self.fc1 = Linear(2048, num_classes)

    def forward(self, x):
        x = self.mixed(self.stem(x))
        return self.fc1(self.drop(torch.flatten(F.adaptive_avg_pool2d(x, 1), 1)))
