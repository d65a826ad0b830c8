"""Small reference workloads: the TF CIFAR-10 tutorial CNN (W4/W5), LeNet (dist_lenet.py), AlexNet, MLP."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..ops.linear import Linear
from ..ops.nn import LocalResponseNorm, local_response_norm


class Cifar10CNN(nn.Module):
    """TFM/tutorials/image/cifar10/cifar10.py:188-270: 2 x (conv5x5-64, pool3/2, LRN) + FC384 + FC192 + FC10,
    24x24 crops."""

    def __init__(self, num_classes: int = 10, image_size: int = 24):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 64, 5, padding=2)
        self.conv2 = nn.Conv2d(64, 64, 5, padding=2)
        side = -(-(-(-image_size // 2)) // 2)          # two SAME-padded stride-2 pools
        self.local3 = Linear(64 * side * side, 384)
        self.local4 = Linear(384, 192)
        self.softmax_linear = Linear(192, num_classes)
        nn.init.trunc_normal_(self.conv1.weight, std=5e-2)
        nn.init.trunc_normal_(self.conv2.weight, std=5e-2)
        nn.init.constant_(self.conv2.bias, 0.1)

    def forward(self, x):
        x = F.relu(self.conv1(x))
        x = F.max_pool2d(x, 3, 2, padding=1)
        x = local_response_norm(x, 9, alpha=0.001 / 9.0, beta=0.75, k=1.0, tf_style=True)   # depth_radius 4 -> size 9
        x = F.relu(self.conv2(x))
        x = local_response_norm(x, 9, alpha=0.001 / 9.0, beta=0.75, k=1.0, tf_style=True)
        x = F.max_pool2d(x, 3, 2, padding=1)
        x = torch.flatten(x, 1)
        x = F.relu(self.local3(x))
        x = F.relu(self.local4(x))
        return self.softmax_linear(x)


class LeNet(nn.Module):
    def __init__(self, num_classes: int = 10, image_size: int = 28, in_ch: int = 1):
        super().__init__()
        self.c1, self.c2 = nn.Conv2d(in_ch, 20, 5), nn.Conv2d(20, 50, 5)
        s = ((image_size - 4) // 2 - 4) // 2
        self.f1, self.f2 = Linear(50 * s * s, 500), Linear(500, num_classes)

    def forward(self, x):
        x = F.max_pool2d(torch.tanh(self.c1(x)), 2)
        x = F.max_pool2d(torch.tanh(self.c2(x)), 2)
        x = torch.flatten(x, 1)
        return self.f2(torch.tanh(self.f1(x)))


class MLP(nn.Module):
    def __init__(self, num_classes: int = 10, in_features: int = 784, hidden=(128, 64)):
        super().__init__()
        dims = [in_features, *hidden]
        self.hidden = nn.ModuleList(Linear(a, b) for a, b in zip(dims, dims[1:]))
        self.out = Linear(dims[-1], num_classes)

    def forward(self, x):
        x = torch.flatten(x, 1)
        for l in self.hidden:
            x = F.relu(l(x))
        return self.out(x)


class AlexNet(nn.Module):
    """MX/example/image-classification/symbols/alexnet.py (one-tower variant)."""

    def __init__(self, num_classes: int = 1000, image_size: int = 224):
        super().__init__()
        self.features = nn.Sequential(
            nn.Conv2d(3, 96, 11, 4), nn.ReLU(inplace=True), LocalResponseNorm(5, alpha=1e-4, beta=0.75, k=2), nn.MaxPool2d(3, 2),
            nn.Conv2d(96, 256, 5, padding=2), nn.ReLU(inplace=True), LocalResponseNorm(5, alpha=1e-4, beta=0.75, k=2), nn.MaxPool2d(3, 2),
            nn.Conv2d(256, 384, 3, padding=1), nn.ReLU(inplace=True), nn.Conv2d(384, 384, 3, padding=1), nn.ReLU(inplace=True),
            nn.Conv2d(384, 256, 3, padding=1), nn.ReLU(inplace=True), nn.MaxPool2d(3, 2))
        with torch.no_grad():
            n = self.features(torch.zeros(1, 3, image_size, image_size)).numel()
        self.classifier = nn.Sequential(Linear(n, 4096), nn.ReLU(inplace=True), nn.Dropout(0.5), Linear(4096, 4096),
                                        nn.ReLU(inplace=True), nn.Dropout(0.5), Linear(4096, num_classes))

    def forward(self, x):
        return self.classifier(torch.flatten(self.features(x), 1))
