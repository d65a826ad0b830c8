"""VGG 11/13/16/19 (+BN) - the model of the reference's README example (`--model vgg11`, README.md:139; gluon
model zoo `vgg.py`: conv stacks [1,1,2,2,2]/[2,2,2,2,2]/[2,2,3,3,3]/[2,2,4,4,4] x [64,128,256,512,512],
two Dense-4096 + dropout 0.5, Dense-classes)."""
from __future__ import annotations

import torch
import torch.nn as nn

from ..ops.batchnorm import BatchNormReLU2d
from ..ops.linear import Linear

_SPEC = {11: [1, 1, 2, 2, 2], 13: [2, 2, 2, 2, 2], 16: [2, 2, 3, 3, 3], 19: [2, 2, 4, 4, 4]}
_FILTERS = [64, 128, 256, 512, 512]


class VGG(nn.Module):
    def __init__(self, num_layers: int = 11, num_classes: int = 1000, batch_norm: bool = False, image_size: int = 224):
        super().__init__()
        layers, cin = [], 3
        for n, f in zip(_SPEC[num_layers], _FILTERS):
            for _ in range(n):
                layers.append(nn.Conv2d(cin, f, 3, padding=1, bias=not batch_norm))
                layers.append(BatchNormReLU2d(f, relu=True) if batch_norm else nn.ReLU(inplace=True))
                cin = f
            layers.append(nn.MaxPool2d(2, 2))
        self.features = nn.Sequential(*layers)
        side = max(1, image_size // 32)
        self.classifier = nn.Sequential(Linear(512 * side * side, 4096), nn.ReLU(inplace=True), nn.Dropout(0.5),
                                        Linear(4096, 4096), nn.ReLU(inplace=True), nn.Dropout(0.5),
                                        Linear(4096, num_classes))
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)

    def forward(self, x):
        return self.classifier(torch.flatten(self.features(x), 1))
