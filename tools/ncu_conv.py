"""Launches a fixed list of ResNet-50 convolution kernels (batch 256) for an ncu capture:
   ncu --set full --clock-control none --import-source on -k regex:kgemm -o gpurun_out/conv_ncu python tools/ncu_conv.py
Order of launches (each op twice, the second one is the interesting, warm one):
   0,1  3x3 64->64 @56 fprop      2,3  3x3 64->64 dgrad     4,5  3x3 64->64 wgrad (+reduce)
   6,7  1x1 64->256 @56 fprop     8,9  1x1 256->64 @56 fprop 10,11 3x3 256->256 @14 fprop  12,13 1x1 64->256 wgrad"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deeplearning_cfn_b200.ops import conv as conv_ops  # noqa: E402

B = int(os.environ.get("PROBE_BATCH", "256"))
dev = torch.device("cuda", 0)
cl = torch.channels_last


def mk(c, k, hw, ks, s=1):
    pad = ks // 2
    x = torch.randn(B, c, hw, hw, device=dev).bfloat16().contiguous(memory_format=cl)
    w = (torch.randn(k, c, ks, ks, device=dev) * 0.05).bfloat16().contiguous(memory_format=cl)
    ho = (hw + 2 * pad - ks) // s + 1
    gy = torch.randn(B, k, ho, ho, device=dev).bfloat16().contiguous(memory_format=cl)
    return conv_ops.ConvOps(x, w, s, pad), gy


o1, g1 = mk(64, 64, 56, 3)
o2, g2 = mk(64, 256, 56, 1)
o3, g3 = mk(256, 64, 56, 1)
o4, g4 = mk(256, 256, 14, 3)
for _ in range(2):
    o1.fprop()
for _ in range(2):
    o1.dgrad(g1)
for _ in range(2):
    o1.wgrad(g1)
for _ in range(2):
    o2.fprop()
for _ in range(2):
    o3.fprop()
for _ in range(2):
    o4.fprop()
for _ in range(2):
    o2.wgrad(g2)
torch.cuda.synchronize()
print("done")
