"""Small driver for `ncu`: one ResNet bottleneck-sized BatchNorm fwd+bwd, one 1x1-conv GEMM set, one softmax-CE -
each kernel family appears a few times so `-k regex:... -c N` can pick them."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deeplearning_cfn_b200.ops import native  # noqa: E402
from deeplearning_cfn_b200.ops.batchnorm import BatchNormReLU2d  # noqa: E402
from deeplearning_cfn_b200.ops.conv import conv1x1  # noqa: E402
from deeplearning_cfn_b200.ops.nn import softmax_cross_entropy  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
x = torch.randn(256, 256, 56, 56, device=dev).bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True)
bn = BatchNormReLU2d(256).to(dev).train()
w = (torch.randn(64, 256, 1, 1, device=dev) * 0.05).bfloat16().requires_grad_(True)
for _ in range(2):
    y = conv1x1(bn(x), w, 1)                      # kbn_stats, kbn_fwd_finalize, kbn_apply, kgemm (fprop)
    y.backward(torch.randn_like(y))               # kgemm dgrad + split-K wgrad, kbn_bwd_reduce/finalize/apply
    x.grad = None
    w.grad = None
logits = torch.randn(4096, 30528, device=dev).bfloat16().requires_grad_(True)
labels = torch.randint(0, 30522, (4096,), device=dev)
softmax_cross_entropy(logits, labels, num_classes=30522).backward()
a = torch.randn(8192, 8192, device=dev).bfloat16()
b = torch.randn(8192, 8192, device=dev).bfloat16()
native.gemm_bf16(a, b)
torch.cuda.synchronize()
print("ncu targets done")
