"""cuDNN's kernels for the 64-channel 3x3 convolution of ResNet-50 (batch 256), for an ncu capture next to tools/ncu_conv.py:
how many bytes does the library move through the L2 -> SM crossbar for the same layer?"""
import torch
import torch.nn.functional as F

dev = torch.device("cuda", 0)
cl = torch.channels_last
torch.backends.cudnn.benchmark = True
x = torch.randn(256, 64, 56, 56, device=dev).bfloat16().contiguous(memory_format=cl)
w = (torch.randn(64, 64, 3, 3, device=dev) * 0.05).bfloat16().contiguous(memory_format=cl)
gy = torch.randn(256, 64, 56, 56, device=dev).bfloat16().contiguous(memory_format=cl)
for _ in range(3):
    y = F.conv2d(x, w, None, 1, 1)
    torch.ops.aten.convolution_backward(gy, x, w, None, [1, 1], [1, 1], [1, 1], False, [0, 0], 1, [True, True, False])
torch.cuda.synchronize()
print("done")
