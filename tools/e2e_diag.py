"""Which part of the end-to-end loop costs time?  Same trainer, 20 steps each:
  A device-resident input, no read-back          (= bench's device-timed loop)
  B device-resident input, pipelined loss read-back
  C prefetcher (pinned H2D + normalise), no read-back
  D prefetcher + pipelined read-back              (= bench's e2e loop)
  E prefetcher input, but the SAME slot every step (no copy in flight during the step)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deeplearning_cfn_b200.models import resnet, to_b200_layout  # noqa: E402
from deeplearning_cfn_b200.parallel.trainer import DataParallelTrainer  # noqa: E402
from deeplearning_cfn_b200.utils.data import DevicePrefetcher, SyntheticImageSource, synthetic_device_batch  # noqa: E402

dev = torch.device("cuda", 0)
model = to_b200_layout(resnet(50).to(dev), torch.bfloat16).train()
tr = DataParallelTrainer(model, mode="local", optimizer="sgd_mom", lr=0.1, momentum=0.9, wd=1e-4)
x, y = synthetic_device_batch(256, 224, 1000, dev)
src = SyntheticImageSource(256, 224, 1000)
pre = DevicePrefetcher(src, dev)
host = [torch.zeros(1).pin_memory() for _ in range(2)]
evs = [torch.cuda.Event(), torch.cuda.Event()]
for _ in range(5):
    tr.step(x, y)


def run(name, use_pre, readback, same_slot=False):
    xb0, yb0 = pre.next()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(20):
        if use_pre and not same_slot:
            xb, yb = pre.next()
        elif use_pre:
            xb, yb = xb0, yb0
        else:
            xb, yb = x, y
        loss = tr.step(xb, yb)
        if readback:
            k = i & 1
            host[k].copy_(loss.detach().float().reshape(1), non_blocking=True)
            evs[k].record()
            if i:
                evs[k ^ 1].synchronize()
    b.record()
    torch.cuda.synchronize()
    print(f"{name}: {a.elapsed_time(b) / 20:.3f} ms/step", flush=True)


run("A device input, no readback     ", False, False)
run("B device input, readback        ", False, True)
run("C prefetcher, no readback       ", True, False)
run("D prefetcher + readback         ", True, True)
run("E prefetcher tensor, fixed slot ", True, False, True)
run("A again                         ", False, False)
