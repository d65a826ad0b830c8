"""Kernel-time breakdown of the flagship ResNet-50 step (torch.profiler / CUPTI), written to gpurun_out/."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deeplearning_cfn_b200.models import resnet, to_b200_layout  # noqa: E402
from deeplearning_cfn_b200.parallel.trainer import DataParallelTrainer  # noqa: E402
from deeplearning_cfn_b200.utils.data import synthetic_device_batch  # noqa: E402


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    dev = torch.device("cuda", 0)
    torch.backends.cudnn.benchmark = True
    model = to_b200_layout(resnet(50).to(dev), torch.bfloat16).train()
    tr = DataParallelTrainer(model, mode="local")
    x, y = synthetic_device_batch(batch, 224, 1000, dev)
    for _ in range(4):
        tr.step(x, y)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            tr.step(x, y)
        torch.cuda.synchronize()
    ev = [e for e in prof.key_averages() if e.device_type == torch.autograd.DeviceType.CUDA or e.self_device_time_total > 0]
    ev.sort(key=lambda e: -e.self_device_time_total)
    total = sum(e.self_device_time_total for e in ev)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/profile_resnet50.txt", "w") as f:
        f.write(f"batch {batch}: total device time over 3 steps {total/1e3:.2f} ms ({total/3e3:.2f} ms/step)\n")
        for e in ev[:60]:
            f.write(f"{e.self_device_time_total/3e3:9.3f} ms/step  {e.count/3:7.1f} calls/step  {e.key[:150]}\n")
    print(open("gpurun_out/profile_resnet50.txt").read()[:6000])


if __name__ == "__main__":
    main()
