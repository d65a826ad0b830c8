"""Where do the aten::copy_ / layout-conversion kernels of the ResNet-50 step come from?  Profiles one step with
shapes + python stacks and prints every copy grouped by (shape, strides in/out, innermost python frame of the repo)."""
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deeplearning_cfn_b200.models import resnet, to_b200_layout  # noqa: E402
from deeplearning_cfn_b200.parallel.trainer import DataParallelTrainer  # noqa: E402
from deeplearning_cfn_b200.utils.data import synthetic_device_batch  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda", 0)
torch.backends.cudnn.benchmark = True
model = to_b200_layout(resnet(50).to(dev), torch.bfloat16).train()
tr = DataParallelTrainer(model, mode="local")
x, y = synthetic_device_batch(batch, 224, 1000, dev)
for _ in range(3):
    tr.step(x, y)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True, with_stack=True) as prof:
    tr.step(x, y)
    torch.cuda.synchronize()
groups = collections.Counter()
times = collections.Counter()
for e in prof.events():
    if e.name in ("aten::copy_", "aten::contiguous", "aten::clone", "aten::_to_copy") and e.device_time_total > 20:
        frames = [f for f in (e.stack or []) if "deeplearning_cfn_b200" in f or "bench" in f]
        where = frames[0] if frames else ((e.stack or ["?"])[0])
        key = (e.name, str(e.input_shapes)[:80], where[-90:])
        groups[key] += 1
        times[key] += e.device_time_total
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/copy_diag.txt", "w") as f:
    for k, t in sorted(times.items(), key=lambda kv: -kv[1])[:25]:
        line = f"{t/1e3:8.3f} ms  x{groups[k]:3d}  {k[0]:16s} {k[1]}  @ {k[2]}"
        print(line)
        f.write(line + "\n")
