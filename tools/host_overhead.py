"""Host-side enqueue cost of one training step: the same model and step at a tiny batch, where the GPU work is
negligible, so ms/step ~ the time python + launches need per step; cProfile of 5 steps shows where it goes."""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deeplearning_cfn_b200.models import resnet, to_b200_layout  # noqa: E402
from deeplearning_cfn_b200.parallel.trainer import DataParallelTrainer  # noqa: E402
from deeplearning_cfn_b200.utils.data import synthetic_device_batch  # noqa: E402

dev = torch.device("cuda", 0)
torch.backends.cudnn.benchmark = True
model = to_b200_layout(resnet(50).to(dev), torch.bfloat16).train()
tr = DataParallelTrainer(model, mode="local")
x, y = synthetic_device_batch(8, 64, 1000, dev)
for _ in range(5):
    tr.step(x, y)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    tr.step(x, y)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {1e3 * (t1 - t0) / 20:.2f} ms/step (drain {1e3 * (t2 - t1):.2f} ms)")
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    tr.step(x, y)
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("cumulative")
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/host_overhead.txt", "w") as f:
    st.stream = f
    st.print_stats(45)
print(open("gpurun_out/host_overhead.txt").read()[-6000:])
