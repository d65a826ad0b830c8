"""Python-side attribution of layout copies: wraps Tensor.contiguous and reports big non-trivial conversions with
the calling frame (custom autograd Functions run their backward in python, so they are visible here)."""
import collections
import os
import sys
import traceback

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deeplearning_cfn_b200.models import resnet, to_b200_layout  # noqa: E402
from deeplearning_cfn_b200.parallel.trainer import DataParallelTrainer  # noqa: E402
from deeplearning_cfn_b200.utils.data import synthetic_device_batch  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda", 0)
model = to_b200_layout(resnet(50).to(dev), torch.bfloat16).train()
tr = DataParallelTrainer(model, mode="local")
x, y = synthetic_device_batch(batch, 224, 1000, dev)
for _ in range(2):
    tr.step(x, y)
torch.cuda.synchronize()
seen = collections.Counter()
orig = torch.Tensor.contiguous


def wrapped(self, *a, **kw):
    mf = kw.get("memory_format", a[0] if a else torch.contiguous_format)
    if self.numel() > (1 << 20) and not self.is_contiguous(memory_format=mf):
        fr = traceback.extract_stack(limit=3)[0]
        seen[(tuple(self.shape), tuple(self.stride()), str(mf), f"{os.path.basename(fr.filename)}:{fr.lineno} {fr.name}")] += 1
    return orig(self, *a, **kw)


torch.Tensor.contiguous = wrapped
tr.step(x, y)
torch.cuda.synchronize()
torch.Tensor.contiguous = orig
for k, n in seen.most_common(20):
    print(n, k)
