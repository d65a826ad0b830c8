#!/bin/bash
mkdir -p gpurun_out
echo "=== failing test detail"; timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -x -k "fused_add" 2>&1 | grep -E "Error|assert|frac|Mismatch|Greatest|tensor" | head -12
echo "=== layout copies"; DLCFN_DEBUG_LAYOUT=1 timeout 300 python - <<'PY' 2>&1 | tail -15
import sys, torch
sys.path.insert(0,'.')
from deeplearning_cfn_b200.models import resnet, to_b200_layout
from deeplearning_cfn_b200.parallel.trainer import DataParallelTrainer
from deeplearning_cfn_b200.utils.data import synthetic_device_batch
from deeplearning_cfn_b200.ops import batchnorm
dev=torch.device('cuda',0)
m=to_b200_layout(resnet(50).to(dev), torch.bfloat16).train()
tr=DataParallelTrainer(m, mode='local')
x,y=synthetic_device_batch(32,224,1000,dev)
tr.step(x,y); batchnorm.layout_copies.clear(); tr.step(x,y); torch.cuda.synchronize()
print("BN-side layout copies:", batchnorm.layout_copies)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True) as prof:
    tr.step(x,y); torch.cuda.synchronize()
for e in prof.key_averages(group_by_input_shape=True):
    if e.key in ("aten::contiguous","aten::copy_","aten::clone") and e.device_time_total > 20:
        print(e.key, e.input_shapes, round(e.device_time_total,1), "us x", e.count)
PY
