#!/bin/bash
mkdir -p gpurun_out
echo "=== resnet bench"; timeout 600 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench1_resnet.log
echo "=== bert-large bench b16"; timeout 900 python bench.py --model bert-large --batch 16 --steps 10 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench1_bert.log
echo "=== bert-large bench b32"; timeout 900 python bench.py --model bert-large --batch 32 --steps 10 --warmup 3 2>&1 | tail -2 | tee gpurun_out/bench1_bert32.log
echo "=== bert profile"; timeout 600 python - <<'PY' 2>&1 | tail -45
import torch, sys
sys.path.insert(0, '.')
from deeplearning_cfn_b200.models import get_model, to_b200_layout, bert_pretrain_loss
from deeplearning_cfn_b200.parallel.trainer import DataParallelTrainer
dev = torch.device('cuda', 0)
m = to_b200_layout(get_model('bert-large').to(dev), torch.bfloat16).train()
tr = DataParallelTrainer(m, mode='local', optimizer='adam', lr=1e-4, loss_fn=bert_pretrain_loss)
B, S = 16, 512
ids = torch.randint(0, 30522, (B, S), device=dev)
x = (ids, torch.zeros_like(ids), torch.randint(0, S, (B, 76), device=dev))
y = (torch.randint(0, 30522, (B, 76), device=dev), torch.randint(0, 2, (B,), device=dev))
for _ in range(3): tr.step(x, y)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(2): tr.step(x, y)
    torch.cuda.synchronize()
ev = sorted(prof.key_averages(), key=lambda e: -e.self_device_time_total)
tot = sum(e.self_device_time_total for e in ev)
print(f"total {tot/2e3:.2f} ms/step")
for e in ev[:35]: print(f"{e.self_device_time_total/2e3:8.3f} ms {e.count/2:6.1f}x {e.key[:120]}")
PY
