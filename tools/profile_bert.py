"""Kernel-time breakdown of the BERT-large pre-training step (torch.profiler / CUPTI), written to gpurun_out/profile_bert.txt."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deeplearning_cfn_b200.models import bert_pretrain_loss, get_model, to_b200_layout  # noqa: E402
from deeplearning_cfn_b200.parallel.trainer import DataParallelTrainer  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    B, S, npred = 16, 512, 76
    model = to_b200_layout(get_model("bert-large", max_seq=S).to(dev), torch.bfloat16).train()
    model.padded_logits = True
    tr = DataParallelTrainer(model, mode="local", optimizer="adam", lr=1e-4, wd=0.01, loss_fn=bert_pretrain_loss)
    g = torch.Generator().manual_seed(0)
    ids = torch.randint(0, 30522, (B, S), generator=g).to(dev)
    x = (ids, torch.zeros_like(ids), torch.randint(0, S, (B, npred), generator=g).to(dev))
    y = (torch.randint(0, 30522, (B, npred), generator=g).to(dev), torch.randint(0, 2, (B,), generator=g).to(dev))
    for _ in range(3):
        tr.step(x, y)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(2):
            tr.step(x, y)
        torch.cuda.synchronize()
    ev = sorted(prof.key_averages(), key=lambda e: -e.self_device_time_total)
    total = sum(e.self_device_time_total for e in ev)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/profile_bert.txt", "w") as f:
        f.write(f"bert-large batch {B} x {S}: total device time over 2 steps {total/1e3:.2f} ms ({total/2e3:.2f} ms/step)\n")
        for e in ev[:60]:
            f.write(f"{e.self_device_time_total/2e3:9.3f} ms/step  {e.count/2:7.1f} calls/step  {e.key[:150]}\n")


if __name__ == "__main__":
    main()
