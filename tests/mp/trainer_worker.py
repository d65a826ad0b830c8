"""torchrun worker: the fused all-reduce path, the fused PS path and the NCCL baseline must produce the
same weights (up to bf16 rounding) after a few steps from the same seed."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from deeplearning_cfn_b200.models import resnet, to_b200_layout  # noqa: E402
from deeplearning_cfn_b200.parallel.trainer import DataParallelTrainer  # noqa: E402
from deeplearning_cfn_b200.utils.data import synthetic_device_batch  # noqa: E402


def run(mode, dev, rank):
    torch.manual_seed(7)
    model = to_b200_layout(resnet(20, 10, 32).to(dev), torch.bfloat16).train()
    tr = DataParallelTrainer(model, mode=mode, optimizer="sgd_mom", lr=0.05, momentum=0.9, wd=1e-4,
                             bucket_bytes=1 << 18)
    x, y = synthetic_device_batch(16, 32, 10, dev, seed=100 + rank)
    losses = []
    for _ in range(4):
        losses.append(float(tr.step(x, y)))
    tr.check()
    torch.cuda.synchronize()
    flat = torch.cat([g.weight.float().reshape(-1) for g in tr.flat.groups])
    return flat.clone(), losses


def main():
    rank, lrank = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lrank)
    dev = torch.device("cuda", lrank)
    dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.deterministic = True
    ok = True
    modes = tuple(os.environ.get("DLCFN_TEST_MODES", "nccl,allreduce,ps_sync,nccl_ps").split(","))
    res = {m: run(m, dev, rank) for m in modes}
    base, base_loss = res["nccl"]
    for m, (w, losses) in res.items():
        # identical across ranks?
        other = w.clone()
        dist.broadcast(other, src=0)
        if not torch.allclose(w, other, rtol=0, atol=0):
            print(f"[rank {rank}] {m}: weights differ across ranks", flush=True)
            ok = False
        rel = (w - base).abs().max().item() / (base.abs().max().item() + 1e-9)
        if rank == 0:
            print(f"{m}: losses {['%.4f' % l for l in losses]} max rel weight diff vs nccl {rel:.2e}", flush=True)
        if rel > 5e-2 or any(l != l for l in losses):
            ok = False
    flag = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(flag)
    if rank == 0:
        print("TRAINER_OK" if flag.item() == 0 else "TRAINER_FAILED", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 0 else 1)


if __name__ == "__main__":
    main()
