"""2-rank CPU (gloo) check of the data-parallel engine and the Horovod-style API."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from deeplearning_cfn_b200.models import resnet  # noqa: E402
from deeplearning_cfn_b200.parallel import horovod as hvd  # noqa: E402
from deeplearning_cfn_b200.parallel.trainer import DataParallelTrainer  # noqa: E402
from deeplearning_cfn_b200.utils.data import synthetic_device_batch  # noqa: E402


def main():
    hvd.init()
    assert hvd.size() == 2 and dist.get_backend() == "gloo"
    r = hvd.rank()
    t = hvd.allreduce(torch.full((5,), float(r + 1)), average=False)
    assert torch.all(t == 3.0)
    torch.manual_seed(1 + r)                       # different init on purpose: the broadcast must fix it
    model = resnet(20, 10, 32).train()
    tr = DataParallelTrainer(model, mode="auto", optimizer="sgd_mom", lr=0.05, bucket_bytes=1 << 16)
    assert tr.mode == "gloo" and len(tr.flat.buckets) > 1
    x, y = synthetic_device_batch(4, 32, 10, torch.device("cpu"), dtype=torch.float32, seed=10 + r)
    losses = [float(tr.step(x, y).detach()) for _ in range(3)]
    w = torch.cat([g.weight.reshape(-1) for g in tr.flat.groups])
    other = w.clone()
    dist.broadcast(other, src=0)
    assert torch.equal(w, other), "ranks diverged"
    # equivalence with one process seeing both batches
    if r == 0:
        torch.manual_seed(1)
        ref = resnet(20, 10, 32).train()
        rt_ = DataParallelTrainer(ref, mode="local", optimizer="sgd_mom", lr=0.05)
        xs = [synthetic_device_batch(4, 32, 10, torch.device("cpu"), dtype=torch.float32, seed=10 + k) for k in range(2)]
        for _ in range(3):
            # BatchNorm statistics are per-rank in data parallel, so average per-rank losses instead of concatenating
            rt_.flat.zero_grad()
            for xb, yb in xs:
                (rt_.loss_fn(ref(xb), yb) / 2).backward()
            rt_.synchronize_and_update()
        wr = torch.cat([g.weight.reshape(-1) for g in rt_.flat.groups])
        assert torch.allclose(w, wr, atol=2e-5), float((w - wr).abs().max())
    dist.barrier()
    if r == 0:
        print("GLOO_TRAINER_OK", losses, flush=True)
    hvd.shutdown()


if __name__ == "__main__":
    main()
