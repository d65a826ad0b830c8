"""Single-process, multi-device synchronous data parallelism ("towers").

Reference: `TFM/tutorials/image/cifar10/cifar10_multi_gpu_train.py` - one tower per GPU inside one process,
`average_gradients` = mean over towers, one shared set of variables (:101-136); the MXNet equivalent is
`split_and_load` + `kv.create('device')`.  Here each local device holds a replica, the batch is sliced with
`split_and_load`, the per-device gradients are pushed as a LIST to a `device` KVStore (reduced on one GPU by the
dlcfn elementwise-sum kernel after P2P copies; host reduce for `local`), the update runs once on the store, and
the new weights are pulled back to every replica.  Works with CPU "devices" too (the reference's unit tests fake
devices the same way).
"""
from __future__ import annotations

import argparse
import copy
import sys
import time
from datetime import datetime

import torch
import torch.nn.functional as F


def main(argv=None) -> int:
    p = argparse.ArgumentParser()
    p.add_argument("--num_gpus", type=int, default=1)
    p.add_argument("--max_steps", type=int, default=100)
    p.add_argument("--batch_size", type=int, default=128)
    p.add_argument("--kvstore", type=str, default="device")
    p.add_argument("--learning_rate", type=float, default=0.1)
    p.add_argument("--log_frequency", type=int, default=10)
    a = p.parse_args(argv)
    import deeplearning_cfn_b200.kv as kv
    from deeplearning_cfn_b200.examples.common import SplitSampler, batches, load_dataset
    from deeplearning_cfn_b200.models import Cifar10CNN
    from deeplearning_cfn_b200.parallel.utils import split_and_load

    ngpu = min(a.num_gpus, torch.cuda.device_count()) if torch.cuda.is_available() else 0
    devices = [torch.device("cuda", i) for i in range(ngpu)] or [torch.device("cpu")] * max(1, a.num_gpus)
    torch.manual_seed(0)
    master = Cifar10CNN(10, 24)
    towers = [copy.deepcopy(master).to(d) for d in devices]
    store = kv.create(a.kvstore)
    store.set_optimizer(kv.optimizer.create("sgd", learning_rate=a.learning_rate, rescale_grad=1.0 / a.batch_size))
    names = [n for n, _ in master.named_parameters()]
    plists = [dict(t.named_parameters()) for t in towers]
    for i, n in enumerate(names):
        store.init(i, plists[0][n].data)
    train = load_dataset("cifar10", None, True, 24, 2048)
    sampler = SplitSampler(len(train), 1, 0)
    step, t0 = 0, time.time()
    while step < a.max_steps:
        for x, y in batches(train, a.batch_size, sampler, torch.device("cpu"), channels_last=False):
            xs, ys = split_and_load(x, devices), split_and_load(y, devices)
            losses = []
            for t, xb, yb in zip(towers, xs, ys):
                t.zero_grad(set_to_none=True)
                l = F.cross_entropy(t(xb), yb, reduction="sum")
                l.backward()
                losses.append(float(l.detach()))
            for i, n in enumerate(names):
                store.push(i, [pl[n].grad for pl in plists], priority=-i)      # list of per-device gradients
                store.pull(i, out=[pl[n].data for pl in plists], priority=-i)
            step += 1
            if step % a.log_frequency == 0:
                dt = (time.time() - t0) / a.log_frequency
                t0 = time.time()
                print(f"{datetime.now()}: step {step}, loss = {sum(losses) / a.batch_size:.2f} "
                      f"({a.batch_size / dt:.1f} examples/sec; {dt:.3f} sec/batch)", flush=True)
            if step >= a.max_steps:
                break
    w0 = torch.cat([p.detach().cpu().reshape(-1) for p in towers[0].parameters()])
    for t in towers[1:]:
        assert torch.allclose(w0, torch.cat([p.detach().cpu().reshape(-1) for p in t.parameters()])), "towers diverged"
    print("TOWERS_IN_SYNC", flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
