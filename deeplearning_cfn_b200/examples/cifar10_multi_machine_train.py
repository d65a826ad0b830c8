"""Asynchronous parameter-server CIFAR-10 training with explicit ps/worker host lists.

Reference: `examples/tensorflow/cifar10_multi_machine_train.py` - between-graph replication, ASYNCHRONOUS PS:
flags --max_steps --ps_hosts --worker_hosts --job_name --task_index --train_dir (+ --batch_size --data_dir
--use_fp16 from cifar10.py) (:21-34); a `ps` process only serves (:72-73); workers train the tutorial CNN with
variables placed on the PS tasks, the chief (task 0) checkpoints every 60 s and restores on start
(:97-107); `_LoggerHook` prints loss and examples/sec every 10 steps (:38-60).  Started on every node by
the scripts that `generate-trainer` writes.

Training recipe of the tutorial model (TFM/tutorials/image/cifar10/cifar10.py:325-378): SGD with an exponentially
decayed learning rate (0.1, x0.1 every 350 epochs, staircase) evaluated ON the parameter servers, and an
exponential moving average (0.9999) of the trained variables kept by the chief and written into the checkpoint.
`--sync_replicas` switches to the synchronous PS recipe of the reference's other TF trainers
(inception_distributed_train.py:197, slim train_image_classifier.py:526 - `SyncReplicasOptimizer` with
`--replicas_to_aggregate`, backup workers allowed): see parallel/sync_replicas.py.

Mapping: ps tasks become dlcfn parameter-server nodes (csrc/rt), ps task 0 also hosts the scheduler at
port+1 of its address, workers use `kv.create('dist_async')` - every push is applied on arrival, no barrier.
"""
from __future__ import annotations

import argparse
import os
import subprocess
import sys
import time
from datetime import datetime

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))   # runnable as a plain script


def parse(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--max_steps", type=int, default=1000000)
    p.add_argument("--ps_hosts", type=str, default="")
    p.add_argument("--worker_hosts", type=str, default="")
    p.add_argument("--job_name", type=str, default="", help="One of 'ps', 'worker'")
    p.add_argument("--task_index", type=int, default=0)
    p.add_argument("--train_dir", type=str, default="/tmp/cifar10_train")
    p.add_argument("--batch_size", type=int, default=128)
    p.add_argument("--data_dir", type=str, default="/tmp/cifar10_data")
    p.add_argument("--use_fp16", action="store_true")
    p.add_argument("--log_frequency", type=int, default=10)
    p.add_argument("--save_checkpoint_secs", type=float, default=60.0)
    p.add_argument("--learning_rate", type=float, default=0.1)
    p.add_argument("--learning_rate_decay_factor", type=float, default=0.1)
    p.add_argument("--num_epochs_per_decay", type=float, default=350.0)
    p.add_argument("--moving_average_decay", type=float, default=0.9999)
    p.add_argument("--sync_replicas", action="store_true", help="synchronous PS with token queue (SyncReplicasOptimizer)")
    p.add_argument("--replicas_to_aggregate", type=int, default=None, help="default: number of workers")
    return p.parse_args(argv)


def _split(hosts: str):
    out = []
    for h in hosts.split(","):
        if h.strip():
            name, port = h.strip().rsplit(":", 1)
            out.append((name, int(port)))
    return out


def main(argv=None) -> int:
    a = parse(argv)
    ps, workers = _split(a.ps_hosts), _split(a.worker_hosts)
    if not ps or not workers:
        raise SystemExit("--ps_hosts and --worker_hosts are required")
    from deeplearning_cfn_b200.launch.hostfile import resolve_alias
    root_host, root_port = resolve_alias(ps[0][0]), ps[0][1] + 1
    os.environ.update(DMLC_PS_ROOT_URI=root_host, DMLC_PS_ROOT_PORT=str(root_port), DMLC_NUM_SERVER=str(len(ps)),
                      DMLC_NUM_WORKER=str(len(workers)))
    sched = None
    if a.job_name == "ps":
        if a.task_index == 0:   # the cluster's rendezvous point lives next to ps task 0
            env = dict(os.environ, DMLC_ROLE="scheduler")
            sched = subprocess.Popen([sys.executable, "-c", "import deeplearning_cfn_b200.kv"], env=env)
        os.environ["DMLC_ROLE"] = "server"
        os.environ["DMLC_NODE_HOST"] = resolve_alias(ps[a.task_index][0])
        import deeplearning_cfn_b200.kv  # noqa: F401  serves until the chief stops the servers, then exits
        return 0
    if a.job_name != "worker":
        raise SystemExit("--job_name must be ps or worker")
    os.environ["DMLC_ROLE"] = "worker"
    os.environ["DMLC_NODE_HOST"] = resolve_alias(workers[a.task_index][0])

    import torch
    import torch.nn.functional as F
    import deeplearning_cfn_b200.kv as kv
    from deeplearning_cfn_b200.examples.common import SplitSampler, batches, load_dataset
    from deeplearning_cfn_b200.models import Cifar10CNN

    use_gpu = torch.cuda.is_available() and os.environ.get("CUDA_VISIBLE_DEVICES", "x") != ""
    dev = torch.device("cuda", 0) if use_gpu else torch.device("cpu")
    if not use_gpu:
        # CPU workers that share a host share its cores: with every process running a full intra-op thread pool a step of
        # this small network takes 50x longer than with the pools sized to the share (measured: 376 ms vs 8 ms backward)
        host_of = lambda h: h[0] if isinstance(h, (tuple, list)) else str(h).split(":")[0]      # noqa: E731
        mine = host_of(workers[a.task_index])
        local = sum(1 for h in list(workers) + list(ps) if host_of(h) == mine)
        torch.set_num_threads(max(1, (os.cpu_count() or 1) // max(1, local)))
    torch.manual_seed(0)
    net = Cifar10CNN(10, 24).to(dev)
    store = kv.create("dist_async")
    is_chief = a.task_index == 0
    ckpt = os.path.join(a.train_dir, "model.ckpt")
    os.makedirs(a.train_dir, exist_ok=True)
    step0 = 0
    if is_chief and os.path.exists(ckpt):          # chief restores, then seeds the servers with the restored values
        sd = torch.load(ckpt, map_location="cpu")
        net.load_state_dict(sd["model"])
        step0 = sd["global_step"]
        print(f"restored {ckpt} at step {step0}", flush=True)
    params = [p for p in net.parameters()]
    train = load_dataset("cifar10", a.data_dir, True, 24, 4096)
    # lr = 0.1 * 0.1^floor(step / decay_steps), decay_steps = batches_per_epoch * 350 (cifar10.py:338-346); the schedule
    # counts updates on the servers, where the optimizer runs
    decay_steps = max(1, int(len(train) / a.batch_size * a.num_epochs_per_decay))
    sched_lr = kv.optimizer.FactorScheduler(step=decay_steps, factor=a.learning_rate_decay_factor)
    sgd = kv.optimizer.create("sgd", learning_rate=a.learning_rate, rescale_grad=1.0 / a.batch_size, lr_scheduler=sched_lr)
    sync = None
    if a.sync_replicas:
        from deeplearning_cfn_b200.parallel.sync_replicas import SyncReplicasTrainer
        sync = SyncReplicasTrainer(store, params, sgd, None, replicas_to_aggregate=a.replicas_to_aggregate or len(workers))
    else:
        for i, p in enumerate(params):
            store.init(i, p.data)
        store.set_optimizer(sgd)
        for i, p in enumerate(params):
            store.pull(i, out=p.data)
    ema = [p.detach().clone() for p in params] if is_chief else None      # shadow variables (cifar10.py:368-372)
    if train.x.shape[-1] != 24:
        train.x = train.x[..., 4:28, 4:28].contiguous()       # the tutorial trains on 24x24 crops
    sampler = SplitSampler(len(train), len(workers), a.task_index)
    step, last_save, t0 = step0, time.time(), time.time()
    while step < a.max_steps:
        for x, y in batches(train, a.batch_size, sampler, dev, channels_last=False):
            net.zero_grad(set_to_none=True)
            loss = F.cross_entropy(net(x), y, reduction="sum")
            loss.backward()
            if sync is not None:
                sync.step()                                     # push + wait for the token of the next global step
            else:
                for i, p in enumerate(params):
                    store.push(i, p.grad, priority=-i)          # applied on the servers as they arrive (async)
                for i, p in enumerate(params):
                    store.pull(i, out=p.data, priority=-i)
            step += 1
            if ema is not None:
                d = min(a.moving_average_decay, (1.0 + step) / (10.0 + step))   # tf.train.ExponentialMovingAverage(num_updates)
                with torch.no_grad():
                    for e, p in zip(ema, params):
                        e.mul_(d).add_(p.detach(), alpha=1.0 - d)
            if step % a.log_frequency == 0:
                dt = time.time() - t0
                t0 = time.time()
                eps = a.log_frequency * a.batch_size / dt
                print(f"{datetime.now()}: step {step}, loss = {float(loss.detach()) / a.batch_size:.2f} ({eps:.1f} examples/sec; "
                      f"{dt / a.log_frequency:.3f} sec/batch)", flush=True)
            if is_chief and time.time() - last_save >= a.save_checkpoint_secs:
                torch.save({"model": net.state_dict(), "global_step": step, "moving_average": [e.cpu() for e in ema]}, ckpt)
                last_save = time.time()
            if step >= a.max_steps:
                break
    if is_chief:
        torch.save({"model": net.state_dict(), "global_step": step, "moving_average": [e.cpu() for e in ema]}, ckpt)
    store.close()
    if sched is not None:
        sched.wait(30)
    return 0


if __name__ == "__main__":
    sys.exit(main())
