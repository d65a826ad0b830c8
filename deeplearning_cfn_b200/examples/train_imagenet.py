"""ImageNet-style training / synthetic benchmark with the bucketed data-parallel engine.

Reference: `MX/example/image-classification/train_imagenet.py` + `common/fit.py` (`add_fit_args` :77-145:
--network --num-layers --gpus --kv-store --num-epochs --lr --lr-factor --lr-step-epochs --optimizer --mom --wd
--batch-size --disp-batches --model-prefix --save-period --load-epoch --dtype --gc-type --gc-threshold
--warmup-epochs ...; `--benchmark 1` => synthetic data, common/data.py:132-138; SGD momentum 0.9 wd 1e-4
multi_precision, fit.py:229-238; per-rank log prefix `Node[rank]`, fit.py:173-174; resume recomputes the LR
schedule, fit.py:29-55).  `--kv-store` picks the communication path:
    device / local      single process                         (one GPU per process here)
    dist_sync[_device]  fused parameter-server kernel, servers sharded over all GPUs
    horovod / allreduce fused bucketed all-reduce
    nccl                NCCL all-reduce baseline
"""
from __future__ import annotations

import argparse
import logging
import os
import sys
import time

import torch


def add_fit_args(p: argparse.ArgumentParser) -> None:
    t = p.add_argument_group("Training", "model training")
    t.add_argument("--network", type=str, default="resnet")
    t.add_argument("--num-layers", type=int, default=50)
    t.add_argument("--gpus", type=str, default=None)
    t.add_argument("--kv-store", type=str, default="device")
    t.add_argument("--num-epochs", type=int, default=1)
    t.add_argument("--lr", type=float, default=0.1)
    t.add_argument("--lr-factor", type=float, default=0.1)
    t.add_argument("--lr-step-epochs", type=str, default="30,60")
    t.add_argument("--initializer", type=str, default="default")
    t.add_argument("--optimizer", type=str, default="sgd")
    t.add_argument("--mom", type=float, default=0.9)
    t.add_argument("--wd", type=float, default=0.0001)
    t.add_argument("--batch-size", type=int, default=128, help="the batch size per worker")
    t.add_argument("--disp-batches", type=int, default=20)
    t.add_argument("--model-prefix", type=str, default=None)
    t.add_argument("--save-period", type=int, default=1)
    t.add_argument("--monitor", type=int, default=0)
    t.add_argument("--load-epoch", type=int, default=None)
    t.add_argument("--top-k", type=int, default=0)
    t.add_argument("--loss", type=str, default="")
    t.add_argument("--test-io", type=int, default=0)
    t.add_argument("--dtype", type=str, default="bfloat16")
    t.add_argument("--gc-type", type=str, default="none")
    t.add_argument("--gc-threshold", type=float, default=0.5)
    t.add_argument("--macrobatch-size", type=int, default=0)
    t.add_argument("--warmup-epochs", type=int, default=5)
    t.add_argument("--warmup-strategy", type=str, default="linear")
    t.add_argument("--profile-worker-suffix", type=str, default="")
    t.add_argument("--profile-server-suffix", type=str, default="")


def parse(argv=None):
    p = argparse.ArgumentParser(description="train imagenet-1k", formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    add_fit_args(p)
    d = p.add_argument_group("Data")
    d.add_argument("--benchmark", type=int, default=0, help="if 1, then feed the network with synthetic data")
    d.add_argument("--num-classes", type=int, default=1000)
    d.add_argument("--num-examples", type=int, default=1281167)
    d.add_argument("--image-shape", type=str, default="3,224,224")
    d.add_argument("--data-train", type=str, default="", help="RecordIO file (.rec with its .idx) of training images")
    d.add_argument("--data-nthreads", type=int, default=4, help="decoder threads of the native record pipeline")
    d.add_argument("--rand-crop", type=int, default=0)
    d.add_argument("--rand-mirror", type=int, default=0)
    d.add_argument("--rgb-mean", type=str, default="123.68,116.779,103.939")
    d.add_argument("--rgb-std", type=str, default="58.393,57.12,57.375")
    d.add_argument("--num-batches", type=int, default=0, help="stop an epoch after this many batches (benchmark runs)")
    d.add_argument("--cuda-graph", type=int, default=0,
                   help="1: replay the training step as one CUDA graph (experimental; the counterpart of the reference's bulk "
                        "execution, MXNET_EXEC_BULK_EXEC_TRAIN) - pays off for launch-bound small batches")
    return p.parse_args(argv)


def main(argv=None) -> int:
    a = parse(argv)
    epoch_seed = 0
    from deeplearning_cfn_b200.models import get_model, to_b200_layout
    from deeplearning_cfn_b200.parallel import horovod as hvd
    from deeplearning_cfn_b200.parallel.trainer import DataParallelTrainer
    from deeplearning_cfn_b200.utils import profiler
    from deeplearning_cfn_b200.utils.callback import Speedometer, latest_checkpoint, load_checkpoint, save_checkpoint
    from deeplearning_cfn_b200.utils.data import synthetic_device_batch

    hvd.init()
    rank, world = hvd.rank(), hvd.size()
    head = "%(asctime)-15s " + (f"Node[{rank}] " if world > 1 else "") + "%(message)s"
    logging.basicConfig(level=logging.INFO, format=head, force=True)
    use_gpu = torch.cuda.is_available()
    dev = torch.device("cuda", torch.cuda.current_device()) if use_gpu else torch.device("cpu")
    size = int(a.image_shape.split(",")[-1])
    try:
        model = get_model(a.network, a.num_classes, size, num_layers=a.num_layers)
    except ValueError as e:           # unknown --network / a depth the family does not have: a usage error
        raise SystemExit(f"train_imagenet: {e}")
    if a.initializer not in ("default", ""):
        # fit.py:200-215: 'xavier' -> Xavier(gaussian, in, 2), 'msra' -> MSRAPrelu, else the named mx.init class
        from deeplearning_cfn_b200.kv import initializer as init_mod
        spec = {"xavier": ("xavier", dict(rnd_type="gaussian", factor_type="in", magnitude=2)),
                "msra": ("msraprelu", {}), "orthogonal": ("orthogonal", {}), "normal": ("normal", {}),
                "uniform": ("uniform", {}), "one": ("one", {}), "zero": ("zero", {})}.get(a.initializer, (a.initializer, {}))
        init_mod.create(spec[0], **spec[1]).init_module(model)
    model = model.to(dev)
    dtype = torch.bfloat16 if (use_gpu and a.dtype in ("float16", "bfloat16")) else torch.float32
    if dtype != torch.float32:
        model = to_b200_layout(model, dtype)
    model.train()
    kvs = a.kv_store.lower()
    if world == 1:
        mode = "local"
    elif kvs in ("horovod", "allreduce", "device", "local"):
        mode = "auto"
    elif kvs == "nccl":
        mode = "nccl"
    elif "dist" in kvs:
        mode = "ps_sync" if use_gpu else "gloo"
    else:
        raise SystemExit(f"unknown --kv-store {a.kv_store}")
    opt_kind = {"sgd": "sgd_mom" if a.mom else "sgd", "adam": "adam", "nag": "sgd_mom"}.get(a.optimizer, "sgd_mom")
    tr = DataParallelTrainer(model, mode=mode, optimizer=opt_kind, lr=a.lr, momentum=a.mom, wd=a.wd)
    begin = 0
    if a.model_prefix and a.load_epoch is not None:
        load_checkpoint(a.model_prefix, a.load_epoch, model, tr, rank)
        begin = a.load_epoch
    elif a.model_prefix and a.load_epoch is None and latest_checkpoint(a.model_prefix) and os.environ.get("DLCFN_AUTO_RESUME") == "1":
        begin = latest_checkpoint(a.model_prefix)
        load_checkpoint(a.model_prefix, begin, model, tr, rank)
    steps = [int(s) for s in a.lr_step_epochs.split(",") if s.strip()]
    lr = a.lr * a.lr_factor ** sum(1 for s in steps if begin >= s)          # resumed runs re-derive the schedule
    tr.set_learning_rate(lr)
    if a.profile_worker_suffix:
        profiler.set_config(filename=f"rank{rank}_{a.profile_worker_suffix}")
        profiler.set_state("run")
    x, y = synthetic_device_batch(a.batch_size, size, a.num_classes, dev, dtype=dtype if use_gpu else torch.float32)
    epoch_batches = a.num_batches or max(1, a.num_examples // (a.batch_size * world))
    loader = None
    if a.data_train and not a.benchmark:
        from deeplearning_cfn_b200.utils.data import RecordBatchLoader
        from deeplearning_cfn_b200.utils.recordio import image_record_iter
        c, hh, ww = (int(v) for v in a.image_shape.split(","))
        rec_it = image_record_iter(a.data_train, a.batch_size, (c, hh, ww), shuffle=True, num_parts=hvd.size(),
                                       part_index=rank, preprocess_threads=a.data_nthreads, rand_crop=bool(a.rand_crop),
                                       rand_mirror=bool(a.rand_mirror), round_batch=False, seed=epoch_seed)
        mean = tuple(float(v) / 255.0 for v in a.rgb_mean.split(","))
        std = tuple(float(v) / 255.0 for v in a.rgb_std.split(","))
        loader = RecordBatchLoader(rec_it, dev, mean, std, dtype if use_gpu else torch.float32)
        logging.info("training data: %s, %d samples in this shard, %d batches/epoch", a.data_train, rec_it.num_samples, len(rec_it))
    speed = Speedometer(a.batch_size, a.disp_batches)
    step_fn = None
    for epoch in range(begin, a.num_epochs):
        if epoch in steps and epoch > begin:          # steps up to `begin` are already in the resumed rate (fit.py:29-55)
            lr *= a.lr_factor
            tr.set_learning_rate(lr)
            logging.info("Update[%d]: Change learning rate to %0.5e", epoch, lr)
        tic = time.time()
        if a.cuda_graph and step_fn is None:
            step_fn = tr.capture(x, y)
            logging.info("training step: %r", step_fn)
        run = step_fn if step_fn is not None else tr.step
        if loader is not None:       # real data: native record pipeline, sharded by rank
            for i, (xb, yb) in enumerate(loader):
                if a.num_batches and i >= a.num_batches:
                    break
                with profiler.scope("batch", "train", device=use_gpu):
                    loss = run(xb, yb)
                speed(epoch, i, None)
        else:
            for i in range(epoch_batches):
                with profiler.scope("batch", "train", device=use_gpu):
                    loss = run(x, y)
                speed(epoch, i, None)
        if use_gpu:
            torch.cuda.synchronize()
        logging.info("Epoch[%d] Train-loss=%f", epoch, float(loss.detach()))
        logging.info("Epoch[%d] Time cost=%.3f", epoch, time.time() - tic)
        if a.model_prefix and (epoch + 1) % a.save_period == 0:
            save_checkpoint(a.model_prefix, epoch + 1, model, tr, rank)
    tr.check()
    if a.profile_worker_suffix:
        profiler.dump()
    hvd.shutdown()
    return 0


if __name__ == "__main__":
    sys.exit(main())
