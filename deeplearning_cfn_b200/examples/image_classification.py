"""Image classification with the KVStore / Trainer API - the reference's headline example.

Reference: `MX/example/gluon/image_classification.py` (CLI :49-102; train loop :214-229:
`split_and_load` -> record/backward -> `trainer.step(batch_size)`; kvstore from `--kvstore`, default `device`;
checkpoint every `--save-frequency` + best :180-189; `--resume` :126-127), launched as
`launch.py -n $DEEPLEARNING_WORKERS_COUNT -H $DEEPLEARNING_WORKERS_PATH python image_classification.py
--dataset cifar10 --model vgg11 --epochs 1 --kvstore dist_device_sync` (reference README.md:139).

Two execution engines:
  --engine kvstore   (default) per-key push/pull through `kv.Trainer`, exactly the reference's structure
  --engine bucketed  the fused bucketed data-parallel engine (`parallel.DataParallelTrainer`) - the fast path
"""
from __future__ import annotations

import argparse
import logging
import os
import sys
import time

import torch
import torch.nn.functional as F


def parse(argv=None):
    p = argparse.ArgumentParser(description="Train a model for image classification.")
    p.add_argument("--dataset", type=str, default="cifar10", help="dataset to use. options are mnist, cifar10, caltech101, imagenet and dummy.")
    p.add_argument("--data-dir", type=str, default="", help="training directory of imagenet images, contains train/val subdirs.")
    p.add_argument("--num-worker", "-j", dest="num_workers", default=4, type=int)
    p.add_argument("--batch-size", type=int, default=32, help="training batch size per device (CPU/GPU).")
    p.add_argument("--gpus", type=str, default="", help='ordinates of gpus to use, can be "0,1,2" or empty for cpu only.')
    p.add_argument("--epochs", type=int, default=120)
    p.add_argument("--lr", type=float, default=0.1)
    p.add_argument("--momentum", type=float, default=0.9)
    p.add_argument("--wd", type=float, default=0.0001)
    p.add_argument("--seed", type=int, default=123)
    p.add_argument("--mode", type=str, default="imperative", help="mode in which to train the model. options are symbolic, imperative, hybrid")
    p.add_argument("--model", type=str, required=True, help="type of model to use. see models.get_model for options.")
    p.add_argument("--use_thumbnail", action="store_true")
    p.add_argument("--batch-norm", action="store_true")
    p.add_argument("--use-pretrained", action="store_true")
    p.add_argument("--prefix", default="", type=str, help="path to checkpoint prefix, default is current working dir")
    p.add_argument("--start-epoch", default=0, type=int)
    p.add_argument("--resume", type=str, default="", help="path to saved weight where you want resume")
    p.add_argument("--lr-factor", default=0.1, type=float)
    p.add_argument("--lr-steps", default="30,60,90", type=str)
    p.add_argument("--dtype", default="float32", type=str, help="float32 | float16 | bfloat16")
    p.add_argument("--save-frequency", default=10, type=int)
    p.add_argument("--kvstore", type=str, default="device")
    p.add_argument("--log-interval", type=int, default=50)
    p.add_argument("--profile", action="store_true")
    p.add_argument("--builtin-profiler", type=int, default=0)
    p.add_argument("--engine", type=str, default="kvstore", choices=["kvstore", "bucketed"])
    p.add_argument("--num-samples", type=int, default=2048, help="size of the synthetic dataset")
    p.add_argument("--max-batches", type=int, default=0, help="stop every epoch after this many batches (0 = all)")
    return p.parse_args(argv)


def main(argv=None) -> int:
    opt = parse(argv)
    logging.basicConfig(level=logging.INFO, format="%(message)s")
    import deeplearning_cfn_b200.kv as kv   # servers / scheduler never return from this import
    from deeplearning_cfn_b200.examples.common import Accuracy, DATASETS, SplitSampler, batches, load_dataset
    from deeplearning_cfn_b200.models import get_model, to_b200_layout
    from deeplearning_cfn_b200.utils import profiler
    from deeplearning_cfn_b200.utils.callback import Speedometer, load_checkpoint, save_checkpoint

    torch.manual_seed(opt.seed)
    use_gpu = (opt.gpus != "" or "CUDA_VISIBLE_DEVICES" in os.environ) and torch.cuda.is_available()
    device = torch.device("cuda", int(opt.gpus.split(",")[0]) if (use_gpu and opt.gpus) else 0) if use_gpu else torch.device("cpu")
    if use_gpu:
        torch.cuda.set_device(device)
    dtype = {"float32": torch.float32, "float16": torch.bfloat16, "bfloat16": torch.bfloat16}[opt.dtype]
    ch, side, classes = DATASETS.get(opt.dataset, DATASETS["dummy"])
    if opt.use_thumbnail and opt.dataset in ("imagenet", "dummy"):
        side = 32
    name = opt.model + ("_bn" if opt.batch_norm and opt.model.startswith("vgg") and "bn" not in opt.model else "")
    try:
        net = get_model(name, classes, side, in_ch=ch).to(device)
    except ValueError as e:           # unknown --model: a usage error
        raise SystemExit(f"image_classification: {e}")
    if dtype != torch.float32 and use_gpu:
        net = to_b200_layout(net, dtype)
    if opt.resume:
        from deeplearning_cfn_b200.utils.callback import load_params_file
        load_params_file(opt.resume, net)          # periodic checkpoints and the -best file use different key layouts

    # ---- distributed identity comes from the kvstore (rank / num_workers), as in the reference
    bucketed = opt.engine == "bucketed"
    if bucketed:
        from deeplearning_cfn_b200.parallel import horovod as hvd
        from deeplearning_cfn_b200.parallel.trainer import DataParallelTrainer
        hvd.init()
        rank, nworker = hvd.rank(), hvd.size()
        mode = {"device": "auto", "local": "auto", "nccl": "nccl"}.get(opt.kvstore, "ps_sync" if "dist" in opt.kvstore else "auto")
        if nworker == 1:
            mode = "local"
        dp = DataParallelTrainer(net, mode=mode, optimizer="sgd_mom", lr=opt.lr, momentum=opt.momentum, wd=opt.wd)
        store = None
    else:
        store = kv.create(opt.kvstore)
        rank, nworker = store.rank, store.num_workers
        trainer = kv.trainer.Trainer(net.named_parameters(), "sgd",
                                     {"learning_rate": opt.lr, "wd": opt.wd, "momentum": opt.momentum,
                                      "multi_precision": dtype != torch.float32}, kvstore=store)
    train = load_dataset(opt.dataset, opt.data_dir, True, side, opt.num_samples)
    val = load_dataset(opt.dataset, opt.data_dir, False, side, max(256, opt.num_samples // 8))
    sampler = SplitSampler(len(train), nworker, rank, opt.seed)
    vsampler = SplitSampler(len(val), 1, 0, 1)
    lr_steps = [int(x) for x in opt.lr_steps.split(",") if x.strip()]
    speed = Speedometer(opt.batch_size, opt.log_interval, rank=rank if nworker > 1 else None)
    if opt.profile or opt.builtin_profiler:
        profiler.set_config(filename=os.path.join(opt.prefix or ".", "image-classifier-profile.json"))
        profiler.set_state("run")
    best = 0.0
    metric = Accuracy()
    in_dtype = dtype if use_gpu else torch.float32
    for epoch in range(opt.start_epoch, opt.epochs):
        if epoch in lr_steps:
            lr = opt.lr * opt.lr_factor ** (lr_steps.index(epoch) + 1)
            dp.set_learning_rate(lr) if bucketed else trainer.set_learning_rate(lr)
        tic = time.time()
        net.train()
        metric.reset()
        for i, (x, y) in enumerate(batches(train, opt.batch_size, sampler, device, in_dtype)):
            with profiler.scope("step", "train"):
                if bucketed:
                    loss = dp.step(x, y)
                else:
                    net.zero_grad(set_to_none=True)
                    out = net(x)
                    loss = F.cross_entropy(out.float(), y, reduction="sum")
                    loss.backward()
                    trainer.step(x.shape[0])
                    metric.update(out.detach(), y)
            speed(epoch, i, {"loss": float(loss.detach()) / (1 if bucketed else x.shape[0])} if i % opt.log_interval == 0 else None)
            if opt.max_batches and i + 1 >= opt.max_batches:
                break
        logging.info("[Epoch %d] training: accuracy=%f", epoch, metric.get())
        logging.info("[Epoch %d] time cost: %f", epoch, time.time() - tic)
        net.eval()
        metric.reset()
        with torch.no_grad():
            for x, y in batches(val, opt.batch_size, vsampler, device, in_dtype):
                metric.update(net(x), y)
        acc = metric.get()
        logging.info("[Epoch %d] validation: accuracy=%f", epoch, acc)
        if rank == 0 and opt.prefix:
            if acc > best:
                best = acc
                torch.save(net.state_dict(), f"{opt.prefix}-best.params")
            if opt.save_frequency and (epoch + 1) % opt.save_frequency == 0:
                save_checkpoint(opt.prefix, epoch + 1, net)
    if opt.profile or opt.builtin_profiler:
        logging.info(profiler.dumps())
        profiler.dump()
    if store is not None and hasattr(store, "close"):
        store.close()
    if bucketed:
        from deeplearning_cfn_b200.parallel import horovod as hvd
        hvd.shutdown()
    return 0


if __name__ == "__main__":
    sys.exit(main())
