"""Mask-RCNN R50-FPN training under the Horovod-style engine (the job `run.sh` launches).

Reference: TensorPack `examples/FasterRCNN/train.py` invoked by `examples/distributed-tensorflow/run.sh:82-95`
with `--logdir <dir> --config MODE_MASK=True MODE_FPN=True DATA.BASEDIR=... TRAIN.STEPS_PER_EPOCH=...
TRAIN.LR_SCHEDULE=[...] BACKBONE.NORM=FreezeBN TRAINER=horovod` (TensorPack itself is cloned at run time by
prepare-s3-bucket.sh:41-43, it is not part of the reference tree).  Kept here: the `--config KEY=VAL` override
surface (SURVEY §5.6 tier 4), the horovod trainer, checkpoints + logs under `--logdir` on the shared store,
data read from `DATA.BASEDIR` (the COCO-shaped shards written by `prepare-bucket`).
The detector is torchvision's Mask R-CNN R50-FPN (a library model, like TensorPack is for the reference) with
frozen BatchNorm (`BACKBONE.NORM=FreezeBN`); gradients are exchanged by the dlcfn fused all-reduce.
"""
from __future__ import annotations

import argparse
import ast
import glob
import json
import logging
import os
import sys
import time

import numpy as np
import torch

DEFAULTS = {"MODE_MASK": True, "MODE_FPN": True, "DATA.BASEDIR": "", "DATA.TRAIN": ["coco_train2017"],
            "DATA.VAL": ("coco_val2017",), "DATA.NUM_CLASS": 81, "TRAIN.EVAL_PERIOD": 1, "TRAIN.STEPS_PER_EPOCH": 500,
            "TRAIN.LR_SCHEDULE": [240000, 320000, 360000], "TRAIN.BASE_LR": 0.01, "TRAIN.WEIGHT_DECAY": 1e-4,
            "TRAIN.NUM_EPOCHS": 1, "TRAIN.BATCH_SIZE_PER_GPU": 1, "BACKBONE.WEIGHTS": "", "BACKBONE.NORM": "FreezeBN",
            "TRAINER": "horovod", "PREPROC.MAX_SIZE": 256}


def parse_config(items: list[str]) -> dict:
    cfg = dict(DEFAULTS)
    for it in items:
        if "=" not in it:
            raise SystemExit(f"--config expects KEY=VALUE, got {it}")
        k, v = it.split("=", 1)
        if k not in cfg:
            raise SystemExit(f"unknown config key {k}")
        try:
            cfg[k] = ast.literal_eval(v)
        except (ValueError, SyntaxError):
            cfg[k] = v
    return cfg


class Shards:
    def __init__(self, basedir: str, rank: int, world: int, max_size: int, num_class: int = 81):
        self.num_class = num_class
        files = sorted(glob.glob(os.path.join(basedir, "coco", "train2017-*.npz")))
        self.samples = []
        for f in files[rank::max(1, world)] or files:
            z = np.load(f, allow_pickle=True)
            for img, boxes, labels in zip(z["images"], z["boxes"], z["labels"]):
                self.samples.append((img, np.asarray(boxes, dtype=np.float32), np.asarray(labels, dtype=np.int64)))
        if not self.samples:   # no staged data: synthesise a few images so the job is still runnable
            rng = np.random.default_rng(rank)
            for _ in range(8):
                self.samples.append((rng.integers(0, 256, (max_size, max_size, 3), dtype=np.uint8),
                                     np.array([[8, 8, 64, 64], [40, 30, 120, 100]], dtype=np.float32), np.array([1, 2])))
        self.i = 0

    def next(self, device, with_masks: bool):
        img, boxes, labels = self.samples[self.i % len(self.samples)]
        self.i += 1
        x = torch.from_numpy(img).permute(2, 0, 1).float().div(255.0).to(device)
        labels = (labels - 1) % (self.num_class - 1) + 1     # keep labels inside [1, NUM_CLASS)
        t = {"boxes": torch.from_numpy(boxes).to(device), "labels": torch.from_numpy(labels).to(device)}
        if with_masks:
            h, w = x.shape[1:]
            m = torch.zeros(len(boxes), h, w, dtype=torch.uint8, device=device)
            for k, b in enumerate(boxes.astype(int)):
                m[k, b[1]:b[3] + 1, b[0]:b[2] + 1] = 1
            t["masks"] = m
        return x, t


def main(argv=None) -> int:
    p = argparse.ArgumentParser()
    p.add_argument("--logdir", required=True)
    p.add_argument("--config", nargs="*", default=[])
    p.add_argument("--load", default="")
    a = p.parse_args(argv)
    cfg = parse_config(a.config)
    from deeplearning_cfn_b200.parallel import horovod as hvd
    hvd.init()
    rank, world = hvd.rank(), hvd.size()
    os.makedirs(a.logdir, exist_ok=True)
    logging.basicConfig(level=logging.INFO, format=f"[rank {rank}] %(message)s", force=True,
                        handlers=[logging.StreamHandler(), logging.FileHandler(os.path.join(a.logdir, f"log.rank{rank}.txt"))])
    if rank == 0:
        json.dump({k: (list(v) if isinstance(v, tuple) else v) for k, v in cfg.items()}, open(os.path.join(a.logdir, "config.json"), "w"), indent=1)
    try:
        from torchvision.models.detection import maskrcnn_resnet50_fpn, fasterrcnn_resnet50_fpn
        from torchvision.ops.misc import FrozenBatchNorm2d
    except Exception as e:
        raise SystemExit(f"torchvision detection models are required for this example: {e}")
    dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    norm = FrozenBatchNorm2d if cfg["BACKBONE.NORM"] == "FreezeBN" else torch.nn.BatchNorm2d
    ctor = maskrcnn_resnet50_fpn if cfg["MODE_MASK"] else fasterrcnn_resnet50_fpn
    size = int(cfg["PREPROC.MAX_SIZE"])
    model = ctor(weights=None, weights_backbone=None, num_classes=int(cfg["DATA.NUM_CLASS"]), norm_layer=norm,
                 min_size=size, max_size=size, trainable_backbone_layers=3).to(dev).train()
    from deeplearning_cfn_b200.ops import roi as roi_ops
    if roi_ops.patch_torchvision():          # DLCFN_ROI_OPS=dlcfn: ROIAlign / NMS on the dlcfn kernels (experimental)
        logging.info("ROIAlign and NMS run on the dlcfn kernels")
    from deeplearning_cfn_b200.parallel.trainer import DataParallelTrainer
    # data-parallel engine: fp32 detector, fused all-reduce of fp32 gradient buckets, fused SGD-momentum
    tr = DataParallelTrainer(model, mode="auto", optimizer="sgd_mom", lr=float(cfg["TRAIN.BASE_LR"]), momentum=0.9,
                             wd=float(cfg["TRAIN.WEIGHT_DECAY"]), loss_fn=lambda out, _y: sum(out.values()))
    data = Shards(cfg["DATA.BASEDIR"], rank, world, size, int(cfg["DATA.NUM_CLASS"]))
    sched = list(cfg["TRAIN.LR_SCHEDULE"])
    ckpt = os.path.join(a.logdir, "checkpoint.pt")
    step = 0
    if os.path.exists(ckpt):
        sd = torch.load(ckpt, map_location="cpu", weights_only=False)
        tr.load_state_dict(sd["trainer"])
        step = sd["step"]
        logging.info("resumed from %s at step %d", ckpt, step)
    bs = int(cfg["TRAIN.BATCH_SIZE_PER_GPU"])
    for epoch in range(int(cfg["TRAIN.NUM_EPOCHS"])):
        t0 = time.time()
        for _ in range(int(cfg["TRAIN.STEPS_PER_EPOCH"])):
            imgs, tgts = zip(*[data.next(dev, cfg["MODE_MASK"]) for _ in range(bs)])
            lr = float(cfg["TRAIN.BASE_LR"]) * 0.1 ** sum(1 for s in sched if step * world >= s)
            tr.set_learning_rate(lr)
            loss = tr.step((list(imgs), list(tgts)), None)
            step += 1
            if step % 10 == 0:
                logging.info("step %d total_loss %.4f lr %.5f", step, float(loss.detach()), lr)
        logging.info("epoch %d done in %.1f s (%.2f img/s per rank)", epoch, time.time() - t0,
                     cfg["TRAIN.STEPS_PER_EPOCH"] * bs / (time.time() - t0))
        if rank == 0:
            torch.save({"trainer": tr.state_dict(), "step": step}, ckpt)
    tr.check()
    hvd.shutdown()
    return 0


if __name__ == "__main__":
    sys.exit(main())
