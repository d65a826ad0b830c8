"""Shared pieces of the example scripts: datasets (synthetic when no local copy exists - there is no network),
batch iteration sharded by rank (reference `SplitSampler`, MX/example/distributed_training/cifar10_dist.py:54-81),
accuracy metric."""
from __future__ import annotations

import os

import numpy as np
import torch

DATASETS = {"mnist": (1, 28, 10), "cifar10": (3, 32, 10), "caltech101": (3, 224, 101), "imagenet": (3, 224, 1000),
            "dummy": (3, 224, 1000)}


class SyntheticDataset:
    """Fixed random images with labels that are a (noisy) function of the image so that training can learn."""

    def __init__(self, name: str, size: int, num: int = 2048, image_size: int | None = None, seed: int = 0):
        ch, side, classes = DATASETS.get(name, DATASETS["dummy"])
        side = image_size or side
        g = torch.Generator().manual_seed(seed)
        self.classes = classes
        protos = torch.randn(classes, ch, 8, 8, generator=g)
        self.y = torch.randint(0, classes, (num,), generator=g)
        base = torch.nn.functional.interpolate(protos[self.y], size=(side, side), mode="nearest")
        self.x = base + 0.5 * torch.randn(num, ch, side, side, generator=g)

    def __len__(self):
        return self.x.shape[0]


class RecordDataset:
    """`<data_dir>/train.rec` (+ .idx) / `val.rec`: the RecordIO form the reference's image-classification examples read for
    ImageNet-style data (`--data-dir` with train.rec / val.rec, image_classification.py:104-140).  Batches come from the
    native decoder pipeline (utils/recordio.py:NativeImageRecordIter), sharded like `SplitSampler` does for in-memory data."""

    def __init__(self, rec_path: str, name: str, train: bool, image_size: int | None):
        from ..utils.recordio import NativeRecordIO
        ch, side, classes = DATASETS.get(name, DATASETS["dummy"])
        self.rec_path, self.train, self.classes = rec_path, train, classes
        self.shape = (ch, image_size or side, image_size or side)
        idx = os.path.splitext(rec_path)[0] + ".idx"
        r = NativeRecordIO(idx if os.path.exists(idx) else None, rec_path, "r")
        self._n = len(r)
        r.close()
        self._iters: dict = {}

    def __len__(self):
        return self._n

    def iterator(self, batch_size: int, num_parts: int, part_index: int, seed: int):
        from ..utils.recordio import image_record_iter
        key = (batch_size, num_parts, part_index)
        if key not in self._iters:
            self._iters[key] = image_record_iter(self.rec_path, batch_size, self.shape, shuffle=self.train,
                                                     num_parts=num_parts, part_index=part_index, seed=seed,
                                                     rand_crop=self.train, rand_mirror=self.train, round_batch=False)
        return self._iters[key]


def load_dataset(name: str, data_dir: str | None, train: bool = True, image_size: int | None = None, num: int = 2048):
    """RecordIO files (train.rec / val.rec) or CIFAR-10 python batches under data_dir are used when present; otherwise
    synthetic data of the same shape."""
    if data_dir:
        rec = os.path.join(data_dir, "train.rec" if train else "val.rec")
        if os.path.exists(rec):
            return RecordDataset(rec, name, train, image_size)
    if name == "cifar10" and data_dir:
        d = os.path.join(data_dir, "cifar-10-batches-py")
        if os.path.isdir(d):
            import pickle
            files = [f"data_batch_{i}" for i in range(1, 6)] if train else ["test_batch"]
            xs, ys = [], []
            for f in files:
                with open(os.path.join(d, f), "rb") as fh:
                    b = pickle.load(fh, encoding="bytes")
                xs.append(np.asarray(b[b"data"], dtype=np.uint8).reshape(-1, 3, 32, 32))
                ys.append(np.asarray(b[b"labels"]))
            ds = SyntheticDataset.__new__(SyntheticDataset)
            ds.x = (torch.from_numpy(np.concatenate(xs)).float() / 255.0 - 0.5) / 0.25
            ds.y = torch.from_numpy(np.concatenate(ys)).long()
            ds.classes = 10
            return ds
    return SyntheticDataset(name, 0, num=num, image_size=image_size, seed=0 if train else 1)


class SplitSampler:
    """Each worker sees its own 1/num_parts slice, reshuffled every epoch (cifar10_dist.py:54-81)."""

    def __init__(self, length: int, num_parts: int = 1, part_index: int = 0, seed: int = 0):
        self.num_parts, self.part_index = num_parts, part_index
        self.part_len = length // num_parts
        self.start = self.part_len * part_index
        self.seed = seed
        self.epoch = 0

    def indices(self) -> torch.Tensor:
        g = torch.Generator().manual_seed(self.seed + self.epoch)
        self.epoch += 1
        return self.start + torch.randperm(self.part_len, generator=g)

    def __len__(self):
        return self.part_len


def batches(ds, batch_size: int, sampler: SplitSampler, device, dtype=torch.float32, channels_last: bool = True):
    if isinstance(ds, RecordDataset):
        from ..utils.data import IMAGENET_MEAN, IMAGENET_STD
        mean = torch.tensor(IMAGENET_MEAN, dtype=torch.float32).view(1, 1, 1, 3)
        std = torch.tensor(IMAGENET_STD, dtype=torch.float32).view(1, 1, 1, 3)
        for x_u8, y in ds.iterator(batch_size, sampler.num_parts, sampler.part_index, sampler.seed):
            x = ((x_u8.float() / 255.0 - mean) / std).permute(0, 3, 1, 2)        # NHWC storage viewed as NCHW = channels_last
            x = x.to(device=device, dtype=dtype, non_blocking=True)
            if not channels_last:
                x = x.contiguous()
            yield x, y.to(device, non_blocking=True)
        return
    idx = sampler.indices()
    for i in range(0, len(idx) - batch_size + 1, batch_size):
        j = idx[i:i + batch_size]
        x = ds.x[j].to(device=device, dtype=dtype, non_blocking=True)
        if channels_last and x.dim() == 4:
            x = x.contiguous(memory_format=torch.channels_last)
        yield x, ds.y[j].to(device, non_blocking=True)


class Accuracy:
    def __init__(self):
        self.correct = self.total = 0

    def update(self, out, y):
        self.correct += int((out.argmax(-1) == y).sum())
        self.total += int(y.numel())

    def get(self) -> float:
        return self.correct / max(1, self.total)

    def reset(self):
        self.correct = self.total = 0
