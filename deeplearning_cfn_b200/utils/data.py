"""Synthetic input pipeline: pinned host batches -> async H2D on a copy stream -> fused uint8->bf16
normalise kernel.  Stands in for the reference's RecordIO iterator + prefetcher (reference E7) on the
benchmark path, which the reference itself replaces by synthetic data under `--benchmark 1`
(common/data.py:98-138)."""
from __future__ import annotations

import torch

from ..ops import native
from ..ops import reference as ref_ops

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


class SyntheticImageSource:
    """A few distinct pinned uint8 NHWC batches, cycled (host side of the input pipeline)."""

    def __init__(self, batch: int, size: int = 224, classes: int = 1000, distinct: int = 4, seed: int = 0,
                 pin: bool = True):
        g = torch.Generator().manual_seed(seed)
        self.images = [torch.randint(0, 256, (batch, size, size, 3), dtype=torch.uint8, generator=g)
                       for _ in range(distinct)]
        self.labels = [torch.randint(0, classes, (batch,), dtype=torch.int64, generator=g) for _ in range(distinct)]
        if pin and torch.cuda.is_available():
            self.images = [t.pin_memory() for t in self.images]
            self.labels = [t.pin_memory() for t in self.labels]
        self._i = 0
        self.bytes_per_batch = self.images[0].numel() + self.labels[0].numel() * 8

    def next(self):
        i = self._i % len(self.images)
        self._i += 1
        return self.images[i], self.labels[i]


class DevicePrefetcher:
    """Double-buffered H2D: batch i+1 is copied on `copy_stream` while step i computes."""

    def __init__(self, source: SyntheticImageSource, device: torch.device, mean=IMAGENET_MEAN, std=IMAGENET_STD):
        self.source, self.device, self.mean, self.std = source, device, mean, std
        self.copy_stream = torch.cuda.Stream()
        self._slots = []
        for _ in range(2):
            img, lab = source.images[0], source.labels[0]
            self._slots.append({"u8": torch.empty(img.shape, dtype=torch.uint8, device=device),
                                "lab": torch.empty(lab.shape, dtype=torch.int64, device=device),
                                "x": torch.empty(img.shape, dtype=torch.bfloat16, device=device),
                                "ev": torch.cuda.Event()})
        self._k = 0
        self._issue()

    def _issue(self):
        s = self._slots[self._k % 2]
        img, lab = self.source.next()
        # the slot's previous consumer ran on the current stream
        self.copy_stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.copy_stream):
            s["u8"].copy_(img, non_blocking=True)
            s["lab"].copy_(lab, non_blocking=True)
            s["ev"].record(self.copy_stream)

    def next(self):
        """Returns (x bf16 NCHW-shaped channels_last view, labels) for the current step."""
        s = self._slots[self._k % 2]
        torch.cuda.current_stream().wait_event(s["ev"])
        native.normalize_u8_nhwc(s["u8"], s["x"], self.mean, self.std)
        self._k += 1
        self._issue()
        # NHWC storage viewed as NCHW = channels_last tensor, no copy
        return s["x"].permute(0, 3, 1, 2), s["lab"]


def synthetic_device_batch(batch: int, size: int, classes: int, device, dtype=torch.bfloat16, seed: int = 0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn(batch, 3, size, size, generator=g).to(device=device, dtype=dtype)
    x = x.contiguous(memory_format=torch.channels_last)
    y = torch.randint(0, classes, (batch,), generator=g).to(device)
    return x, y


class RecordBatchLoader:
    """Feeds a trainer from a RecordIO file: native decoder ring (utils/recordio.py:NativeImageRecordIter) -> asynchronous
    H2D out of the page-locked ring slot on a copy stream -> fused uint8->bf16 normalise kernel.  On CPU the batch is
    normalised with the reference expression.  Yields (x [B,C,H,W] channels_last, y int64)."""

    def __init__(self, it, device: torch.device, mean=IMAGENET_MEAN, std=IMAGENET_STD, dtype=torch.bfloat16):
        self.it, self.device, self.mean, self.std, self.dtype = it, device, mean, std, dtype
        self.cuda = device.type == "cuda"
        self.copy_stream = torch.cuda.Stream() if self.cuda else None
        self._bufs: list = []          # two device staging sets, alternated
        self._k = 0

    def __len__(self):
        return len(self.it)

    def __iter__(self):
        for x_u8, y in self.it:
            if not self.cuda:
                x = ref_ops.normalize_u8_nhwc(x_u8, self.mean, self.std).to(
                    self.dtype if self.dtype != torch.bfloat16 else torch.float32)
                yield x.permute(0, 3, 1, 2), y.clone()
                continue
            if len(self._bufs) < 2:
                self._bufs.append({"u8": torch.empty(x_u8.shape, dtype=torch.uint8, device=self.device),
                                   "y": torch.empty(y.shape, dtype=torch.int64, device=self.device),
                                   "x": torch.empty(x_u8.shape, dtype=torch.bfloat16, device=self.device),
                                   "ev": torch.cuda.Event()})
            s = self._bufs[self._k % 2]
            n = x_u8.shape[0]
            if n > s["u8"].shape[0] or x_u8.shape[1:] != s["u8"].shape[1:]:      # a larger / differently shaped batch: new staging set
                s.update(u8=torch.empty(x_u8.shape, dtype=torch.uint8, device=self.device),
                         y=torch.empty(y.shape, dtype=torch.int64, device=self.device),
                         x=torch.empty(x_u8.shape, dtype=torch.bfloat16, device=self.device))
            u8, yd, xd = s["u8"][:n], s["y"][:n], s["x"][:n]                      # the last batch of an epoch may be short
            self._k += 1
            self.copy_stream.wait_stream(torch.cuda.current_stream())     # the slot's previous consumer has been enqueued
            with torch.cuda.stream(self.copy_stream):
                u8.copy_(x_u8, non_blocking=True)                          # ring slot is page-locked: true async DMA
                yd.copy_(y.pin_memory() if not y.is_pinned() else y, non_blocking=True)
                s["ev"].record(self.copy_stream)
            s["ev"].synchronize()      # the ring slot may be recycled `keep` batches later; the copy is short
            torch.cuda.current_stream().wait_event(s["ev"])
            native.normalize_u8_nhwc(u8, xd, self.mean, self.std)
            yield xd.permute(0, 3, 1, 2), yd
