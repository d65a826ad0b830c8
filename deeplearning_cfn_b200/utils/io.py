"""Data iterators with the `mx.io` contract (reference MX/python/mxnet/io.py; the C++ side is E7 src/io):
`DataDesc`, `DataBatch`, `DataIter` (reset / next / iteration protocol, `provide_data`, `provide_label`),
`NDArrayIter` (in-memory arrays, shuffle, last_batch_handle pad | discard | roll_over), `ResizeIter`,
`PrefetchingIter` (background thread + bounded queue, optionally staging batches in pooled pinned memory so that
the H2D copy that follows is asynchronous).  RecordIO-backed iteration lives in utils/recordio.py."""
from __future__ import annotations

import queue
import threading
from collections import namedtuple

import torch


class DataDesc(namedtuple("DataDesc", ["name", "shape"])):
    def __new__(cls, name, shape, dtype=torch.float32, layout="NCHW"):
        ret = super().__new__(cls, name, tuple(shape))
        ret.dtype, ret.layout = dtype, layout
        return ret

    @staticmethod
    def get_batch_axis(layout: str | None) -> int:
        return 0 if layout is None else layout.find("N")


class DataBatch:
    def __init__(self, data, label=None, pad: int = 0, index=None, provide_data=None, provide_label=None):
        self.data = list(data) if isinstance(data, (list, tuple)) else [data]
        self.label = [] if label is None else (list(label) if isinstance(label, (list, tuple)) else [label])
        self.pad, self.index = pad, index
        self.provide_data, self.provide_label = provide_data, provide_label

    def __iter__(self):          # lets `x, y = batch` work for the single-input / single-label case
        yield self.data[0]
        yield self.label[0] if self.label else None

    def __str__(self):
        return f"DataBatch: data shapes: {[tuple(d.shape) for d in self.data]} label shapes: {[tuple(l.shape) for l in self.label]}"


class DataIter:
    def __init__(self, batch_size: int = 0):
        self.batch_size = batch_size

    def __iter__(self):
        return self

    def reset(self) -> None:
        pass

    def next(self) -> DataBatch:
        if self.iter_next():
            return DataBatch(self.getdata(), self.getlabel(), self.getpad(), self.getindex())
        raise StopIteration

    def __next__(self):
        return self.next()

    def iter_next(self) -> bool:
        raise NotImplementedError

    def getdata(self):
        raise NotImplementedError

    def getlabel(self):
        raise NotImplementedError

    def getindex(self):
        return None

    def getpad(self) -> int:
        return 0


def _named(arrs, default):
    if arrs is None:
        return []
    if isinstance(arrs, torch.Tensor):
        return [(default, arrs)]
    if isinstance(arrs, dict):
        return list(arrs.items())
    arrs = list(arrs)
    if len(arrs) == 1:
        return [(default, arrs[0])]
    return [(f"_{i}_{default}", a) for i, a in enumerate(arrs)]


class NDArrayIter(DataIter):
    def __init__(self, data, label=None, batch_size: int = 1, shuffle: bool = False, last_batch_handle: str = "pad",
                 data_name: str = "data", label_name: str = "softmax_label", seed: int = 0):
        super().__init__(batch_size)
        if last_batch_handle not in ("pad", "discard", "roll_over"):
            raise ValueError("last_batch_handle must be pad, discard or roll_over")
        self.data, self.label = _named(data, data_name), _named(label, label_name)
        self.num_data = self.data[0][1].shape[0]
        for _, a in self.data + self.label:
            if a.shape[0] != self.num_data:
                raise ValueError("all arrays need the same number of samples")
        self.shuffle, self.last_batch_handle = shuffle, last_batch_handle
        self._gen = torch.Generator().manual_seed(seed)
        self._carry = None                  # roll_over: indices left from the previous epoch
        self.reset()

    @property
    def provide_data(self):
        return [DataDesc(n, (self.batch_size,) + tuple(a.shape[1:]), a.dtype) for n, a in self.data]

    @property
    def provide_label(self):
        return [DataDesc(n, (self.batch_size,) + tuple(a.shape[1:]), a.dtype) for n, a in self.label]

    def reset(self) -> None:
        order = torch.randperm(self.num_data, generator=self._gen) if self.shuffle else torch.arange(self.num_data)
        if self.last_batch_handle == "roll_over" and self._carry is not None and len(self._carry):
            order = torch.cat([self._carry, order])
        self._carry = None
        self._order, self._cursor = order, -self.batch_size

    def iter_next(self) -> bool:
        self._cursor += self.batch_size
        n = len(self._order)
        if self._cursor >= n:
            return False
        if self._cursor + self.batch_size > n:
            if self.last_batch_handle == "discard":
                return False
            if self.last_batch_handle == "roll_over":
                self._carry = self._order[self._cursor:]
                return False
        return True

    def _take(self, arrs):
        n = len(self._order)
        idx = self._order[self._cursor:self._cursor + self.batch_size]
        if len(idx) < self.batch_size:                     # pad with samples from the start of the epoch
            idx = torch.cat([idx, self._order[:self.batch_size - len(idx)]])
        return [a[idx.to(a.device)] for _, a in arrs], idx, n

    def getdata(self):
        return self._take(self.data)[0]

    def getlabel(self):
        return self._take(self.label)[0]

    def getindex(self):
        return self._take(self.data)[1]

    def getpad(self) -> int:
        n = len(self._order)
        if self.last_batch_handle == "pad" and self._cursor + self.batch_size > n:
            return self._cursor + self.batch_size - n
        return 0


class ResizeIter(DataIter):
    """Fix the number of batches per epoch of another iterator (wraps around / cuts short)."""

    def __init__(self, data_iter: DataIter, size: int, reset_internal: bool = True):
        super().__init__(data_iter.batch_size)
        self.data_iter, self.size, self.reset_internal = data_iter, size, reset_internal
        self.cur = 0
        self.current_batch = None
        self.provide_data = getattr(data_iter, "provide_data", None)
        self.provide_label = getattr(data_iter, "provide_label", None)

    def reset(self) -> None:
        self.cur = 0
        if self.reset_internal:
            self.data_iter.reset()

    def iter_next(self) -> bool:
        if self.cur == self.size:
            return False
        try:
            self.current_batch = self.data_iter.next()
        except StopIteration:
            self.data_iter.reset()
            self.current_batch = self.data_iter.next()
        self.cur += 1
        return True

    def getdata(self):
        return self.current_batch.data

    def getlabel(self):
        return self.current_batch.label

    def getindex(self):
        return self.current_batch.index

    def getpad(self) -> int:
        return self.current_batch.pad


class PrefetchingIter(DataIter):
    """Runs the wrapped iterator on a background thread, `depth` batches ahead (reference PrefetchingIter / the C++
    PrefetcherIter, src/io/iter_prefetcher.h).  pin=True stages every tensor in a pooled page-locked block
    (utils/storage.py) so that a following `.to(device, non_blocking=True)` is a true asynchronous copy."""

    _END = object()

    def __init__(self, data_iter: DataIter, depth: int = 2, pin: bool = False):
        super().__init__(data_iter.batch_size)
        self.data_iter, self.depth, self.pin = data_iter, depth, pin
        self.provide_data = getattr(data_iter, "provide_data", None)
        self.provide_label = getattr(data_iter, "provide_label", None)
        self._q: queue.Queue = queue.Queue(maxsize=depth)
        self._thread = None
        self._stop = threading.Event()
        self._pinned: list = []
        self.current_batch = None
        self._start()

    def _stage(self, t: torch.Tensor) -> torch.Tensor:
        if not self.pin or t.is_cuda:
            return t
        from . import storage
        buf = storage.host_empty(tuple(t.shape), t.dtype, pin=torch.cuda.is_available())
        buf.copy_(t)
        self._pinned.append(buf)
        return buf

    def _worker(self):
        try:
            while not self._stop.is_set():
                try:
                    b = self.data_iter.next()
                except StopIteration:
                    self._q.put(self._END)
                    return
                b.data = [self._stage(t) for t in b.data]
                b.label = [self._stage(t) for t in b.label]
                self._q.put(b)
        except Exception as e:        # surface iterator errors on the consumer side
            self._q.put(e)

    def _start(self):
        self._stop.clear()
        self._thread = threading.Thread(target=self._worker, daemon=True)
        self._thread.start()

    def _release_pinned(self, keep: int = 0):
        if not self._pinned:
            return
        from . import storage
        while len(self._pinned) > keep:
            storage.release(self._pinned.pop(0))

    def reset(self) -> None:
        self._stop.set()
        while self._thread is not None and self._thread.is_alive():
            try:
                self._q.get(timeout=0.05)
            except queue.Empty:
                pass
        while not self._q.empty():
            self._q.get_nowait()
        self._release_pinned()
        self.data_iter.reset()
        self._start()

    def iter_next(self) -> bool:
        item = self._q.get()
        if item is self._END:
            return False
        if isinstance(item, Exception):
            raise item
        self.current_batch = item
        # blocks handed out more than `depth + 2` batches ago have been consumed: recycle them
        per = len(item.data) + len(item.label)
        self._release_pinned(keep=(self.depth + 2) * per)
        return True

    def getdata(self):
        return self.current_batch.data

    def getlabel(self):
        return self.current_batch.label

    def getindex(self):
        return self.current_batch.index

    def getpad(self) -> int:
        return self.current_batch.pad

    def close(self) -> None:
        self._stop.set()
        while self._thread is not None and self._thread.is_alive():
            try:
                self._q.get(timeout=0.05)
            except queue.Empty:
                pass
        self._release_pinned()
