"""RecordIO files and a prefetching image-record iterator (reference E7: MX/src/io/*, dmlc-core recordio,
`mx.recordio.MXRecordIO / MXIndexedRecordIO / pack / unpack`, `ImageRecordIter` with its prefetcher thread).

File format (dmlc recordio): every record = uint32 magic 0xced7230a, uint32 lrecord (upper 3 bits: continuation
flag, lower 29 bits: length), payload padded to 4 bytes.  `pack`/`unpack` prepend the IRHeader
(uint32 flag, float32 label, uint64 id, uint64 id2).  Images are stored raw (uint8 HWC + shape) here - there is no
JPEG codec dependency in the image; the layout is otherwise interchangeable.
"""
from __future__ import annotations

import os
import queue
import struct
import threading
from collections import namedtuple

import numpy as np
import torch

MAGIC = 0xCED7230A
IRHeader = namedtuple("IRHeader", ["flag", "label", "id", "id2"])
_IR = struct.Struct("<IfQQ")


class MXRecordIO:
    def __init__(self, uri: str, flag: str):
        assert flag in ("r", "w")
        self.uri, self.flag = uri, flag
        self.f = open(uri, "rb" if flag == "r" else "wb")

    def write(self, buf: bytes) -> None:
        n = len(buf)
        if n >= 1 << 29:
            raise ValueError("record too large")
        self.f.write(struct.pack("<II", MAGIC, n))
        self.f.write(buf)
        self.f.write(b"\0" * ((4 - n % 4) % 4))

    def read(self) -> bytes | None:
        head = self.f.read(8)
        if len(head) < 8:
            return None
        magic, lrec = struct.unpack("<II", head)
        if magic != MAGIC:
            raise IOError(f"bad record magic in {self.uri}")
        n = lrec & ((1 << 29) - 1)
        buf = self.f.read(n)
        self.f.read((4 - n % 4) % 4)
        return buf

    def tell(self) -> int:
        return self.f.tell()

    def seek(self, pos: int) -> None:
        self.f.seek(pos)

    def reset(self) -> None:
        self.f.seek(0)

    def close(self) -> None:
        self.f.close()


class MXIndexedRecordIO(MXRecordIO):
    def __init__(self, idx_path: str, uri: str, flag: str):
        super().__init__(uri, flag)
        self.idx_path, self.idx, self.keys = idx_path, {}, []
        if flag == "r" and os.path.exists(idx_path):
            for line in open(idx_path):
                k, pos = line.split("\t")
                self.idx[int(k)] = int(pos)
                self.keys.append(int(k))

    def write_idx(self, idx: int, buf: bytes) -> None:
        self.idx[idx] = self.tell()
        self.keys.append(idx)
        self.write(buf)

    def read_idx(self, idx: int) -> bytes:
        self.seek(self.idx[idx])
        return self.read()

    def close(self) -> None:
        if self.flag == "w":
            with open(self.idx_path, "w") as f:
                for k in self.keys:
                    f.write(f"{k}\t{self.idx[k]}\n")
        super().close()


def pack(header: IRHeader, s: bytes) -> bytes:
    return _IR.pack(int(header.flag), float(header.label), int(header.id), int(header.id2)) + s


def unpack(s: bytes):
    h = IRHeader(*_IR.unpack(s[:_IR.size]))
    return h, s[_IR.size:]


def pack_img(header: IRHeader, img: np.ndarray) -> bytes:
    img = np.ascontiguousarray(img, dtype=np.uint8)
    return pack(header, struct.pack("<III", *img.shape) + img.tobytes())


def unpack_img(s: bytes):
    h, body = unpack(s)
    hh, ww, cc = struct.unpack("<III", body[:12])
    return h, np.frombuffer(body, dtype=np.uint8, offset=12).reshape(hh, ww, cc)


class ImageRecordIter:
    """Batches of (uint8 NHWC pinned tensor, int64 labels) read from a .rec file by `preprocess_threads` workers
    into a bounded prefetch queue (`prefetch_buffer`), sharded by (part_index, num_parts) like the reference."""

    def __init__(self, path_imgrec: str, batch_size: int, data_shape=(3, 224, 224), path_imgidx: str | None = None,
                 shuffle: bool = False, num_parts: int = 1, part_index: int = 0, preprocess_threads: int = 2,
                 prefetch_buffer: int = 4, seed: int = 0, round_batch: bool = True):
        self.batch_size, self.shape = batch_size, tuple(data_shape)
        idx_path = path_imgidx or os.path.splitext(path_imgrec)[0] + ".idx"
        self.rec = MXIndexedRecordIO(idx_path, path_imgrec, "r")
        if not self.rec.keys:
            raise IOError(f"index file {idx_path} missing or empty")
        keys = self.rec.keys[part_index::num_parts]
        self.keys, self.shuffle, self.seed, self.epoch = keys, shuffle, seed, 0
        self.threads, self.depth = preprocess_threads, prefetch_buffer
        self._lock = threading.Lock()
        self._q: queue.Queue | None = None
        self.reset()

    def reset(self) -> None:
        order = list(self.keys)
        if self.shuffle:
            rng = np.random.default_rng(self.seed + self.epoch)
            rng.shuffle(order)
        self.epoch += 1
        batches = [order[i:i + self.batch_size] for i in range(0, len(order) - self.batch_size + 1, self.batch_size)]
        self._q = queue.Queue(maxsize=self.depth)
        self._n = len(batches)
        it = iter(enumerate(batches))
        done = [0]

        def worker():
            while True:
                with self._lock:
                    nxt = next(it, None)
                if nxt is None:
                    break
                _, ks = nxt
                c, h, w = self.shape
                x = torch.empty((len(ks), h, w, c), dtype=torch.uint8)
                y = torch.empty(len(ks), dtype=torch.int64)
                for j, k in enumerate(ks):
                    with self._lock:
                        raw = self.rec.read_idx(k)
                    hd, img = unpack_img(raw)
                    x[j] = torch.from_numpy(img[:h, :w, :c].copy())
                    y[j] = int(hd.label)
                if torch.cuda.is_available():
                    x, y = x.pin_memory(), y.pin_memory()
                self._q.put((x, y))
            with self._lock:
                done[0] += 1
                if done[0] == self.threads:
                    self._q.put(None)
        for _ in range(self.threads):
            threading.Thread(target=worker, daemon=True).start()

    def __iter__(self):
        return self

    def __next__(self):
        item = self._q.get()
        if item is None:
            raise StopIteration
        return item

    def __len__(self):
        return self._n
