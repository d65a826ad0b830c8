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
        """Next record; records that a dmlc writer split at embedded magic words (cflag 1 / 2 / 3) are re-joined."""
        out, parts = b"", 0
        while True:
            head = self.f.read(8)
            if len(head) < 8:
                return None if parts == 0 else out
            magic, lrec = struct.unpack("<II", head)
            if magic != MAGIC:
                raise IOError(f"bad record magic in {self.uri}")
            cflag, n = lrec >> 29, lrec & ((1 << 29) - 1)
            buf = self.f.read(n)
            self.f.read((4 - n % 4) % 4)
            if cflag == 0:
                return buf
            out += buf
            parts += 1
            if cflag == 3:
                return out
            out += struct.pack("<I", MAGIC)

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


def pack_img(header: IRHeader, img: np.ndarray, quality: int = 95, img_fmt: str = ".raw") -> bytes:
    """Pack an HWC uint8 image.  img_fmt '.jpg' / '.jpeg' / '.png' store the encoded image as the reference does
    (`mx.recordio.pack_img(header, img, quality, img_fmt)`; readable by any RecordIO consumer), '.raw' (default) stores
    uint32 h, w, c + the pixels, which is what the native decoder ring reads without a codec."""
    img = np.array(img, dtype=np.uint8, order="C")
    if img.ndim == 2:
        img = img[:, :, None]
    fmt = img_fmt.lower()
    if fmt == ".raw":
        return pack(header, struct.pack("<III", *img.shape) + img.tobytes())
    import torchvision.io as tvio
    chw = torch.from_numpy(img).permute(2, 0, 1).contiguous()
    if fmt in (".jpg", ".jpeg"):
        enc = tvio.encode_jpeg(chw, quality=int(quality))
    elif fmt == ".png":
        enc = tvio.encode_png(chw, compression_level=max(0, min(9, int(quality))) if quality <= 9 else 6)
    else:
        raise ValueError(f"unsupported image format {img_fmt!r}")
    return pack(header, enc.numpy().tobytes())


def image_format(body: bytes) -> str:
    if body[:3] == b"\xff\xd8\xff":
        return ".jpg"
    if body[:8] == b"\x89PNG\r\n\x1a\n":
        return ".png"
    return ".raw"


def unpack_img(s: bytes, device: str = "cpu"):
    """(header, HWC uint8 numpy image); JPEG / PNG payloads are decoded (torchvision.io: libjpeg-turbo on the host,
    nvJPEG with device='cuda'), raw payloads are viewed in place."""
    h, body = unpack(s)
    off = 4 * int(h.flag) if h.flag > 0 else 0           # multi-label records carry `flag` floats first
    if h.flag > 0:
        labels = np.frombuffer(body, dtype=np.float32, count=int(h.flag))
        h = IRHeader(h.flag, labels, h.id, h.id2)
    body = body[off:]
    fmt = image_format(body)
    if fmt == ".raw":
        hh, ww, cc = struct.unpack("<III", body[:12])
        return h, np.frombuffer(body, dtype=np.uint8, offset=12, count=hh * ww * cc).reshape(hh, ww, cc)
    import torchvision.io as tvio
    data = torch.frombuffer(bytearray(body), dtype=torch.uint8)
    img = tvio.decode_jpeg(data, device=device) if fmt == ".jpg" else tvio.decode_png(data)
    return h, img.permute(1, 2, 0).contiguous().cpu().numpy()


class ImageRecordIter:
    """Batches of (uint8 NHWC pinned tensor, int64 labels) read from a .rec file by `preprocess_threads` workers
    into a bounded prefetch queue (`prefetch_buffer`), sharded by (part_index, num_parts) like the reference."""

    def __init__(self, path_imgrec: str, batch_size: int, data_shape=(3, 224, 224), path_imgidx: str | None = None,
                 shuffle: bool = False, num_parts: int = 1, part_index: int = 0, preprocess_threads: int = 2,
                 prefetch_buffer: int = 4, seed: int = 0, round_batch: bool = True, rand_crop: bool = False,
                 rand_mirror: bool = False):
        self.batch_size, self.shape = batch_size, tuple(data_shape)
        self.rand_crop, self.rand_mirror, self.errors = rand_crop, rand_mirror, 0
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
        self.close()                                        # an unfinished epoch's workers stop first
        self._done = False
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

        stop = self._stop = threading.Event()
        q = self._q

        def worker():
            try:
                work(q, stop)
            except Exception as e:                          # surfaced by __next__ instead of a silent hang
                q.put(e)

        def work(q, stop):
            while not stop.is_set():
                with self._lock:
                    nxt = next(it, None)
                if nxt is None:
                    break
                bi, ks = nxt
                aug = np.random.default_rng((self.seed, self.epoch, bi))     # per batch: independent of thread timing
                c, h, w = self.shape
                x = torch.empty((len(ks), h, w, c), dtype=torch.uint8)
                y = torch.empty(len(ks), dtype=torch.int64)
                for j, k in enumerate(ks):
                    with self._lock:
                        raw = self.rec.read_idx(k)
                    hd, img = unpack_img(raw)
                    sh, sw = img.shape[0], img.shape[1]
                    y0, x0 = max(0, (sh - h) // 2), max(0, (sw - w) // 2)          # center crop, zero pad when smaller
                    if self.rand_crop:
                        y0, x0 = int(aug.integers(0, max(0, sh - h) + 1)), int(aug.integers(0, max(0, sw - w) + 1))
                    crop = img[y0:y0 + h, x0:x0 + w, :c]
                    if self.rand_mirror and aug.integers(0, 2):
                        crop = crop[:, ::-1]
                    x[j].zero_()
                    dy0, dx0 = (h - crop.shape[0]) // 2, (w - crop.shape[1]) // 2
                    x[j, dy0:dy0 + crop.shape[0], dx0:dx0 + crop.shape[1], :crop.shape[2]] = torch.from_numpy(crop.copy())
                    y[j] = int(hd.label if np.isscalar(hd.label) else hd.label[0])
                if torch.cuda.is_available():
                    x, y = x.pin_memory(), y.pin_memory()
                while not stop.is_set():
                    try:
                        q.put((x, y), timeout=0.1)
                        break
                    except queue.Full:
                        pass
            with self._lock:
                done[0] += 1
                last = done[0] == self.threads
            if last and not stop.is_set():
                q.put(None)
        self._workers = [threading.Thread(target=worker, daemon=True) for _ in range(self.threads)]
        for t in self._workers:
            t.start()

    def close(self) -> None:
        """Stop the prefetch threads (an abandoned epoch otherwise keeps reading the file)."""
        if getattr(self, "_stop", None) is not None:
            self._stop.set()
        for t in getattr(self, "_workers", []):
            t.join(timeout=5)
        self._workers = []

    def __iter__(self):
        if getattr(self, "_done", False):        # a finished epoch: iterating again starts the next one (like the native reader)
            self.reset()
        return self

    def __next__(self):
        if getattr(self, "_done", False):
            raise StopIteration
        item = self._q.get()
        if item is None:
            self._done = True
            raise StopIteration
        if isinstance(item, Exception):
            self._done = True
            raise item
        return item

    def __len__(self):
        return self._n

    @property
    def num_samples(self) -> int:
        return len(self.keys)


# ----------------------------------------------------------------------------------------- native pipeline (csrc/rt/recordio.cpp)
def _rtlib():
    import ctypes
    from ctypes import POINTER, c_char_p, c_int, c_long, c_longlong, c_size_t, c_ulonglong, c_void_p
    from . import rt
    L = rt.lib()
    if not getattr(L, "_recio_declared", False):
        L.dlcfn_rec_open.argtypes = [c_char_p, c_char_p]
        L.dlcfn_rec_open.restype = c_void_p
        L.dlcfn_rec_num.argtypes = [c_void_p]
        L.dlcfn_rec_num.restype = c_long
        L.dlcfn_rec_key.argtypes = [c_void_p, c_long]
        L.dlcfn_rec_key.restype = c_longlong
        L.dlcfn_rec_read.argtypes = [c_void_p, c_long, c_void_p, c_size_t]
        L.dlcfn_rec_read.restype = c_long
        L.dlcfn_rec_close.argtypes = [c_void_p]
        L.dlcfn_rec_writer_open.argtypes = [c_char_p, c_char_p]
        L.dlcfn_rec_writer_open.restype = c_void_p
        L.dlcfn_rec_write.argtypes = [c_void_p, c_longlong, c_char_p, c_size_t]
        L.dlcfn_rec_write.restype = c_int
        L.dlcfn_rec_writer_close.argtypes = [c_void_p]
        L.dlcfn_rec_writer_close.restype = c_int
        L.dlcfn_imgiter_create.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_ulonglong, c_int, c_int, c_int,
                                           c_int, c_int, c_int, c_int, c_int]
        L.dlcfn_imgiter_create.restype = c_void_p
        for f in ("dlcfn_imgiter_num_batches", "dlcfn_imgiter_num_samples", "dlcfn_imgiter_errors"):
            getattr(L, f).argtypes = [c_void_p]
            getattr(L, f).restype = c_long
        L.dlcfn_imgiter_num_slots.argtypes = [c_void_p]
        L.dlcfn_imgiter_num_slots.restype = c_int
        L.dlcfn_imgiter_slot_data.argtypes = [c_void_p, c_int]
        L.dlcfn_imgiter_slot_data.restype = c_void_p
        L.dlcfn_imgiter_next.argtypes = [c_void_p, POINTER(c_void_p), POINTER(c_void_p), POINTER(c_int)]
        L.dlcfn_imgiter_next.restype = c_int
        L.dlcfn_imgiter_reset.argtypes = [c_void_p]
        L.dlcfn_imgiter_destroy.argtypes = [c_void_p]
        L._recio_declared = True
    return L, ctypes


class NativeRecordIO:
    """The same .rec / .idx files through the C++ reader / writer (mmap reads, dmlc continuation framing)."""

    def __init__(self, idx_path: str | None, uri: str, flag: str):
        assert flag in ("r", "w")
        self.L, self._ct = _rtlib()
        self.flag = flag
        enc = lambda p: p.encode() if p else None      # noqa: E731
        self.h = (self.L.dlcfn_rec_open if flag == "r" else self.L.dlcfn_rec_writer_open)(enc(uri), enc(idx_path))
        if not self.h:
            raise IOError(f"cannot open {uri}")

    def __len__(self):
        return int(self.L.dlcfn_rec_num(self.h))

    def key(self, i: int) -> int:
        return int(self.L.dlcfn_rec_key(self.h, i))

    def read(self, i: int) -> bytes:
        n = self.L.dlcfn_rec_read(self.h, i, None, 0)
        if n < 0:
            raise IOError(f"bad record {i}")
        buf = self._ct.create_string_buffer(int(n))
        self.L.dlcfn_rec_read(self.h, i, buf, n)
        return buf.raw

    def write(self, key: int, buf: bytes) -> None:
        if self.L.dlcfn_rec_write(self.h, key, buf, len(buf)) != 0:
            raise ValueError("record too large")

    def close(self) -> None:
        if self.h:
            if self.flag == "r":
                self.L.dlcfn_rec_close(self.h)
            elif self.L.dlcfn_rec_writer_close(self.h) != 0:
                raise IOError("could not finish the record file / index")
            self.h = None


class NativeImageRecordIter:
    """`ImageRecordIter` on the native pipeline: decoder threads fill fixed-shape uint8 NHWC batches in a ring of pooled
    host buffers, in order, `prefetch_buffer` batches ahead.  Yields (x uint8 [B,H,W,C], y int64 [B]) views over the
    ring; a view stays valid for `keep` further `next()` calls (enough for an asynchronous H2D copy that is
    synchronised one step later).  With CUDA present the ring is page-locked once."""

    def __init__(self, path_imgrec: str, batch_size: int, data_shape=(3, 224, 224), path_imgidx: str | None = None,
                 shuffle: bool = False, num_parts: int = 1, part_index: int = 0, preprocess_threads: int = 4,
                 prefetch_buffer: int = 4, seed: int = 0, round_batch: bool = True, rand_crop: bool = False,
                 rand_mirror: bool = False, label_width: int = 1, keep: int = 2, pin: bool | None = None):
        self.L, self._ct = _rtlib()
        self.batch_size, self.shape, self.label_width = batch_size, tuple(data_shape), label_width
        idx = path_imgidx or os.path.splitext(path_imgrec)[0] + ".idx"
        self._rec = self.L.dlcfn_rec_open(path_imgrec.encode(), idx.encode() if os.path.exists(idx) else None)
        if not self._rec:
            raise IOError(f"cannot open {path_imgrec}")
        n0 = self.L.dlcfn_rec_read(self._rec, 0, None, 0)
        if n0 > 0:
            buf = self._ct.create_string_buffer(int(n0))
            self.L.dlcfn_rec_read(self._rec, 0, buf, n0)
            hd, body = unpack(buf.raw)
            if image_format(body[4 * int(hd.flag) if hd.flag > 0 else 0:]) != ".raw":
                self.L.dlcfn_rec_close(self._rec)
                raise ValueError(f"{path_imgrec} holds encoded (JPEG / PNG) images: the native decoder ring reads raw pixel "
                                 "records (im2rec --encoding .raw); use ImageRecordIter / image_record_iter() for encoded files")
        c, h, w = self.shape
        self._it = self.L.dlcfn_imgiter_create(self._rec, batch_size, h, w, c, label_width, int(shuffle), seed, part_index,
                                               num_parts, preprocess_threads, prefetch_buffer, int(rand_crop),
                                               int(rand_mirror), int(round_batch), keep)
        if not self._it:
            self.L.dlcfn_rec_close(self._rec)
            raise IOError("image iterator could not be created (empty shard or bad arguments)")
        self._pinned = []
        pin = torch.cuda.is_available() if pin is None else pin
        if pin and torch.cuda.is_available():
            from . import storage
            nbytes = storage.size_class(batch_size * h * w * c)
            for i in range(self.L.dlcfn_imgiter_num_slots(self._it)):
                p = self.L.dlcfn_imgiter_slot_data(self._it, i)
                if int(torch.cuda.cudart().cudaHostRegister(p, nbytes, 0)) == 0:
                    self._pinned.append(p)
        self.pad = 0
        self._first = True

    def __len__(self):
        return int(self.L.dlcfn_imgiter_num_batches(self._it))

    @property
    def num_samples(self) -> int:
        return int(self.L.dlcfn_imgiter_num_samples(self._it))

    def reset(self) -> None:
        if not self._first:
            self.L.dlcfn_imgiter_reset(self._it)
        self._first = False

    def __iter__(self):
        self.reset()
        return self

    def __next__(self):
        self._first = False
        d, l, nv = self._ct.c_void_p(), self._ct.c_void_p(), self._ct.c_int()
        if self.L.dlcfn_imgiter_next(self._it, self._ct.byref(d), self._ct.byref(l), self._ct.byref(nv)) == 0:
            raise StopIteration
        c, h, w = self.shape
        B = self.batch_size
        xb = (self._ct.c_uint8 * (B * h * w * c)).from_address(d.value)
        lb = (self._ct.c_float * (B * self.label_width)).from_address(l.value)
        x = torch.frombuffer(xb, dtype=torch.uint8).view(B, h, w, c)
        lab = torch.frombuffer(lb, dtype=torch.float32).view(B, self.label_width)
        self.pad = B - nv.value
        y = lab[:, 0].to(torch.int64) if self.label_width == 1 else lab.clone()
        return x, y

    @property
    def errors(self) -> int:
        return int(self.L.dlcfn_imgiter_errors(self._it))

    def close(self) -> None:
        if self._it:
            for p in self._pinned:
                torch.cuda.cudart().cudaHostUnregister(p)
            self._pinned = []
            self.L.dlcfn_imgiter_destroy(self._it)
            self.L.dlcfn_rec_close(self._rec)
            self._it = self._rec = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def image_record_iter(path_imgrec: str, batch_size: int, data_shape=(3, 224, 224), **kw):
    """The native pipeline when the file holds raw pixel records, the python iterator (torchvision codecs) otherwise."""
    try:
        return NativeImageRecordIter(path_imgrec, batch_size, data_shape, **kw)
    except ValueError:
        allowed = {"path_imgidx", "shuffle", "num_parts", "part_index", "preprocess_threads", "prefetch_buffer", "seed",
                   "round_batch", "rand_crop", "rand_mirror"}
        return ImageRecordIter(path_imgrec, batch_size, data_shape, **{k: v for k, v in kw.items() if k in allowed})


def rec2idx(rec_path: str, idx_path: str | None = None) -> int:
    """Rebuild the `.idx` of a `.rec` file (reference MX/tools/rec2idx.py): one `key\toffset` line per record, the key
    taken from the record's image header (`id`) when it has one, the ordinal otherwise.  Returns the record count."""
    idx_path = idx_path or os.path.splitext(rec_path)[0] + ".idx"
    r = MXRecordIO(rec_path, "r")
    n = 0
    with open(idx_path, "w") as f:
        while True:
            pos = r.tell()
            buf = r.read()
            if buf is None:
                break
            key = n
            if len(buf) >= _IR.size:
                try:
                    hid = int(unpack(buf)[0].id)
                    key = hid if hid >= 0 else n
                except Exception:
                    key = n
            f.write(f"{key}\t{pos}\n")
            n += 1
    r.close()
    return n


def main(argv=None) -> int:
    import argparse
    ap = argparse.ArgumentParser(prog="rec2idx", description="Create the index file of a RecordIO file")
    ap.add_argument("record", help="path to the .rec file")
    ap.add_argument("index", nargs="?", default=None, help="path of the .idx file to write (default: next to the .rec)")
    a = ap.parse_args(argv)
    n = rec2idx(a.record, a.index)
    print(f"{a.record}: {n} records indexed")
    return 0
