"""`im2rec` - turn a directory of images into `.lst` / `.rec` / `.idx` files (reference MX/tools/im2rec.py).

    python -m deeplearning_cfn_b200 im2rec --list prefix root          # write prefix.lst: index \t label \t relative path
    python -m deeplearning_cfn_b200 im2rec prefix root --resize 256     # read prefix.lst, write prefix.rec + prefix.idx

One sub-directory per class (sorted -> label ids) like `--recursive` in the reference; `--train-ratio` / `--test-ratio`
split the list; `--shuffle`; `--resize N` scales the shorter edge to N and `--center-crop` crops the square.  `--encoding
.raw` (default) stores uint8 HWC pixels preceded by their shape - what the native decoder ring (csrc/rt/recordio.cpp) reads
with no codec in the loop (this image has no libjpeg / libpng headers to link it against); `--encoding .jpg --quality Q` /
`.png` write the reference's encoded payloads, read back through torchvision's codecs by `ImageRecordIter`
(`recordio.image_record_iter` picks the reader from the file's first record).
"""
from __future__ import annotations

import argparse
import os
import random
import sys

import numpy as np

from . import recordio

EXTS = (".jpg", ".jpeg", ".png", ".bmp", ".npy")


def list_images(root: str, recursive: bool = True, exts=EXTS):
    """Yields (index, relative path, label); labels are the sorted class-directory ids (im2rec.py:38-73)."""
    i = 0
    if recursive:
        cat = {}
        for path, dirs, files in os.walk(root, followlinks=True):
            dirs.sort()
            files.sort()
            for fname in files:
                if os.path.splitext(fname)[1].lower() in exts:
                    rel_dir = os.path.relpath(path, root)
                    if rel_dir not in cat:
                        cat[rel_dir] = len(cat)
                    yield i, os.path.relpath(os.path.join(path, fname), root), cat[rel_dir]
                    i += 1
    else:
        for fname in sorted(os.listdir(root)):
            if os.path.splitext(fname)[1].lower() in exts:
                yield i, fname, 0
                i += 1


def write_list(path_out: str, image_list) -> None:
    with open(path_out, "w") as f:
        for idx, rel, label in image_list:
            f.write(f"{idx}\t{float(label)}\t{rel}\n")


def read_list(path_in: str):
    with open(path_in) as f:
        for line in f:
            parts = line.rstrip("\n").split("\t")
            if len(parts) < 3:
                continue
            yield int(parts[0]), parts[-1], [float(v) for v in parts[1:-1]]


def make_lists(prefix: str, root: str, shuffle: bool = True, train_ratio: float = 1.0, test_ratio: float = 0.0,
               recursive: bool = True, seed: int = 100) -> list[str]:
    items = list(list_images(root, recursive))
    if shuffle:
        random.Random(seed).shuffle(items)
    n = len(items)
    n_test = int(n * test_ratio)
    n_train = int(n * train_ratio)
    out = []
    if test_ratio > 0:
        write_list(prefix + "_test.lst", items[:n_test])
        out.append(prefix + "_test.lst")
    if train_ratio + test_ratio < 1.0:
        write_list(prefix + "_train.lst", items[n_test:n_test + n_train])
        write_list(prefix + "_val.lst", items[n_test + n_train:])
        out += [prefix + "_train.lst", prefix + "_val.lst"]
    elif test_ratio > 0:
        write_list(prefix + "_train.lst", items[n_test:])
        out.append(prefix + "_train.lst")
    else:
        write_list(prefix + ".lst", items)
        out.append(prefix + ".lst")
    return out


def load_image(path: str, resize: int = 0, center_crop: bool = False, color: int = 1) -> np.ndarray:
    if path.endswith(".npy"):
        img = np.load(path)
    else:
        from PIL import Image
        im = Image.open(path).convert("RGB" if color == 1 else "L")
        if resize:
            w, h = im.size
            s = resize / min(w, h)
            im = im.resize((max(1, round(w * s)), max(1, round(h * s))), Image.BILINEAR)
        img = np.asarray(im, dtype=np.uint8)
    if img.ndim == 2:
        img = img[:, :, None]
    if center_crop:
        h, w = img.shape[:2]
        m = min(h, w)
        y0, x0 = (h - m) // 2, (w - m) // 2
        img = img[y0:y0 + m, x0:x0 + m]
    return np.ascontiguousarray(img, dtype=np.uint8)


def write_record(lst_path: str, root: str, resize: int = 0, center_crop: bool = False, color: int = 1,
                 pack_label: bool = False, encoding: str = ".raw", quality: int = 95) -> tuple[str, int]:
    base = os.path.splitext(lst_path)[0]
    rec = recordio.MXIndexedRecordIO(base + ".idx", base + ".rec", "w")
    n = 0
    for idx, rel, labels in read_list(lst_path):
        img = load_image(os.path.join(root, rel), resize, center_crop, color)
        if pack_label and len(labels) > 1:
            import struct
            header = recordio.IRHeader(len(labels), 0.0, idx, 0)
            enc = recordio.pack_img(recordio.IRHeader(0, 0.0, idx, 0), img, quality, encoding)[recordio._IR.size:]
            rec.write_idx(idx, recordio.pack(header, struct.pack(f"<{len(labels)}f", *labels) + enc))
        else:
            rec.write_idx(idx, recordio.pack_img(recordio.IRHeader(0, labels[0] if labels else 0.0, idx, 0), img, quality,
                                                 encoding))
        n += 1
    rec.close()
    return base + ".rec", n


def main(argv=None) -> int:
    p = argparse.ArgumentParser(prog="im2rec", description="Create an image list or a RecordIO file from an image folder")
    p.add_argument("prefix", help="prefix of the .lst / .rec / .idx files")
    p.add_argument("root", help="folder that holds the images (one sub-folder per class)")
    p.add_argument("--list", action="store_true", help="create the image list instead of the record file")
    p.add_argument("--no-recursive", action="store_true")
    p.add_argument("--no-shuffle", action="store_true")
    p.add_argument("--train-ratio", type=float, default=1.0)
    p.add_argument("--test-ratio", type=float, default=0.0)
    p.add_argument("--resize", type=int, default=0, help="resize the shorter edge to this many pixels")
    p.add_argument("--center-crop", action="store_true")
    p.add_argument("--color", type=int, default=1, choices=[0, 1])
    p.add_argument("--pack-label", action="store_true", help="store every label column of the list (multi-label records)")
    p.add_argument("--encoding", default=".raw", choices=[".raw", ".jpg", ".png"],
                   help=".raw feeds the native decoder ring; .jpg / .png are the reference's formats (decoded through torchvision)")
    p.add_argument("--quality", type=int, default=95, help="JPEG quality 1-100 / PNG compression 1-9")
    a = p.parse_args(argv)
    if a.list:
        for f in make_lists(a.prefix, a.root, not a.no_shuffle, a.train_ratio, a.test_ratio, not a.no_recursive):
            print(f)
        return 0
    d, b = os.path.split(a.prefix)
    lists = sorted(os.path.join(d or ".", f) for f in os.listdir(d or ".") if f.startswith(b) and f.endswith(".lst"))
    if not lists:
        print(f"no {a.prefix}*.lst found - run with --list first", file=sys.stderr)
        return 1
    for lst in lists:
        rec, n = write_record(lst, a.root, a.resize, a.center_crop, a.color, a.pack_label, a.encoding, a.quality)
        print(f"{rec}: {n} records")
    return 0


if __name__ == "__main__":
    sys.exit(main())
