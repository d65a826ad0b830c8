"""Flat, bucketed parameter/gradient storage.

The reference moves one message per parameter (157 keys for ResNet-50, 104 of them <= 4096 elements;
SURVEY Appendix E) and relies on `priority=-index` for ordering (model.py:145-155).  Here all
parameters of one dtype live in ONE contiguous weight buffer and ONE contiguous gradient buffer
(inside the symmetric peer-HBM heap when there is more than one rank), laid out in reverse
registration order so that the gradients that become ready first sit at the lowest offsets, and cut
into buckets (default 16 MiB; `HOROVOD_FUSION_THRESHOLD` is honoured, run.sh:71).  A bucket is a
contiguous byte range, so a single fused kernel handles it with no pack/unpack copies.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import torch
import torch.nn as nn

ALIGN_BYTES = 128


@dataclass
class Bucket:
    index: int
    group: "Group"
    start: int          # element offset in the group's flat buffers
    end: int
    params: list = field(default_factory=list)
    pending: int = 0
    gathered: bool = False   # this step's gradients have been moved into the flat buffer

    @property
    def numel(self) -> int:
        return self.end - self.start


@dataclass
class Group:
    dtype: torch.dtype
    params: list
    offsets: list       # element offset of each param
    numel: int          # padded total
    weight: torch.Tensor = None
    grad: torch.Tensor = None
    weight_off: int = -1   # byte offsets in the symmetric heap (or -1)
    grad_off: int = -1
    buckets: list = field(default_factory=list)

    @property
    def esize(self) -> int:
        return 2 if self.dtype in (torch.bfloat16, torch.float16) else 4


def default_bucket_bytes() -> int:
    v = os.environ.get("DLCFN_BUCKET_BYTES") or os.environ.get("HOROVOD_FUSION_THRESHOLD")
    return int(v) if v else 16 << 20


class FlatModel:
    """Re-homes a module's parameters and gradients into flat per-dtype buffers."""

    def __init__(self, model: nn.Module, heap=None, bucket_bytes: int | None = None):
        self.model = model
        self.heap = heap
        bucket_bytes = bucket_bytes or default_bucket_bytes()
        params = [p for p in model.parameters() if p.requires_grad]
        if not params:
            raise ValueError("model has no trainable parameters")
        self.device = params[0].device
        by_dtype: dict = {}
        for p in reversed(params):  # reverse registration order ~ gradient-ready order
            by_dtype.setdefault(p.dtype, []).append(p)
        self.groups: list[Group] = []
        self.buckets: list[Bucket] = []
        self._grad_view: dict = {}
        # largest group first so bucket 0.. belong to the bulk (bf16) weights
        for dtype, plist in sorted(by_dtype.items(), key=lambda kv: -sum(p.numel() for p in kv[1])):
            esz = 2 if dtype in (torch.bfloat16, torch.float16) else 4
            align = ALIGN_BYTES // esz
            offs, cur = [], 0
            for p in plist:
                offs.append(cur)
                cur += (p.numel() + align - 1) // align * align
            g = Group(dtype, plist, offs, cur)
            if heap is not None:
                g.weight_off, g.weight = heap.alloc_tensor(cur, dtype)
                g.grad_off, g.grad = heap.alloc_tensor(cur, dtype)
            else:
                g.weight = torch.empty(cur, dtype=dtype, device=self.device)
                g.grad = torch.empty(cur, dtype=dtype, device=self.device)
            g.weight.zero_()
            g.grad.zero_()
            for p, o in zip(plist, offs):
                wv = g.weight[o:o + p.numel()].as_strided(p.shape, p.stride())
                wv.copy_(p.data)
                p.data = wv
                gv = g.grad[o:o + p.numel()].as_strided(p.shape, p.stride())
                p.grad = gv
                self._grad_view[id(p)] = gv
            # buckets: contiguous runs of parameters
            bstart, bparams = 0, []
            for i, (p, o) in enumerate(zip(plist, offs)):
                bparams.append(p)
                nxt = offs[i + 1] if i + 1 < len(plist) else cur
                if (nxt - bstart) * esz >= bucket_bytes or i + 1 == len(plist):
                    b = Bucket(len(self.buckets), g, bstart, nxt, bparams, len(bparams))
                    g.buckets.append(b)
                    self.buckets.append(b)
                    bstart, bparams = nxt, []
            self.groups.append(g)
        self._param_bucket = {id(p): b for b in self.buckets for p in b.params}

    def bucket_of(self, p) -> Bucket:
        return self._param_bucket[id(p)]

    def zero_grad(self) -> None:
        for g in self.groups:
            g.grad.zero_()

    def reset_pending(self) -> None:
        for b in self.buckets:
            b.pending = len(b.params)
            b.gathered = False

    def release_grads(self) -> None:
        """Start of a step: give the `.grad` slots back to autograd.  With `p.grad is None` the engine hands each
        parameter its freshly computed gradient tensor as is (no `grad += g` read-modify-write kernel per parameter, no
        memset of the flat buffer); `gather_bucket` then moves a whole bucket into the flat buffer with one fused
        multi-tensor copy when its last gradient has arrived."""
        for g in self.groups:
            for p in g.params:
                p.grad = None

    def gather_bucket(self, b: Bucket) -> None:
        dst, src = [], []
        for p in b.params:
            v = self._grad_view[id(p)]
            if p.grad is None:
                v.zero_()                       # parameter took no part in this step
            elif p.grad.data_ptr() != v.data_ptr():
                dst.append(v)
                src.append(p.grad)
        if dst:
            torch._foreach_copy_(dst, src)
        b.gathered = True

    def total_bytes(self) -> int:
        return sum(g.numel * g.esize for g in self.groups)

    def state_tensors(self) -> dict:
        return {f"group{i}_{str(g.dtype).split('.')[-1]}": g.weight for i, g in enumerate(self.groups)}
