"""Single-process multi-device helpers and the remaining parallelism strategies of the reference inventory
(SURVEY §2.8): batch slicing across local devices (`gluon.utils.split_and_load`, utils.py:42-91;
`DataParallelExecutorGroup.decide_slices`, executor_group.py:281), layer-wise model-parallel placement
(`ctx_group` / `group2ctx` with automatic cross-device copies, graph_executor.cc:908-1638) and cross-device
synchronised BatchNorm (contrib/sync_batch_norm-inl.h)."""
from __future__ import annotations

import torch
import torch.distributed as dist
import torch.nn as nn


def split_data(data: torch.Tensor, num_slice: int, batch_axis: int = 0, even_split: bool = True) -> list[torch.Tensor]:
    size = data.shape[batch_axis]
    if even_split and size % num_slice != 0:
        raise ValueError(f"data with shape {tuple(data.shape)} cannot be evenly split into {num_slice} slices along axis "
                         f"{batch_axis}. Use a batch size that's multiple of {num_slice} or set even_split=False")
    step = size // num_slice
    if step == 0:
        raise ValueError("too many slices for this batch")
    out = []
    for i in range(num_slice):
        lo, hi = i * step, (i + 1) * step if i < num_slice - 1 else size
        out.append(data.narrow(batch_axis, lo, hi - lo))
    return out


def split_and_load(data: torch.Tensor, ctx_list: list, batch_axis: int = 0, even_split: bool = True) -> list[torch.Tensor]:
    """Split a batch along `batch_axis` and put slice i on device ctx_list[i]."""
    if len(ctx_list) == 1:
        return [data.to(ctx_list[0], non_blocking=True)]
    return [s.to(d, non_blocking=True) for s, d in zip(split_data(data, len(ctx_list), batch_axis, even_split), ctx_list)]


class ModelParallelSequential(nn.Module):
    """Layer-wise model parallelism: stage i lives on devices[i]; activations are copied across devices between
    stages (the reference inserts `_CrossDeviceCopy` nodes for `ctx_group` boundaries).  Single machine only, like
    the reference (docs/faq/distributed_training.md:13-15)."""

    def __init__(self, stages: list[nn.Module], devices: list):
        super().__init__()
        if len(stages) != len(devices):
            raise ValueError("one device per stage")
        self.stages = nn.ModuleList(s.to(d) for s, d in zip(stages, devices))
        self.devices = [torch.device(d) for d in devices]

    def forward(self, x):
        for s, d in zip(self.stages, self.devices):
            x = s(x.to(d, non_blocking=True))
        return x


_SYNC_HEAP = {"heap": None, "off": 0, "buf": None}


def set_sync_bn_heap(heap) -> None:
    """Route the SyncBatchNorm statistics exchange (reference C12: 2C+1 floats per layer, forward and backward) through the
    fused one-shot all-reduce kernel on the symmetric peer-HBM heap (csrc/cuda/collectives.cu:kallreduce_oneshot - push into
    every peer's staging slot, ONE flag barrier, local reduce) instead of an NCCL call.  `heap` = a parallel.symm.SymmetricHeap
    shared by all ranks of the process group (pass None to go back to torch.distributed).

    One heap serves ONE stream of small all-reduces: the one-shot kernel has a single staging area and flag channel per heap, so
    give SyncBatchNorm a heap of its own (`SymmetricHeap(1 << 20)`) rather than the heap of a `DataParallelTrainer` whose
    communication stream may be all-reducing a small bucket at the same time."""
    if heap is None:
        _SYNC_HEAP.update(heap=None, off=0, buf=None)
        return
    off, buf = heap.alloc_tensor(heap.ONESHOT_MAX // 4, torch.float32)
    _SYNC_HEAP.update(heap=heap, off=off, buf=buf)


def _exchange_sum(vec: torch.Tensor, group) -> torch.Tensor:
    """Sum of a small fp32 vector over the ranks."""
    h = _SYNC_HEAP["heap"]
    if h is not None and vec.device == _SYNC_HEAP["buf"].device and vec.numel() <= _SYNC_HEAP["buf"].numel() and group is None:
        buf = _SYNC_HEAP["buf"][: vec.numel()]
        buf.copy_(vec)
        h.allreduce_(_SYNC_HEAP["off"], vec.numel(), torch.float32, 1.0)
        return buf.clone()
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(vec, group=group)
    return vec


class _SyncBNFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps, group):
        dims = [0] + list(range(2, x.dim()))
        xf = x.float()
        n = torch.tensor([xf.numel() / xf.shape[1]], device=x.device)
        stats = _exchange_sum(torch.cat([xf.sum(dims), (xf * xf).sum(dims), n]), group)
        C = x.shape[1]
        cnt = stats[-1]
        mean = stats[:C] / cnt
        var = (stats[C:2 * C] / cnt - mean * mean).clamp_min(0)
        invstd = torch.rsqrt(var + eps)
        shape = [1, C] + [1] * (x.dim() - 2)
        xhat = (xf - mean.view(shape)) * invstd.view(shape)
        ctx.save_for_backward(xhat, weight, invstd)
        ctx.group, ctx.cnt = group, cnt
        y = xhat * weight.float().view(shape) + bias.float().view(shape)
        return y.to(x.dtype), mean, var

    @staticmethod
    def backward(ctx, dy, _dm, _dv):
        xhat, weight, invstd = ctx.saved_tensors
        dims = [0] + list(range(2, dy.dim()))
        C = dy.shape[1]
        shape = [1, C] + [1] * (dy.dim() - 2)
        dyf = dy.float()
        sums = torch.cat([dyf.sum(dims), (dyf * xhat).sum(dims)])
        local = sums.clone()
        sums = _exchange_sum(sums, ctx.group)
        mean_dy, mean_dyx = sums[:C] / ctx.cnt, sums[C:] / ctx.cnt
        dx = (weight.float() * invstd).view(shape) * (dyf - mean_dy.view(shape) - xhat * mean_dyx.view(shape))
        return dx.to(dy.dtype), local[C:].to(weight.dtype), local[:C].to(weight.dtype), None, None


class SyncBatchNorm(nn.BatchNorm2d):
    """BatchNorm whose statistics are reduced over all ranks (per-channel sums travel, 2C+1 floats per layer)."""

    def __init__(self, num_features: int, eps: float = 1e-5, momentum: float = 0.1, process_group=None):
        super().__init__(num_features, eps=eps, momentum=momentum)
        self.process_group = process_group

    def forward(self, x):
        if not self.training:
            return super().forward(x)
        y, mean, var = _SyncBNFn.apply(x, self.weight, self.bias, self.eps, self.process_group)
        with torch.no_grad():
            self.running_mean.mul_(1 - self.momentum).add_(mean, alpha=self.momentum)
            self.running_var.mul_(1 - self.momentum).add_(var, alpha=self.momentum)
        return y
