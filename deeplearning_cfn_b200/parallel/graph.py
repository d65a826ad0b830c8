"""A whole training step as ONE CUDA graph (validated on B200 in round 2: 22.18 -> 21.11 ms per ResNet-50 step at batch 256 on
one GPU, 22.56 -> 21.53 ms on eight; bench.py replays it by default).

The reference hides per-operator launch cost with its dependency engine's "bulk execution" of graph segments
(`MXNET_EXEC_BULK_EXEC_TRAIN`, MX/src/executor/graph_executor.cc:1339-1398) and CachedOp static allocation
(MX/src/imperative/cached_op.cc).  The B200 equivalent is a captured CUDA graph: forward, backward (autograd runs under
capture), the per-bucket gradient gathers, the fused collectives on the communication stream (their flag epochs live in
device memory, so a replay continues the protocol where the previous one stopped) and the fused optimizer are recorded once
and replayed with a single launch.  A ResNet-50 step at batch 256 is device-bound (8.9 ms of host work under 21.6 ms of
kernels, tools/host_overhead.py), yet the replay still removes ~1 ms of gaps between its ~600 kernels; the small-batch CIFAR
recipes of the reference (W1 / W3 / W4: batch 32 - 128 of 32x32 images) are launch-bound and gain the most.

    step = trainer.capture(x, y)          # 3 warm-up steps on a side stream (they DO train), then capture
    for x, y in loader: loss = step(x, y) # copies into the static inputs, replays; `loss` is a static tensor

What a graph freezes, and how it is handled:
  * kernel ARGUMENTS computed on the host - the learning rate, weight decay, momentum ...: `GraphedStep` keeps the tuple
    it was captured with and re-captures when `trainer.set_learning_rate` (or any hyper-parameter) changed it.  Optimizers
    whose effective rate changes EVERY step (Adam / Adamax fold the bias correction into lr on the host) are not graphed:
    `capture` returns an eager pass-through for them.
  * shapes and addresses - inputs are copied into static buffers; a batch of another shape runs eagerly.
  * host-side counters (step_count, dlcfn launch counter) - advanced by hand on every replay.
On CPU tensors, or with the debugging engine (MXNET_ENGINE_TYPE=NaiveEngine synchronises inside the step), `capture`
returns the eager pass-through as well, so callers need no second code path.
"""
from __future__ import annotations

import torch

from ..ops import native

_UNGRAPHED_OPTIMIZERS = ("adam", "adamax")


def _map(x, fn):
    if isinstance(x, (tuple, list)):
        return type(x)(_map(v, fn) for v in x)
    return fn(x) if torch.is_tensor(x) else x


def _flatten(x, out):
    if isinstance(x, (tuple, list)):
        for v in x:
            _flatten(v, out)
    elif torch.is_tensor(x):
        out.append(x)
    return out


class GraphedStep:
    def __init__(self, trainer, x, y, warmup: int = 3):
        self.trainer = trainer
        self.warmup = int(warmup)
        self.graph = None
        self.static_loss = None
        self.replays = 0
        self.captures = 0
        self.reason = None                      # why the step is not graphed (None when it is)
        if not trainer.on_gpu:
            self.reason = "parameters are not on a CUDA device"
        elif trainer.naive:
            self.reason = "NaiveEngine debugging mode synchronises inside the step"
        elif trainer.opt_kind in _UNGRAPHED_OPTIMIZERS:
            self.reason = f"{trainer.opt_kind}: the bias-corrected learning rate is a host-computed kernel argument that changes every step"
        if self.reason is None:
            self.sx = _map(x, lambda t: t.clone())
            self.sy = _map(y, lambda t: t.clone())
            self._sig_in = self._signature_inputs(x, y)
            self._capture()

    # ------------------------------------------------------------------ helpers
    @property
    def graphed(self) -> bool:
        return self.graph is not None

    def _hp_signature(self):
        t = self.trainer
        return (t._lr(),) + tuple(sorted(t.hp.items())) + (t.grad_scale,)

    @staticmethod
    def _signature_inputs(x, y):
        return tuple((tuple(t.shape), t.dtype, t.stride()) for t in _flatten((x, y), []))

    def _capture(self, warmup: int | None = None) -> None:
        t = self.trainer
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):           # warm-up off the capture stream: cuDNN autotune, lazy context binds in the
            for _ in range(self.warmup if warmup is None else warmup):   # autograd threads, tensor-map caches, first-use
                t.step(self.sx, self.sy)                                 # cudaFuncSetAttribute
        cur.wait_stream(side)
        torch.cuda.synchronize(t.device)
        if t.world > 1:
            import torch.distributed as dist
            dist.barrier()                      # every rank captures the same step sequence
        launches0, step0 = native.launch_count, t.step_count
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.static_loss = t.step(self.sx, self.sy)
        # capture records, it does not execute: undo its host-side bookkeeping
        self.launches_per_step = native.launch_count - launches0
        native.launch_count = launches0
        t.step_count = step0
        self.graph = g
        self._sig_hp = self._hp_signature()
        self.captures += 1

    # ------------------------------------------------------------------ call
    def __call__(self, x=None, y=None):
        t = self.trainer
        if self.graph is None:
            return t.step(x, y)
        if t.is_worker and self._signature_inputs(x, y) != self._sig_in:
            return t.step(x, y)                 # e.g. the last, smaller batch of an epoch
        if self._hp_signature() != self._sig_hp:
            self.graph = None                   # learning-rate schedule moved: record again with the new arguments
            self._capture(warmup=0)             # (everything is warm: a re-capture must not train extra steps)
        if t.is_worker:
            for dst, src in zip(_flatten((self.sx, self.sy), []), _flatten((x, y), [])):
                if dst.data_ptr() != src.data_ptr():
                    dst.copy_(src, non_blocking=True)
        self.graph.replay()
        t.step_count += 1
        native.launch_count += self.launches_per_step
        self.replays += 1
        return self.static_loss

    def __repr__(self):
        state = f"graphed, {self.launches_per_step} dlcfn launches/step, {self.replays} replays, {self.captures} captures" \
            if self.graph is not None else f"eager ({self.reason})"
        return f"GraphedStep({state})"
