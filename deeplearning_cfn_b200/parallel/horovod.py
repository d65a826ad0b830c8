"""Horovod-style API: `hvd.init()`, `rank/size/local_rank`, `allreduce`, `broadcast_parameters`,
`DistributedOptimizer`.

The reference's third example family (README.md:145-163; examples/distributed-tensorflow/run.sh:70-95) runs
TensorFlow under Horovod with `HOROVOD_FUSION_THRESHOLD=67108864`, `HOROVOD_CYCLE_TIME=0.5`; Horovod itself
is not in the tree.  This module gives the same programming model on B200: gradients are fused into
buckets (threshold honoured) and each bucket is all-reduced by ONE fused scale+all-reduce kernel over peer
HBM (csrc/cuda/collectives.cu:kallreduce), on a side stream, as soon as its last gradient is produced -
no negotiation thread and no cycle time are needed inside one NVSwitch domain.
Rendezvous: torchrun / dlcfn mpirun env (RANK, WORLD_SIZE, MASTER_*), OpenMPI env (OMPI_COMM_WORLD_*), or the
dlcfn all-reduce tracker (DMLC_TRACKER_URI/PORT, launch -s 0).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

_state = {"initialized": False, "tracker": None, "tracker_rank": None}


def init(backend: str | None = None) -> None:
    if _state["initialized"]:
        return
    env = os.environ
    if "RANK" not in env and "OMPI_COMM_WORLD_RANK" in env:
        env["RANK"], env["WORLD_SIZE"] = env["OMPI_COMM_WORLD_RANK"], env["OMPI_COMM_WORLD_SIZE"]
        env.setdefault("LOCAL_RANK", env.get("OMPI_COMM_WORLD_LOCAL_RANK", "0"))
    if "RANK" not in env and "DMLC_TRACKER_URI" in env:
        from ..launch.tracker import TrackerClient
        c = TrackerClient()
        info = c.start(world=int(env.get("DMLC_NUM_WORKER", "-1")), jobid=env.get("DMLC_TASK_ID", "NULL"))
        env["RANK"], env["WORLD_SIZE"] = str(info["rank"]), str(info["world"])
        env["MASTER_ADDR"], env["MASTER_PORT"] = info["master_addr"], str(info["master_port"])
        env.setdefault("LOCAL_RANK", str(info["rank"]))
        _state["tracker"], _state["tracker_rank"] = c, info["rank"]
    world = int(env.get("WORLD_SIZE", "1"))
    use_cuda = torch.cuda.is_available()
    if use_cuda:
        lr = int(env.get("LOCAL_RANK", "0"))
        ndev = torch.cuda.device_count()
        torch.cuda.set_device(lr % ndev)
    if world > 1 and not dist.is_initialized():
        env.setdefault("MASTER_ADDR", "127.0.0.1")
        env.setdefault("MASTER_PORT", "29500")
        dist.init_process_group(backend or ("nccl" if use_cuda else "gloo"))
    _state["initialized"] = True


def shutdown() -> None:
    if _state["tracker"] is not None:
        try:
            _state["tracker"].shutdown(_state["tracker_rank"])
        except OSError:
            pass
        _state["tracker"] = None
    if dist.is_initialized():
        dist.destroy_process_group()
    _state["initialized"] = False


def size() -> int:
    return dist.get_world_size() if dist.is_initialized() else 1


def rank() -> int:
    return dist.get_rank() if dist.is_initialized() else 0


def local_rank() -> int:
    return int(os.environ.get("LOCAL_RANK", "0"))


def local_size() -> int:
    return int(os.environ.get("LOCAL_WORLD_SIZE", str(torch.cuda.device_count() or 1)))


_heap = None


def _scratch_heap(nbytes: int):
    global _heap
    from .symm import SymmetricHeap
    if _heap is None or _heap.heap_bytes < nbytes:
        _heap = SymmetricHeap(max(nbytes, 256 << 20))
    return _heap


def allreduce(tensor: torch.Tensor, average: bool = True, name: str | None = None) -> torch.Tensor:
    """Out-of-place all-reduce of one tensor (hvd.allreduce).  CUDA float32/bfloat16 tensors go through the
    fused peer-HBM kernel, everything else through the process group."""
    out = tensor.clone()
    if size() == 1:
        return out
    if out.is_cuda and out.dtype in (torch.float32, torch.bfloat16):
        h = _scratch_heap(out.numel() * out.element_size() + 4096)
        view = h.view(0, out.numel(), out.dtype)
        view.copy_(out.reshape(-1))
        h.allreduce_(0, out.numel(), out.dtype, 1.0 / size() if average else 1.0)
        out.copy_(view.view(out.shape))
        return out
    dist.all_reduce(out)
    if average:
        out /= size()
    return out


def broadcast_parameters(params, root_rank: int = 0) -> None:
    """hvd.broadcast_parameters(model.state_dict(), root_rank=0)."""
    if size() == 1:
        return
    items = params.items() if hasattr(params, "items") else params
    for _, p in items:
        if isinstance(p, torch.Tensor):
            dist.broadcast(p.data if hasattr(p, "data") else p, src=root_rank)


def broadcast_optimizer_state(optimizer, root_rank: int = 0) -> None:
    if size() == 1:
        return
    sd = [optimizer.state_dict()]
    dist.broadcast_object_list(sd, src=root_rank)
    if rank() != root_rank:
        optimizer.load_state_dict(sd[0])


class DistributedOptimizer:
    """hvd.DistributedOptimizer(model=..., optimizer='sgd_mom', lr=...): fused bucketed all-reduce overlapped
    with backward + one fused multi-precision optimizer launch per dtype group."""

    def __init__(self, model: torch.nn.Module, optimizer: str = "sgd_mom", fusion_threshold: int | None = None, **hp):
        from .trainer import DataParallelTrainer
        thr = fusion_threshold or int(os.environ.get("HOROVOD_FUSION_THRESHOLD", str(16 << 20)))
        self.trainer = DataParallelTrainer(model, mode="auto", optimizer=optimizer, bucket_bytes=thr, **hp)

    def step(self, x, y):
        return self.trainer.step(x, y)

    def __getattr__(self, name):
        return getattr(self.trainer, name)
