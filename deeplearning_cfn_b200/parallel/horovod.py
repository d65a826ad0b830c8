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
    if "RANK" not in env and env.get("DMLC_ROLE") == "worker" and "DMLC_PS_ROOT_URI" in env:
        # Worker of a parameter-server job (`launch -n W -s S script`): the PS rendezvous is the only thing that knows the
        # ranks.  Join it (the servers and the scheduler are waiting for every worker anyway), take rank / world from it and
        # build the process group next to the scheduler's port.  The store is kept: `shutdown()` runs its closing protocol
        # (barrier, rank 0 stops the servers), without which the job would never end.
        from .. import kv as _kv
        store = _kv.create("dist_sync")
        env["RANK"], env["WORLD_SIZE"] = str(store.rank), str(store.num_workers)
        env.setdefault("LOCAL_RANK", str(store.rank))
        env.setdefault("MASTER_ADDR", env["DMLC_PS_ROOT_URI"])
        env.setdefault("MASTER_PORT", str(20000 + (int(env.get("DMLC_PS_ROOT_PORT", "9091")) * 7) % 20000))
        _state["ps_store"] = store
    world = int(env.get("WORLD_SIZE", "1"))
    use_cuda = torch.cuda.is_available()
    if use_cuda:
        lr = int(env.get("LOCAL_RANK", "0"))
        ndev = torch.cuda.device_count()
        torch.cuda.set_device(lr % ndev)
    if not use_cuda:
        _share_cpu_threads(env, world)
    if world > 1 and not dist.is_initialized():
        env.setdefault("MASTER_ADDR", "127.0.0.1")
        env.setdefault("MASTER_PORT", "29500")
        dist.init_process_group(backend or ("nccl" if use_cuda else "gloo"))
    _state["initialized"] = True


def _share_cpu_threads(env, world: int) -> None:
    from ..utils.cpu import share_cpu_threads
    share_cpu_threads(world, env)


def shutdown() -> None:
    if _state["tracker"] is not None:
        try:
            _state["tracker"].shutdown(_state["tracker_rank"])
        except OSError:
            pass
        _state["tracker"] = None
    if dist.is_initialized():
        dist.destroy_process_group()
    if _state.get("ps_store") is not None:
        _state.pop("ps_store").close()
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


def allreduce_(tensor: torch.Tensor, average: bool = True, name: str | None = None) -> torch.Tensor:
    """In-place form (hvd.allreduce_)."""
    tensor.copy_(allreduce(tensor, average, name))
    return tensor


def allgather(tensor: torch.Tensor, name: str | None = None) -> torch.Tensor:
    """Concatenate every rank's tensor along dim 0; the first dimension may differ between ranks (hvd.allgather)."""
    if size() == 1:
        return tensor.clone()
    n = torch.tensor([tensor.shape[0]], device=tensor.device, dtype=torch.int64)
    sizes = [torch.zeros_like(n) for _ in range(size())]
    dist.all_gather(sizes, n)
    sizes = [int(s) for s in sizes]
    mx = max(sizes)
    pad = tensor
    if tensor.shape[0] < mx:
        pad = torch.cat([tensor, tensor.new_zeros((mx - tensor.shape[0],) + tuple(tensor.shape[1:]))])
    parts = [torch.empty_like(pad) for _ in range(size())]
    dist.all_gather(parts, pad.contiguous())
    return torch.cat([p[:k] for p, k in zip(parts, sizes)])


def broadcast(tensor: torch.Tensor, root_rank: int = 0, name: str | None = None) -> torch.Tensor:
    """Out-of-place broadcast of one tensor (hvd.broadcast)."""
    out = tensor.clone()
    if size() > 1:
        dist.broadcast(out, src=root_rank)
    return out


def broadcast_(tensor: torch.Tensor, root_rank: int = 0, name: str | None = None) -> torch.Tensor:
    if size() > 1:
        dist.broadcast(tensor, src=root_rank)
    return tensor


def alltoall(tensor: torch.Tensor, splits: list[int] | None = None, name: str | None = None) -> torch.Tensor:
    """Rank r sends slice j of dim 0 to rank j (hvd.alltoall); `splits` = rows sent to each rank (default: equal)."""
    w = size()
    if w == 1:
        return tensor.clone()
    if splits is None:
        assert tensor.shape[0] % w == 0, "dim 0 must divide by the number of ranks when no splits are given"
        splits = [tensor.shape[0] // w] * w
    send = torch.tensor(splits, device=tensor.device, dtype=torch.int64)
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send) if dist.get_backend() != "gloo" else _gloo_all_to_all_counts(recv, send)
    recv_splits = [int(v) for v in recv]
    out = tensor.new_empty((sum(recv_splits),) + tuple(tensor.shape[1:]))
    if dist.get_backend() == "gloo":
        _gloo_all_to_all(out, tensor.contiguous(), recv_splits, splits)
    else:
        dist.all_to_all_single(out, tensor.contiguous(), recv_splits, splits)
    return out


def _gloo_all_to_all_counts(recv: torch.Tensor, send: torch.Tensor) -> None:
    gathered = [torch.empty_like(send) for _ in range(size())]
    dist.all_gather(gathered, send)
    for j in range(size()):
        recv[j] = gathered[j][rank()]


def _gloo_all_to_all(out: torch.Tensor, inp: torch.Tensor, recv_splits: list[int], send_splits: list[int]) -> None:
    """gloo has no all_to_all: every rank broadcasts its buffer in turn and the others cut their slice out of it."""
    all_send = [torch.empty(size(), dtype=torch.int64) for _ in range(size())]
    dist.all_gather(all_send, torch.tensor(send_splits, dtype=torch.int64))
    lens = torch.tensor([inp.shape[0]], dtype=torch.int64)
    all_len = [torch.empty_like(lens) for _ in range(size())]
    dist.all_gather(all_len, lens)
    off_out = 0
    for src in range(size()):
        buf = inp if src == rank() else inp.new_empty((int(all_len[src]),) + tuple(inp.shape[1:]))
        dist.broadcast(buf, src=src)
        start = int(all_send[src][:rank()].sum())
        k = int(all_send[src][rank()])
        out[off_out:off_out + k] = buf[start:start + k]
        off_out += k


def barrier() -> None:
    if size() > 1:
        dist.barrier()


def join() -> int:
    """hvd.join(): wait for every rank; returns the last rank to arrive (here: always size() - 1)."""
    barrier()
    return size() - 1


class Compression:
    """hvd.Compression.none / .fp16 (here bf16): wire dtype of `allreduce` for fp32 tensors."""

    class none:
        @staticmethod
        def compress(t):
            return t, None

        @staticmethod
        def decompress(t, ctx):
            return t

    class fp16:
        @staticmethod
        def compress(t):
            return (t.to(torch.bfloat16), t.dtype) if t.dtype == torch.float32 else (t, None)

        @staticmethod
        def decompress(t, ctx):
            return t.to(ctx) if ctx is not None else t


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
