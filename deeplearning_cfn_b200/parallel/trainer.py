"""Data-parallel training engine: one process per GPU, bucketed gradient exchange overlapped with
backward, fused optimizer.  It is the common machinery under the Horovod-style API
(`parallel.horovod`), the KVStore/Trainer API (`kv`) and `bench.py`.

Modes (the reference's three communication shapes + their NCCL baselines):
  allreduce  Horovod-style (reference C11, run.sh:70-95): per bucket ONE fused scale+all-reduce kernel
             over peer HBM (csrc/cuda/collectives.cu:kallreduce) on a side stream while backward keeps
             running; then ONE fused multi-precision optimizer launch per dtype group.
  ps_sync    MXNet dist_sync / dist_device_sync (reference §3.3): per bucket ONE kernel doing
             push -> server-side sum over workers -> optimizer on the fp32 master -> pull
             (collectives.cu:kps_sync).  Servers are GPUs; every bucket is split across all servers
             (the reference's big-array rule, kvstore_dist.h:547-570, applied to buckets).  Servers may
             be colocated with workers (launch.py default: num_servers == num_workers) or dedicated
             (e.g. 1 server + 7 workers).
  nccl       baseline: torch.distributed NCCL all-reduce per bucket + the same fused optimizer.
  nccl_ps    baseline with the reference KVStoreNCCL shape (kvstore_nccl.h:221-405): ncclReduce to the
             bucket's root, optimizer on the root, ncclBcast.
  hier       multi-node data parallelism (the reference's actual deployment: N nodes x G GPUs behind CloudFormation):
             per bucket, (1) the node-local fused all-reduce kernel over NVLink / NVSwitch, (2) every local rank l
             all-reduces ITS 1/G slice of the bucket with the ranks l of the other nodes (G concurrent "rail" rings over
             the inter-node fabric, NCCL), (3) node-local broadcasts hand the slices back.  Inter-node traffic per GPU is
             1/G of the bucket.  On CPU tensors the node-local steps run over gloo groups (tests).  Node size =
             LOCAL_WORLD_SIZE (torchrun) or DLCFN_VIRTUAL_NODE_SIZE (several virtual nodes on one box, for testing).
  local      single GPU / CPU, no communication.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from ..ops import native
from ..ops.nn import softmax_cross_entropy
from ..ops import reference as ref_ops
from .flat import Bucket, FlatModel
from .symm import SymmetricHeap

_HEAP_SLACK = 8 << 20


def _dist_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


class DataParallelTrainer:
    def __init__(self, model: nn.Module, *, mode: str = "auto", optimizer: str = "sgd_mom", lr: float = 0.1,
                 momentum: float = 0.9, wd: float = 1e-4, beta1: float = 0.9, beta2: float = 0.999,
                 eps: float = 1e-8, clip: float = -1.0, bucket_bytes: int | None = None,
                 server_ranks: list[int] | None = None, worker_ranks: list[int] | None = None,
                 comm_blocks: int = 0, loss_fn=None, heap: SymmetricHeap | None = None,
                 broadcast_from: int | None = 0):
        self.rank, self.world = _dist_info()
        self.device = next(model.parameters()).device
        self.on_gpu = self.device.type == "cuda"
        if mode == "auto":
            mode = "local" if self.world == 1 else ("allreduce" if self.on_gpu else "gloo")
            per_node = int(os.environ.get("DLCFN_VIRTUAL_NODE_SIZE") or os.environ.get("LOCAL_WORLD_SIZE") or self.world)
            if self.on_gpu and 0 < per_node < self.world and self.world % per_node == 0:
                mode = "hier"         # the job spans several nodes: peer HBM only reaches the GPUs of one NVSwitch domain
        if self.world == 1 and mode != "local":
            mode = "local"
        if mode == "local":
            self.rank, self.world = 0, 1      # explicit single-process trainer inside a larger job
        self.mode = mode
        self.model = model
        self.opt_kind = optimizer
        self.hp = dict(lr=lr, wd=wd, momentum=momentum, beta1=beta1, beta2=beta2, eps=eps, clip=clip)
        self.comm_blocks = comm_blocks
        self.loss_fn = loss_fn or softmax_cross_entropy     # fused loss kernels on CUDA, torch on CPU
        self.step_count = 0
        self.check_every = int(os.environ.get("DLCFN_CHECK_EVERY", "25"))     # steps between polls of the peers' error word

        self.node_size, self.node_group, self.rail_group = self.world, None, None
        if mode == "hier":
            self._init_node_topology()
            if self.node_size >= self.world:        # a single node: the plain fused all-reduce is the same thing
                mode = self.mode = "allreduce" if self.on_gpu else "gloo"
        fused = mode in ("allreduce", "ps_sync") or (mode == "hier" and self.on_gpu)
        self.heap = heap
        if mode == "hier" and self.on_gpu and heap is None:
            nbytes = sum((p.numel() * p.element_size() + 127) // 128 * 128 for p in model.parameters() if p.requires_grad)
            self.heap = SymmetricHeap(2 * nbytes + _HEAP_SLACK, group=self.node_group)
        if fused and self.heap is None:
            nbytes = 0
            for p in model.parameters():
                if p.requires_grad:
                    nbytes += (p.numel() * p.element_size() + 127) // 128 * 128
            self.heap = SymmetricHeap(2 * nbytes + _HEAP_SLACK)
        self.flat = FlatModel(model, self.heap if fused else None, bucket_bytes)

        self.workers = sorted(worker_ranks) if worker_ranks is not None else list(range(self.world))
        self.servers = sorted(server_ranks) if server_ranks is not None else list(range(self.world))
        if mode != "ps_sync":
            self.workers = list(range(self.world))
        self.is_worker = self.rank in self.workers
        self.worker_mask = sum(1 << r for r in self.workers)
        self.grad_scale = 1.0 / len(self.workers)

        # ---- initial parameter broadcast (hvd.broadcast_parameters / kv.init + first pull) ----
        if self.world > 1 and broadcast_from is not None:
            self._broadcast_initial(broadcast_from)

        # ---- optimizer state ----
        self._init_state()

        # Debugging mode (the reference's MXNET_ENGINE_TYPE=NaiveEngine): no side stream, no overlap - every
        # bucket's collective is launched after backward, in order, on the main stream, followed by a device
        # synchronise and a check of the flag-protocol error word, so a failure is attributed to the bucket that caused it.
        self.naive = (os.environ.get("MXNET_ENGINE_TYPE", "") == "NaiveEngine" or
                      os.environ.get("DLCFN_NAIVE_COLLECTIVES", "0") == "1")
        self.comm_stream = torch.cuda.Stream(priority=-1) if (self.on_gpu and self.world > 1 and not self.naive) else None
        self._ready: list[bool] = [False] * len(self.flat.buckets)
        self._next = 0
        self._works: list = []
        if self.is_worker:
            for b in self.flat.buckets:
                for p in b.params:
                    p.register_post_accumulate_grad_hook(self._make_hook(b))

    # ------------------------------------------------------------------ setup
    def _init_node_topology(self) -> None:
        """Ranks are laid out node by node (torchrun's order): node = rank // G, local rank = rank % G.  Two families of
        groups: the G ranks of a node, and for every local rank l the "rail" {l, G + l, 2G + l, ...} across nodes."""
        G = int(os.environ.get("DLCFN_VIRTUAL_NODE_SIZE") or os.environ.get("LOCAL_WORLD_SIZE") or self.world)
        if G <= 0 or self.world % G:
            raise ValueError(f"world size {self.world} is not a multiple of the node size {G}")
        self.node_size = G
        self.num_nodes = self.world // G
        self.node_index, self.local_rank = self.rank // G, self.rank % G
        if G >= self.world:
            return
        for n in range(self.num_nodes):                # every rank creates every group, in the same order
            grp = dist.new_group([n * G + l for l in range(G)])
            if n == self.node_index:
                self.node_group = grp
        for l in range(G):
            grp = dist.new_group([n * G + l for n in range(self.num_nodes)])
            if l == self.local_rank:
                self.rail_group = grp

    def _rail_slice(self, b: Bucket, l: int) -> tuple[int, int]:
        """Element range (relative to the bucket) that local rank l carries across nodes; 16-byte granules."""
        vec = 16 // b.group.esize
        nvec = (b.numel + vec - 1) // vec
        per = (nvec + self.node_size - 1) // self.node_size
        return min(l * per * vec, b.numel), min((l + 1) * per * vec, b.numel)

    def _broadcast_initial(self, root: int) -> None:
        if self.mode == "hier":
            # global root -> its rail (one rank per node), then every node fans out from that local rank
            l0, n0 = root % self.node_size, root // self.node_size
            for g in self.flat.groups:
                if self.local_rank == l0:
                    dist.broadcast(g.weight, src=root, group=self.rail_group)
                if self.heap is not None:
                    torch.cuda.synchronize(self.device)
                    self.heap.broadcast_(g.weight_off, g.numel * g.esize, l0, nblocks=self.comm_blocks or 32)
                else:
                    dist.broadcast(g.weight, src=self.node_index * self.node_size + l0, group=self.node_group)
            if self.heap is not None:
                torch.cuda.synchronize(self.device)
            for buf in self.model.buffers():
                if buf.dtype.is_floating_point and buf.numel() > 0:
                    dist.broadcast(buf, src=root)
            return
        if self.mode in ("allreduce", "ps_sync"):
            for g in self.flat.groups:
                self.heap.broadcast_(g.weight_off, g.numel * g.esize, root, nblocks=self.comm_blocks or 32)
            torch.cuda.synchronize(self.device)
        else:
            for g in self.flat.groups:
                dist.broadcast(g.weight, src=root)
        for buf in self.model.buffers():
            if buf.dtype.is_floating_point and buf.numel() > 0:
                dist.broadcast(buf, src=root)

    def _shard(self, b: Bucket) -> tuple[int, int]:
        """This rank's server shard of bucket b (element range relative to the bucket)."""
        if self.rank not in self.servers:
            return 0, 0
        vec = 16 // b.group.esize
        nvec = (b.numel + vec - 1) // vec
        per = (nvec + len(self.servers) - 1) // len(self.servers)
        j = self.servers.index(self.rank)
        lo = min(j * per, nvec) * vec
        hi = min((j + 1) * per, nvec) * vec
        return lo, hi

    def _init_state(self) -> None:
        need1 = self.opt_kind in native.OPT_USES_S1
        need2 = self.opt_kind in native.OPT_USES_S2
        self.state: dict = {}
        if self.mode in ("ps_sync", "nccl_ps"):
            for b in self.flat.buckets:
                if self.mode == "ps_sync":
                    lo, hi = self._shard(b)
                else:  # nccl_ps: whole bucket lives on its root (key -> server hash)
                    lo, hi = (0, b.numel) if self._nccl_root(b) == self.rank else (0, 0)
                n = hi - lo
                w = b.group.weight[b.start + lo: b.start + hi]
                st = {"lo": lo, "hi": hi, "master": w.float().clone() if n else None,
                      "s1": torch.zeros(n, device=self.device) if (need1 and n) else None,
                      "s2": torch.zeros(n, device=self.device) if (need2 and n) else None}
                self.state[b.index] = st
        else:
            for gi, g in enumerate(self.flat.groups):
                master = g.weight if g.dtype == torch.float32 else g.weight.float().clone()
                self.state[gi] = {"master": master,
                                  "s1": torch.zeros(g.numel, device=self.device) if need1 else None,
                                  "s2": torch.zeros(g.numel, device=self.device) if need2 else None}

    def _nccl_root(self, b: Bucket) -> int:
        return (b.index * 9973) % self.world  # reference small-key rule (key*9973) % S, kvstore_dist.h:547

    # ------------------------------------------------------------------ comm
    def _make_hook(self, b: Bucket):
        def hook(_p):
            b.pending -= 1
            if b.pending == 0:
                self.flat.gather_bucket(b)          # one fused copy: autograd-owned gradients -> flat bucket
                if self.world > 1 and not self.naive:
                    self._ready[b.index] = True
                    self._drain()
        return hook

    def _drain(self) -> None:
        while self._next < len(self._ready) and self._ready[self._next]:
            self._launch_bucket(self.flat.buckets[self._next])
            self._next += 1

    def _lr(self) -> float:
        lr = self.hp["lr"]
        if self.opt_kind == "adam":  # bias correction folded into lr (optimizer.py:1080-1082)
            t = self.step_count + 1
            lr = lr * (1.0 - self.hp["beta2"] ** t) ** 0.5 / (1.0 - self.hp["beta1"] ** t)
        elif self.opt_kind == "adamax":
            lr = lr / (1.0 - self.hp["beta1"] ** (self.step_count + 1))
        return lr

    def _launch_bucket(self, b: Bucket) -> None:
        g = b.group
        if self.is_worker and not b.gathered:
            self.flat.gather_bucket(b)              # flushed before its last hook fired (an unused parameter)
        if self.comm_stream is not None and self.is_worker:
            self.comm_stream.wait_stream(torch.cuda.current_stream())
        if self.mode == "allreduce":
            self.heap.allreduce_(g.grad_off + b.start * g.esize, b.numel, g.dtype, self.grad_scale,
                                 nblocks=self.comm_blocks, stream=self.comm_stream)
        elif self.mode == "ps_sync":
            st = self.state[b.index]
            hp = self.hp
            self.heap.ps_sync_step(g.grad_off + b.start * g.esize, g.weight_off + b.start * g.esize, b.numel,
                                   st["lo"], st["hi"], self.worker_mask, st["master"], st["s1"], st["s2"], g.dtype,
                                   self.opt_kind, lr=self._lr(), wd=hp["wd"], momentum=hp["momentum"],
                                   beta1=hp["beta1"], beta2=hp["beta2"], eps=hp["eps"], rescale=self.grad_scale,
                                   clip=hp["clip"], nblocks=self.comm_blocks, stream=self.comm_stream)
        elif self.mode == "hier":
            lo, hi = self._rail_slice(b, self.local_rank)
            bview = g.grad[b.start:b.end]
            if self.heap is not None:
                off = g.grad_off + b.start * g.esize
                # (1) node-local sum, scaled by 1 / world: the fused kernel over NVLink
                self.heap.allreduce_(off, b.numel, g.dtype, self.grad_scale, nblocks=self.comm_blocks, stream=self.comm_stream)
                with torch.cuda.stream(self.comm_stream):
                    # (2) this rank's slice across the nodes (G concurrent rails)
                    if hi > lo:
                        dist.all_reduce(bview[lo:hi], op=dist.ReduceOp.SUM, group=self.rail_group)
                # (3) the slices come back node-locally; the flag barrier inside each broadcast orders it after its
                #     root's rail all-reduce (same stream on the root)
                for l in range(self.node_size):
                    slo, shi = self._rail_slice(b, l)
                    if shi > slo:
                        self.heap.broadcast_(off + slo * g.esize, (shi - slo) * g.esize, l, nblocks=self.comm_blocks or 16,
                                             stream=self.comm_stream)
            else:                                   # host tensors (tests): the same three steps over gloo groups
                dist.all_reduce(bview, op=dist.ReduceOp.SUM, group=self.node_group)
                bview.mul_(self.grad_scale)
                if hi > lo:
                    dist.all_reduce(bview[lo:hi], op=dist.ReduceOp.SUM, group=self.rail_group)
                base = self.node_index * self.node_size
                for l in range(self.node_size):
                    slo, shi = self._rail_slice(b, l)
                    if shi > slo:
                        dist.broadcast(bview[slo:shi], src=base + l, group=self.node_group)
        elif self.mode in ("nccl", "gloo"):
            view = g.grad[b.start:b.end]
            if self.comm_stream is not None:
                with torch.cuda.stream(self.comm_stream):
                    op = dist.ReduceOp.AVG if self.mode == "nccl" else dist.ReduceOp.SUM
                    self._works.append(dist.all_reduce(view, op=op, async_op=True))
            else:
                dist.all_reduce(view, op=dist.ReduceOp.SUM)
                view.mul_(self.grad_scale)
        elif self.mode == "nccl_ps":
            root = self._nccl_root(b)
            gview = g.grad[b.start:b.end]
            wview = g.weight[b.start:b.end]
            with torch.cuda.stream(self.comm_stream):
                dist.reduce(gview, dst=root, op=dist.ReduceOp.SUM)
                if root == self.rank:
                    self._apply_opt(gview, self.state[b.index], wview if g.dtype != torch.float32 else None,
                                    rescale=self.grad_scale)
                    if g.dtype == torch.float32:
                        wview.copy_(self.state[b.index]["master"])
                dist.broadcast(wview, src=root)
        if self.naive and self.on_gpu:
            torch.cuda.synchronize(self.device)
            self.check()

    def _apply_opt(self, grad, st, wout, rescale: float = 1.0) -> None:
        hp = self.hp
        if grad.is_cuda:
            native.fused_optimizer(self.opt_kind, grad, st["master"], st["s1"], st["s2"], wout, lr=self._lr(),
                                   wd=hp["wd"], momentum=hp["momentum"], beta1=hp["beta1"], beta2=hp["beta2"],
                                   eps=hp["eps"], rescale=rescale, clip=hp["clip"])
        else:
            ref_ops.fused_optimizer(self.opt_kind, grad, st["master"], st["s1"], st["s2"], wout, lr=self._lr(),
                                    wd=hp["wd"], momentum=hp["momentum"], beta1=hp["beta1"], beta2=hp["beta2"],
                                    eps=hp["eps"], rescale=rescale, clip=hp["clip"])

    # ------------------------------------------------------------------ step
    def forward_backward(self, x, y):
        self.flat.reset_pending()
        self._ready = [False] * len(self.flat.buckets)
        self._next = 0
        self.flat.release_grads()
        out = self.model(*x) if isinstance(x, (tuple, list)) else self.model(x)
        loss = self.loss_fn(out, y)
        loss.backward()
        return loss

    def synchronize_and_update(self) -> None:
        """Flush buckets whose hooks never fired, wait for the wire, apply the optimizer."""
        if self.is_worker:
            for b in self.flat.buckets:
                if not b.gathered and self.world == 1:
                    self.flat.gather_bucket(b)
        if self.world > 1:
            for b in self.flat.buckets[self._next:]:
                self._launch_bucket(b)
            self._next = len(self.flat.buckets)
            for w in self._works:
                w.wait()
            self._works.clear()
            if self.comm_stream is not None:
                torch.cuda.current_stream().wait_stream(self.comm_stream)
        if self.mode in ("allreduce", "hier", "nccl", "gloo", "local"):
            for gi, g in enumerate(self.flat.groups):
                st = self.state[gi]
                self._apply_opt(g.grad, st, g.weight if g.dtype != torch.float32 else None)
        self.step_count += 1
        # dead / stalled peer: the device-side barriers set an error word after their spin timeout instead of hanging; poll it
        # without draining the queue (SymmetricHeap.poll_error) so a broken job stops within `check_every` steps
        if self.heap is not None and self.check_every > 0 and self.step_count % self.check_every == 0 \
                and not torch.cuda.is_current_stream_capturing():
            self.heap.poll_error()

    def step(self, x=None, y=None):
        """One full training step.  Dedicated parameter-server ranks pass no data: they only serve."""
        if not self.is_worker:
            self._next = 0
            self.synchronize_and_update()
            return None
        loss = self.forward_backward(x, y)
        self.synchronize_and_update()
        return loss

    def capture(self, x=None, y=None, warmup: int = 3):
        """The whole step as one CUDA graph (parallel/graph.py, experimental): `step = trainer.capture(x, y)`, then
        `loss = step(x, y)` per batch.  Falls back to the eager step (and says why in `step.reason`) on CPU tensors, under
        the debugging engine and for optimizers whose learning rate is recomputed on the host every step."""
        from .graph import GraphedStep
        return GraphedStep(self, x, y, warmup)

    def set_learning_rate(self, lr: float) -> None:
        self.hp["lr"] = lr

    def check(self) -> None:
        if self.heap is not None:
            self.heap.check_error()

    # ------------------------------------------------------------------ checkpoint
    def state_dict(self) -> dict:
        sd = {"step": self.step_count, "hp": dict(self.hp), "opt_kind": self.opt_kind, "mode": self.mode,
              "model": self.model.state_dict(), "opt": {}}
        for k, st in self.state.items():
            sd["opt"][k] = {n: (t.detach().cpu() if isinstance(t, torch.Tensor) else t) for n, t in st.items()}
        return sd

    def load_state_dict(self, sd: dict) -> None:
        self.step_count = sd["step"]
        self.hp.update(sd["hp"])
        self.model.load_state_dict(sd["model"])
        for k, st in sd["opt"].items():
            for n, t in st.items():
                if isinstance(t, torch.Tensor) and self.state[k].get(n) is not None:
                    self.state[k][n].copy_(t)
