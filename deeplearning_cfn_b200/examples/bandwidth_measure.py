"""Push/pull and all-reduce bandwidth micro-benchmark.

Reference: `MX/tools/bandwidth/measure.py` - init one key per weight of a chosen network, push random
gradients from every GPU, pull, time it, verify against the expected sum and report
`size * 2 * (n-1)/n / t` GB/s per GPU (:76-143); CLI flags `--network --gpus --num-layers --kv-store
--num-batches --disp-batches --test-results --image-shape --num-classes --optimizer --gc-type` (:34-60).

Here one process drives one GPU (start with torchrun / `dlcfn mpirun`), and `--kv-store` selects the path:
  allreduce   fused scale+all-reduce kernel over peer HBM      (Horovod path)
  ps          fused push/reduce/optimizer/pull kernel          (dist_device_sync path, sharded servers)
  nccl        torch.distributed all_reduce                     (baseline)
  nccl_ps     NCCL reduce to root + broadcast                  (baseline with the reference KVStoreNCCL shape)
`--sweep` measures power-of-4 sizes from 1 KiB to 1 GiB instead of a network's parameter list (BASELINE config 5).
Every number is device-timed (CUDA events) and reduced with MAX over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch
import torch.distributed as dist


def parse(argv=None):
    p = argparse.ArgumentParser(description="command for benchmark kv-store")
    p.add_argument("--network", type=str, default="resnet")
    p.add_argument("--gpus", type=str, default="", help="ignored: one process per GPU, world size from the launcher")
    p.add_argument("--num-layers", type=int, default=50)
    p.add_argument("--kv-store", type=str, default="allreduce")
    p.add_argument("--num-batches", type=int, default=5)
    p.add_argument("--disp-batches", type=int, default=1)
    p.add_argument("--test-results", type=int, default=1)
    p.add_argument("--image-shape", type=str, default="3,224,224")
    p.add_argument("--num-classes", type=int, default=1000)
    p.add_argument("--optimizer", type=str, default="None")
    p.add_argument("--gc-type", type=str, default="none")
    p.add_argument("--dtype", type=str, default="bfloat16", choices=["bfloat16", "float32"])
    p.add_argument("--sweep", action="store_true")
    p.add_argument("--max-bytes", type=int, default=1 << 30)
    p.add_argument("--out", type=str, default="")
    return p.parse_args(argv)


def network_shapes(name: str, num_layers: int, num_classes: int, image_shape: str) -> list[int]:
    from ..models import get_model
    size = int(image_shape.split(",")[-1])
    m = get_model(name, num_layers=num_layers, num_classes=num_classes, image_size=size)
    return [p.numel() for p in m.parameters() if p.requires_grad]


class Runner:
    def __init__(self, kind: str, dtype: torch.dtype, max_elems: int, total_elems: int | None = None):
        from ..parallel.symm import SymmetricHeap
        self.kind, self.dtype = kind, dtype
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.dev = torch.device("cuda", torch.cuda.current_device())
        esz = 2 if dtype == torch.bfloat16 else 4
        self.esz = esz
        self.heap = None
        cap = (total_elems or max_elems) + 64
        if kind in ("allreduce", "ps"):
            self.heap = SymmetricHeap(2 * (cap * esz + 4096) + (1 << 20))
            self.goff, self.g = self.heap.alloc_tensor(cap, dtype)
            self.woff, self.w = self.heap.alloc_tensor(cap, dtype)
            self.master = torch.zeros(max_elems // self.world + 64, device=self.dev)
        else:
            self.g = torch.empty(cap, dtype=dtype, device=self.dev)
            self.w = torch.empty(cap, dtype=dtype, device=self.dev)

    def fill(self, n: int, it: int, at: int = 0) -> None:
        if at + n > self.g.numel():
            raise IndexError(f"array [{at}, {at + n}) does not fit the {self.g.numel()}-element arena")
        self.g[at:at + n] = float(self.rank + 1 + it % 3)

    def expected(self, n: int, it: int) -> float:
        return float(sum(r + 1 + it % 3 for r in range(self.world)))

    def run(self, n: int, at: int = 0) -> torch.Tensor:
        """One push+pull (or all-reduce) of n elements at element offset `at` (a multiple of 64); returns a view of
        the reduced values."""
        k = self.kind
        b = at * self.esz
        if at % 64 or at + n > self.g.numel():           # a region past the arena would be reduced out of bounds on the peers
            raise IndexError(f"array [{at}, {at + n}) does not fit the {self.g.numel()}-element arena (offsets are multiples of 64)")
        if k == "allreduce":
            self.heap.allreduce_(self.goff + b, n, self.dtype, 1.0)
            return self.g[at:at + n]
        if k == "ps":
            vec = 16 // self.esz
            nvec = (n + vec - 1) // vec
            per = (nvec + self.world - 1) // self.world
            lo, hi = min(self.rank * per, nvec) * vec, min((self.rank + 1) * per, nvec) * vec
            self.heap.ps_sync_step(self.goff + b, self.woff + b, n, lo, hi, (1 << self.world) - 1, self.master, None,
                                   None, self.dtype, "none", lr=0.0)
            return self.w[at:at + n]
        if k == "nccl":
            dist.all_reduce(self.g[at:at + n])
            return self.g[at:at + n]
        if k == "nccl_ps":
            dist.reduce(self.g[at:at + n], dst=0)
            if self.rank == 0:
                self.w[at:at + n].copy_(self.g[at:at + n])
            dist.broadcast(self.w[at:at + n], src=0)
            return self.w[at:at + n]
        raise ValueError(k)


def measure(r: Runner, sizes: list[int], num_batches: int, check: bool, repeat: int = 1) -> tuple[float, float]:
    """Returns (seconds per batch [max over ranks], max abs error)."""
    err = 0.0
    offs, cur = [], 0
    for n in sizes:                          # every array gets its own 128-byte aligned region
        offs.append(cur)
        cur += (n + 63) // 64 * 64
    for it in range(2):                      # warm-up
        for n, at in zip(sizes, offs):
            r.fill(n, it, at)
            r.run(n, at)
    torch.cuda.synchronize()
    dist.barrier()
    total_ms = 0.0
    for it in range(num_batches):
        for n, at in zip(sizes, offs):
            r.fill(n, it, at)
        torch.cuda.synchronize()
        dist.barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = repeat if len(sizes) == 1 else 1        # a single small array: `repeat` back-to-back launches between the events,
        s.record()                                       # so that the figure is the kernel, not the launch + event overhead
        for _ in range(reps):
            outs = [(n, r.run(n, at)) for n, at in zip(sizes, offs)]
        e.record()
        torch.cuda.synchronize()
        total_ms += s.elapsed_time(e) / reps
        if reps > 1:                                     # in-place all-reduce applied `reps` times: W^(reps-1) x the single result
            check = False
        if check:
            for n, out in outs[-3:] + outs[:1]:
                err = max(err, float((out.float() - r.expected(n, it)).abs().max()))
    t = torch.tensor([total_ms / num_batches / 1e3, err], device=r.dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0]), float(t[1])


def main(argv=None) -> int:
    a = parse(argv)
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        # like the reference tool (measure.py:62-75 builds its kvstore over mx.gpu(i) contexts) this measures device stores
        raise SystemExit("bandwidth_measure: no CUDA device on this host - the tool measures the GPU communication paths "
                         "(`--kv-store device|nccl|dist_sync...` between GPUs); use `diagnose` to see what the box has")
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    dtype = torch.bfloat16 if a.dtype == "bfloat16" else torch.float32
    esz = 2 if dtype == torch.bfloat16 else 4
    results = []
    if a.sweep:
        sizes_b = []
        b = 1024
        while b <= a.max_bytes:
            sizes_b.append(b)
            b *= 4
        r = Runner(a.kv_store, dtype, max(sizes_b) // esz)
        for nb in sizes_b:
            n = nb // esz
            iters = a.num_batches if nb >= (1 << 24) else a.num_batches * 4
            sec, err = measure(r, [n], iters, bool(a.test_results))                      # verified, one launch per timing
            if nb <= (1 << 20):                                                          # latency-bound: time 20 back-to-back launches
                sec, _ = measure(r, [n], max(2, iters // 4), False, repeat=20)
            bw = nb * 2 * (world - 1) / world / sec / 1e9
            results.append({"bytes": nb, "us": sec * 1e6, "busbw_GBps_per_gpu": bw, "error": err})
            if rank == 0:
                print(f"{a.kv_store:9s} {nb:>11d} B  {sec * 1e6:10.1f} us  {bw:8.2f} GB/s per GPU  error {err:g}", flush=True)
    else:
        shapes = network_shapes(a.network, a.num_layers, a.num_classes, a.image_shape)
        r = Runner(a.kv_store, dtype, max(shapes), sum((n + 63) // 64 * 64 for n in shapes))
        nbytes = sum(shapes) * esz
        if rank == 0:
            print(f"num of arrays = {len(shapes)}, total size = {nbytes / 1e6:f} MB", flush=True)
        for i in range(a.num_batches // max(1, a.disp_batches)):
            sec, err = measure(r, shapes, a.disp_batches, bool(a.test_results))
            bw = nbytes * 2 * (world - 1) / world / sec / 1e9
            results.append({"iter": i, "sec": sec, "busbw_GBps_per_gpu": bw, "error": err})
            if rank == 0:
                print(f"iter {i + 1}, {sec:f} sec, {bw:f} GB/sec per gpu, error {err:f}", flush=True)
    if rank == 0 and a.out:
        with open(a.out, "w") as f:
            json.dump({"kv_store": a.kv_store, "world": world, "dtype": a.dtype, "results": results}, f, indent=1)
    dist.barrier()
    dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
