"""All-reduce kernel variants x grid sizes (run under torchrun): peer ld/st in-place, peer ld/st out-of-place,
NVLS multimem; prints bus bandwidth per GPU."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deeplearning_cfn_b200.parallel.symm import SymmetricHeap  # noqa: E402


def main():
    rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lrank)
    dev = torch.device("cuda", lrank)
    dist.init_process_group("nccl", device_id=dev)
    sizes = [int(x) << 20 for x in os.environ.get("TUNE_MIB", "1,16,64,256").split(",")]
    blocks = [int(b) for b in os.environ.get("TUNE_BLOCKS", "32,64,128").split(",")]
    heap = SymmetricHeap(2 * max(sizes) + (8 << 20))
    if rank == 0:
        print(f"backend={heap.backend} multicast={heap.has_multicast} why={getattr(heap, '_why_not_torch', '')}", flush=True)
    a_off, a = heap.alloc_tensor(max(sizes) // 2, torch.bfloat16)
    b_off, b = heap.alloc_tensor(max(sizes) // 2, torch.bfloat16)
    out = []
    variants = [("p2p-inplace", 0, False)]
    if heap.has_multicast:
        variants += [("nvls-inplace", 1, False)]
    for nbytes in sizes:
        n = nbytes // 2
        for name, algo, oop in variants:
            for nb in blocks:
                a[:n] = 1.0
                torch.cuda.synchronize()
                dist.barrier()
                # correctness
                heap.allreduce_(a_off, n, torch.bfloat16, 0.5, nblocks=nb, dst_off=b_off if oop else None, algo=algo)
                torch.cuda.synchronize()
                res = (b if oop else a)[:n]
                ok = bool(torch.all(res == world * 0.5))
                a[:n] = 0.0
                torch.cuda.synchronize()
                dist.barrier()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                iters = 10
                s.record()
                for _ in range(iters):
                    heap.allreduce_(a_off, n, torch.bfloat16, 1.0, nblocks=nb, dst_off=b_off if oop else None, algo=algo)
                e.record()
                torch.cuda.synchronize()
                t = torch.tensor([s.elapsed_time(e) / iters], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                us = float(t[0]) * 1e3
                bw = nbytes * 2 * (world - 1) / world / (us * 1e-6) / 1e9
                out.append({"bytes": nbytes, "variant": name, "nblocks": nb, "us": us, "busbw": bw, "ok": ok})
                if rank == 0:
                    print(f"{nbytes >> 20:5d} MiB {name:16s} nblocks {nb:4d}  {us:9.1f} us  {bw:7.1f} GB/s  ok={ok}", flush=True)
    # fused parameter-server step (no optimizer: pure push + reduce + pull), peer ld/st vs multimem
    vec = 8
    for nbytes in sizes:
        n = nbytes // 2
        nvec = n // vec
        per = (nvec + world - 1) // world
        lo, hi = min(rank * per, nvec) * vec, min((rank + 1) * per, nvec) * vec
        master = torch.zeros(hi - lo + 64, device=dev)
        for name, algo in [("ps-p2p", 0)] + ([("ps-nvls", 1)] if heap.has_multicast else []):
            for nb in blocks:
                a[:n] = float(rank + 1)
                torch.cuda.synchronize()
                dist.barrier()
                heap.ps_sync_step(a_off, b_off, n, lo, hi, (1 << world) - 1, master, None, None, torch.bfloat16, "none",
                                  lr=0.0, nblocks=nb, algo=algo)
                torch.cuda.synchronize()
                ok = bool(torch.all(b[:n] == world * (world + 1) / 2))
                dist.barrier()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                for _ in range(10):
                    heap.ps_sync_step(a_off, b_off, n, lo, hi, (1 << world) - 1, master, None, None, torch.bfloat16, "none",
                                      lr=0.0, nblocks=nb, algo=algo)
                e.record()
                torch.cuda.synchronize()
                t = torch.tensor([s.elapsed_time(e) / 10], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                us = float(t[0]) * 1e3
                bw = nbytes * 2 * (world - 1) / world / (us * 1e-6) / 1e9
                out.append({"bytes": nbytes, "variant": name, "nblocks": nb, "us": us, "busbw": bw, "ok": ok})
                if rank == 0:
                    print(f"{nbytes >> 20:5d} MiB {name:16s} nblocks {nb:4d}  {us:9.1f} us  {bw:7.1f} GB/s  ok={ok}", flush=True)
    heap.check_error()
    if rank == 0:
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(out, open(f"gpurun_out/allreduce_tune_{world}.json", "w"), indent=1)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
