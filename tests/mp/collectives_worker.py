"""Run under torchrun (one rank per GPU): checks the fused peer-HBM kernels against closed-form answers.
The closed forms follow the reference's nightly dist_sync_kvstore.py:101-112 pattern (every worker pushes
ones*(rank+1) under the `Test` optimizer, so w after i rounds = 1 + i * rate * n(n+1)/2)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from deeplearning_cfn_b200.parallel.symm import SymmetricHeap  # noqa: E402


def main():
    rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lrank)
    dev = torch.device("cuda", lrank)
    dist.init_process_group("nccl", device_id=dev)
    heap = SymmetricHeap(512 << 20)
    if rank == 0:
        print(f"symmetric heap backend={heap.backend} multicast={heap.has_multicast}", flush=True)
    ok = True

    def fail(msg):
        nonlocal ok
        ok = False
        print(f"[rank {rank}] FAIL {msg}", flush=True)

    # ---- all-reduce: sizes incl. non-multiples of 16 B, both dtypes, fused scale ----
    for dtype in (torch.float32, torch.bfloat16):
        for n in (1, 3, 8, 1000, 65_537, 4_000_003):
            off, t = heap.alloc_tensor(n + 16, dtype)
            for it in range(3):
                t.zero_()
                t[:n] = (torch.arange(n, device=dev) % 13 + rank + it).to(dtype)
                torch.cuda.synchronize()
                heap.allreduce_(off, n, dtype, scale=0.5)
                torch.cuda.synchronize()
                want = 0.5 * sum((torch.arange(n, device=dev) % 13 + r + it).float() for r in range(world))
                if not torch.allclose(t[:n].float(), want.to(dtype).float(), rtol=1e-2, atol=1e-2):
                    fail(f"allreduce {dtype} n={n} it={it}: {t[:4].tolist()} vs {want[:4].tolist()}")
    # bit-identical across ranks
    n = 1_000_000
    off, t = heap.alloc_tensor(n, torch.bfloat16)
    torch.manual_seed(rank)
    t.copy_(torch.randn(n, device=dev).bfloat16())
    mine_before = t.clone()
    gathered = [torch.empty_like(mine_before) for _ in range(world)]
    dist.all_gather(gathered, mine_before)
    heap.allreduce_(off, n, torch.bfloat16, scale=1.0 / world)
    torch.cuda.synchronize()
    acc = torch.zeros(n, device=dev)
    for gbuf in gathered:
        acc += gbuf.float()
    want = (acc / world).bfloat16()
    if heap.allreduce_algo == 0:
        if not torch.equal(t, want):
            fail(f"allreduce not bit-exact vs fp32-accumulate reference: max diff {(t.float() - want.float()).abs().max().item()}")
    else:   # NVLS: the switch chooses the summation order, so allow the last bf16 bit to differ
        if not torch.allclose(t.float(), want.float(), rtol=2 ** -7, atol=1e-6):
            fail(f"nvls allreduce too far from the fp32-accumulate reference: {(t.float() - want.float()).abs().max().item()}")
    other = t.clone()
    dist.broadcast(other, src=0)
    if not torch.equal(t, other):
        fail("allreduce result differs between ranks")

    # ---- broadcast ----
    off, t = heap.alloc_tensor(100_003, torch.float32)
    t.fill_(float(rank))
    torch.cuda.synchronize()
    heap.broadcast_(off, 100_003 * 4, root=world - 1)
    torch.cuda.synchronize()
    if not torch.all(t == float(world - 1)):
        fail("broadcast")

    # ---- barrier ----
    for _ in range(5):
        heap.barrier()
    torch.cuda.synchronize()

    # ---- PS sync, colocated servers, Test optimizer closed form ----
    for dtype in (torch.float32, torch.bfloat16):
        vec = 4 if dtype == torch.float32 else 8
        n = 64 * 1024 + 64
        goff, g = heap.alloc_tensor(n, dtype)
        woff, w = heap.alloc_tensor(n, dtype)
        nvec = n // vec
        per = (nvec + world - 1) // world
        lo, hi = min(rank * per, nvec) * vec, min((rank + 1) * per, nvec) * vec
        master = torch.ones(hi - lo, device=dev)
        w.fill_(1.0)
        rate = 0.125
        for i in range(3):
            g.fill_(float(rank + 1))
            torch.cuda.synchronize()
            heap.ps_sync_step(goff, woff, n, lo, hi, (1 << world) - 1, master, None, None, dtype, "test", lr=0.0,
                              rescale=rate)
            torch.cuda.synchronize()
            want = 1.0 + (i + 1) * rate * world * (world + 1) / 2
            if not torch.allclose(w.float(), torch.full_like(w, want).float(), rtol=1e-2):
                fail(f"ps_sync {dtype} round {i}: got {w[:3].tolist()} .. {w[-3:].tolist()} want {want}")

    # ---- PS sync, dedicated server (rank 0 serves, others work), SGD momentum vs reference ----
    if world >= 2:
        from deeplearning_cfn_b200.ops import reference as ref
        n = 8192
        goff, g = heap.alloc_tensor(n, torch.bfloat16)
        woff, w = heap.alloc_tensor(n, torch.bfloat16)
        workers = list(range(1, world))
        mask = sum(1 << r for r in workers)
        lo, hi = (0, n) if rank == 0 else (0, 0)
        w0 = torch.linspace(-1, 1, n, device=dev)
        master = w0.clone() if rank == 0 else None
        mom = torch.zeros(n, device=dev) if rank == 0 else None
        mref, sref = w0.clone(), torch.zeros(n, device=dev)
        for i in range(3):
            g.fill_(0.01 * (rank + i))
            torch.cuda.synchronize()
            heap.ps_sync_step(goff, woff, n, lo, hi, mask, master, mom, None, torch.bfloat16, "sgd_mom", lr=0.1,
                              wd=1e-3, momentum=0.9, rescale=1.0 / len(workers))
            torch.cuda.synchronize()
            gsum = sum(torch.full((n,), 0.01 * (r + i), device=dev).bfloat16().float() for r in workers)
            ref.fused_optimizer("sgd_mom", gsum, mref, sref, None, None, lr=0.1, wd=1e-3, momentum=0.9,
                                rescale=1.0 / len(workers))
            if not torch.allclose(w.float(), mref.bfloat16().float(), rtol=2e-2, atol=2e-2):
                fail(f"ps dedicated round {i}: {w[:3].tolist()} vs {mref[:3].tolist()}")

    # ---- PS async, dedicated server ----
    if world >= 2:
        n = 16 * 1024
        goff, g = heap.alloc_tensor(n, torch.float32)
        woff, w = heap.alloc_tensor(n, torch.float32)
        workers = list(range(1, world))
        mask = sum(1 << r for r in workers)
        master = torch.zeros(n, device=dev) if rank == 0 else None
        rounds = 4
        for i in range(rounds):
            if rank == 0:
                heap.ps_async_server(goff, woff, 0, n, mask, master, None, None, torch.float32, "test", lr=0.0,
                                     rescale=1.0, nblocks=8)
            else:
                g.fill_(float(rank))
                heap.ps_async_worker(1, nblocks=8)
            torch.cuda.synchronize()
        dist.barrier()
        if rank == 0:
            want = rounds * sum(workers)
            if not torch.allclose(master, torch.full_like(master, float(want))):
                fail(f"ps_async master {master[:3].tolist()} want {want}")
        else:
            # weights returned to a worker reflect at least its own pushes and at most everybody's
            lo_b, hi_b = rounds * rank, rounds * sum(workers)
            if not (w.min().item() >= rank - 1e-3 and w.max().item() <= hi_b + 1e-3):
                fail(f"ps_async worker weights out of range: {w.min().item()} {w.max().item()} [{lo_b},{hi_b}]")

    # ---- both all-reduce algorithms explicitly (peer ld/st two-shot, NVLS multimem) on the same data ----
    n = 3_000_001
    off, t = heap.alloc_tensor(n + 16, torch.bfloat16)
    src = (torch.arange(n, device=dev) % 17 + rank).to(torch.bfloat16)
    want = sum((torch.arange(n, device=dev) % 17 + r).float() for r in range(world)).bfloat16()
    for algo in ([0, 1] if heap.has_multicast else [0]):
        t[:n].copy_(src)
        torch.cuda.synchronize()
        dist.barrier()
        heap.allreduce_(off, n, torch.bfloat16, scale=1.0, algo=algo)
        torch.cuda.synchronize()
        if not torch.allclose(t[:n].float(), want.float(), rtol=2 ** -7, atol=1e-6):
            fail(f"allreduce algo {algo}: max diff {(t[:n].float() - want.float()).abs().max().item()}")
    # ---- one-shot (single barrier) against two-shot on latency-bound sizes, many back-to-back calls (staging parity reuse) ----
    for dtype, n in ((torch.bfloat16, 1), (torch.bfloat16, 4096), (torch.float32, 4096), (torch.bfloat16, 8192), (torch.float32, 777)):
        off, t = heap.alloc_tensor(n + 16, dtype)
        for it in range(6):
            vals = ((torch.arange(n, device=dev) * 7 + rank * 3 + it) % 23).to(dtype)
            res = {}
            algos_small = (0, 2, 3) if os.environ.get("DLCFN_EXPERIMENTAL") == "1" else (0, 2)     # 3 = flag-in-data form, not yet run on a GPU
            for algo in algos_small:
                t[:n].copy_(vals)
                torch.cuda.synchronize()
                dist.barrier()
                heap.allreduce_(off, n, dtype, scale=0.25, algo=algo)
                heap.allreduce_(off, n, dtype, scale=1.0, algo=algo)      # back to back, in place: sum of equal values = W * x
                torch.cuda.synchronize()
                res[algo] = t[:n].clone()
            for algo in algos_small[1:]:
                if not torch.equal(res[0], res[algo]):
                    fail(f"all-reduce algo {algo} differs from two-shot: {dtype} n={n} it={it} max diff {(res[0].float() - res[algo].float()).abs().max().item()}")
    # ---- SyncBatchNorm statistics exchange through the one-shot kernel vs torch.distributed (reference C12) ----
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from _syncbn_check import check_syncbn_over_heap
    bad = check_syncbn_over_heap(heap, dev, rank)
    if bad:
        fail(bad)
    # ---- shared staging: rank 0 publishes "initial weights", everyone reads them out of rank 0's HBM ----
    w0 = torch.arange(10_000, device=dev, dtype=torch.float32) * (1.0 if rank == 0 else -1.0)
    heap.stage_put("init/weights", w0)
    torch.cuda.synchronize()
    heap.barrier()
    got = heap.stage_get("init/weights", 0)
    torch.cuda.synchronize()
    if not torch.equal(got, torch.arange(10_000, device=dev, dtype=torch.float32)):
        fail("stage_get did not return rank 0's staged tensor")
    backend_want = os.environ.get("DLCFN_EXPECT_SYMM_BACKEND")
    if backend_want and heap.backend != backend_want:
        fail(f"symmetric heap backend is {heap.backend}, expected {backend_want} ({getattr(heap, '_why_not_vmm', '')})")

    heap.check_error()
    flag = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(flag)
    if rank == 0:
        print("COLLECTIVES_OK" if flag.item() == 0 else "COLLECTIVES_FAILED", flush=True)
    dist.barrier()
    heap.close()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 0 else 1)


if __name__ == "__main__":
    main()
