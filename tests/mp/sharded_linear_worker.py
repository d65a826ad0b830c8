"""torchrun worker: the weight of a Linear layer stays on its parameter-server shards; forward / dgrad pull the tiles
through TMA inside the tcgen05 GEMM, wgrad + fused PS update keep the shards current.  Checked against a dense fp32
model that sees the same data."""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from deeplearning_cfn_b200.ops.sharded_linear import PSResidentLinear  # noqa: E402
from deeplearning_cfn_b200.parallel.symm import SymmetricHeap  # noqa: E402


def main():
    rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lrank)
    dev = torch.device("cuda", lrank)
    dist.init_process_group("nccl", device_id=dev)
    heap = SymmetricHeap(256 << 20)
    K, N, M = 2048, 1000, 256
    lin = PSResidentLinear(K, N, heap, bias=True, seed=3)
    # dense reference weight = concatenation of all shards
    shards = [torch.empty_like(lin.weight_shard) for _ in range(world)]
    dist.all_gather(shards, lin.weight_shard)
    Wd = torch.cat(shards).view(lin.out_padded, K).float()
    ok = True
    mom = torch.zeros_like(Wd)
    for it in range(3):
        torch.manual_seed(100 + it * world + rank)
        x = torch.randn(M, K, device=dev).bfloat16().requires_grad_(True)
        gy = torch.randn(M, N, device=dev).bfloat16() * 0.1
        y = lin(x)
        y.backward(gy)
        yr = x.detach().float() @ Wd[:N].t() + lin.bias_f32[:N]
        dxr = gy.float() @ Wd[:N]
        e1 = (y.float() - yr).abs().max().item() / (yr.abs().max().item() + 1e-6)
        e2 = (x.grad.float() - dxr).abs().max().item() / (dxr.abs().max().item() + 1e-6)
        # reference update: average of every rank's dW, SGD momentum
        dW = torch.zeros_like(Wd)
        dW[:N] = gy.float().t() @ x.detach().float()
        dW = dW.bfloat16().float()
        dist.all_reduce(dW)
        mom = 0.9 * mom - 0.05 * (dW / world)
        Wd = Wd + mom
        lin.step(lr=0.05, momentum=0.9)
        shards = [torch.empty_like(lin.weight_shard) for _ in range(world)]
        dist.all_gather(shards, lin.weight_shard)
        Wn = torch.cat(shards).view(lin.out_padded, K).float()
        e3 = (Wn - Wd).abs().max().item() / (Wd.abs().max().item() + 1e-6)
        if rank == 0:
            print(f"iter {it}: fwd rel err {e1:.3e}, dgrad rel err {e2:.3e}, weights after PS update rel err {e3:.3e}", flush=True)
        if e1 > 2e-2 or e2 > 2e-2 or e3 > 2e-2:
            ok = False
        Wd = Wn.clone()          # keep following the bf16-rounded trajectory
    # timing: fused pull+GEMM vs NCCL all-gather followed by cuBLAS
    x = torch.randn(M, K, device=dev).bfloat16()
    full = torch.empty(lin.out_padded, K, device=dev, dtype=torch.bfloat16)

    def fused():
        return lin(x)

    def baseline():
        dist.all_gather_into_tensor(full.view(-1), lin.weight_shard)
        return torch.nn.functional.linear(x, full)
    res = {}
    for name, fn in (("fused", fused), ("allgather+cublas", baseline)):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(50):
            fn()
        e.record()
        torch.cuda.synchronize()
        t = torch.tensor([s.elapsed_time(e) / 50], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[name] = float(t[0]) * 1e3
    if rank == 0:
        print(f"FC {M}x{K} -> {N}, weights sharded over {world} ranks ({'staged NVLS push' if lin.staged else 'per-tile TMA pull'}): "
              f"fused pull+GEMM {res['fused']:.1f} us, NCCL all-gather + cuBLAS {res['allgather+cublas']:.1f} us", flush=True)
    # a BERT-large projection (tokens 8192, 1024 -> 4096): numerics of the fused forward / dgrad, then the same timing
    M2, K2, N2 = 8192, 1024, 4096
    lin2 = PSResidentLinear(K2, N2, heap, bias=False, seed=5)
    shards2 = [torch.empty_like(lin2.weight_shard) for _ in range(world)]
    dist.all_gather(shards2, lin2.weight_shard)
    W2 = torch.cat(shards2).view(lin2.out_padded, K2)
    torch.manual_seed(7 + rank)
    x2 = torch.randn(M2, K2, device=dev).bfloat16().requires_grad_(True)
    gy2 = (torch.randn(M2, N2, device=dev) * 0.1).bfloat16()
    y2 = lin2(x2)
    y2.backward(gy2)
    yr2 = x2.detach().float() @ W2[:N2].float().t()
    dxr2 = gy2.float() @ W2[:N2].float()
    e1 = (y2.float() - yr2).abs().max().item() / (yr2.abs().max().item() + 1e-6)
    e2 = (x2.grad.float() - dxr2).abs().max().item() / (dxr2.abs().max().item() + 1e-6)
    if e1 > 2e-2 or e2 > 2e-2:
        ok = False
        print(f"[rank {rank}] FAIL projection fwd rel err {e1:.3e} dgrad rel err {e2:.3e}", flush=True)
    full2 = torch.empty(lin2.out_padded, K2, device=dev, dtype=torch.bfloat16)
    xin = x2.detach()

    def fused2():
        return lin2(xin)

    def baseline2():
        dist.all_gather_into_tensor(full2.view(-1), lin2.weight_shard)
        return torch.nn.functional.linear(xin, full2)
    res2 = {}
    for name, fn in (("fused", fused2), ("allgather+cublas", baseline2)):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(30):
            fn()
        e.record()
        torch.cuda.synchronize()
        t = torch.tensor([s.elapsed_time(e) / 30], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res2[name] = float(t[0]) * 1e3
    if rank == 0:
        print(f"projection {M2}x{K2} -> {N2}, weights sharded over {world} ranks ({'staged NVLS push' if lin2.staged else 'per-tile TMA pull'}): "
              f"fused pull+GEMM {res2['fused']:.1f} us, NCCL all-gather + cuBLAS {res2['allgather+cublas']:.1f} us "
              f"(fwd rel err {e1:.2e}, dgrad rel err {e2:.2e})", flush=True)
    heap.check_error()
    flag = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(flag)
    if rank == 0:
        print("SHARDED_LINEAR_OK" if flag.item() == 0 else "SHARDED_LINEAR_FAILED", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 0 else 1)


if __name__ == "__main__":
    main()
