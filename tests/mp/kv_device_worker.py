"""torchrun worker: `kv.create('dist_device_sync')` on GPUs = keys in peer HBM + one fused kernel per push."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import deeplearning_cfn_b200.kv as kv  # noqa: E402


def main():
    rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lrank)
    dev = torch.device("cuda", lrank)
    dist.init_process_group("nccl", device_id=dev)
    os.environ["DLCFN_KV_HEAP_BYTES"] = str(256 << 20)
    os.environ["MXNET_KVSTORE_BIGARRAY_BOUND"] = "100000"
    store = kv.create("dist_device_sync")
    assert isinstance(store, kv.KVStoreDistDevice) and store.rank == rank and store.num_workers == world
    ok = True
    rate = 2.0
    shapes = {3: (4, 5), "big": (300_000,), 7: (33, 8)}
    store.init(3, torch.ones(shapes[3], device=dev) * (1 if rank == 0 else 99))      # only rank 0's value counts
    store.init(99, torch.ones(shapes["big"], device=dev, dtype=torch.bfloat16))
    store.init(7, torch.arange(33 * 8, device=dev, dtype=torch.float32).view(33, 8))
    # no optimizer: the pushed sum replaces the value
    store.push(3, torch.ones(shapes[3], device=dev) * (rank + 1))
    o = torch.zeros(shapes[3], device=dev)
    store.pull(3, out=o)
    if not torch.allclose(o, torch.full_like(o, world * (world + 1) / 2)):
        ok = False
        print(f"[{rank}] sum-without-updater wrong: {o.flatten()[:3].tolist()}", flush=True)
    store.set_optimizer(kv.optimizer.create("test", rescale_grad=rate))
    base = world * (world + 1) / 2
    for i in range(3):
        store.push(3, [torch.ones(shapes[3], device=dev) * (rank + 1)] * 2)          # two local "devices"
        store.push(99, torch.ones(shapes["big"], device=dev, dtype=torch.bfloat16) * (rank + 1))
        ob = torch.zeros(shapes["big"], device=dev, dtype=torch.bfloat16)
        store.pull([3, 99], out=[o, ob])
        w3 = base + rate * (i + 1) * 2 * world * (world + 1) / 2
        wb = 1 + rate * (i + 1) * world * (world + 1) / 2
        if not torch.allclose(o, torch.full_like(o, w3)) or not torch.allclose(ob.float(), torch.full_like(ob.float(), wb), rtol=1e-2):
            ok = False
            print(f"[{rank}] round {i}: {o.flatten()[0].item()} vs {w3}; {ob[0].item()} vs {wb}", flush=True)
    rows = torch.full((33, 8), -1.0, device=dev)
    store.row_sparse_pull(7, out=rows, row_ids=torch.tensor([5, 1, 5, 32], device=dev))
    want = torch.arange(33 * 8, device=dev, dtype=torch.float32).view(33, 8)
    if not (torch.equal(rows[5], want[5]) and torch.equal(rows[32], want[32]) and float(rows[0].abs().sum()) == 0):
        ok = False
        print(f"[{rank}] row_sparse_pull wrong", flush=True)
    try:
        store.push(12345, torch.ones(2, device=dev))
        ok = False
    except kv.KVStoreError:
        pass
    store._barrier()
    store.heap.check_error()
    flag = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(flag)
    if rank == 0:
        print("KV_DEVICE_OK" if flag.item() == 0 else "KV_DEVICE_FAILED", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 0 else 1)


if __name__ == "__main__":
    main()
