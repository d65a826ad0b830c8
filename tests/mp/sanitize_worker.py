"""Small 2-rank run of every cross-GPU kernel for compute-sanitizer (racecheck / synccheck see one process = one GPU each; both
ranks run under their own sanitizer):   tools/sanitize_collectives.sh   -> gpurun_out/sanitize_world2_<tool>_rank<r>.log"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from deeplearning_cfn_b200.parallel.symm import SymmetricHeap  # noqa: E402

rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lrank)
dev = torch.device("cuda", lrank)
dist.init_process_group("nccl", device_id=dev)
heap = SymmetricHeap(16 << 20)
n = 40_000
off, t = heap.alloc_tensor(n, torch.float32)
woff, w = heap.alloc_tensor(n, torch.float32)
master = torch.zeros(n // world + 64, device=dev)
for it in range(3):
    for algo in [0, 2] + ([3] if os.environ.get("DLCFN_EXPERIMENTAL") == "1" else []) + ([1] if heap.has_multicast else []):
        m = n if algo not in (2, 3) else 2048
        t[:m] = float(rank + 1)
        heap.allreduce_(off, m, torch.float32, 1.0, algo=algo)
        torch.cuda.synchronize()
        assert torch.all(t[:m] == float(sum(range(1, world + 1)))), (algo, t[:3])
    t.fill_(float(rank + 1))
    per = (n // 4 + world - 1) // world * 4
    lo, hi = min(rank * per, n), min((rank + 1) * per, n)
    heap.ps_sync_step(off, woff, n, lo, hi, (1 << world) - 1, master, None, None, torch.float32, "none", lr=0.0)
    heap.broadcast_(off, n * 4, it % world)
    heap.barrier()
torch.cuda.synchronize()
heap.check_error()
dist.barrier()
print(f"rank {rank} SANITIZE_RUN_OK", flush=True)
dist.destroy_process_group()
