"""Host run of the SyncBatchNorm-over-heap check (tests/mp/_syncbn_check.py) with a stand-in symmetric heap whose all-reduce is
gloo's: the routing of `parallel.utils._exchange_sum` (staging buffer slice, in-place reduce, clone) and the check itself are
executed before the GPU box runs them over the one-shot kernel."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _syncbn_check import check_syncbn_over_heap  # noqa: E402

rank = int(os.environ.get("RANK", os.environ.get("OMPI_COMM_WORLD_RANK", "0")))
world = int(os.environ.get("WORLD_SIZE", os.environ.get("OMPI_COMM_WORLD_SIZE", "1")))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("MASTER_PORT", "29645")
dist.init_process_group("gloo", rank=rank, world_size=world)


class HostHeap:
    """The three things set_sync_bn_heap / _exchange_sum use of a SymmetricHeap."""
    ONESHOT_MAX = 16 << 10

    def __init__(self):
        self.mem = torch.zeros(1 << 16, dtype=torch.uint8)
        self.top, self.calls = 0, 0

    def alloc_tensor(self, numel, dtype):
        off, nb = self.top, numel * torch.empty((), dtype=dtype).element_size()
        self.top += nb
        return off, self.mem[off:off + nb].view(dtype)

    def allreduce_(self, off, count, dtype, scale=1.0):
        t = self.mem[off:off + count * 4].view(torch.float32)
        dist.all_reduce(t)
        t.mul_(scale)
        self.calls += 1


heap = HostHeap()
bad = check_syncbn_over_heap(heap, torch.device("cpu"), rank)
assert bad is None, bad
assert heap.calls == 2, heap.calls          # one exchange in forward, one in backward went through the heap
dist.barrier()
print(f"rank {rank} SYNCBN_HEAP_OK", flush=True)
dist.destroy_process_group()
