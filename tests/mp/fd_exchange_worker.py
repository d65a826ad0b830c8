"""File-descriptor exchange of the VMM symmetric-heap backend (parallel/symm.py:exchange_fds) on the host: every rank shares
a temporary file's descriptor, every peer reads the owner's rank out of the received descriptor; then the "only rank 0 has
something to share" form used for the multicast object; then a failure on one rank, which every rank must report."""
import os
import sys
import tempfile

import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from deeplearning_cfn_b200.parallel.symm import exchange_fds  # noqa: E402

rank = int(os.environ.get("RANK", os.environ.get("OMPI_COMM_WORLD_RANK", "0")))
world = int(os.environ.get("WORLD_SIZE", os.environ.get("OMPI_COMM_WORLD_SIZE", "1")))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("MASTER_PORT", "29641")
dist.init_process_group("gloo", rank=rank, world_size=world)
f = tempfile.TemporaryFile()
f.write(f"arena-of-rank-{rank}".encode())
f.flush()
fds = exchange_fds(f.fileno(), rank, world, None, "test")
assert fds[rank] == f.fileno() and all(x >= 0 for x in fds)
for p, fd in enumerate(fds):
    assert os.pread(fd, 64, 0) == f"arena-of-rank-{p}".encode(), (p, os.pread(fd, 64, 0))
    if p != rank:
        os.close(fd)
one = exchange_fds(f.fileno() if rank == 0 else -1, rank, world, None, "mc")
assert one[0] >= 0
assert os.pread(one[0], 64, 0) == b"arena-of-rank-0"
assert all(one[p] == -1 for p in range(1, world))
# a failure on ONE rank (the last rank offers a descriptor that is not open: sendmsg fails there, its peers wait in vain) must
# raise on EVERY rank - a rank that gave up alone would leave the others in the next collective
os.environ["DLCFN_FD_EXCHANGE_TIMEOUT"] = "3"
bad = 1_000_000          # no such descriptor (a just-closed number would be reused by the exchange's own sockets)
try:
    exchange_fds(bad if rank == world - 1 else f.fileno(), rank, world, None, "broken")
    raise SystemExit(f"rank {rank}: the broken exchange did not raise")
except RuntimeError as e:
    assert "file-descriptor exchange 'broken' failed" in str(e) and f"rank {world - 1}:" in str(e), str(e)
dist.barrier()
print(f"rank {rank} FD_EXCHANGE_OK", flush=True)
dist.destroy_process_group()
