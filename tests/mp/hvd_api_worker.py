"""Horovod-style collectives over gloo (3 ranks): allreduce(_), allgather with uneven first dims, broadcast(_), alltoall."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from deeplearning_cfn_b200.parallel import horovod as hvd  # noqa: E402


def main():
    hvd.init()
    r, n = hvd.rank(), hvd.size()
    assert n == 3
    t = torch.full((4,), float(r + 1))
    assert torch.allclose(hvd.allreduce(t, average=True), torch.full((4,), 2.0)) and torch.equal(t, torch.full((4,), float(r + 1)))
    hvd.allreduce_(t, average=False)
    assert torch.allclose(t, torch.full((4,), 6.0))
    g = hvd.allgather(torch.full((r + 1, 2), float(r)))                    # 1 + 2 + 3 rows
    assert g.shape == (6, 2) and g[:, 0].tolist() == [0.0, 1.0, 1.0, 2.0, 2.0, 2.0]
    b = hvd.broadcast(torch.full((3,), float(r)), root_rank=2)
    assert b.tolist() == [2.0, 2.0, 2.0]
    x = torch.full((2,), float(r))
    hvd.broadcast_(x, root_rank=1)
    assert x.tolist() == [1.0, 1.0]
    a = hvd.alltoall(torch.arange(6, dtype=torch.float32) + 10 * r)       # 2 rows to each rank
    assert a.tolist() == [2.0 * r, 2.0 * r + 1, 10 + 2.0 * r, 11 + 2.0 * r, 20 + 2.0 * r, 21 + 2.0 * r], a
    u = hvd.alltoall(torch.arange(3, dtype=torch.float32) + 10 * r, splits=[0, 1, 2])    # uneven: rank j receives j rows from everyone
    want = {0: [], 1: [0.0, 10.0, 20.0], 2: [1.0, 2.0, 11.0, 12.0, 21.0, 22.0]}[r]
    assert u.tolist() == want, (r, u)
    c, ctx = hvd.Compression.fp16.compress(torch.ones(3))
    assert c.dtype == torch.bfloat16 and hvd.Compression.fp16.decompress(c, ctx).dtype == torch.float32
    assert hvd.join() == n - 1
    print(f"HVD_API_OK {r}", flush=True)
    hvd.shutdown()


if __name__ == "__main__":
    main()
