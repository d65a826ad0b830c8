"""Regression worker for the round-1 advisor findings on the distributed store (started by `launch -n W -s S`):
big-key 2-bit compression with shard boundaries that are not multiples of 16, row-sparse rows on a big key with an odd
row count, a bf16 gradient pushed to an fp32 key, back-to-back pushes of one key (staging-buffer reuse), and the gluon
Trainer with batch_size > 1 and a learning-rate change under update_on_kvstore."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
os.environ.setdefault("MXNET_KVSTORE_BIGARRAY_BOUND", "1000")
import deeplearning_cfn_b200.kv as kv  # noqa: E402
from deeplearning_cfn_b200.ops import reference as ref_ops  # noqa: E402

mode = sys.argv[1]


def close(a, b, what, tol=1e-4):
    if not torch.allclose(a.float(), b.float(), rtol=tol, atol=tol):
        bad = (a.float() - b.float()).abs() > tol
        raise AssertionError(f"{what}: {int(bad.sum())} wrong elements, first at {int(bad.reshape(-1).nonzero()[0])}: "
                             f"got {a.reshape(-1)[bad.reshape(-1)][:4].tolist()} want {b.reshape(-1)[bad.reshape(-1)][:4].tolist()}")


def store_cases():
    store = kv.create("dist_sync")
    r, n = store.rank, store.num_workers
    rate = 2.0
    store.set_optimizer(kv.optimizer.create("test", rescale_grad=rate))
    # ---- row-sparse on a big key whose rows straddle the naive element split: (7, 300), rows 0 3 4 6
    store.init(21, torch.zeros(7, 300))
    rows = torch.tensor([0, 3, 4, 6])
    store.push_row_sparse(21, torch.ones(4, 300) * (r + 1), rows)
    out = torch.full((7, 300), -1.0)
    store.row_sparse_pull(21, out=out, row_ids=torch.arange(7))
    want = torch.zeros(7, 300)
    want[rows] = rate * n * (n + 1) / 2
    close(out, want, "row-sparse big key")
    # ---- bf16 gradient into an fp32 key
    store.init(22, torch.ones(33))
    store.push(22, torch.ones(33, dtype=torch.bfloat16) * (r + 1))
    o = torch.zeros(33)
    store.pull(22, out=o)
    close(o, torch.full((33,), 1 + rate * n * (n + 1) / 2), "bf16 gradient, fp32 key")
    # ---- five back-to-back pushes of a 200k-element key, then one pull
    store.init(23, torch.zeros(200_000))
    for i in range(1, 6):
        store.push(23, torch.full((200_000,), float(i)))
    o = torch.zeros(200_000)
    store.pull(23, out=o)
    close(o, torch.full((200_000,), rate * 15 * n), "push x5 then pull")
    # ---- 2-bit compression of a big key, non-uniform gradient, shards not multiples of 16 under the naive split
    store.set_gradient_compression({"type": "2bit", "threshold": 0.5})
    N = 3600
    store.init(24, torch.zeros(N))
    gen = torch.Generator().manual_seed(7)
    gs = [torch.randn(N, generator=gen) * 0.7 for _ in range(n)]           # same stream on every worker
    store.push(24, gs[r])
    o = torch.zeros(N)
    store.pull(24, out=o)
    want = torch.zeros(N)
    for g in gs:
        res = torch.zeros(N)
        want += ref_ops.dequantize_2bit(ref_ops.quantize_2bit(g.clone(), res, 0.5), N, 0.5)
    close(o, rate * want, "2-bit big key")
    store._barrier()
    print(f"worker {r} REGRESS_OK", flush=True)
    store.close()


def trainer_case():
    w = torch.nn.Parameter(torch.zeros(6))
    tr = kv.Trainer([w], "sgd", {"learning_rate": 1.0}, kvstore="dist_sync")
    n = None
    for i in range(2):
        w.grad = torch.full((6,), 8.0)
        tr.step(8)                                  # rescale 1/8: every worker contributes -1 per step
        n = tr._kvstore.num_workers
        close(w.data, torch.full((6,), -float(n) * (i + 1)), f"step {i} with batch_size 8")
    tr.set_learning_rate(0.0)
    before = w.data.clone()
    w.grad = torch.full((6,), 8.0)
    tr.step(8)
    close(w.data, before, "lr = 0 reaches the servers")
    tr.set_learning_rate(0.5)
    w.grad = torch.full((6,), 8.0)
    tr.step(4)                                      # rescale 1/4, lr 0.5 -> -1 per worker again
    close(w.data, before - float(n), "lr and batch size change")
    print(f"worker {tr._kvstore.rank} TRAINER_OK", flush=True)
    tr._kvstore.close()


if __name__ == "__main__":
    trainer_case() if mode == "trainer" else store_cases()
