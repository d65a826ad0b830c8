"""Worker body of the SyncReplicas test: 3 replicas, 2 aggregated per step, replica 2 is a slow backup worker."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import deeplearning_cfn_b200.kv as kv  # noqa: E402
from deeplearning_cfn_b200.parallel.sync_replicas import SyncReplicasTrainer  # noqa: E402


def main():
    store = kv.create("dist_async")
    r, n = store.rank, store.num_workers
    params = [torch.ones(50, requires_grad=True), torch.ones(7, 3, requires_grad=True)]
    t = SyncReplicasTrainer(store, params, "sgd", {"learning_rate": 0.1}, replicas_to_aggregate=n - 1)
    assert len(t.keys) >= 2, "test wants more than one chunk (MXNET_KVSTORE_BIGARRAY_BOUND small)"
    target = 8
    while t.global_step < target:
        for p in params:
            p.grad = torch.ones_like(p)            # constant gradient: every applied update is exactly -lr
        if r == n - 1:
            time.sleep(0.05)                       # the backup worker: most of its gradients arrive stale
        t.step()
        flat = torch.cat([p.detach().reshape(-1) for p in params])
        for (lo, hi), s in zip(t.chunks, t.last_steps):
            want = 1.0 - 0.1 * s
            if not torch.allclose(flat[lo:hi], torch.full((hi - lo,), want), atol=1e-5):
                raise AssertionError(f"rank {r}: chunk {lo}:{hi} at step {s}: {flat[lo:lo + 3].tolist()} want {want}")
    store._barrier()
    print(f"worker {r} SYNCREP_OK steps_taken={t.steps_taken} global_step={t.global_step}", flush=True)
    store.close()


if __name__ == "__main__":
    main()
