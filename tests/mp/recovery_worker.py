"""A worker dies in the middle of a dist_sync job; the local launcher restarts it (DMLC_NUM_ATTEMPT), the scheduler hands
the restarted process the dead node's id once its heartbeats have timed out (is_recovery), it skips the start-up barriers,
and the interrupted synchronous round completes with its push.  Reference: ps-lite van.cc:86-156, kvstore_dist.h:52,206,
dmlc-core tracker/local.py:25-42."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import deeplearning_cfn_b200.kv as kv  # noqa: E402
from deeplearning_cfn_b200.utils import rt  # noqa: E402


def main():
    attempt = int(os.environ.get("DMLC_NUM_ATTEMPT", "0"))
    store = kv.create("dist_sync")
    r = store.rank
    recovered = bool(rt.lib().dlcfn_ps_is_recovery())
    assert recovered == (attempt > 0), (recovered, attempt)
    store.init(3, torch.ones(4))
    if not recovered:
        store.set_optimizer(kv.optimizer.create("test", rescale_grad=1.0))     # ends with a worker barrier
        if r == 1:
            print("worker 1 crashing", flush=True)
            os._exit(17)                                                       # dies before its push of this round
    out = torch.zeros(4)
    store.push(3, torch.ones(4) * (r + 1))
    store.pull(3, out=out)                  # rank 0 waits here until the restarted rank 1 has pushed
    assert torch.allclose(out, torch.full((4,), 1.0 + 1.0 + 2.0)), out
    print(f"worker {r} attempt {attempt} RECOVERY_OK", flush=True)
    store.close()


if __name__ == "__main__":
    main()
