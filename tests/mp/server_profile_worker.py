"""Remote profiling of the parameter servers from worker rank 0 (reference tests/nightly/test_server_profiling.py)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import deeplearning_cfn_b200.kv as kv  # noqa: E402
from deeplearning_cfn_b200.utils import profiler  # noqa: E402


def main():
    out = sys.argv[1]
    store = kv.create("dist_sync")
    r = store.rank
    if r == 0:
        profiler.set_config(filename=os.path.join(out, "server_profile.json"), profile_process="server")
        profiler.set_state("run", profile_process="server")
    store._barrier()
    store.init(3, torch.ones(8))
    store.init(99, torch.ones(1200, 3))                 # crosses MXNET_KVSTORE_BIGARRAY_BOUND: lands on every server
    for _ in range(3):
        store.push(3, torch.ones(8))
        store.push(99, torch.ones(1200, 3))
        a, b = torch.zeros(8), torch.zeros(1200, 3)
        store.pull(3, out=a)
        store.pull(99, out=b)
    store._barrier()
    if r == 0:
        profiler.set_state("stop", profile_process="server")
        profiler.dump(profile_process="server")
    store._barrier()
    print(f"worker {r} PROFILED", flush=True)
    store.close()


if __name__ == "__main__":
    main()
