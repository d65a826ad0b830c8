"""Per-key mode of the bandwidth tool for all four stores (reference bar: MX/tools/bandwidth/test_measure.py - error == 0)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from deeplearning_cfn_b200.examples import bandwidth_measure as bm  # noqa: E402

torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
shapes = bm.network_shapes("resnet", 18, 100, "3,64,64")
assert any(n % 64 for n in shapes), "the parameter list must contain sizes that need padding"
for kind in ("allreduce", "ps", "nccl", "nccl_ps"):
    r = bm.Runner(kind, torch.bfloat16, max(shapes), sum((n + 63) // 64 * 64 for n in shapes))
    sec, err = bm.measure(r, shapes, 2, True)
    assert err == 0.0, (kind, err)
    sec, err = bm.measure(r, [4096], 2, True)
    assert err == 0.0, (kind, "single", err)
    try:
        r.run(shapes[0], at=r.g.numel())
        raise SystemExit("out-of-arena region must be refused")
    except IndexError:
        pass
    del r
dist.barrier()
if dist.get_rank() == 0:
    print("BANDWIDTH_OK", flush=True)
dist.destroy_process_group()
