"""4-rank CPU (gloo) check of the hierarchical multi-node mode: 2 virtual nodes x 2 local ranks
(DLCFN_VIRTUAL_NODE_SIZE=2).  The three-step exchange (node-local sum, rail all-reduce of the slices, node-local
broadcast of the slices) must train exactly like one flat all-reduce over the four ranks."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from deeplearning_cfn_b200.models import resnet  # noqa: E402
from deeplearning_cfn_b200.parallel import horovod as hvd  # noqa: E402
from deeplearning_cfn_b200.parallel.trainer import DataParallelTrainer  # noqa: E402
from deeplearning_cfn_b200.utils.data import synthetic_device_batch  # noqa: E402


def flat_weights(tr):
    return torch.cat([g.weight.reshape(-1).float() for g in tr.flat.groups])


def main():
    hvd.init()
    r, w = hvd.rank(), hvd.size()
    assert w == 4 and os.environ.get("DLCFN_VIRTUAL_NODE_SIZE") == "2"
    x, y = synthetic_device_batch(4, 32, 10, torch.device("cpu"), dtype=torch.float32, seed=20 + r)
    results = {}
    for mode in ("hier", "gloo"):
        torch.manual_seed(7 + r)                   # different init per rank on purpose: the initial broadcast must fix it
        model = resnet(20, 10, 32).train()
        tr = DataParallelTrainer(model, mode=mode, optimizer="sgd_mom", lr=0.05, bucket_bytes=1 << 16, broadcast_from=3)
        assert tr.mode == mode and len(tr.flat.buckets) > 1
        if mode == "hier":
            assert (tr.node_size, tr.num_nodes, tr.node_index, tr.local_rank) == (2, 2, r // 2, r % 2)
            b = tr.flat.buckets[0]
            lo0, hi0 = tr._rail_slice(b, 0)
            lo1, hi1 = tr._rail_slice(b, 1)
            assert lo0 == 0 and hi0 == lo1 and hi1 == b.numel and lo1 % 4 == 0      # 16-byte granules, full cover
        w0 = flat_weights(tr)
        ref0 = w0.clone()
        dist.broadcast(ref0, src=3)
        assert torch.equal(w0, ref0), f"{mode}: initial broadcast from rank 3 did not reach rank {r}"
        losses = [float(tr.step(x, y).detach()) for _ in range(3)]
        results[mode] = (flat_weights(tr), losses)
    wh, wg = results["hier"][0], results["gloo"][0]
    assert torch.allclose(wh, wg, atol=1e-6), float((wh - wg).abs().max())
    same = wh.clone()
    dist.broadcast(same, src=0)
    assert torch.equal(wh, same), "ranks diverged in hier mode"
    # a single-node job degenerates to the flat mode
    os.environ["DLCFN_VIRTUAL_NODE_SIZE"] = "4"
    tr1 = DataParallelTrainer(resnet(20, 10, 32).train(), mode="hier", optimizer="sgd_mom", lr=0.05)
    assert tr1.mode == "gloo"
    dist.barrier()
    if r == 0:
        print("HIER_TRAINER_OK", results["hier"][1], flush=True)
    hvd.shutdown()


if __name__ == "__main__":
    main()
