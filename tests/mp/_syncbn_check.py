"""SyncBatchNorm statistics exchange through a symmetric heap vs torch.distributed (reference C12), shared by the GPU worker
(tests/mp/collectives_worker.py: real heap, one-shot kernel) and the host worker (tests/mp/syncbn_heap_worker.py: stand-in heap
over gloo), so the code the GPU box runs has been executed on the host."""
import torch


def check_syncbn_over_heap(heap, dev, rank: int):
    """None, or a description of the disagreement."""
    from deeplearning_cfn_b200.parallel import utils as putils
    torch.manual_seed(50 + rank)
    xs = torch.randn(4, 32, 6, 6, device=dev)
    outs = {}
    try:
        for mode in ("dist", "heap"):
            putils.set_sync_bn_heap(heap if mode == "heap" else None)
            bn = putils.SyncBatchNorm(32).to(dev).train()
            xi = xs.clone().requires_grad_(True)
            yb = bn(xi)
            (yb * torch.arange(32, device=dev).view(1, 32, 1, 1)).sum().backward()
            outs[mode] = (yb.detach(), xi.grad.clone(), bn.running_var.clone())
    finally:
        putils.set_sync_bn_heap(None)
    for a, b in zip(outs["dist"], outs["heap"]):
        if not torch.allclose(a, b, rtol=1e-5, atol=1e-5):
            return f"SyncBatchNorm over the heap differs from torch.distributed: {(a - b).abs().max().item()}"
    return None
