import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.timeout(600)          # with >= 2 GPUs smoke() also runs a 2-rank collectives check (own limit: 300 s)
def test_smoke_entry():
    import __graft_entry__ as g
    g.smoke()


def test_local_trainer_reduces_loss():
    from deeplearning_cfn_b200.models import resnet, to_b200_layout
    from deeplearning_cfn_b200.ops import native
    from deeplearning_cfn_b200.parallel.trainer import DataParallelTrainer
    from deeplearning_cfn_b200.utils.data import synthetic_device_batch
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    model = to_b200_layout(resnet(20, 10, 32).to(dev), torch.bfloat16).train()
    tr = DataParallelTrainer(model, mode="local", optimizer="sgd_mom", lr=0.05)
    x, y = synthetic_device_batch(32, 32, 10, dev)
    n0 = native.launch_count
    losses = [float(tr.step(x, y)) for _ in range(12)]
    assert native.launch_count > n0
    assert losses[-1] < losses[0], losses


def test_prefetcher_pipeline():
    from deeplearning_cfn_b200.utils.data import DevicePrefetcher, SyntheticImageSource
    dev = torch.device("cuda", 0)
    src = SyntheticImageSource(4, 32, 10, distinct=2)
    pre = DevicePrefetcher(src, dev)
    for i in range(5):
        x, y = pre.next()
        assert x.shape == (4, 3, 32, 32) and x.dtype == torch.bfloat16 and y.shape == (4,)
        assert x.is_contiguous(memory_format=torch.channels_last)
    torch.cuda.synchronize()
