"""`trainer.capture` (parallel/graph.py): the whole training step as one CUDA graph - the bench's default since round 2 (validated
on B200: 22.18 -> 21.11 ms per ResNet-50 step).  Runs after the kernel tests in the GPU tier (file name) because a capture that went wrong can leave
the process in capture mode."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda", 0)


def _frac_close(a, b, rtol, atol, frac=0.999):
    ok = (a.float() - b.float()).abs() <= atol + rtol * b.float().abs()
    assert ok.float().mean().item() >= frac, f"only {ok.float().mean().item():.5f} close; max err {(a.float() - b.float()).abs().max().item()}"


@pytest.mark.parametrize("net", ["cifar10_cnn", "resnet18"])
def test_graphed_step_matches_eager_training(net):
    """trainer.capture: the replayed graph must train exactly like the eager step (same kernels, same order), follow a
    learning-rate change by re-capturing, and run other batch shapes eagerly."""
    from deeplearning_cfn_b200.models import get_model, to_b200_layout
    from deeplearning_cfn_b200.parallel.trainer import DataParallelTrainer
    size = 32
    g = torch.Generator().manual_seed(0)
    xs = [torch.randn(16, 3, size, size, generator=g).to(dev()).bfloat16().contiguous(memory_format=torch.channels_last) for _ in range(4)]
    ys = [torch.randint(0, 10, (16,), generator=g).to(dev()) for _ in range(4)]

    def make():
        torch.manual_seed(5)
        m = get_model(net if net != "resnet18" else "resnet", 10, size, **({"num_layers": 20} if net == "resnet18" else {}))
        return DataParallelTrainer(to_b200_layout(m.to(dev())).train(), mode="local", optimizer="sgd_mom", lr=0.05)
    eager, graphed = make(), make()
    for _ in range(3):                               # the warm-up steps of capture() train on the example batch
        eager.step(xs[0], ys[0])
    step = graphed.capture(xs[0], ys[0], warmup=3)
    assert step.graphed and step.launches_per_step > 0, repr(step)
    for i in range(6):
        if i == 3:
            eager.set_learning_rate(0.01)
            graphed.set_learning_rate(0.01)
        le = float(eager.step(xs[i % 4], ys[i % 4]))
        lg = float(step(xs[i % 4], ys[i % 4]))
        assert abs(le - lg) <= 2e-2 * abs(le) + 1e-3, (i, le, lg)
    assert step.captures == 2 and step.replays == 6 and graphed.step_count == eager.step_count
    for pe, pg in zip(eager.model.parameters(), graphed.model.parameters()):
        _frac_close(pg, pe, 2e-2, 2e-2 * float(pe.float().abs().max()) + 1e-4, frac=0.99)
    small = step(xs[1][:8], ys[1][:8])               # another shape: eager
    assert torch.isfinite(small) and step.replays == 6
