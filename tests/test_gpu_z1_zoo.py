"""Model-zoo families on the dlcfn kernels (bf16, channels_last) against the same weights in fp32 on the host.  Since round 2 the
1x1 convolutions of these models run on the tcgen05 GEMM by default (DLCFN_CONV1X1); the per-layer numerics are pinned in
tests/test_gpu_conv.py / test_gpu_nn.py, every layer shape of every family walks past the C entry points' argument checks on the
host (tests/test_mock_gpu.py) - this is the whole-model check."""
import pytest
import torch

from deeplearning_cfn_b200.models import get_model, to_b200_layout


def dev():
    return torch.device("cuda", 0)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["mobilenet", "mobilenetv2", "resnext", "densenet121", "inception-bn"])
def test_zoo_models_on_cuda_kernels_match_fp32_cpu(name):
    """bf16 channels_last on the dlcfn kernels (depthwise, 1x1 GEMM-capable convs, fused BN+ReLU, NHWC pools) against the
    same weights in fp32 on the host: logits agree to bf16 accuracy and every parameter receives a finite gradient."""
    torch.manual_seed(1)
    ref = get_model(name, 16, 64)
    m = get_model(name, 16, 64)
    m.load_state_dict(ref.state_dict())
    m = to_b200_layout(m.to(dev()))
    x = torch.randn(8, 3, 64, 64)
    y_ref = ref(x)
    y = m(x.to(dev()).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)).float().cpu()
    cos = torch.nn.functional.cosine_similarity(y.flatten(), y_ref.flatten(), dim=0).item()
    assert torch.isfinite(y).all() and cos > 0.95, (cos, (y - y_ref).abs().max().item())
    m(x.to(dev()).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)).float().logsumexp(1).sum().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters() if p.requires_grad)
