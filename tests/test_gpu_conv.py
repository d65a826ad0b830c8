"""Numerics of the tcgen05 convolution kernels (csrc/cuda/gemm_tcgen05.cu: plain GEMM for 1x1 / stride 1, im2col-mode TMA
implicit GEMM for everything else, parity-decomposed strided dgrad) against plain PyTorch fp32 references (GPU only)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda", 0)


def _frac_close(a, b, rtol, atol, frac=0.999):
    ok = (a.float() - b.float()).abs() <= atol + rtol * b.float().abs()
    assert ok.float().mean().item() >= frac, f"only {ok.float().mean().item():.5f} close; max err {(a.float() - b.float()).abs().max().item()}"


# ------------------------------------------------------------------------------------------ implicit-GEMM convolution
def _conv_case(N, C, H, W, K, R, stride, pad, seed=0):
    torch.manual_seed(seed)
    x = torch.randn(N, C, H, W, device=dev()).bfloat16().contiguous(memory_format=torch.channels_last)
    w = (torch.randn(K, C, R, R, device=dev()) / (C * R * R) ** 0.5).bfloat16().contiguous(memory_format=torch.channels_last)
    return x, w


# (N, C, H, W, K, R, stride, pad): ResNet-50 conv2 shapes at small batch, M / N / K tails, row / image wrap inside one tile
_CONV_SHAPES = [(1, 64, 7, 9, 128, 3, 1, 1), (1, 192, 5, 5, 64, 5, 1, 2), (2, 64, 56, 56, 64, 3, 1, 1), (3, 128, 28, 28, 128, 3, 1, 1), (4, 256, 14, 14, 256, 3, 1, 1),
                (8, 512, 7, 7, 512, 3, 1, 1), (1, 64, 9, 13, 72, 3, 1, 1), (5, 64, 5, 5, 200, 3, 1, 1),
                (2, 128, 56, 56, 128, 3, 2, 1), (2, 64, 15, 15, 64, 3, 2, 1), (2, 64, 12, 12, 96, 5, 1, 2),
                (2, 64, 10, 10, 64, 1, 1, 0), (2, 64, 16, 16, 128, 1, 2, 0), (2, 64, 8, 8, 64, 3, 1, 0)]


@pytest.mark.parametrize("shape", _CONV_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_conv_fprop_matches_fp32(shape):
    from deeplearning_cfn_b200.ops import native
    N, C, H, W, K, R, stride, pad = shape
    x, w = _conv_case(*shape)
    y = native.conv_fprop(x, w, stride, pad)
    yr = F.conv2d(x.float(), w.float(), None, stride, pad)
    assert y.shape == yr.shape and y.is_contiguous(memory_format=torch.channels_last)
    _frac_close(y, yr, 2e-2, 2e-2)


# strided gradients run one implicit GEMM per output-parity class: odd / even extents, classes without taps (1x1 / stride 2,
# 3x3 / stride 3 -> zero-filled), classes with 1, 2 and 4 taps, a negative lower corner (pad < R - 1)
_DGRAD_STRIDED = [(2, 64, 7, 7, 64, 1, 2, 0), (2, 128, 14, 14, 64, 5, 2, 2), (2, 64, 13, 11, 64, 3, 3, 1), (1, 64, 20, 20, 64, 7, 2, 3),
                  (3, 64, 9, 9, 128, 3, 2, 0), (2, 256, 28, 28, 256, 3, 2, 1), (2, 512, 14, 14, 512, 3, 2, 1)]


@pytest.mark.parametrize("shape", [s for s in _CONV_SHAPES if s[4] % 64 == 0] + _DGRAD_STRIDED, ids=lambda s: "x".join(map(str, s)))
def test_conv_dgrad_matches_fp32(shape):
    from deeplearning_cfn_b200.ops import native
    N, C, H, W, K, R, stride, pad = shape
    x, w = _conv_case(*shape)
    P, Q = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - R) // stride + 1
    dy = torch.randn(N, K, P, Q, device=dev()).bfloat16().contiguous(memory_format=torch.channels_last)
    dx = native.conv_dgrad(dy, w, pad, (H, W), stride=stride)
    dxr = torch.nn.grad.conv2d_input((N, C, H, W), w.float(), dy.float(), stride, pad)
    assert dx.shape == dxr.shape
    _frac_close(dx, dxr, 2e-2, 2e-2 * float(dxr.abs().max()))


@pytest.mark.parametrize("shape", _CONV_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_conv_wgrad_matches_fp32(shape):
    from deeplearning_cfn_b200.ops import native
    N, C, H, W, K, R, stride, pad = shape
    x, w = _conv_case(*shape)
    P, Q = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - R) // stride + 1
    dy = torch.randn(N, K, P, Q, device=dev()).bfloat16().contiguous(memory_format=torch.channels_last)
    dw = native.conv_wgrad(x, dy, (R, R), stride, pad)
    dwr = torch.nn.grad.conv2d_weight(x.float(), (K, C, R, R), dy.float(), stride, pad)
    assert dw.shape == dwr.shape and dw.is_contiguous(memory_format=torch.channels_last)
    _frac_close(dw, dwr, 2e-2, 2e-2 * float(dwr.abs().max()))
    assert torch.equal(dw, native.conv_wgrad(x, dy, (R, R), stride, pad))          # split-K reduction is deterministic


def test_convkxk_module_autograd_and_resnet_step():
    """ConvKxK on the tcgen05 backend inside autograd, against the same module on cuDNN; then one ResNet-50 training step
    with every 1x1 and 3x3 convolution on the dlcfn kernels."""
    from deeplearning_cfn_b200.ops import conv
    torch.manual_seed(0)
    m = conv.ConvKxK(64, 128, 3, 2, 1).to(dev()).to(torch.bfloat16).to(memory_format=torch.channels_last)
    x = torch.randn(4, 64, 28, 28, device=dev()).bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    outs = {}
    for be in ("cudnn", "tcgen05"):
        conv.set_convkxk_backend(be)
        try:
            m.zero_grad(); x.grad = None
            y = m(x)
            y.float().square().mean().backward()
            outs[be] = (y.detach().float(), x.grad.float().clone(), m.weight.grad.float().clone())
        finally:
            conv.set_convkxk_backend("tcgen05")
    for a, b in zip(outs["tcgen05"], outs["cudnn"]):
        _frac_close(a, b, 3e-2, 3e-2 * float(b.abs().max()))

    from deeplearning_cfn_b200.models import resnet, to_b200_layout
    from deeplearning_cfn_b200.ops import native
    net = to_b200_layout(resnet(50, 16, 64).to(dev()))
    xb = torch.randn(8, 3, 64, 64, device=dev()).bfloat16().contiguous(memory_format=torch.channels_last)
    losses = {}
    for be in ("cudnn", "tcgen05"):
        conv.set_convkxk_backend(be)
        conv.set_conv1x1_backend(be)
        try:
            net.zero_grad()
            n0 = native.launch_count
            loss = net(xb).float().logsumexp(1).mean()
            loss.backward()
            losses[be] = (float(loss), native.launch_count - n0, net.fc1.weight.grad.float().clone())
        finally:
            conv.set_convkxk_backend("tcgen05")
            conv.set_conv1x1_backend("tcgen05")
    assert abs(losses["tcgen05"][0] - losses["cudnn"][0]) < 0.05 * abs(losses["cudnn"][0]) + 0.05
    assert losses["tcgen05"][1] > losses["cudnn"][1] + 40            # the 3x3 layers really ran on the dlcfn kernel
    cos = F.cosine_similarity(losses["tcgen05"][2].flatten(), losses["cudnn"][2].flatten(), dim=0).item()
    assert cos > 0.98, cos


# ------------------------------------------------------------------ weights-stationary pipeline, 64-wide tiles, fused statistics
# (N, C, H, W, K, R, stride, pad): enough 128-pixel tiles (>= 2 per SM) for the weights-stationary mode; K = 64 -> BLOCK_N 64,
# K = 192 -> two N tiles of 128 (the resident panel is swapped mid-kernel), K = 256 -> BLOCK_N 256
_BIG_SHAPES = [(16, 64, 56, 56, 64, 3, 1, 1), (16, 64, 56, 56, 192, 3, 1, 1), (16, 64, 56, 56, 256, 1, 1, 0), (16, 256, 56, 56, 64, 1, 1, 0),
               (32, 128, 28, 28, 128, 3, 1, 1), (16, 64, 56, 56, 64, 1, 1, 0), (24, 128, 56, 56, 128, 3, 2, 1), (16, 128, 28, 28, 512, 1, 1, 0)]


@pytest.mark.parametrize("shape", _BIG_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_conv_ops_large_tiles_and_fused_statistics(shape):
    from deeplearning_cfn_b200.ops import native
    from deeplearning_cfn_b200.ops.conv import ConvOps
    N, C, H, W, K, R, stride, pad = shape
    x, w = _conv_case(*shape)
    ops = ConvOps(x, w, stride, pad)
    stats = native.new_stats(K, x.device)
    y = ops.fprop(stats)
    yr = F.conv2d(x.float(), w.float(), None, stride, pad)
    _frac_close(y, yr, 2e-2, 2e-2)
    assert torch.equal(y, ops.fprop())                                   # same kernel without the statistics epilogue
    tot = stats.sum(0)                                                   # [2, K]
    yf = y.float()
    s_ref, q_ref = yf.sum((0, 2, 3)), (yf * yf).sum((0, 2, 3))
    assert torch.allclose(tot[0], s_ref, rtol=1e-3, atol=1e-3 * float(q_ref.max()) ** 0.5), (tot[0][:4], s_ref[:4])
    assert torch.allclose(tot[1], q_ref, rtol=1e-3), (tot[1][:4], q_ref[:4])
    dy = torch.randn_like(y)
    dx = ops.dgrad(dy)
    dxr = torch.nn.grad.conv2d_input(x.shape, w.float(), dy.float(), stride, pad)
    _frac_close(dx, dxr, 2e-2, 2e-2 * float(dxr.abs().max()))
    dw = ops.wgrad(dy)
    dwr = torch.nn.grad.conv2d_weight(x.float(), w.shape, dy.float(), stride, pad)
    _frac_close(dw, dwr, 2e-2, 2e-2 * float(dwr.abs().max()))


def test_batchnorm_takes_statistics_from_the_convolution_epilogue():
    from deeplearning_cfn_b200.ops.batchnorm import BatchNormReLU2d
    from deeplearning_cfn_b200.ops.conv import Conv1x1, ConvKxK
    torch.manual_seed(1)
    for conv in (Conv1x1(256, 64), ConvKxK(64, 128, 3, 1, 1)):
        conv = conv.to(dev()).bfloat16().to(memory_format=torch.channels_last)
        x = torch.randn(16, conv.in_channels, 28, 28, device=dev()).bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True)
        bn_a, bn_b = BatchNormReLU2d(conv.out_channels).to(dev()).train(), BatchNormReLU2d(conv.out_channels).to(dev()).train()
        ya, st = conv(x, stats=True)
        assert st is not None and st.shape[1:] == (2, conv.out_channels)
        a = bn_a((ya, st))
        b = bn_b(conv(x))
        _frac_close(a, b, 1e-2, 1e-2)
        assert torch.allclose(bn_a.running_mean, bn_b.running_mean, atol=1e-4) and torch.allclose(bn_a.running_var, bn_b.running_var, rtol=1e-3, atol=1e-5)
        ga = torch.autograd.grad(a.float().square().mean(), x)[0]
        gb = torch.autograd.grad(b.float().square().mean(), x)[0]
        _frac_close(ga, gb, 2e-2, 2e-2 * float(gb.abs().max()))


def test_fused_stem_matches_the_unfused_modules():
    """bn_data + 7x7/2 stem on the overlapping-view implicit GEMM with the analytic d beta, against the same modules run one by
    one (BatchNorm kernels + space-to-depth + cuDNN convolution + autograd through all of them) and against fp32 torch."""
    import importlib
    R = importlib.import_module("deeplearning_cfn_b200.models.resnet")     # (the package also exports a function `resnet`)
    torch.manual_seed(2)
    net = R.to_b200_layout(R.resnet(50, 16, 64).to(dev())).train()
    x = torch.randn(8, 3, 64, 64, device=dev()).bfloat16().contiguous(memory_format=torch.channels_last)
    outs = {}
    for fused in (True, False):
        R._FUSED_STEM = fused
        try:
            net.zero_grad(set_to_none=True)
            rm0 = net.bn_data.running_mean.clone()
            if fused:
                y, st = R._FusedStemFn.apply(x, net.stem[0].weight, net.bn_data.bias, net.bn_data, True)
                assert st is not None
            else:
                y = net.stem[0](net.bn_data(x))
            g = torch.Generator(device="cpu").manual_seed(3)
            dy = torch.randn(y.shape, generator=g).to(dev()).bfloat16().contiguous(memory_format=torch.channels_last)
            y.backward(dy)
            outs[fused] = (y.detach().float(), net.stem[0].weight.grad.float().clone(), net.bn_data.bias.grad.float().clone())
            net.bn_data.running_mean.copy_(rm0)
        finally:
            R._FUSED_STEM = True
    for a, b in zip(outs[True], outs[False]):
        _frac_close(a, b, 3e-2, 3e-2 * float(b.abs().max()), frac=0.995)
    # statistics of the fused path against its own output
    tot = st.sum(0)
    assert torch.allclose(tot[0], outs[True][0].sum((0, 2, 3)), rtol=1e-3, atol=1.0)
