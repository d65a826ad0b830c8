"""Numerics of every hand-written sm_100a kernel against a plain PyTorch fp32 reference of the same op."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from deeplearning_cfn_b200.ops import native, reference as ref


def dev():
    return torch.device("cuda", 0)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("n", [1, 7, 4096, 1_000_003])
@pytest.mark.parametrize("k", [1, 2, 5, 8])
def test_elementwise_sum(dtype, n, k):
    torch.manual_seed(0)
    xs = [torch.randn(n, device=dev()).to(dtype) for _ in range(k)]
    out = torch.empty(n, device=dev(), dtype=dtype)
    native.elementwise_sum(xs, out, 0.5)
    want = torch.empty(n, device=dev(), dtype=dtype)
    ref.elementwise_sum(xs, want, 0.5)
    tol = 1e-6 if dtype == torch.float32 else 2e-2
    torch.testing.assert_close(out.float(), want.float(), rtol=tol, atol=tol)


@pytest.mark.parametrize("n", [5, 1024, 777_777])
def test_cast_scale(n):
    x = torch.randn(n, device=dev())
    y = torch.empty(n, device=dev(), dtype=torch.bfloat16)
    native.cast_scale(x, y, 0.25)
    torch.testing.assert_close(y.float(), (x * 0.25).bfloat16().float())
    z = torch.empty(n, device=dev())
    native.cast_scale(y, z, 2.0)
    torch.testing.assert_close(z, y.float() * 2.0)


ALL_KINDS = ["sgd", "sgd_mom", "adam", "test", "none", "nag", "signum", "signsgd", "adagrad", "rmsprop", "adadelta",
             "ftrl", "adamax"]


def _mostly_close(a, b, rtol, atol, frac=0.9999):
    """sign()/threshold optimizers may legitimately differ on an element that sits within one ulp of the switch point."""
    ok = (a.float() - b.float()).abs() <= atol + rtol * b.float().abs()
    assert ok.float().mean().item() >= frac, f"only {ok.float().mean().item():.6f} close"


@pytest.mark.parametrize("kind", ALL_KINDS)
@pytest.mark.parametrize("gdtype,wdtype", [(torch.bfloat16, torch.bfloat16), (torch.float32, torch.float32),
                                           (torch.float32, torch.bfloat16)])
@pytest.mark.parametrize("n", [3, 4099, 300_001])
def test_fused_optimizer(kind, gdtype, wdtype, n):
    torch.manual_seed(1)
    g = torch.randn(n, device=dev()).to(gdtype)
    hp = dict(lr=0.1, wd=1e-2, momentum=0.9, beta1=0.9, beta2=0.99, eps=1e-6, rescale=0.5, clip=1.5)
    if kind in ("signum", "signsgd"):
        hp["beta1"] = 0.01                       # wd_lh
    if kind == "ftrl":
        hp.update(beta1=0.05, beta2=1.0)         # lamda1, beta
    if kind == "rmsprop":
        hp["beta2"] = 2.0                        # clip_weights
    m0 = torch.randn(n, device=dev())
    a0, b0 = torch.randn(n, device=dev()) * 0.1, torch.rand(n, device=dev()) * 0.1
    if kind in ("adagrad", "rmsprop", "adadelta"):
        a0 = a0.abs()
    m, a, b = m0.clone(), a0.clone(), b0.clone()
    w = torch.empty(n, device=dev(), dtype=wdtype)
    mr, ar, br = m0.clone(), a0.clone(), b0.clone()
    wr = torch.empty(n, device=dev(), dtype=wdtype)
    for _ in range(3):
        native.fused_optimizer(kind, g, m, a, b, w, **hp)
        ref.fused_optimizer(kind, g, mr, ar, br, wr, **hp)
    _mostly_close(m, mr, 1e-4, 1e-5)
    wt = 1e-2 if wdtype == torch.bfloat16 else 1e-4
    _mostly_close(w, wr, wt, wt)
    if kind in native.OPT_USES_S1:
        _mostly_close(a, ar, 1e-4, 1e-5)
    if kind in native.OPT_USES_S2:
        _mostly_close(b, br, 1e-4, 1e-5)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("n", [1, 16, 17, 1000, 65_537])
def test_quantize_2bit_roundtrip(dtype, n):
    torch.manual_seed(2)
    thr = 0.5
    g = (torch.randn(n, device=dev()) * 0.6).to(dtype)
    res = torch.randn(n, device=dev()) * 0.2
    res_ref = res.clone()
    words = torch.zeros((n + 15) // 16, dtype=torch.int32, device=dev())
    native.quantize_2bit(g, res, words, thr)
    want = ref.quantize_2bit(g, res_ref, thr)
    assert torch.equal(words, want)
    torch.testing.assert_close(res, res_ref)
    out = torch.zeros(n, device=dev())
    native.dequantize_2bit(words, out, thr)
    torch.testing.assert_close(out, ref.dequantize_2bit(want, n, thr))
    native.dequantize_2bit(words, out, thr, accumulate=True)
    torch.testing.assert_close(out, 2 * ref.dequantize_2bit(want, n, thr))


def test_unique_rows():
    ids = torch.tensor([5, 3, 5, 9, 0, 3, 3, 11], device=dev())
    got = native.unique_rows(ids, 12)
    assert got.tolist() == [0, 3, 5, 9, 11]
    big = torch.randint(0, 50_000, (200_000,), device=dev())
    assert torch.equal(native.unique_rows(big, 50_000), torch.unique(big))


def test_normalize_u8():
    from deeplearning_cfn_b200.utils.data import IMAGENET_MEAN, IMAGENET_STD
    x = torch.randint(0, 256, (3, 17, 19, 3), dtype=torch.uint8, device=dev())
    y = torch.empty(x.shape, dtype=torch.bfloat16, device=dev())
    native.normalize_u8_nhwc(x, y, IMAGENET_MEAN, IMAGENET_STD)
    want = ref.normalize_u8_nhwc(x, IMAGENET_MEAN, IMAGENET_STD)
    torch.testing.assert_close(y.float(), want.float(), rtol=2e-2, atol=2e-2)


@pytest.mark.parametrize("relu", [True, False])
@pytest.mark.parametrize("N,C,H,W", [(4, 64, 14, 14), (8, 256, 7, 9), (2, 2048, 7, 7), (3, 96, 5, 5), (16, 8, 33, 31),
                                      (32, 128, 28, 28)])
def test_fused_batchnorm_relu(N, C, H, W, relu):
    from deeplearning_cfn_b200.ops.batchnorm import batch_norm_relu
    torch.manual_seed(3)
    x = (torch.randn(N, C, H, W, device=dev()) * 1.5 + 0.3).bfloat16().contiguous(memory_format=torch.channels_last)
    gamma = (torch.rand(C, device=dev()) + 0.5).requires_grad_(True)
    beta = (torch.randn(C, device=dev()) * 0.2).requires_grad_(True)
    rm, rv = torch.zeros(C, device=dev()), torch.ones(C, device=dev())
    xg = x.clone().requires_grad_(True)
    y = batch_norm_relu(xg, gamma, beta, rm, rv, True, 0.1, 2e-5, relu)
    dy = torch.randn_like(y)
    y.backward(dy)
    # fp32 reference
    xr = x.float().requires_grad_(True)
    gr, br = gamma.detach().clone().requires_grad_(True), beta.detach().clone().requires_grad_(True)
    rmr, rvr = torch.zeros(C, device=dev()), torch.ones(C, device=dev())
    yr = torch.nn.functional.batch_norm(xr, rmr, rvr, gr, br, True, 0.1, 2e-5)
    if relu:
        yr = torch.relu(yr)
    yr.backward(dy.float())
    torch.testing.assert_close(y.float(), yr.detach(), rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(rm, rmr, rtol=1e-3, atol=1e-3)
    torch.testing.assert_close(rv, rvr, rtol=1e-3, atol=1e-3)
    torch.testing.assert_close(xg.grad.float(), xr.grad, rtol=5e-2, atol=5e-2)
    torch.testing.assert_close(gamma.grad, gr.grad, rtol=2e-2, atol=2e-2 * (N * H * W) ** 0.5)
    torch.testing.assert_close(beta.grad, br.grad, rtol=2e-2, atol=2e-2 * (N * H * W) ** 0.5)
    # eval mode
    ye = batch_norm_relu(x, gamma.detach(), beta.detach(), rm, rv, False, 0.1, 2e-5, relu)
    yer = torch.nn.functional.batch_norm(x.float(), rm, rv, gamma.detach(), beta.detach(), False, 0.1, 2e-5)
    if relu:
        yer = torch.relu(yer)
    torch.testing.assert_close(ye.float(), yer, rtol=2e-2, atol=2e-2)


@pytest.mark.parametrize("N,C,H,W,k,s,p", [(2, 64, 12, 12, 3, 2, 1), (4, 8, 7, 9, 3, 2, 1), (2, 16, 8, 8, 2, 2, 0),
                                            (1, 128, 15, 13, 3, 1, 1), (8, 64, 56, 56, 3, 2, 1)])
def test_maxpool_nhwc(N, C, H, W, k, s, p):
    from deeplearning_cfn_b200.ops.pool import max_pool2d
    torch.manual_seed(4)
    # distinct values so arg-max ties cannot make the two implementations disagree
    x = torch.randperm(N * C * H * W, device=dev()).float().view(N, C, H, W).mul_(1e-3).bfloat16().float()
    x = x + torch.randn_like(x) * 4
    x = x.bfloat16().contiguous(memory_format=torch.channels_last)
    xa = x.clone().requires_grad_(True)
    y = max_pool2d(xa, k, s, p)
    xr = x.float().requires_grad_(True)
    yr = torch.nn.functional.max_pool2d(xr, k, s, p)
    assert y.shape == yr.shape and torch.equal(y.float(), yr)
    dy = torch.randn_like(yr)
    y.backward(dy.bfloat16())
    yr.backward(dy.bfloat16().float())
    torch.testing.assert_close(xa.grad.float().sum(), xr.grad.sum(), rtol=1e-2, atol=1e-1)
    close = torch.isclose(xa.grad.float(), xr.grad, rtol=2e-2, atol=2e-2).float().mean().item()
    assert close > 0.995, close


@pytest.mark.parametrize("k,s,p,size", [(7, 2, 3, 64), (3, 1, 1, 32)])
def test_stem_conv_tensor_core_form_matches_plain_conv(k, s, p, size):
    from deeplearning_cfn_b200.models.resnet import StemConv
    torch.manual_seed(5)
    conv = StemConv(3, 64, k, s, p, bias=False).to(dev()).bfloat16().to(memory_format=torch.channels_last)
    x = torch.randn(4, 3, size, size, device=dev()).bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    y = conv(x)
    gy = torch.randn_like(y)
    y.backward(gy)
    xr = x.detach().float().requires_grad_(True)
    wr = conv.weight.detach().float().requires_grad_(True)
    yr = torch.nn.functional.conv2d(xr, wr, None, s, p)
    yr.backward(gy.float())
    assert y.shape == yr.shape
    torch.testing.assert_close(y.float(), yr, rtol=3e-2, atol=3e-2)
    torch.testing.assert_close(conv.weight.grad.float(), wr.grad, rtol=5e-2, atol=5e-1)
    torch.testing.assert_close(x.grad.float(), xr.grad, rtol=5e-2, atol=5e-2)


@pytest.mark.parametrize("relu", [True, False])
def test_fused_add_batchnorm_relu(relu):
    from deeplearning_cfn_b200.ops.batchnorm import add_batch_norm_relu
    torch.manual_seed(6)
    N, C, H, W = 8, 128, 14, 14
    mk = lambda: (torch.randn(N, C, H, W, device=dev()) * 1.2).bfloat16().contiguous(memory_format=torch.channels_last)  # noqa: E731
    a0, b0 = mk(), mk()
    gamma = (torch.rand(C, device=dev()) + 0.5).requires_grad_(True)
    beta = (torch.randn(C, device=dev()) * 0.2).requires_grad_(True)
    rm, rv = torch.zeros(C, device=dev()), torch.ones(C, device=dev())
    a, b = a0.clone().requires_grad_(True), b0.clone().requires_grad_(True)
    z, y = add_batch_norm_relu(a, b, gamma, beta, rm, rv, True, 0.1, 2e-5, relu)
    gz, gy = torch.randn_like(z), torch.randn_like(y)
    (z * gz).sum().add((y * gy).sum()).backward()
    ar, br = a0.float().requires_grad_(True), b0.float().requires_grad_(True)
    gr, ber = gamma.detach().clone().requires_grad_(True), beta.detach().clone().requires_grad_(True)
    zr = (ar + br)
    zq = zr + (zr.detach().bfloat16().float() - zr.detach())   # the kernel normalises the bf16-rounded sum (straight-through)
    yr = torch.nn.functional.batch_norm(zq, torch.zeros(C, device=dev()), torch.ones(C, device=dev()), gr, ber, True, 0.1, 2e-5)
    if relu:
        yr = torch.relu(yr)
    ((zr * gz.float()).sum() + (yr * gy.float()).sum()).backward()
    torch.testing.assert_close(z.float(), zr.detach().bfloat16().float())
    torch.testing.assert_close(y.float(), yr.detach(), rtol=2e-2, atol=2e-2)
    # a ReLU mask may flip for values that round to ~0: require almost-everywhere agreement
    for got, want in ((a.grad.float(), ar.grad), (b.grad.float(), br.grad)):
        frac = torch.isclose(got, want, rtol=5e-2, atol=5e-2).float().mean().item()
        assert frac > 0.995, frac
    torch.testing.assert_close(gamma.grad, gr.grad, rtol=5e-2, atol=2.0)
    torch.testing.assert_close(beta.grad, ber.grad, rtol=5e-2, atol=2.0)


def test_stem_space_to_depth_and_pad_kernels():
    from deeplearning_cfn_b200.ops.pool import pad3to8, stem_space_to_depth
    import torch.nn.functional as F
    torch.manual_seed(7)
    x3 = torch.randn(3, 3, 20, 16, device=dev()).bfloat16().contiguous(memory_format=torch.channels_last)
    x8 = pad3to8(x3)
    assert x8.shape == (3, 8, 20, 16) and torch.equal(x8[:, :3], x3) and float(x8[:, 3:].abs().sum()) == 0
    xg = x8.clone().requires_grad_(True)
    s = stem_space_to_depth(xg)
    N, C, H, W = 3, 3, 20, 16
    xp = F.pad(x3, (3, 3, 3, 3))
    ref = xp.reshape(N, C, (H + 6) // 2, 2, (W + 6) // 2, 2).permute(0, 1, 3, 5, 2, 4).reshape(N, 12, (H + 6) // 2, (W + 6) // 2)
    assert s.shape == (3, 16, 13, 11) and torch.equal(s[:, :12], ref) and float(s[:, 12:].abs().sum()) == 0
    g = torch.randn_like(s)
    s.backward(g)
    xr = x3.clone().float().requires_grad_(True)
    xpr = F.pad(xr, (3, 3, 3, 3))
    refr = xpr.reshape(N, C, (H + 6) // 2, 2, (W + 6) // 2, 2).permute(0, 1, 3, 5, 2, 4).reshape(N, 12, (H + 6) // 2, (W + 6) // 2)
    refr.backward(g[:, :12].float())
    assert torch.equal(xg.grad[:, :3].float(), xr.grad) and float(xg.grad[:, 3:].abs().sum()) == 0
