"""Numerics of csrc/cuda/nn_ops.cu + ops/conv.py against plain PyTorch fp32 references (GPU only)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from deeplearning_cfn_b200.ops import native  # noqa: E402


def dev():
    return torch.device("cuda", 0)


def _cl(t):
    return t.contiguous(memory_format=torch.channels_last)


def _frac_close(a, b, rtol, atol, frac=0.999):
    ok = (a.float() - b.float()).abs() <= atol + rtol * b.float().abs()
    assert ok.float().mean().item() >= frac, f"only {ok.float().mean().item():.5f} close; max err {(a.float() - b.float()).abs().max().item()}"


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("M,C", [(7, 10), (256, 1000), (33, 2048), (5, 30522), (64, 3)])
def test_softmax_cross_entropy(dtype, M, C):
    from deeplearning_cfn_b200.ops.nn import softmax_cross_entropy
    torch.manual_seed(0)
    x = (torch.randn(M, C, device=dev()) * 3).to(dtype).requires_grad_(True)
    y = torch.randint(0, C, (M,), device=dev())
    y[0] = -100                                                   # ignored row
    n0 = native.launch_count
    loss = softmax_cross_entropy(x, y)
    loss.backward()
    assert native.launch_count - n0 == 2
    xr = x.detach().float().requires_grad_(True)
    ref = F.cross_entropy(xr, y, ignore_index=-100)
    ref.backward()
    torch.testing.assert_close(loss, ref, rtol=1e-4, atol=1e-4)
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    _frac_close(x.grad, xr.grad, tol, tol * 0.05)
    assert float(x.grad[0].abs().sum()) == 0.0
    # reduction='none' + upstream gradient
    x2 = x.detach().clone().requires_grad_(True)
    rows = softmax_cross_entropy(x2, y, reduction="none")
    up = torch.rand(M, device=dev())
    (rows * up).sum().backward()
    x3 = x.detach().float().requires_grad_(True)
    (F.cross_entropy(x3, y, ignore_index=-100, reduction="none") * up).sum().backward()
    _frac_close(x2.grad, x3.grad, tol, tol * 0.05)


@pytest.mark.parametrize("N,C,H,W,size,tf", [(2, 64, 12, 12, 9, True), (3, 96, 7, 5, 5, False), (1, 256, 3, 3, 5, False),
                                             (2, 8, 4, 4, 4, False)])
def test_lrn(N, C, H, W, size, tf):
    from deeplearning_cfn_b200.ops.nn import local_response_norm
    torch.manual_seed(1)
    alpha, beta, k = (0.001 / 9.0, 0.75, 1.0) if tf else (1e-2, 0.75, 2.0)
    x = _cl(torch.randn(N, C, H, W, device=dev()) * 4).bfloat16().requires_grad_(True)
    y = local_response_norm(x, size, alpha, beta, k, tf_style=tf)
    gy = _cl(torch.randn_like(y))
    y.backward(gy)
    xr = x.detach().float().requires_grad_(True)
    a_torch = alpha * size if tf else alpha                        # torch divides alpha by size
    yr = F.local_response_norm(xr, size, a_torch, beta, k)
    yr.backward(gy.float())
    _frac_close(y, yr, 1e-2, 1e-2)
    _frac_close(x.grad, xr.grad, 2e-2, 2e-2)


@pytest.mark.parametrize("N,C,H,W,K,S,P", [(2, 32, 14, 14, 3, 1, 1), (3, 64, 15, 13, 3, 2, 1), (1, 8, 9, 9, 5, 1, 2),
                                           (2, 96, 7, 7, 3, 1, 0), (2, 264, 10, 10, 3, 2, 1)])
def test_depthwise_conv(N, C, H, W, K, S, P):
    from deeplearning_cfn_b200.ops.nn import depthwise_conv2d
    torch.manual_seed(2)
    x = _cl(torch.randn(N, C, H, W, device=dev())).bfloat16().requires_grad_(True)
    w = (torch.randn(C, 1, K, K, device=dev()) * 0.3).requires_grad_(True)
    y = depthwise_conv2d(x, w, S, P)
    gy = _cl(torch.randn_like(y))
    y.backward(gy)
    xr = x.detach().float().requires_grad_(True)
    wr = w.detach().clone().requires_grad_(True)
    yr = F.conv2d(xr, wr, None, S, P, 1, C)
    yr.backward(gy.float())
    assert y.shape == yr.shape
    _frac_close(y, yr, 1e-2, 1e-2)
    _frac_close(x.grad, xr.grad, 1e-2, 1e-2)
    _frac_close(w.grad, wr.grad, 1e-2, 3e-2)


@pytest.mark.parametrize("N,C,H,W,S", [(2, 64, 8, 8, 2), (3, 16, 7, 9, 2), (1, 8, 5, 5, 3)])
def test_subsample(N, C, H, W, S):
    from deeplearning_cfn_b200.ops.nn import subsample
    x = _cl(torch.randn(N, C, H, W, device=dev())).bfloat16().requires_grad_(True)
    y = subsample(x, S)
    assert torch.equal(y, x.detach()[:, :, ::S, ::S])
    gy = _cl(torch.randn_like(y))
    y.backward(gy)
    want = torch.zeros_like(x.detach())
    want[:, :, ::S, ::S] = gy
    assert torch.equal(x.grad, want)


@pytest.mark.parametrize("N,Cin,Cout,H,W,S", [(4, 64, 256, 14, 14, 1), (2, 256, 64, 9, 7, 1), (3, 256, 512, 14, 14, 2),
                                              (2, 1024, 2048, 7, 7, 2), (32, 64, 64, 56, 56, 1)])
def test_conv1x1_on_tcgen05_matches_conv2d(N, Cin, Cout, H, W, S):
    from deeplearning_cfn_b200.ops.conv import Conv1x1, set_conv1x1_backend, conv1x1_backend
    torch.manual_seed(3)
    old = conv1x1_backend()
    set_conv1x1_backend("tcgen05")
    try:
        m = Conv1x1(Cin, Cout, S).to(dev()).to(memory_format=torch.channels_last).bfloat16()
        x = _cl(torch.randn(N, Cin, H, W, device=dev())).bfloat16().requires_grad_(True)
        n0 = native.launch_count
        y = m(x)
        assert native.launch_count > n0, "tcgen05 path not taken"
        assert y.is_contiguous(memory_format=torch.channels_last)
        gy = _cl(torch.randn_like(y))
        y.backward(gy)
    finally:
        set_conv1x1_backend(old)
    xr = x.detach().float().requires_grad_(True)
    wr = m.weight.detach().float().requires_grad_(True)
    yr = F.conv2d(xr, wr, None, S)
    yr.backward(gy.float())
    _frac_close(y, yr, 2e-2, 2e-2)
    _frac_close(x.grad, xr.grad, 2e-2, 2e-2)
    _frac_close(m.weight.grad, wr.grad, 2e-2, 0.02 * float(wr.grad.abs().max()))
    assert m.weight.grad.shape == m.weight.shape


def test_resnet_step_same_loss_on_both_conv1x1_backends():
    """Whole-model agreement of the two 1x1-convolution backends.  A 50-layer bf16 network amplifies rounding
    differences layer by layer, so the check is on what the loss sees directly (loss value, classifier gradient) plus a
    loose direction check on the first convolution's gradient; exact per-layer numerics are pinned by
    test_conv1x1_on_tcgen05_matches_conv2d."""
    from deeplearning_cfn_b200.models import get_model
    from deeplearning_cfn_b200.models.resnet import to_b200_layout
    from deeplearning_cfn_b200.ops.conv import set_conv1x1_backend, conv1x1_backend
    torch.manual_seed(4)
    model = to_b200_layout(get_model("resnet50", num_classes=100, image_size=128)).to(dev())
    x = _cl(torch.randn(32, 3, 128, 128, device=dev())).bfloat16()
    y = torch.randint(0, 100, (32,), device=dev())
    old = conv1x1_backend()
    out = {}
    try:
        for be in ("cudnn", "tcgen05"):
            set_conv1x1_backend(be)
            model.zero_grad(set_to_none=True)
            n0 = native.launch_count
            loss = F.cross_entropy(model(x).float(), y)
            loss.backward()
            out[be] = (loss.item(), model.fc1.weight.grad.float().clone(), model.stem[0].weight.grad.float().clone(),
                       native.launch_count - n0)
    finally:
        set_conv1x1_backend(old)
    # 36 stride-1 1x1 convolutions x 3 GEMMs, minus one launch for each of the 16 BatchNorms that take their statistics from a
    # tcgen05 convolution's epilogue instead of their own pass (104 on the stand-in GPU): well above 60 either way
    assert out["tcgen05"][3] > out["cudnn"][3] + 60, "the tcgen05 backend should add ~3 GEMM launches per 1x1 conv"
    assert abs(out["cudnn"][0] - out["tcgen05"][0]) < 0.05 * abs(out["cudnn"][0]) + 0.05, (out["cudnn"][0], out["tcgen05"][0])
    cos_fc = F.cosine_similarity(out["cudnn"][1].flatten(), out["tcgen05"][1].flatten(), dim=0).item()
    cos_stem = F.cosine_similarity(out["cudnn"][2].flatten(), out["tcgen05"][2].flatten(), dim=0).item()
    assert cos_fc > 0.99, cos_fc
    assert cos_stem > 0.7, cos_stem


@pytest.mark.parametrize("kind", ["sgd_mom", "adam", "adagrad", "ftrl"])
def test_sparse_row_update(kind):
    from deeplearning_cfn_b200.ops import sparse as sp, reference as ref
    torch.manual_seed(5)
    R, L = 1000, 48
    w = torch.randn(R, L, device=dev())
    s1 = torch.rand(R, L, device=dev()) * 0.1
    s2 = torch.rand(R, L, device=dev()) * 0.1
    rows = torch.randperm(R, device=dev())[:137].sort().values
    g = torch.randn(137, L, device=dev())
    hp = dict(lr=0.1, wd=1e-2, momentum=0.9, beta1=0.9, beta2=0.99, eps=1e-6, rescale=0.5, clip=1.5)
    if kind == "ftrl":
        hp.update(beta1=0.05, beta2=1.0)
    wr, s1r, s2r = w.clone(), s1.clone(), s2.clone()
    wb = torch.zeros(R, L, device=dev(), dtype=torch.bfloat16)
    sp.sparse_row_update(kind, rows, g, w, s1, s2, wb, **hp)
    ws, a, b = wr[rows].clone(), s1r[rows].clone(), s2r[rows].clone()
    ref.fused_optimizer(kind, g.reshape(-1), ws.view(-1), a.view(-1), b.view(-1), None, **hp)
    mask = torch.ones(R, dtype=torch.bool, device=dev())
    mask[rows] = False
    assert torch.equal(w[mask], wr[mask]) and torch.equal(s1[mask], s1r[mask])
    _frac_close(w[rows], ws, 1e-4, 1e-5, frac=0.9999)
    _frac_close(wb[rows], ws, 1e-2, 1e-2)
    assert float(wb[mask].abs().sum()) == 0.0


def test_group_adagrad_dense_and_sparse():
    import deeplearning_cfn_b200.kv as kv
    torch.manual_seed(6)
    w = torch.randn(50, 33, device=dev())
    wc = w.cpu().clone()
    o, oc = kv.optimizer.create("groupadagrad", learning_rate=0.1), kv.optimizer.create("groupadagrad", learning_rate=0.1)
    u, uc = kv.optimizer.get_updater(o), kv.optimizer.get_updater(oc)
    for _ in range(3):
        g = torch.randn(50, 33, device=dev())
        u(0, g, w)
        uc(0, g.cpu(), wc)
    torch.testing.assert_close(w.cpu(), wc, rtol=1e-4, atol=1e-5)
    ids = torch.tensor([4, 9, 4, 30], device=dev())
    gr = torch.randn(4, 33, device=dev())
    u.update_row_sparse(0, ids, gr, w)
    uc.update_row_sparse(0, ids.cpu(), gr.cpu(), wc)
    torch.testing.assert_close(w.cpu(), wc, rtol=1e-4, atol=1e-5)


def test_optimizer_update_row_sparse_cuda_matches_cpu():
    import deeplearning_cfn_b200.kv as kv
    torch.manual_seed(7)
    for name, kw in [("sgd", dict(momentum=0.9, wd=0.01)), ("adam", {}), ("adagrad", {})]:
        w = torch.randn(40, 16, device=dev())
        wc = w.cpu().clone()
        u = kv.optimizer.get_updater(kv.optimizer.create(name, learning_rate=0.1, **kw))
        uc = kv.optimizer.get_updater(kv.optimizer.create(name, learning_rate=0.1, **kw))
        for _ in range(3):
            ids = torch.randint(0, 40, (9,), device=dev())
            g = torch.randn(9, 16, device=dev())
            u.update_row_sparse(0, ids, g, w)
            uc.update_row_sparse(0, ids.cpu(), g.cpu(), wc)
        torch.testing.assert_close(w.cpu(), wc, rtol=1e-4, atol=1e-5)


def test_device_pool_and_pinned_host_pool():
    from deeplearning_cfn_b200.utils import storage
    pool = storage.DevicePool()
    pool.set_strategy("Naive")
    pool.release_all()
    s0 = pool.stats()
    p = pool.alloc(10_000)
    t = torch.empty(0, device=dev())        # make sure a context exists
    assert p % 256 == 0
    pool.free(p)
    q = pool.alloc(12_000)                   # same 12 KiB page class
    assert q == p
    s1 = pool.stats()
    assert s1["reuses"] == s0["reuses"] + 1 and s1["fresh_allocations"] == s0["fresh_allocations"] + 1
    pool.free(q)
    with pytest.raises(RuntimeError):
        pool.free(q)
    pool.release_all()
    assert pool.stats()["bytes_pooled"] == 0
    # pinned host block: async H2D out of the pool, block recycled with its registration
    h = storage.pinned_empty((1 << 20,), torch.float32)
    h.copy_(torch.arange(1 << 20, dtype=torch.float32))
    d = torch.empty(1 << 20, device=dev())
    d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    assert float(d[-1]) == float((1 << 20) - 1)
    addr = h.data_ptr()
    storage.release(h)
    h2 = storage.pinned_empty((1 << 20,), torch.float32)
    assert h2.data_ptr() == addr
    storage.release(h2)
    assert storage.release_all() >= 1
