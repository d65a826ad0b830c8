"""Model zoo: every family the reference's `--network` / `--model` switches accept builds with the published parameter count,
runs forward + backward on CPU and keeps its shape contract (the GPU run of the depthwise / 1x1 / fused-BN variants is
tests/test_gpu_z1_zoo.py)."""
import pytest
import torch

from deeplearning_cfn_b200.models import get_model, to_b200_layout


def dev():
    return torch.device("cuda", 0)


def _params(m):
    return sum(p.numel() for p in m.parameters())


# (name, kwargs, parameters at 1000 classes in millions: the published numbers of the papers / model zoos)
PUBLISHED = [("resnet-v1", {"num_layers": 50}, 25.56), ("resnet18_v1", {}, 11.69), ("resnext", {"num_layers": 50}, 25.03),
             ("resnext", {"num_layers": 101}, 44.18), ("resnet", {"num_layers": 50}, 25.55), ("mobilenet", {}, 4.23),
             ("mobilenetv2", {}, 3.50), ("googlenet", {}, 6.99), ("inception-v4", {}, 42.68), ("inception-resnet-v2", {}, 55.84),
             ("squeezenet1.0", {}, 1.25), ("squeezenet1.1", {}, 1.24), ("densenet121", {}, 7.98), ("densenet169", {}, 14.15),
             ("vgg11", {}, 132.86), ("alexnet", {}, None), ("inception-v3", {}, None), ("inception-bn", {}, None)]


@pytest.mark.parametrize("name,kw,millions", PUBLISHED, ids=[f"{n}-{k.get('num_layers', '')}" for n, k, _ in PUBLISHED])
def test_parameter_counts_match_published(name, kw, millions):
    m = get_model(name, 1000, 224, **kw)
    n = _params(m) / 1e6
    assert n > 0.5
    if millions is not None:
        assert abs(n - millions) / millions < 0.012, f"{name}: {n:.2f} M parameters, published {millions} M"


@pytest.mark.parametrize("name,size", [("resnet-v1", 64), ("resnext", 64), ("mobilenet0.5", 64), ("mobilenetv2_0.5", 64),
                                       ("googlenet", 64), ("inception-bn", 64), ("squeezenet1.1", 64), ("densenet121", 64),
                                       ("inception-v4", 139), ("inception-resnet-v2", 139)])
def test_forward_backward_cpu(name, size):
    torch.manual_seed(0)
    m = get_model(name, 7, size)
    x = torch.randn(2, 3, size, size)
    y = m(x)
    assert y.shape == (2, 7) and torch.isfinite(y).all()
    y.logsumexp(1).sum().backward()
    missing = [n for n, p in m.named_parameters() if p.requires_grad and p.grad is None]
    assert not missing, f"no gradient for {missing[:3]}"
    m.eval()
    with torch.no_grad():
        assert m(x).shape == (2, 7)


def test_cifar_variants_and_names():
    assert type(get_model("resnet-v1", 10, 32, num_layers=20)).__name__ == "ResNetV1"
    assert type(get_model("resnet", 10, 32, num_layers=20)).__name__ == "ResNetV2"
    assert type(get_model("resnet50_v2", 10)).__name__ == "ResNetV2"
    assert type(get_model("resnet101_v1", 10)).__name__ == "ResNetV1"
    with pytest.raises(ValueError):
        get_model("resnet-v1", 10, 32, num_layers=21)
    with pytest.raises(ValueError):
        get_model("no-such-net")
    # the 0.25 mobilenet keeps channel counts the NHWC kernels accept (multiples of 8)
    from deeplearning_cfn_b200.ops.nn import DepthwiseConv2d
    assert all(d.channels % 8 == 0 for d in get_model("mobilenet0.25", 10).modules() if isinstance(d, DepthwiseConv2d))
    assert all(d.channels % 8 == 0 for d in get_model("mobilenetv2_0.25", 10).modules() if isinstance(d, DepthwiseConv2d))


@pytest.mark.parametrize("case", [(1, 64, 7, 9, 128, 3, 1, 1), (2, 64, 9, 9, 72, 3, 2, 1), (1, 128, 5, 5, 64, 1, 1, 0)],
                         ids=lambda c: "x".join(map(str, c)))
def test_implicit_gemm_conv_producer_index_math(case):
    """tools/conv_im2col_model.py replays the index arithmetic of the convolution producer warp (k-block -> tap / channel
    slice, tile -> first pixel, reversed taps for dgrad, column panels for wgrad) on a host model of im2col-mode TMA."""
    import importlib.util
    import pathlib
    spec = importlib.util.spec_from_file_location("conv_im2col_model",
                                                  pathlib.Path(__file__).resolve().parents[1] / "tools" / "conv_im2col_model.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    errs = mod.check(*case)
    assert "fprop" in errs and "wgrad" in errs and all(v < 1e-9 for v in errs.values()), errs


def test_bert_qkv_planes_function_matches_the_packed_projection(monkeypatch):
    """`models.bert._QKVFn` (q / k / v as three dense planes; opt-in DLCFN_BERT_QKV_PLANES=1, not yet run on a GPU) against the
    packed projection it replaces, with the tcgen05 GEMM and the column-sum kernel replaced by host references of their
    contracts (operand storage flags, `out=` views, `accumulate`): forward planes, dx (three accumulating GEMMs), dW (row blocks
    of one gradient) and db agree."""
    import torch.nn.functional as F
    from deeplearning_cfn_b200.models import bert
    from deeplearning_cfn_b200.ops import native

    def gemm_ref(a, b, *, a_mn=False, b_mn=False, out=None, out_dtype=torch.bfloat16, bias=None, accumulate=False, **_):
        A = a.float().t() if a_mn else a.float()              # [M, K]
        Bt = b.float() if b_mn else b.float().t()             # [K, N]
        d = A @ Bt
        if bias is not None:
            d = d + bias.float()[: d.shape[1]]
        if out is None:
            return d.to(out_dtype)
        assert out.shape == d.shape and out.stride(1) == 1
        out.copy_((out.float() + d) if accumulate else d)
        return out

    monkeypatch.setattr(native, "gemm_bf16", gemm_ref)
    monkeypatch.setattr(native, "colsum", lambda t: t.float().sum(0))
    torch.manual_seed(0)
    B, S, H = 2, 8, 32
    w = (torch.randn(3 * H, H) * 0.2).bfloat16().requires_grad_(True)
    b = (torch.randn(3 * H) * 0.1).requires_grad_(True)
    x = torch.randn(B, S, H).bfloat16().requires_grad_(True)
    g = torch.randn(3, B * S, H).bfloat16()
    planes = bert._QKVFn.apply(x, w, b)
    assert planes.shape == (3, B * S, H) and planes.is_contiguous()
    planes.backward(g)
    got = (planes.detach().float(), x.grad.float().clone(), w.grad.float().clone(), b.grad.float().clone())
    xr, wr, br = (t.detach().float().requires_grad_(True) for t in (x, w, b))
    ref = F.linear(xr, wr, br).reshape(-1, 3, H).permute(1, 0, 2)
    ref.backward(g.float())
    for a, r in zip(got, (ref.detach(), xr.grad, wr.grad, br.grad)):
        assert a.shape == r.shape
        assert torch.allclose(a, r, rtol=3e-2, atol=3e-2 * float(r.abs().max())), (a - r).abs().max()
    # and the module's view chain hands attention the same [3, B, h, S, d] tensor either way
    heads = 4
    v1 = planes.detach().reshape(3, B, S, heads, H // heads).transpose(2, 3)
    v2 = F.linear(x.detach().float(), w.detach().float(), b.detach()).view(B, S, 3, heads, H // heads).permute(2, 0, 3, 1, 4)
    assert torch.allclose(v1.float(), v2, rtol=3e-2, atol=3e-2)
