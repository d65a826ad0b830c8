"""The CUDA branches of the Python layer, walked on the host with a stand-in GPU (tests/mockgpu.py): argument tables,
constructors, shapes, strides and the set of kernels each path launches.  Numerics are the GPU tier's job."""
import pytest
import torch

from mockgpu import mock_gpu

BF = torch.bfloat16
CL = torch.channels_last


def _cl(*shape, dtype=BF):
    return torch.randn(*shape).to(dtype).contiguous(memory_format=CL)


def test_experimental_wrappers_conv_roi_layernorm():
    from deeplearning_cfn_b200.ops import conv, layernorm as ln, native, roi
    with mock_gpu() as log:
        x, w = _cl(2, 64, 8, 8), _cl(128, 64, 3, 3)
        y = native.conv_fprop(x, w, 1, 1)
        assert y.shape == (2, 128, 8, 8) and y.is_contiguous(memory_format=CL)
        assert native.conv_dgrad(y, w, 1, (8, 8)).shape == x.shape
        dw = native.conv_wgrad(x, y, (3, 3), 1, 1)
        assert dw.shape == w.shape and dw.is_contiguous(memory_format=CL)
        assert native.conv_fprop(x, w, 2, 1).shape == (2, 128, 4, 4)
        conv.set_convkxk_backend("tcgen05")
        try:
            m = conv.ConvKxK(64, 128, 3, 2, 1).to(BF).to(memory_format=CL)
            xx = x.clone().requires_grad_(True)
            out = m(xx)
            out.sum().backward()                                  # strided: fprop, wgrad and the parity-class dgrad on the kernel
            assert out.shape == (2, 128, 4, 4) and xx.grad.shape == x.shape and m.weight.grad.shape == (128, 64, 3, 3)
            m1 = conv.ConvKxK(64, 64, 3, 1, 1).to(BF).to(memory_format=CL)
            n0 = len(log)
            m1(xx).sum().backward()
            assert log[n0:].count("dlcfn_conv_bf16") + log[n0:].count("dlcfn_conv_bf16_ex") == 3   # fprop + dgrad + wgrad
            odd = conv.ConvKxK(24, 64, 3, 1, 1).to(BF).to(memory_format=CL)   # Cin % 64 != 0: stays on aten
            n0 = len(log)
            odd(_cl(1, 24, 8, 8))
            assert "dlcfn_conv_bf16" not in log[n0:] and "dlcfn_conv_bf16_ex" not in log[n0:]
        finally:
            conv.set_convkxk_backend("tcgen05")

        roi.set_backend("dlcfn")
        try:
            f = torch.randn(2, 16, 10, 12).contiguous(memory_format=CL).requires_grad_(True)
            rois = torch.tensor([[0, 1., 1., 8., 8.], [1, 2., 2., 9., 7.]])
            o = roi.roi_align(f, rois, 7, 0.5, 2, True)
            o.sum().backward()
            assert o.shape == (2, 16, 7, 7) and f.grad.shape == f.shape and f.grad.dtype == torch.float32
            ob = roi.roi_align(_cl(2, 16, 10, 12).requires_grad_(True), [rois[:1, 1:], rois[1:, 1:]], (3, 5), 1.0, -1, False)
            assert ob.shape == (2, 16, 3, 5) and ob.dtype == BF
            keep = roi.nms(torch.rand(10, 4), torch.rand(10), 0.5)
            assert keep.dtype == torch.long and keep.dim() == 1
            assert roi.nms(torch.zeros(0, 4), torch.zeros(0), 0.5).numel() == 0
        finally:
            roi.set_backend("torchvision")

        h, r = torch.randn(16, 256).to(BF).requires_grad_(True), torch.randn(16, 256).to(BF).requires_grad_(True)
        g, b = torch.ones(256).to(BF).requires_grad_(True), torch.zeros(256).to(BF).requires_grad_(True)
        z = ln.dropout_add_layer_norm(h, r, g, b, p=0.1, eps=1e-12, training=True, force=True)
        z.sum().backward()
        assert z.shape == h.shape and h.grad.shape == h.shape and g.grad.shape == g.shape
    launched = set(log)
    assert {"dlcfn_conv_bf16", "dlcfn_conv_bf16_ex", "dlcfn_roi_align_fwd", "dlcfn_roi_align_bwd", "dlcfn_nms", "dlcfn_dropout_add_ln_fwd",
            "dlcfn_dropout_add_ln_bwd"} <= launched, launched


@pytest.mark.parametrize("backend", ["cudnn", "tcgen05"])
def test_resnet50_training_step_walks_every_cuda_branch(backend):
    """bf16 channels_last ResNet-50 through DataParallelTrainer with every dlcfn kernel stubbed: the stem (pad 3->8, space to
    depth), fused BN / add+BN, NHWC max-pool, 1x1 / 3x3 convolutions on the chosen backend, tcgen05 classifier, fused loss,
    bucket gather, fused multi-precision optimizer."""
    from deeplearning_cfn_b200.models import resnet, to_b200_layout
    from deeplearning_cfn_b200.ops import conv
    from deeplearning_cfn_b200.parallel.trainer import DataParallelTrainer
    torch.manual_seed(0)
    conv.set_conv1x1_backend(backend)
    conv.set_convkxk_backend(backend)
    try:
        with mock_gpu() as log:
            model = to_b200_layout(resnet(50, 16, 64)).train()
            tr = DataParallelTrainer(model, mode="local", optimizer="sgd_mom", lr=0.05)
            x, y = _cl(4, 3, 64, 64), torch.randint(0, 16, (4,))
            loss = tr.step(x, y)
            tr.step(x, y)
            assert loss.dim() == 0
    finally:
        conv.set_conv1x1_backend("tcgen05")
        conv.set_convkxk_backend("tcgen05")
    per_step = {k: log.count(k) // 2 for k in set(log)}
    gemms = per_step.get("dlcfn_gemm_bf16", 0) + per_step.get("dlcfn_gemm_bf16_ex", 0)
    convs = per_step.get("dlcfn_conv_bf16", 0) + per_step.get("dlcfn_conv_bf16_ex", 0)
    assert gemms >= 3                                                    # classifier fwd / dgrad / wgrad at least
    assert any(k.startswith("dlcfn_bn") for k in per_step) and any("softmax_ce" in k for k in per_step)
    assert any("maxpool" in k for k in per_step) and per_step.get("dlcfn_fused_optimizer", 0) >= 1
    if backend == "tcgen05":
        assert convs >= 16 * 3 + 3 * 3 + 2                              # 16 3x3 layers + 3 strided 1x1 shortcuts (fprop, dgrad, wgrad) + stem
        assert gemms >= 3 * 30                                          # the 1x1 layers as GEMMs
        assert per_step.get("dlcfn_bn_forward_pre", 0) >= 16 * 2 + 1    # bn2 / bn3 of every unit + bn0 take the epilogue statistics
    else:
        assert convs == 2                                               # the fused stem (fprop + wgrad) is independent of the backend switch


def test_bert_step_and_zoo_models_walk_their_cuda_branches():
    from deeplearning_cfn_b200.models import Bert, bert_pretrain_loss, get_model, to_b200_layout
    from deeplearning_cfn_b200.ops import layernorm as ln
    from deeplearning_cfn_b200.parallel.trainer import DataParallelTrainer
    torch.manual_seed(0)
    with mock_gpu() as log:
        model = to_b200_layout(Bert(hidden=256, layers=2, heads=4, ffn=512, max_seq=16), BF).train()
        model.padded_logits = True
        tr = DataParallelTrainer(model, mode="local", optimizer="adam", lr=1e-4, loss_fn=bert_pretrain_loss)
        ids = torch.randint(0, 30522, (2, 16))
        x = (ids, torch.zeros_like(ids), torch.randint(0, 16, (2, 5)))
        y = (torch.randint(0, 30522, (2, 5)), torch.randint(0, 2, (2,)))
        old = ln.enabled
        ln.enabled = lambda: True
        try:
            assert tr.step(x, y).dim() == 0
        finally:
            ln.enabled = old
        n_bert = len(log)
        assert "dlcfn_dropout_add_ln_fwd" in log and "dlcfn_gemm_bf16_ex" in log and any("softmax_ce" in k for k in log)
        for name in ("mobilenet0.5", "mobilenetv2_0.5", "densenet121", "inception-bn", "cifar10_cnn", "alexnet"):
            size = {"cifar10_cnn": 32, "alexnet": 160}.get(name, 64)
            m = to_b200_layout(get_model(name, 8, size)).train()
            out = m(_cl(2, 3, size, size))
            assert out.shape == (2, 8), name
            out.float().logsumexp(1).sum().backward()
    tail = set(log[n_bert:])
    assert any("depthwise" in k for k in tail) and any("lrn" in k for k in tail) and any(k.startswith("dlcfn_bn") for k in tail)


@pytest.mark.parametrize("name,size", [("vgg11", 32), ("vgg16_bn", 32), ("inception-v3", 139), ("googlenet", 64), ("squeezenet1.1", 64),
                                       ("resnext", 64), ("resnet-v1", 64), ("resnet18_v1", 64), ("inception-v4", 139),
                                       ("inception-resnet-v2", 139), ("lenet", 28), ("mlp", 28)])
def test_every_model_family_runs_its_cuda_branches(name, size):
    from deeplearning_cfn_b200.models import get_model, to_b200_layout
    torch.manual_seed(0)
    with mock_gpu() as log:
        m = to_b200_layout(get_model(name, 8, size)).train()
        ch = 1 if name in ("lenet", "mlp") else 3
        x = _cl(2, ch, size, size)
        if name == "mlp":
            x = x.reshape(2, -1)
        out = m(x)
        assert out.shape == (2, 8)
        out.float().logsumexp(1).sum().backward()
        missing = [n for n, p in m.named_parameters() if p.requires_grad and p.grad is None]
        assert not missing, missing[:3]
    assert "dlcfn_gemm_bf16_ex" in log or name.startswith("squeezenet")     # every family but SqueezeNet (conv classifier) ends in the tcgen05 Linear


def test_elementwise_compression_and_sparse_wrappers():
    from deeplearning_cfn_b200.ops import native, sparse
    with mock_gpu() as log:
        a, b = torch.randn(1000).to(BF), torch.randn(1000).to(BF)
        out = torch.empty(1000, dtype=BF)
        native.elementwise_sum([a, b], out, scale=0.5)
        native.cast_scale(a, torch.empty(1000), 2.0)
        g, master, mom = torch.randn(1024).to(BF), torch.randn(1024), torch.zeros(1024)
        for kind in native.OPT_KINDS:
            s1 = mom if kind in native.OPT_USES_S1 else None
            s2 = torch.zeros(1024) if kind in native.OPT_USES_S2 else None
            native.fused_optimizer(kind, g, master, s1, s2, torch.empty(1024, dtype=BF), lr=0.1, wd=1e-4, momentum=0.9)
        grad, res = torch.randn(160), torch.zeros(160)
        words = torch.empty(10, dtype=torch.int32)
        native.quantize_2bit(grad, res, words, 0.5)
        native.dequantize_2bit(words, torch.empty(160), 0.5)
        img = torch.randint(0, 255, (2, 8, 8, 3), dtype=torch.uint8)
        native.normalize_u8_nhwc(img, torch.empty(2, 8, 8, 3, dtype=BF), (0.4, 0.4, 0.4), (0.2, 0.2, 0.2))
        w = torch.randn(50, 16)
        rows = torch.tensor([1, 7, 9])
        sparse.sparse_row_update("sgd_mom", rows, torch.randn(3, 16), w, torch.zeros(50, 16), None, torch.empty(50, 16, dtype=BF),
                                 lr=0.1, momentum=0.9)
        sparse.group_adagrad_update(rows, torch.randn(3, 16), w, torch.zeros(50), lr=0.1)
        sparse.group_adagrad_update(None, torch.randn(50, 16), w, torch.zeros(50), lr=0.1)
    assert {"dlcfn_elementwise_sum", "dlcfn_cast_scale", "dlcfn_fused_optimizer", "dlcfn_quantize_2bit", "dlcfn_dequantize_2bit",
            "dlcfn_normalize_u8_nhwc", "dlcfn_sparse_row_update", "dlcfn_group_adagrad_update"} <= set(log)


def test_kvstore_compression_on_device_gradients_uses_the_2bit_kernels():
    """`local` store, 2-bit compression, gradients that live "on the GPU": quantize / dequantize go through the kernels
    (csrc/cuda/compress.cu) with a per-(key, device) residual, the reduce falls back to the host reducer."""
    import deeplearning_cfn_b200.kv as kv
    with mock_gpu() as log:
        store = kv.create("local")
        store.set_gradient_compression({"type": "2bit", "threshold": 0.5})
        store.init(3, torch.zeros(100))
        store.push(3, [torch.randn(100), torch.randn(100).to(BF)])
        out = torch.empty(100)
        store.pull(3, out=out)
        assert out.shape == (100,) and len(store._residuals) == 2
        assert all(r.dtype == torch.float32 and r.numel() == 100 for r in store._residuals.values())
    assert log.count("dlcfn_quantize_2bit") == 2 and log.count("dlcfn_dequantize_2bit") == 2


_FAMILIES = [("resnet18", 64), ("resnet50", 64), ("resnet50-v1", 64), ("resnext50", 64), ("vgg16-bn", 64), ("alexnet", 224),
             ("googlenet", 96), ("inception-bn", 96), ("inception-v3", 299), ("inception-v4", 299), ("inception-resnet-v2", 299),
             ("mobilenet0.25", 64), ("mobilenetv2", 64), ("mobilenetv2-0.25", 64), ("squeezenet1.1", 96), ("densenet121", 64),
             ("cifar10-cnn", 32), ("resnet20", 32)]


@pytest.mark.parametrize("name,size", _FAMILIES, ids=[n for n, _ in _FAMILIES])
def test_every_model_family_walks_forward_and_backward_past_the_entry_point_checks(name, size):
    """bf16 channels_last forward + backward of every model family on the stand-in GPU, whose GEMM / convolution stand-ins
    apply the argument checks of the C entry points (16-byte alignment, leading dimensions, 64-channel TMA rows, geometry): a
    layer shape that the box would answer with cudaErrorInvalidValue fails here.  ResNets must actually take the tcgen05 path."""
    from deeplearning_cfn_b200.models import get_model, to_b200_layout
    torch.manual_seed(0)
    with mock_gpu() as log:
        m = to_b200_layout(get_model(name, 16, size)).train()
        x = torch.randn(2, 3, size, size).bfloat16().contiguous(memory_format=torch.channels_last)
        y = m(x)
        assert y.shape == (2, 16)
        y.float().logsumexp(1).sum().backward()
    assert all(p.grad is not None for p in m.parameters() if p.requires_grad)
    if name in ("resnet18", "resnet50"):             # (the CIFAR ResNets have 16- / 32-channel stages that stay on the library)
        assert sum(log.count(f) for f in ("dlcfn_conv_bf16_ex", "dlcfn_conv_bf16", "dlcfn_gemm_bf16_ex")) >= 60, len(log)
