"""Numerics of the fused dropout + residual add + LayerNorm kernels (csrc/cuda/layernorm.cu, ops/layernorm.py) against plain
PyTorch fp32 references (GPU only).  Validated on B200 in round 2; the fused path is the default of the BERT blocks."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda", 0)


def _frac_close(a, b, rtol, atol, frac=0.999):
    ok = (a.float() - b.float()).abs() <= atol + rtol * b.float().abs()
    assert ok.float().mean().item() >= frac, f"only {ok.float().mean().item():.5f} close; max err {(a.float() - b.float()).abs().max().item()}"


def _unpack_mask(bits: torch.Tensor, rows: int, H: int) -> torch.Tensor:
    b = bits.view(rows, H // 8, 1).to(torch.int32)
    k = torch.arange(8, device=bits.device, dtype=torch.int32).view(1, 1, 8)
    return ((b >> k) & 1).bool().view(rows, H)


@pytest.mark.parametrize("rows,H", [(8, 256), (1000, 1024), (4096, 1024), (77, 768), (33, 2048)])
@pytest.mark.parametrize("with_res", [True, False])
def test_fused_add_layernorm_without_dropout(rows, H, with_res):
    from deeplearning_cfn_b200.ops.layernorm import dropout_add_layer_norm
    torch.manual_seed(0)
    h = torch.randn(rows, H, device=dev()).bfloat16().requires_grad_(True)
    x = torch.randn(rows, H, device=dev()).bfloat16().requires_grad_(True) if with_res else None
    w = (1 + 0.1 * torch.randn(H, device=dev())).bfloat16().requires_grad_(True)
    b = (0.1 * torch.randn(H, device=dev())).bfloat16().requires_grad_(True)
    y = dropout_add_layer_norm(h, x, w, b, p=0.1, eps=1e-12, training=False, force=True)     # eval: no dropout
    gy = torch.randn_like(y)
    y.backward(gy)
    hr = h.detach().float().requires_grad_(True)
    xr = x.detach().float().requires_grad_(True) if with_res else None
    wr, br = w.detach().float().requires_grad_(True), b.detach().float().requires_grad_(True)
    s = hr + xr if with_res else hr
    s = s + (s.detach().bfloat16().float() - s.detach())  # the kernel normalises the bf16-rounded sum (straight-through: value rounded, gradient 1)
    yr = F.layer_norm(s, (H,), wr, br, 1e-12)
    yr.backward(gy.float())
    _frac_close(y, yr, 2e-2, 2e-2)
    _frac_close(h.grad, hr.grad, 3e-2, 3e-2)
    if with_res:
        _frac_close(x.grad, xr.grad, 3e-2, 3e-2)
        assert torch.equal(x.grad, h.grad)               # p = 0: both branches receive the same gradient
    _frac_close(w.grad, wr.grad, 3e-2, 0.03 * float(wr.grad.abs().max()), frac=0.99)
    _frac_close(b.grad, br.grad, 3e-2, 0.03 * float(br.grad.abs().max()), frac=0.99)


@pytest.mark.parametrize("p", [0.1, 0.5])
def test_fused_dropout_add_layernorm_mask_and_gradients(p):
    from deeplearning_cfn_b200.ops import layernorm as ln
    rows, H = 2048, 1024
    torch.manual_seed(1)
    h = torch.randn(rows, H, device=dev()).bfloat16().requires_grad_(True)
    x = torch.randn(rows, H, device=dev()).bfloat16().requires_grad_(True)
    w = torch.ones(H, device=dev()).bfloat16().requires_grad_(True)
    b = torch.zeros(H, device=dev()).bfloat16().requires_grad_(True)
    ln._counter[0] = 1234
    y = ln.dropout_add_layer_norm(h, x, w, b, p=p, eps=1e-12, training=True, force=True)
    mask_bits = y.grad_fn.mask_out
    keep = _unpack_mask(mask_bits, rows, H)
    frac = keep.float().mean().item()
    assert abs(frac - (1 - p)) < 0.01, frac                                  # Bernoulli(1 - p)
    assert abs(keep[:, ::2].float().mean().item() - keep[:, 1::2].float().mean().item()) < 0.01
    # same (seed, offset) -> same mask; next call -> different mask
    ln._counter[0] = 1234
    y2 = ln.dropout_add_layer_norm(h, x, w, b, p=p, eps=1e-12, training=True, force=True)
    assert torch.equal(y, y2)
    y3 = ln.dropout_add_layer_norm(h, x, w, b, p=p, eps=1e-12, training=True, force=True)
    assert not torch.equal(y, y3)
    # forward / backward against torch with the kernel's own mask
    gy = torch.randn_like(y)
    y.backward(gy)
    hr, xr = h.detach().float().requires_grad_(True), x.detach().float().requires_grad_(True)
    s = xr + hr * keep / (1 - p)
    s = s + (s.detach().bfloat16().float() - s.detach())
    yr = F.layer_norm(s, (H,), w.detach().float(), b.detach().float(), 1e-12)
    yr.backward(gy.float())
    _frac_close(y, yr, 2e-2, 2e-2)
    _frac_close(x.grad, xr.grad, 3e-2, 3e-2)
    _frac_close(h.grad, hr.grad, 3e-2, 3e-2)
    assert float(h.grad[~keep].abs().sum()) == 0.0                           # dropped elements get no gradient


def test_bert_layer_same_loss_with_fused_layernorm(monkeypatch):
    from deeplearning_cfn_b200.models import bert_pretrain_loss, get_model
    from deeplearning_cfn_b200.models.resnet import to_b200_layout
    torch.manual_seed(2)
    model = to_b200_layout(get_model("bert-base", max_seq=128, dropout=0.0).to(dev()), torch.bfloat16).train()
    ids = torch.randint(0, 30522, (4, 128), device=dev())
    pos = torch.randint(0, 128, (4, 20), device=dev())
    tgt = (torch.randint(0, 30522, (4, 20), device=dev()), torch.randint(0, 2, (4,), device=dev()))
    out = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("DLCFN_FUSED_LN", flag)
        model.zero_grad(set_to_none=True)
        loss = bert_pretrain_loss(model(ids, torch.zeros_like(ids), pos), tgt)
        loss.backward()
        out[flag] = (loss.item(), model.layers[0].qkv.weight.grad.float().clone())
    assert abs(out["0"][0] - out["1"][0]) < 0.02 * abs(out["0"][0]) + 0.02
    cos = F.cosine_similarity(out["0"][1].flatten(), out["1"][1].flatten(), dim=0).item()
    assert cos > 0.98, cos
