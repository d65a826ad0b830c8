"""tcgen05/TMEM/TMA GEMM vs torch fp32 reference (all operand-major combinations, both tile widths)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from deeplearning_cfn_b200.ops import native
from deeplearning_cfn_b200.ops.linear import Linear, linear


def _mk(rows, cols, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(rows, cols, generator=g) * 0.5).to(device="cuda", dtype=torch.bfloat16)


def _check(out, want, k):
    err = (out.float() - want).abs().max().item()
    scale = want.abs().max().item() + 1e-6
    assert err / scale < 2e-2, f"max rel err {err / scale:.4f} (K={k})"


@pytest.mark.parametrize("block_n", [128, 256])
@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 512, 256), (300, 1000, 2048), (1000, 264, 72), (4096, 1024, 1024)])
def test_gemm(M, N, K, a_mn, b_mn, block_n):
    if a_mn and M % 8:
        M = (M + 7) // 8 * 8   # MN-major operands need a 16-byte row stride (TMA)
    if b_mn and N % 8:
        N = (N + 7) // 8 * 8
    a = _mk(M, K, 1)
    b = _mk(N, K, 2)
    want = a.float() @ b.float().t()
    A = a.t().contiguous() if a_mn else a
    B = b.t().contiguous() if b_mn else b
    out = native.gemm_bf16(A, B, a_mn=a_mn, b_mn=b_mn, block_n=block_n)
    torch.cuda.synchronize()
    _check(out, want, K)


def test_gemm_bias_fp32_accumulate_relu():
    M, N, K = 384, 520, 320
    a, b = _mk(M, K, 3), _mk(N, K, 4)
    bias = torch.randn(N, device="cuda")
    want = a.float() @ b.float().t() + bias
    out = native.gemm_bf16(a, b, bias=bias, out_dtype=torch.float32)
    _check(out, want, K)
    acc = torch.ones(M, N, device="cuda")
    native.gemm_bf16(a, b, out=acc, accumulate=True)
    _check(acc, a.float() @ b.float().t() + 1.0, K)
    r = native.gemm_bf16(a, b, relu=True, out_dtype=torch.float32)
    _check(r, (a.float() @ b.float().t()).clamp_min(0), K)


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (True, True), (False, True)])
@pytest.mark.parametrize("M,N,K", [(64, 64, 8192), (128, 256, 4096), (72, 520, 16384 + 40), (1024, 1024, 8192)])
def test_gemm_split_k(M, N, K, a_mn, b_mn):
    """Few output tiles + long reduction -> the split-K path (fp32 partial tiles + ordered reduce), incl. the
    conv-wgrad shape dW[64,64] = dY^T X with a very long K."""
    a = _mk(M, K, 11) * 0.2
    b = _mk(N, K, 12) * 0.2
    bias = torch.randn(N, device="cuda")
    want = a.float() @ b.float().t()
    A = a.t().contiguous() if a_mn else a
    B = b.t().contiguous() if b_mn else b
    out = native.gemm_bf16(A, B, a_mn=a_mn, b_mn=b_mn)
    _check(out, want, K)
    out32 = native.gemm_bf16(A, B, a_mn=a_mn, b_mn=b_mn, bias=bias, relu=True, out_dtype=torch.float32)
    _check(out32, (want + bias).clamp_min(0), K)
    acc = torch.full((M, N), 2.0, device="cuda")
    native.gemm_bf16(A, B, a_mn=a_mn, b_mn=b_mn, out=acc, accumulate=True)
    _check(acc, want + 2.0, K)
    again = native.gemm_bf16(A, B, a_mn=a_mn, b_mn=b_mn)
    assert torch.equal(out, again), "split-K reduce must be deterministic"


def test_linear_autograd_matches_torch():
    torch.manual_seed(0)
    lin = Linear(2048, 1000).cuda().bfloat16()
    x = torch.randn(256, 2048, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    y = lin(x)
    gy = torch.randn_like(y)
    y.backward(gy)
    xr = x.detach().float().requires_grad_(True)
    wr = lin.weight.detach().float().requires_grad_(True)
    br = lin.bias.detach().float().requires_grad_(True)
    yr = torch.nn.functional.linear(xr, wr, br)
    yr.backward(gy.float())
    _check(y, yr.detach(), 2048)
    _check(x.grad, xr.grad, 1000)
    _check(lin.weight.grad, wr.grad, 256)
    _check(lin.bias.grad, br.grad, 256)
