"""Column-sum (bias gradients of tall Linear layers) and batch-sum (fused stem backward) kernels of csrc/cuda/elementwise.cu
against torch.  The column-sum kernel was written after the round-2 GPU budget was spent (its index arithmetic is checked on
the host, tests/test_kernel_models.py; it ran inside 8-GPU BERT steps without a device error): this is its first numerics run,
so the file sorts late in the GPU tier."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda", 0)


@pytest.mark.parametrize("M,N", [(8192, 1024), (8192, 4096), (77, 1000), (1, 8), (100003, 64)])
def test_colsum_and_batch_sum_match_torch(M, N):
    from deeplearning_cfn_b200.ops import native
    torch.manual_seed(M + N)
    x = torch.randn(M, N, device=dev()).bfloat16()
    got = native.colsum(x)
    want = x.float().sum(0)
    # two fp32 summation orders of M bf16 values: the difference scales with sqrt(M) ulps of the running magnitude
    assert torch.allclose(got, want, rtol=1e-3, atol=2e-3 * max(1.0, M ** 0.5)), (got[:4], want[:4])
    assert torch.equal(got, native.colsum(x))                      # deterministic
    if M <= 8192:
        b = torch.randn(M, 4, 6, 8, device=dev()).bfloat16().contiguous(memory_format=torch.channels_last)
        assert torch.allclose(native.batch_sum(b), b.float().sum(0), rtol=1e-3, atol=2e-3 * max(1.0, M ** 0.5))
