"""Numerics of kernels that were written without GPU access at the end of round 1 and are OFF by default.
Skipped unless DLCFN_EXPERIMENTAL=1:   DLCFN_EXPERIMENTAL=1 python -m pytest tests/test_gpu_experimental.py -m gpu -q
Once these pass on a B200 the flags (DLCFN_FUSED_LN) can become defaults and the tests move to test_gpu_nn.py."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("DLCFN_EXPERIMENTAL", "0") != "1", reason="experimental kernels: set DLCFN_EXPERIMENTAL=1")]


def dev():
    return torch.device("cuda", 0)


def _frac_close(a, b, rtol, atol, frac=0.999):
    ok = (a.float() - b.float()).abs() <= atol + rtol * b.float().abs()
    assert ok.float().mean().item() >= frac, f"only {ok.float().mean().item():.5f} close; max err {(a.float() - b.float()).abs().max().item()}"


_MC_SCRIPT = r"""
import sys, torch
sys.path.insert(0, %r)
from deeplearning_cfn_b200.ops import native
torch.manual_seed(0)
ok = True
for (M, N, K, a_mn, b_mn) in [(4096, 4096, 512, 0, 0), (8192, 2048, 1024, 0, 1), (4224, 4096, 320, 1, 1), (19072, 512, 256, 1, 0),
                              (8192, 8192, 8192, 0, 0)]:
    a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    b = (torch.randn(N, K, device="cuda") * 0.5).bfloat16()
    A = a.t().contiguous() if a_mn else a
    B = b.t().contiguous() if b_mn else b
    out = native.gemm_bf16(A, B, a_mn=bool(a_mn), b_mn=bool(b_mn), block_n=256)
    ref = native.gemm_bf16(A, B, a_mn=bool(a_mn), b_mn=bool(b_mn), block_n=128)     # validated single-CTA kernel
    torch.cuda.synchronize()
    want = a.float() @ b.float().t()
    err = ((out.float() - want).abs().max() / (want.abs().max() + 1e-6)).item()
    same = torch.equal(out, ref)
    print(M, N, K, a_mn, b_mn, "rel err %%.4f" %% err, "bitwise equal to the single-CTA kernel:", same, flush=True)
    ok = ok and err < 2e-2
print("MC_OK" if ok else "MC_WRONG", flush=True)
"""


def test_multicast_pair_gemm_in_a_subprocess():
    """The 2-CTA-cluster multicast GEMM (DLCFN_GEMM_MULTICAST=1) against torch and against the validated single-CTA kernel.
    Runs in a child under a timeout: a protocol bug in a cluster kernel shows up as a hang, not as a wrong number."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run(["timeout", "120", sys.executable, "-c", _MC_SCRIPT % root],
                       env=dict(os.environ, DLCFN_GEMM_MULTICAST="1"), capture_output=True, text=True)
    assert "MC_OK" in r.stdout, f"rc={r.returncode}\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}"


# ----------------------------------------------------------- GEMM variants selected by environment (read once by the .so)
_GEMM_VARIANT = "2cta" if os.environ.get("DLCFN_GEMM_2CTA") == "1" else ("multicast" if os.environ.get("DLCFN_GEMM_MULTICAST") == "1" else None)


@pytest.mark.skipif(_GEMM_VARIANT is None, reason="run with DLCFN_GEMM_2CTA=1 or DLCFN_GEMM_MULTICAST=1 in the environment")
@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K,epi", [(4096, 4096, 1024, False), (4224, 2048, 512, True), (8192, 1000, 256, False),
                                       (16384, 512, 64, True), (5000, 3000, 1096, False)])
def test_gemm_pair_variants_match_fp32(M, N, K, epi, a_mn, b_mn):
    """Shapes with enough 256-wide tile pairs to take the CTA-pair / multicast kernel: odd tile-row counts (phantom second
    tile), N tails, K tails, all four operand layouts, bias + relu epilogue."""
    from deeplearning_cfn_b200.ops import native
    torch.manual_seed(M + N + K)
    if a_mn and M % 8:
        M = (M + 7) // 8 * 8
    if b_mn and N % 8:
        N = (N + 7) // 8 * 8
    a = torch.randn(M, K, device=dev()).bfloat16()
    b = torch.randn(N, K, device=dev()).bfloat16()
    bias = torch.randn(N, device=dev()) if epi else None
    A = a.t().contiguous() if a_mn else a
    B = b.t().contiguous() if b_mn else b
    d = native.gemm_bf16(A, B, a_mn=a_mn, b_mn=b_mn, bias=bias, relu=epi, block_n=256)
    ref = a.float() @ b.float().t()
    if epi:
        ref = torch.relu(ref + bias)
    assert d.shape == (M, N)
    _frac_close(d, ref, 2e-2, 2e-2 * float(ref.abs().max()), frac=0.9999)
    assert torch.equal(d, native.gemm_bf16(A, B, a_mn=a_mn, b_mn=b_mn, bias=bias, relu=epi, block_n=256))


# (the CUDA-graph step test moved to tests/test_gpu_y1_graph.py: validated on B200 in round 2, the bench replays the graph by default)


# (the ROIAlign / NMS kernel tests moved to tests/test_gpu_y2_roi.py: they passed on B200 at the start of round 2)


# -------------------------------------------------------------------------- hierarchical (multi-node) trainer mode on GPUs
@pytest.mark.multigpu
def test_hier_mode_virtual_nodes_agree_with_nccl():
    """`hier` with virtual nodes on one box: node-local fused all-reduce + NCCL rails + node-local broadcasts must train like
    the flat NCCL all-reduce.  2 GPUs -> 2 nodes x 1 GPU (degenerate node-local steps), >= 4 GPUs -> 2 nodes x 2 GPUs."""
    import sys
    from conftest import run_job
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    n = 4 if torch.cuda.device_count() >= 4 else 2
    env = dict(os.environ, DLCFN_TEST_MODES="nccl,hier", DLCFN_VIRTUAL_NODE_SIZE=str(n // 2))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29641", os.path.join(root, "tests", "mp", "trainer_worker.py")]
    r = run_job(cmd, env=env, timeout=600, cwd=root)
    assert r.returncode == 0 and "TRAINER_OK" in r.stdout and "hier:" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
