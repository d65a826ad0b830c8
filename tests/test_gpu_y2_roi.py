"""ROIAlign (forward / backward) and NMS kernels of csrc/cuda/roi_ops.cu (`ops/roi.py`, backend "dlcfn"; torchvision's operators
stay the default of the Mask-RCNN example) against torchvision.  Passed on B200 at the start of round 2 (gpurun call A)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
pytest.importorskip("torchvision")


def dev():
    return torch.device("cuda", 0)


def _frac_close(a, b, rtol, atol, frac=0.999):
    ok = (a.float() - b.float()).abs() <= atol + rtol * b.float().abs()
    assert ok.float().mean().item() >= frac, f"only {ok.float().mean().item():.5f} close; max err {(a.float() - b.float()).abs().max().item()}"


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("aligned,sr", [(True, 2), (False, 2), (True, -1)])
def test_roi_align_kernels_match_torchvision(dtype, aligned, sr):
    import torchvision.ops as tvo
    from deeplearning_cfn_b200.ops import roi
    torch.manual_seed(0)
    x = torch.randn(3, 64, 50, 68, device=dev()).to(dtype).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    g = torch.Generator().manual_seed(1)
    xy = torch.rand(200, 2, generator=g) * torch.tensor([250.0, 180.0]) - 10
    wh = torch.rand(200, 2, generator=g) * 120 + 0.5
    rois = torch.cat([torch.randint(0, 3, (200, 1), generator=g).float(), xy, xy + wh], 1).to(dev())
    roi.set_backend("dlcfn")
    try:
        y = roi.roi_align(x, rois, (7, 7), 0.25, sr, aligned)
        gy = torch.randn_like(y)
        y.backward(gy)
    finally:
        roi.set_backend("torchvision")
    xr = x.detach().float().requires_grad_(True)
    yr = tvo.roi_align(xr, rois, (7, 7), 0.25, sr, aligned)
    yr.backward(gy.float())
    tol = 1e-4 if dtype == torch.float32 else 2e-2
    assert y.shape == yr.shape
    _frac_close(y, yr, tol, tol * float(yr.abs().max()), frac=0.9999)
    _frac_close(x.grad, xr.grad, tol, tol * float(xr.grad.abs().max()) + 1e-6, frac=0.999)


@pytest.mark.parametrize("n", [1, 63, 64, 65, 1000, 6000])
def test_nms_kernel_matches_torchvision(n):
    import torchvision.ops as tvo
    from deeplearning_cfn_b200.ops import roi
    g = torch.Generator().manual_seed(n)
    xy = torch.rand(n, 2, generator=g) * 300
    boxes = torch.cat([xy, xy + 4 + torch.rand(n, 2, generator=g) * 60], 1).to(dev())
    scores = torch.rand(n, generator=g).to(dev())
    roi.set_backend("dlcfn")
    try:
        for thr in (0.3, 0.5, 0.7):
            keep = roi.nms(boxes, scores, thr)
            assert torch.equal(keep.cpu(), tvo.nms(boxes.cpu(), scores.cpu(), thr)), (n, thr)
        cls = torch.randint(0, 5, (n,), generator=g).to(dev())
        assert torch.equal(roi.batched_nms(boxes, scores, cls, 0.5).cpu(), tvo.batched_nms(boxes.cpu(), scores.cpu(), cls.cpu(), 0.5))
    finally:
        roi.set_backend("torchvision")
