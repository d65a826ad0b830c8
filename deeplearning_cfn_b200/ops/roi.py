"""ROIAlign and NMS for the Mask R-CNN workload (W8) on the dlcfn NHWC kernels (csrc/cuda/roi_ops.cu, EXPERIMENTAL).

Reference: MX/src/operator/contrib/roi_align.cu (ROIAlign forward / backward), multi_proposal.cu:275-400 (NMS bitmask);
TensorPack's Mask R-CNN (the job `run.sh` launches) uses the same two primitives from TensorFlow.

    y = roi_align(x, rois, output_size=7, spatial_scale=0.25, sampling_ratio=2, aligned=True)     # x [N,C,H,W]
    keep = nms(boxes, scores, iou_threshold=0.5)                                                  # indices, best first

`roi_align` / `nms` use the CUDA kernels when `DLCFN_ROI_OPS=dlcfn` (or `set_backend("dlcfn")`) and the tensors live on a
GPU; the default, and the only path on CPU tensors, is torchvision's operators (the library yardstick).  `roi_align_ref` /
`nms_ref` are plain-torch statements of the semantics the kernels implement; CPU tests hold them against torchvision, GPU
tests hold the kernels against them.
"""
from __future__ import annotations

import math
import os

import torch

from . import native

_BACKEND = os.environ.get("DLCFN_ROI_OPS", "torchvision").lower()


def set_backend(name: str) -> None:
    global _BACKEND
    if name not in ("dlcfn", "torchvision"):
        raise ValueError("backend must be 'dlcfn' or 'torchvision'")
    _BACKEND = name


def backend() -> str:
    return _BACKEND


def _declare():
    import ctypes
    L = native.lib()
    if getattr(L, "_roi_declared", False):
        return L
    vp, i, f = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
    L.dlcfn_roi_align_fwd.argtypes = [vp, vp, vp, i, i, i, i, i, i, i, i, f, i, i, vp]
    L.dlcfn_roi_align_bwd.argtypes = [vp, vp, vp, i, i, i, i, i, i, i, i, f, i, i, vp]
    L.dlcfn_nms.argtypes = [vp, i, f, vp, vp, vp, vp]
    for fn in ("dlcfn_roi_align_fwd", "dlcfn_roi_align_bwd", "dlcfn_nms"):
        getattr(L, fn).restype = i
    L._roi_declared = True
    return L


def _as_rois(boxes, dtype=torch.float32) -> torch.Tensor:
    """Tensor[K,5] (batch index first) or a list of per-image Tensor[L,4] -> Tensor[K,5] (fp32 for the kernels)."""
    if isinstance(boxes, (list, tuple)):
        parts = [torch.cat([torch.full((b.shape[0], 1), float(i), dtype=b.dtype, device=b.device), b], 1)
                 for i, b in enumerate(boxes)]
        boxes = torch.cat(parts, 0) if parts else torch.zeros(0, 5)
    return boxes.to(dtype).contiguous()


def _pair(v):
    return (int(v), int(v)) if isinstance(v, int) else (int(v[0]), int(v[1]))


# ------------------------------------------------------------------------------------------------ plain-torch semantics
def roi_align_ref(x: torch.Tensor, rois: torch.Tensor, output_size, spatial_scale: float = 1.0, sampling_ratio: int = -1,
                  aligned: bool = False) -> torch.Tensor:
    """Loop-level statement of ROIAlign (what kroi_align_fwd computes); differentiable, slow - for tests."""
    PH, PW = _pair(output_size)
    N, C, H, W = x.shape
    rois = _as_rois(rois, torch.float64)
    out = []
    for r in range(rois.shape[0]):
        b = int(rois[r, 0])
        off = 0.5 if aligned else 0.0
        sw, sh = float(rois[r, 1]) * spatial_scale - off, float(rois[r, 2]) * spatial_scale - off
        rw, rh = float(rois[r, 3]) * spatial_scale - off - sw, float(rois[r, 4]) * spatial_scale - off - sh
        if not aligned:
            rw, rh = max(rw, 1.0), max(rh, 1.0)
        bw, bh = rw / PW, rh / PH
        gw = sampling_ratio if sampling_ratio > 0 else int(math.ceil(rw / PW))
        gh = sampling_ratio if sampling_ratio > 0 else int(math.ceil(rh / PH))
        cnt = max(gw * gh, 1)
        res = x.new_zeros(C, PH, PW, dtype=torch.float32)
        for ph in range(PH):
            for pw in range(PW):
                acc = x.new_zeros(C, dtype=torch.float32)
                for iy in range(gh):
                    yy = sh + ph * bh + (iy + 0.5) * bh / gh
                    for ix in range(gw):
                        xx = sw + pw * bw + (ix + 0.5) * bw / gw
                        if yy < -1.0 or yy > H or xx < -1.0 or xx > W:
                            continue
                        y0, x0 = max(yy, 0.0), max(xx, 0.0)
                        yl, xl = int(y0), int(x0)
                        if yl >= H - 1:
                            yl = yh = H - 1
                            y0 = float(yl)
                        else:
                            yh = yl + 1
                        if xl >= W - 1:
                            xl = xh = W - 1
                            x0 = float(xl)
                        else:
                            xh = xl + 1
                        ly, lx = y0 - yl, x0 - xl
                        hy, hx = 1.0 - ly, 1.0 - lx
                        f = x[b].float()
                        acc = acc + hy * hx * f[:, yl, xl] + hy * lx * f[:, yl, xh] + ly * hx * f[:, yh, xl] + ly * lx * f[:, yh, xh]
                res[:, ph, pw] = acc / cnt
        out.append(res)
    return torch.stack(out).to(x.dtype) if out else x.new_zeros(0, C, PH, PW)


def nms_ref(boxes: torch.Tensor, scores: torch.Tensor, iou_threshold: float) -> torch.Tensor:
    """Greedy NMS: visit boxes by descending score, drop the ones whose IoU with a kept box exceeds the threshold."""
    order = torch.argsort(scores, descending=True, stable=True)
    b = boxes[order].float()
    area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    removed = torch.zeros(b.shape[0], dtype=torch.bool)
    keep = []
    for i in range(b.shape[0]):
        if removed[i]:
            continue
        keep.append(i)
        lt, rb = torch.maximum(b[i, :2], b[i + 1:, :2]), torch.minimum(b[i, 2:], b[i + 1:, 2:])
        wh = (rb - lt).clamp(min=0)
        inter = wh[:, 0] * wh[:, 1]
        removed[i + 1:] |= inter > iou_threshold * (area[i] + area[i + 1:] - inter)
    return order[torch.tensor(keep, dtype=torch.long)] if keep else order[:0]


# ------------------------------------------------------------------------------------------------------- CUDA path
class _RoiAlignFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, rois, PH, PW, scale, sampling_ratio, aligned):
        L = _declare()
        if not x.is_contiguous(memory_format=torch.channels_last):
            x = x.contiguous(memory_format=torch.channels_last)
        N, C, H, W = x.shape
        R = rois.shape[0]
        y = torch.empty((R, C, PH, PW), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
        native.check(L.dlcfn_roi_align_fwd(x.data_ptr(), rois.data_ptr(), y.data_ptr(), 0 if x.dtype == torch.float32 else 1,
                                           N, H, W, C, R, PH, PW, float(scale), int(sampling_ratio), int(aligned),
                                           native.stream_ptr()), "roi_align_fwd")
        native._count()
        ctx.save_for_backward(rois)
        ctx.geom = (N, C, H, W, PH, PW, float(scale), int(sampling_ratio), int(aligned), x.dtype)
        return y

    @staticmethod
    def backward(ctx, dy):
        L = _declare()
        (rois,) = ctx.saved_tensors
        N, C, H, W, PH, PW, scale, sr, aligned, dt = ctx.geom
        if dy.dtype != dt:
            dy = dy.to(dt)
        if not dy.is_contiguous(memory_format=torch.channels_last):
            dy = dy.contiguous(memory_format=torch.channels_last)
        dx32 = torch.zeros((N, H, W, C), dtype=torch.float32, device=dy.device).permute(0, 3, 1, 2)   # channels_last view
        native.check(L.dlcfn_roi_align_bwd(dy.data_ptr(), rois.data_ptr(), dx32.data_ptr(), 0 if dt == torch.float32 else 1,
                                           N, H, W, C, rois.shape[0], PH, PW, scale, sr, aligned, native.stream_ptr()),
                     "roi_align_bwd")
        native._count()
        return (dx32 if dt == torch.float32 else dx32.to(dt)), None, None, None, None, None, None


def roi_align(x: torch.Tensor, boxes, output_size, spatial_scale: float = 1.0, sampling_ratio: int = -1,
              aligned: bool = False) -> torch.Tensor:
    """torchvision.ops.roi_align signature.  x [N,C,H,W]; boxes Tensor[K,5] or list of Tensor[L,4]."""
    PH, PW = _pair(output_size)
    if _BACKEND == "dlcfn" and x.is_cuda and x.dtype in (torch.float32, torch.bfloat16):
        rois = _as_rois(boxes).to(x.device)
        return _RoiAlignFn.apply(x, rois, PH, PW, spatial_scale, sampling_ratio, aligned)
    import torchvision.ops as tvo
    if isinstance(boxes, (list, tuple)):
        boxes = [b.to(x.dtype) for b in boxes]
    else:
        boxes = boxes.to(x.dtype)
    return tvo.roi_align(x, boxes, (PH, PW), spatial_scale, sampling_ratio, aligned)


def nms(boxes: torch.Tensor, scores: torch.Tensor, iou_threshold: float) -> torch.Tensor:
    """Indices of the kept boxes, highest score first (torchvision.ops.nms signature)."""
    if _BACKEND == "dlcfn" and boxes.is_cuda:
        L = _declare()
        n = boxes.shape[0]
        if n == 0:
            return torch.zeros(0, dtype=torch.long, device=boxes.device)
        order = torch.argsort(scores, descending=True, stable=True)
        sb = boxes[order].float().contiguous()
        col_blocks = (n + 63) // 64
        mask = torch.empty(n * col_blocks, dtype=torch.int64, device=boxes.device)
        keep = torch.empty(n, dtype=torch.uint8, device=boxes.device)
        cnt = torch.empty(1, dtype=torch.int32, device=boxes.device)
        native.check(L.dlcfn_nms(sb.data_ptr(), n, float(iou_threshold), mask.data_ptr(), keep.data_ptr(), cnt.data_ptr(),
                                 native.stream_ptr()), "nms")
        native._count(2)
        return order[keep.bool()]
    import torchvision.ops as tvo
    return tvo.nms(boxes.float(), scores.float(), iou_threshold)


def batched_nms(boxes: torch.Tensor, scores: torch.Tensor, idxs: torch.Tensor, iou_threshold: float) -> torch.Tensor:
    """Per-class NMS by the coordinate-offset trick (boxes of different classes never overlap after the shift)."""
    if boxes.numel() == 0:
        return torch.zeros(0, dtype=torch.long, device=boxes.device)
    shift = idxs.to(boxes) * (boxes.max() + 1)
    return nms(boxes + shift[:, None], scores, iou_threshold)


def patch_torchvision() -> bool:
    """Route torchvision's detection models (MultiScaleRoIAlign, RPN / RoIHeads NMS) through this module.  Only acts when the
    backend is `dlcfn`; returns whether it patched.  torchvision looks both functions up as module globals, so replacing
    them there reaches every call site of the Mask R-CNN / Faster R-CNN graphs."""
    if _BACKEND != "dlcfn":
        return False
    import torchvision.ops.boxes as tvb
    import torchvision.ops.poolers as tvp
    tvp.roi_align = roi_align
    tvb.nms = nms
    return True
