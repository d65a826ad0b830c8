"""Per-layer, per-op timing of EVERY convolution shape of ResNet-50 v2 (batch 256): cuDNN vs the dlcfn tcgen05 kernels,
fprop / dgrad / wgrad separately, weighted by how often the shape occurs -> the convolution milliseconds of one training
step for each backend.  Writes gpurun_out/conv_probe[_TAG].json (PROBE_TAG).  Device-timed with CUDA events, 3 warm-up +
10 timed iterations (median), a 256 MiB buffer is rewritten between iterations to flush L2.

  PROBE_ONLY=1x1|3x3|stem   restrict the layer set;   PROBE_CUDNN=0  skip the cuDNN arm (own kernels only)
"""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deeplearning_cfn_b200.ops import native  # noqa: E402
from deeplearning_cfn_b200.ops import conv as conv_ops  # noqa: E402

B = int(os.environ.get("PROBE_BATCH", "256"))
ONLY = os.environ.get("PROBE_ONLY", "")
CUDNN = os.environ.get("PROBE_CUDNN", "1") != "0"
TAG = os.environ.get("PROBE_TAG", "")
# (kernel, Cin, Cout, H of the input, stride, count in ResNet-50 v2, needs dgrad)
LAYERS = [
    (1, 64, 64, 56, 1, 1, True), (1, 64, 256, 56, 1, 4, True), (1, 256, 64, 56, 1, 2, True), (1, 256, 128, 56, 1, 1, True),
    (1, 128, 512, 28, 1, 4, True), (1, 256, 512, 56, 2, 1, True), (1, 512, 128, 28, 1, 3, True), (1, 512, 256, 28, 1, 1, True),
    (1, 256, 1024, 14, 1, 6, True), (1, 512, 1024, 28, 2, 1, True), (1, 1024, 256, 14, 1, 5, True),
    (1, 1024, 512, 14, 1, 1, True), (1, 512, 2048, 7, 1, 3, True), (1, 1024, 2048, 14, 2, 1, True), (1, 2048, 512, 7, 1, 2, True),
    (3, 64, 64, 56, 1, 3, True), (3, 128, 128, 56, 2, 1, True), (3, 128, 128, 28, 1, 3, True), (3, 256, 256, 28, 2, 1, True),
    (3, 256, 256, 14, 1, 5, True), (3, 512, 512, 14, 2, 1, True), (3, 512, 512, 7, 1, 2, True),
]


def timeit(fn, flush):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(10):
        flush.add_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    dev = torch.device("cuda", 0)
    cl = torch.channels_last
    flush = torch.zeros(64 * 1024 * 1024, device=dev)
    torch.backends.cudnn.benchmark = True
    rows, tot = [], {"cudnn": {"fprop": 0.0, "dgrad": 0.0, "wgrad": 0.0}, "tcgen05": {"fprop": 0.0, "dgrad": 0.0, "wgrad": 0.0}}
    for ks, c, k, hw, s, cnt, need_dx in LAYERS:
        if ONLY and ONLY != f"{ks}x{ks}":
            continue
        pad = ks // 2
        x = torch.randn(B, c, hw, hw, device=dev).bfloat16().contiguous(memory_format=cl)
        w = (torch.randn(k, c, ks, ks, device=dev) * 0.05).bfloat16().contiguous(memory_format=cl)
        ho = (hw + 2 * pad - ks) // s + 1
        gy = torch.randn(B, k, ho, ho, device=dev).bfloat16().contiguous(memory_format=cl)
        res = {"kernel": ks, "c": c, "k": k, "hw": hw, "stride": s, "count": cnt}
        flops = 2.0 * B * ho * ho * c * k * ks * ks
        if CUDNN:
            bwd = lambda mask: torch.ops.aten.convolution_backward(gy, x, w, None, [s, s], [pad, pad], [1, 1], False, [0, 0], 1, mask)   # noqa: E731
            res["cudnn"] = {"fprop": timeit(lambda: F.conv2d(x, w, None, s, pad), flush),
                            "dgrad": timeit(lambda: bwd([True, False, False]), flush),
                            "wgrad": timeit(lambda: bwd([False, True, False]), flush)}
        ours = conv_ops.ConvOps(x, w, s, pad)
        res["tcgen05"] = {"fprop": timeit(lambda: ours.fprop(), flush),
                          "dgrad": timeit(lambda: ours.dgrad(gy), flush),
                          "wgrad": timeit(lambda: ours.wgrad(gy), flush)}
        for name in ("cudnn", "tcgen05"):
            if name not in res:
                continue
            res[name] = {op: round(v, 4) for op, v in res[name].items()}
            for op in ("fprop", "dgrad", "wgrad"):
                tot[name][op] += res[name][op] * cnt
            res[name]["fprop_tflops"] = round(flops / (res[name]["fprop"] * 1e-3) / 1e12, 1)
        rows.append(res)
        print(json.dumps(res), flush=True)
    out = {"batch": B, "layers": rows,
           "total_ms_weighted": {n: dict({op: round(v, 3) for op, v in t.items()}, all=round(sum(t.values()), 3)) for n, t in tot.items()}}
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open(f"gpurun_out/conv_probe{('_' + TAG) if TAG else ''}.json", "w"), indent=1)
    print(json.dumps(out["total_ms_weighted"]))


if __name__ == "__main__":
    main()
