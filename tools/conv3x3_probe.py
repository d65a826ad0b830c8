"""Per-layer timing of ResNet-50's 3x3 convolutions (batch 256): cuDNN vs the dlcfn implicit-GEMM kernel (fprop, stride-1
dgrad, split-K wgrad separately).  Run only after the conv tests of tests/test_gpu_experimental.py pass.
Writes gpurun_out/conv3x3_probe.json.  Device-timed with CUDA events, 3 warm-up + 10 timed iterations (median), a 256 MiB
buffer is rewritten between iterations to flush L2."""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deeplearning_cfn_b200.ops import native  # noqa: E402

B = int(os.environ.get("PROBE_BATCH", "256"))
# (C, K, H of the input, stride, count in ResNet-50 v2: conv2 of every bottleneck)
LAYERS = [(64, 64, 56, 1, 3), (128, 128, 56, 2, 1), (128, 128, 28, 1, 3), (256, 256, 28, 2, 1), (256, 256, 14, 1, 5),
          (512, 512, 14, 2, 1), (512, 512, 7, 1, 2)]


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
    rows, tot = [], {"cudnn": 0.0, "tcgen05": 0.0}
    for c, k, hw, s, cnt in LAYERS:
        x = torch.randn(B, c, hw, hw, device=dev).bfloat16().contiguous(memory_format=cl)
        w = (torch.randn(k, c, 3, 3, device=dev) * 0.05).bfloat16().contiguous(memory_format=cl)
        ho = (hw + 2 - 3) // s + 1
        gy = torch.randn(B, k, ho, ho, device=dev).bfloat16().contiguous(memory_format=cl)
        res = {"c": c, "k": k, "hw": hw, "stride": s, "count": cnt}
        bwd = lambda mask: torch.ops.aten.convolution_backward(gy, x, w, None, [s, s], [1, 1], [1, 1], False, [0, 0], 1, mask)   # noqa: E731
        t = {"cudnn": {"fprop": timeit(lambda: F.conv2d(x, w, None, s, 1), flush),
                       "dgrad": timeit(lambda: bwd([True, False, False]), flush),
                       "wgrad": timeit(lambda: bwd([False, True, False]), flush)},
             "tcgen05": {"fprop": timeit(lambda: native.conv_fprop(x, w, s, 1), flush),
                         "dgrad": timeit(lambda: native.conv_dgrad(gy, w, 1, (hw, hw)), flush) if s == 1 else None,
                         "wgrad": timeit(lambda: native.conv_wgrad(x, gy, (3, 3), s, 1), flush)}}
        if t["tcgen05"]["dgrad"] is None:
            t["tcgen05"]["dgrad"] = t["cudnn"]["dgrad"]          # strided dgrad stays on cuDNN
        flops = 2.0 * B * ho * ho * c * k * 9
        for name in ("cudnn", "tcgen05"):
            res[name] = {op: round(v, 4) for op, v in t[name].items()}
            res[name]["fprop_tflops"] = round(flops / (t[name]["fprop"] * 1e-3) / 1e12, 1)
            tot[name] += sum(t[name].values()) * cnt
        rows.append(res)
        print(json.dumps(res), flush=True)
    out = {"batch": B, "layers": rows, "total_ms_weighted": {n: round(v, 3) for n, v in tot.items()}}
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/conv3x3_probe.json", "w"), indent=1)
    print(json.dumps(out["total_ms_weighted"]))


if __name__ == "__main__":
    main()
