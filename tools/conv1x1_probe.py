"""Per-layer timing of ResNet-50's 1x1 convolutions (batch 256): cuDNN vs the dlcfn tcgen05 GEMM path.
Writes gpurun_out/conv1x1_probe.json.  Device-timed with CUDA events, 3 warm-up + 10 timed iterations, a 256 MiB
buffer is rewritten between iterations to flush L2."""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deeplearning_cfn_b200.ops.conv import conv1x1  # noqa: E402

B = int(os.environ.get("PROBE_BATCH", "256"))
# (Cin, Cout, H, stride, count in ResNet-50)
LAYERS = [(64, 64, 56, 1, 1), (64, 256, 56, 1, 4), (256, 64, 56, 1, 2), (256, 128, 56, 1, 1), (128, 512, 28, 1, 4),
          (256, 512, 56, 2, 1), (512, 128, 28, 1, 3), (512, 256, 28, 1, 1), (256, 1024, 14, 1, 6), (512, 1024, 28, 2, 1),
          (1024, 256, 14, 1, 5), (1024, 512, 14, 1, 1), (512, 2048, 7, 1, 3), (1024, 2048, 14, 2, 1), (2048, 512, 7, 1, 2)]


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
    flush = torch.zeros(64 * 1024 * 1024, device=dev)
    rows = []
    tot = {"cudnn": 0.0, "tcgen05": 0.0}
    torch.backends.cudnn.benchmark = True
    for cin, cout, hw, s, cnt in LAYERS:
        x = torch.randn(B, cin, hw, hw, device=dev).bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True)
        w = (torch.randn(cout, cin, 1, 1, device=dev) * 0.05).bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True)
        ho = (hw - 1) // s + 1
        gy = torch.randn(B, cout, ho, ho, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
        res = {"cin": cin, "cout": cout, "hw": hw, "stride": s, "count": cnt}
        for name, f in (("cudnn", lambda: F.conv2d(x, w, None, s)), ("tcgen05", lambda: conv1x1(x, w, s))):
            def fwd():
                with torch.no_grad():
                    f()

            def fwdbwd():
                y = f()
                torch.autograd.grad(y, (x, w), gy)

            tf = timeit(fwd, flush)
            tfb = timeit(fwdbwd, flush)
            res[name] = {"fwd_ms": round(tf, 4), "fwd_bwd_ms": round(tfb, 4)}
            tot[name] += tfb * cnt
        flops = 2.0 * B * ho * ho * cin * cout
        res["fwd_tflops"] = {k: round(flops / (res[k]["fwd_ms"] * 1e-3) / 1e12, 1) for k in ("cudnn", "tcgen05")}
        rows.append(res)
        print(json.dumps(res), flush=True)
    out = {"batch": B, "layers": rows, "total_fwd_bwd_ms_weighted": {k: round(v, 3) for k, v in tot.items()}}
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/conv1x1_probe.json", "w"), indent=1)
    print(json.dumps(out["total_fwd_bwd_ms_weighted"]))


if __name__ == "__main__":
    main()
