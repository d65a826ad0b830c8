"""One-process throughput check of the tcgen05 GEMM on a few shapes (CUDA events, 3 warm-up + 20 timed)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deeplearning_cfn_b200.ops import native  # noqa: E402


def t(fn):
    for _ in range(3):
        fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(20):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / 20


rows = []
for (M, N, K, a_mn, b_mn, bn) in [(8192, 8192, 8192, 0, 0, 256), (8192, 8192, 8192, 0, 0, 128), (4096, 4096, 1024, 0, 0, 256),
                                  (16384, 1024, 1024, 0, 0, 128), (8192, 4096, 1024, 0, 1, 128), (1024, 1024, 8192, 1, 1, 128),
                                  (1024, 4096, 8192, 1, 1, 128), (64, 64, 802816, 1, 1, 128), (256, 64, 802816, 1, 1, 128)]:
    a = torch.randn(K if a_mn else M, M if a_mn else K, device="cuda").bfloat16()
    b = torch.randn(K if b_mn else N, N if b_mn else K, device="cuda").bfloat16()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ms = t(lambda: native.gemm_bf16(a, b, a_mn=bool(a_mn), b_mn=bool(b_mn), block_n=bn, out=out))
    A = a.t() if a_mn else a
    B = b if b_mn else b.t()
    ms2 = t(lambda: torch.matmul(A, B, out=out))
    rows.append({"M": M, "N": N, "K": K, "a_mn": a_mn, "b_mn": b_mn, "block_n": bn, "ms": round(ms, 4),
                 "tflops": round(2 * M * N * K / ms / 1e9, 1), "cublas_ms": round(ms2, 4),
                 "cublas_tflops": round(2 * M * N * K / ms2 / 1e9, 1)})
    print(json.dumps(rows[-1]), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/gemm_quick_perf.json", "w"), indent=1)
