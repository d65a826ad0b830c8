"""`ncu` targets for round 2 (run ONLY after tools/round2_a.sh showed the corresponding tests passing): the new kernels a
few times each, so `-k regex:<name> -c N` can pick one instance.

    DLCFN_GEMM_2CTA=1 ncu --set full --clock-control none --import-source on -k regex:kgemm_pair -c 1 \
        -o gpurun_out/r2_gemm_pair python tools/ncu_round2.py gemm
    ncu --set full --clock-control none --import-source on -k regex:kgemm -c 3 -o gpurun_out/r2_conv python tools/ncu_round2.py conv
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deeplearning_cfn_b200.ops import native  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "gemm"
dev = torch.device("cuda", 0)
torch.manual_seed(0)
if what == "gemm":
    a = torch.randn(8192, 8192, device=dev).bfloat16()
    b = torch.randn(8192, 8192, device=dev).bfloat16()
    out = torch.empty(8192, 8192, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        native.gemm_bf16(a, b, out=out, block_n=256)
elif what == "conv":
    cl = torch.channels_last
    for (n, c, hw, k) in ((256, 64, 56, 64), (256, 256, 14, 256)):       # ResNet-50 stage-1 / stage-3 3x3 layers
        x = torch.randn(n, c, hw, hw, device=dev).bfloat16().contiguous(memory_format=cl)
        w = (torch.randn(k, c, 3, 3, device=dev) * 0.05).bfloat16().contiguous(memory_format=cl)
        y = native.conv_fprop(x, w, 1, 1)           # kgemm<128|256, 0, 0, 0, 1>
        native.conv_dgrad(y, w, 1, (hw, hw))        # kgemm<..., 0, 1, 0, 1>
        native.conv_wgrad(x, y, (3, 3), 1, 1)       # kgemm<..., 1, 1, 0, 1> + ksplitk_reduce
torch.cuda.synchronize()
print("ncu targets done:", what)
