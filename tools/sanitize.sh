#!/bin/bash
# Race / sync checking of the hand-written kernels (reference analogue: the ASAN CI job + NaiveEngine debugging,
# SURVEY 5.2).  compute-sanitizer sees one GPU; cross-GPU flag protocols are exercised by the stress loops in
# tests/mp/collectives_worker.py instead.  Run on a GPU box:  bash tools/sanitize.sh [memcheck|racecheck|synccheck]
TOOL=${1:-racecheck}
mkdir -p gpurun_out
compute-sanitizer --tool "$TOOL" --error-exitcode 1 python - <<'PY' 2>&1 | tail -25 | tee gpurun_out/sanitize_$TOOL.log
import sys, torch
sys.path.insert(0, '.')
from deeplearning_cfn_b200.ops import native
from deeplearning_cfn_b200.ops.batchnorm import batch_norm_relu
from deeplearning_cfn_b200.ops.pool import max_pool2d
d = torch.device('cuda', 0)
a = torch.randn(256, 512, device=d).bfloat16(); b = torch.randn(384, 512, device=d).bfloat16()
native.gemm_bf16(a, b)                                           # tcgen05 / TMA / mbarrier pipeline
x = torch.randn(8, 64, 14, 14, device=d).bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True)
g, bt = torch.ones(64, device=d, requires_grad=True), torch.zeros(64, device=d, requires_grad=True)
y = batch_norm_relu(x, g, bt, torch.zeros(64, device=d), torch.ones(64, device=d), True, 0.1, 1e-5, True)
y = max_pool2d(y, 3, 2, 1); y.sum().backward()                    # smem block reductions, gather backward
m = torch.randn(10007, device=d); s1 = torch.zeros_like(m)
native.fused_optimizer('sgd_mom', torch.randn(10007, device=d).bfloat16(), m, s1, None, torch.empty(10007, device=d, dtype=torch.bfloat16), lr=0.1, momentum=0.9)
from deeplearning_cfn_b200.parallel.symm import SymmetricHeap
h = SymmetricHeap(8 << 20); off, t = h.alloc_tensor(100000, torch.float32); t.fill_(1.0)
h.allreduce_(off, 100000, torch.float32, 1.0); h.barrier(); torch.cuda.synchronize(); h.check_error()
print('sanitize run finished')
PY
