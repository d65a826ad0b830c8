#!/bin/bash
mkdir -p gpurun_out
echo "=== trainer"; timeout 600 python -m pytest tests/test_gpu_trainer.py tests/test_gpu_gemm.py -q -m gpu --timeout 300 2>&1 | tail -8 | tee gpurun_out/t_trainer.log
echo "=== profile"; timeout 600 python tools/profile_step.py 256 2>&1 | tail -70
echo "=== ncu gemm"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:kgemm -s 2 -c 1 -o gpurun_out/gemm_8192 python -c "
import torch,sys
sys.path.insert(0,'.')
from deeplearning_cfn_b200.ops import native
a=torch.randn(8192,8192,device='cuda').bfloat16(); b=torch.randn(8192,8192,device='cuda').bfloat16()
for _ in range(4): native.gemm_bf16(a,b,block_n=256)
torch.cuda.synchronize()
" 2>&1 | tail -5
