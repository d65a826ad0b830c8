#!/bin/bash
mkdir -p gpurun_out
echo "=== tests"; timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_gemm.py tests/test_gpu_trainer.py -q -m gpu --timeout 300 -x 2>&1 | tail -12 | tee gpurun_out/t_all.log
echo "=== bench"; timeout 900 python bench.py --steps 20 --warmup 5 2>&1 | tail -3 | tee gpurun_out/bench1.log
echo "=== profile"; timeout 600 python tools/profile_step.py 256 2>&1 | grep -v "aten::" | head -40
