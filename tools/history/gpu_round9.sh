#!/bin/bash
mkdir -p gpurun_out
echo "=== tests"; timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 300 -x -k "stem or maxpool or batchnorm" 2>&1 | tail -4
echo "=== bench"; timeout 600 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench1_resnet_v4.log | cut -c1-330
echo "=== profile"; timeout 600 python tools/profile_step.py 256 2>&1 | grep -v "aten::\|_BNReLU\|_MaxPool\|_LinearFn" | head -14
cp gpurun_out/profile_resnet50.txt gpurun_out/profile_resnet50_v4.txt
echo "=== sanitize"; timeout 600 bash tools/sanitize.sh memcheck 2>&1 | tail -6
