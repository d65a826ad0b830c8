#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_nn.py -x -q -m gpu --timeout 600 -k "batchnorm or conv1x1 or resnet" 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_20.log
timeout 400 python bench.py --steps 20 --warmup 5 --conv1x1 auto 2>&1 | tail -1 | tee gpurun_out/bench_auto.json
timeout 300 python bench.py --steps 20 --warmup 5 --conv1x1 cudnn --no-e2e 2>&1 | tail -1 | tee gpurun_out/bench_cudnn.json
timeout 300 python tools/profile_step.py 256 > gpurun_out/profile_step.log 2>&1; head -45 gpurun_out/profile_resnet50.txt | cut -c1-150
