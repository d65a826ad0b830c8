#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_nn.py tests/test_gpu_trainer.py -x -q -m gpu --timeout 600 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_21.log
timeout 400 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench_final_1gpu.json
timeout 300 python bench.py --steps 20 --warmup 5 --conv1x1 auto --no-e2e 2>&1 | tail -1 | tee gpurun_out/bench_auto.json
timeout 300 python tools/copy_diag2.py 64 2>&1 | tail -8
