#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_nn.py tests/test_gpu_kernels.py -x -q -m gpu --timeout 600 2>&1 | tail -12 | tee gpurun_out/pytest_gpu_19.log
timeout 400 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench_final_1gpu.json
timeout 300 python bench.py --steps 20 --warmup 5 --conv1x1 tcgen05 2>&1 | tail -1 | tee gpurun_out/bench_tcgen1x1.json
