#!/bin/bash
# first GPU round: kernel numerics, GEMM probe, trainer, short bench
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt 2>&1
echo "=== kernels"; timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 120 2>&1 | tail -15 | tee gpurun_out/t_kernels.log
echo "=== gemm probe"; timeout 900 python tools/gemm_probe.py --perf 2>&1 | tee gpurun_out/gemm_probe.log
echo "=== trainer"; timeout 600 python -m pytest tests/test_gpu_trainer.py -q -m gpu --timeout 300 2>&1 | tail -15 | tee gpurun_out/t_trainer.log
echo "=== bench"; timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -5 | tee gpurun_out/bench1.log
