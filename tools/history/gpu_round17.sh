#!/bin/bash
# GEMM after split-K / 8-warp epilogue: numerics, 8192^3 throughput, 1x1-conv probe, bench on both conv backends
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm.py -x -q -m gpu --timeout 300 2>&1 | tail -8 | tee gpurun_out/pytest_gemm.log
timeout 200 python tools/gemm_quick_perf.py 2>&1 | tail -10 | tee gpurun_out/gemm_perf.log
timeout 300 python tools/conv1x1_probe.py > gpurun_out/conv1x1_probe.log 2>&1; tail -2 gpurun_out/conv1x1_probe.log
timeout 300 python bench.py --steps 20 --warmup 5 --conv1x1 tcgen05 --no-e2e 2>&1 | tail -1 | tee gpurun_out/bench_tcgen1x1.json
timeout 300 python bench.py --model bert-large --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_bert_splitk.json
