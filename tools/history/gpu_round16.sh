#!/bin/bash
# new kernels' numerics, the 1x1-conv backend probe, and the bench on both backends (1 GPU)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_nn.py tests/test_gpu_kernels.py -x -q -m gpu --timeout 300 2>&1 | tail -15 | tee gpurun_out/pytest_nn.log
timeout 300 python tools/conv1x1_probe.py > gpurun_out/conv1x1_probe.log 2>&1; tail -3 gpurun_out/conv1x1_probe.log
timeout 300 python bench.py --steps 20 --warmup 5 --conv1x1 cudnn --no-e2e 2>&1 | tail -1 | tee gpurun_out/bench_cudnn1x1.json
timeout 300 python bench.py --steps 20 --warmup 5 --conv1x1 tcgen05 --no-e2e 2>&1 | tail -1 | tee gpurun_out/bench_tcgen1x1.json
timeout 300 python bench.py --model bert-large --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_bert_fusedloss.json
