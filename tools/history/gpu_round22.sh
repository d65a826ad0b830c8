#!/bin/bash
# 2 GPUs: trainer hook changes at world 2 (all modes agree with the single-process reference), then N=2 and N=1 bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_trainer.py tests/test_gpu_kernels.py -x -q -m gpu --timeout 600 -k "trainer or batchnorm" 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_22.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench_final_2gpu.json
timeout 400 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench_final_1gpu.json
