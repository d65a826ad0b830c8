#!/bin/bash
# what the driver runs at round end: the whole GPU test suite (2 GPUs so the multi-GPU tests are not skipped)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu --timeout 600 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_2.log
