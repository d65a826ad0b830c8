#!/bin/bash
# Round 2 call E: validate BLOCK_N=64 / weights-stationary / fused statistics / fused stem, then probe + bench + profile.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== conv tests"; timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_nn.py tests/test_gpu_gemm.py tests/test_gpu_kernels.py -x -q -m gpu --timeout 300 2>&1 | tail -12
echo "== graph test"; DLCFN_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_experimental.py -x -q -m gpu --timeout 200 -k graphed 2>&1 | tail -3
echo "== conv probe"; PROBE_TAG=e PROBE_CUDNN=0 timeout 600 python tools/conv_probe.py > gpurun_out/conv_probe_e.log 2>&1; tail -1 gpurun_out/conv_probe_e.log
echo "== bench"; timeout 600 python bench.py --steps 50 2> gpurun_out/bench_e.err | tail -1 | tee gpurun_out/bench_e.json | cut -c1-1500
tail -5 gpurun_out/bench_e.err
echo "== profile"; timeout 300 python tools/profile_step.py 256 > gpurun_out/profile_step_e.log 2>&1; head -50 gpurun_out/profile_resnet50.txt | cut -c1-160
