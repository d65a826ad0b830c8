#!/bin/bash
# Round 2 call C: validate the strided dgrad + unified conv path, per-op conv probe vs cuDNN, benches, step profile.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== conv tests"; timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_nn.py tests/test_gpu_gemm.py -x -q -m gpu --timeout 300 2>&1 | tail -8
echo "== conv probe"; PROBE_TAG=c timeout 600 python tools/conv_probe.py > gpurun_out/conv_probe_c.log 2>&1; tail -1 gpurun_out/conv_probe_c.log
echo "== bench tcgen05"; timeout 400 python bench.py --no-e2e 2> gpurun_out/bench_c_tc.err | tail -1 | tee gpurun_out/bench_c_tc.json | cut -c1-400
echo "== bench cudnn"; DLCFN_CONV1X1=cudnn DLCFN_CONVKXK=cudnn timeout 400 python bench.py --no-e2e 2> gpurun_out/bench_c_cudnn.err | tail -1 | tee gpurun_out/bench_c_cudnn.json | cut -c1-400
echo "== profile"; timeout 300 python tools/profile_step.py 256 > gpurun_out/profile_step_c.log 2>&1; head -45 gpurun_out/profile_resnet50.txt | cut -c1-160
