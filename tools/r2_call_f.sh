#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== conv tests"; timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_gemm.py -x -q -m gpu --timeout 300 2>&1 | tail -8
echo "== ncu cudnn"; timeout 600 ncu --metrics gpu__time_duration.sum,l1tex__m_xbar2l1tex_read_bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum,launch__grid_size,launch__cluster_size,launch__block_size,smsp__inst_executed_op_shared_ld.sum --clock-control none -k regex:"cutlass|cudnn|gemm|conv" --csv --log-file gpurun_out/ncu_cudnn.csv python tools/ncu_cudnn.py > gpurun_out/ncu_cudnn.log 2>&1; tail -2 gpurun_out/ncu_cudnn.log
echo "== conv probe"; PROBE_TAG=f PROBE_CUDNN=0 timeout 600 python tools/conv_probe.py > gpurun_out/conv_probe_f.log 2>&1; tail -1 gpurun_out/conv_probe_f.log
echo "== bench"; timeout 600 python bench.py --steps 50 2> gpurun_out/bench_f.err | tail -1 | tee gpurun_out/bench_f.json | cut -c1-1800
tail -3 gpurun_out/bench_f.err
echo "== profile"; timeout 300 python tools/profile_step.py 256 > gpurun_out/profile_step_f.log 2>&1; head -42 gpurun_out/profile_resnet50.txt | cut -c1-150
