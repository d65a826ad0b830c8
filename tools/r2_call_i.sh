#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_gemm.py tests/test_gpu_layernorm.py -x -q -m gpu --timeout 300 2>&1 | tail -3
echo "== gemm"; timeout 300 python tools/gemm_quick_perf.py 2>&1 | head -5
echo "== gemm raster0"; DLCFN_GEMM_RASTER=0 timeout 300 python tools/gemm_quick_perf.py 2>&1 | head -2
echo "== conv probe"; PROBE_TAG=i PROBE_CUDNN=0 timeout 600 python tools/conv_probe.py > gpurun_out/conv_probe_i.log 2>&1; tail -1 gpurun_out/conv_probe_i.log
echo "== bench eager"; timeout 600 python bench.py --steps 50 --no-e2e --stock 0 2> gpurun_out/bench_i_eager.err | tail -1 | tee gpurun_out/bench_i_eager.json | cut -c1-330
echo "== bench graph"; timeout 600 python bench.py --steps 50 --no-e2e --stock 0 --cuda-graph 2> gpurun_out/bench_i_graph.err | tail -1 | tee gpurun_out/bench_i_graph.json | cut -c1-330
echo "== bert"; timeout 600 python bench.py --model bert-large --steps 10 --warmup 3 2> gpurun_out/bench_i_bert.err | tail -1 | tee gpurun_out/bench_i_bert.json | cut -c1-300
echo "== bert profile"; timeout 300 python tools/profile_bert.py > gpurun_out/profile_bert.log 2>&1; head -40 gpurun_out/profile_bert.txt | cut -c1-150
