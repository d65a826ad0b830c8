#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== conv tests"; timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_gemm.py -x -q -m gpu --timeout 300 2>&1 | tail -3
echo "== conv probe"; PROBE_TAG=h PROBE_CUDNN=0 timeout 600 python tools/conv_probe.py > gpurun_out/conv_probe_h.log 2>&1; tail -1 gpurun_out/conv_probe_h.log
echo "== bench eager"; timeout 600 python bench.py --steps 50 --no-e2e --stock 0 2> gpurun_out/bench_h_eager.err | tail -1 | tee gpurun_out/bench_h_eager.json | cut -c1-330
echo "== bench graph"; timeout 600 python bench.py --steps 50 --no-e2e --stock 0 --cuda-graph 2> gpurun_out/bench_h_graph.err | tail -1 | tee gpurun_out/bench_h_graph.json | cut -c1-330
echo "== bert"; timeout 600 python bench.py --model bert-large --steps 10 --warmup 3 2> gpurun_out/bench_h_bert.err | tail -1 | tee gpurun_out/bench_h_bert.json | cut -c1-300
echo "== gemm"; timeout 300 python tools/gemm_quick_perf.py 2>&1 | tail -9
