#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== stem test"; timeout 300 python -m pytest tests/test_gpu_conv.py -x -q -m gpu --timeout 200 -k "stem or batchnorm_takes" 2>&1 | tail -3
echo "== bench eager"; timeout 600 python bench.py --steps 50 --no-e2e --stock 0 2> gpurun_out/bench_g_eager.err | tail -1 | tee gpurun_out/bench_g_eager.json | cut -c1-330
echo "== bench graph"; timeout 600 python bench.py --steps 50 --no-e2e --stock 0 --cuda-graph 2> gpurun_out/bench_g_graph.err | tail -1 | tee gpurun_out/bench_g_graph.json | cut -c1-330
tail -2 gpurun_out/bench_g_graph.err
echo "== bert"; timeout 600 python bench.py --model bert-large --steps 10 --warmup 3 2> gpurun_out/bench_g_bert.err | tail -1 | tee gpurun_out/bench_g_bert.json | cut -c1-400
echo "== ncu pair"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:kgemm_pair -c 2 -o gpurun_out/pair_ncu python tools/ncu_conv.py > gpurun_out/pair_ncu.log 2>&1; tail -1 gpurun_out/pair_ncu.log
