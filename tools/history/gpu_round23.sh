#!/bin/bash
# round-end rehearsal (1 GPU): whole GPU suite as the driver runs it, smoke, default bench, BERT bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/ -x -q -m gpu --timeout 600 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/smoke.log
timeout 400 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_final_1gpu.json
timeout 300 python bench.py --model bert-large --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_bert_final.json
