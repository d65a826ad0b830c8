#!/bin/bash
# 1-GPU: the whole GPU test tier on the final tree, the headline bench exactly as the driver runs it, BERT-large bench + profile.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -x -q -m gpu --timeout 600 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_r2_final.log
echo "== experimental"; DLCFN_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_experimental.py -q -m gpu --timeout 300 2>&1 | tail -3
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench (driver form)"; timeout 900 python bench.py 2> gpurun_out/bench_n.err | tail -1 | tee gpurun_out/bench_n_1gpu.json | cut -c1-2200
tail -2 gpurun_out/bench_n.err
echo "== bert"; timeout 600 python bench.py --model bert-large --steps 20 --warmup 3 2> gpurun_out/bench_n_bert.err | tail -1 | tee gpurun_out/bench_bert_large_1gpu.json | cut -c1-300
echo "== bert profile"; timeout 300 python tools/profile_bert.py > gpurun_out/profile_bert.log 2>&1; head -24 gpurun_out/profile_bert.txt | cut -c1-150
