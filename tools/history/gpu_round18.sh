#!/bin/bash
# round-end rehearsal on 1 GPU: whole GPU suite, smoke, bench (default backend, with e2e), conv probe, ncu captures
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/ -x -q -m gpu --timeout 600 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_1.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.log
timeout 400 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench_final_1gpu.json
timeout 300 python tools/conv1x1_probe.py > gpurun_out/conv1x1_probe.log 2>&1; tail -1 gpurun_out/conv1x1_probe.log
timeout 300 python bench.py --steps 20 --warmup 5 --conv1x1 tcgen05 --no-e2e 2>&1 | tail -1 | tee gpurun_out/bench_tcgen1x1.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'kgemm|kbn_stats|kbn_apply|kbn_bwd_reduce|kbn_bwd_apply|ksoftmax_ce|ksplitk' -c 14 -f -o gpurun_out/r1_kernels python tools/ncu_targets.py 2>&1 | tail -3
ls -la gpurun_out/*.ncu-rep 2>/dev/null
