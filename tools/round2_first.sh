#!/bin/bash
# First GPU call of the next round (1 GPU, ~3 min): validate what round 1 wrote blind, then re-measure.
#   gpurun --timeout 900 -- 'bash tools/round2_first.sh'
mkdir -p gpurun_out
echo "== experimental kernels (fused dropout+add+LayerNorm)"
DLCFN_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_experimental.py -q -m gpu --timeout 200 2>&1 | tail -15 | tee gpurun_out/r2_experimental.log
echo "== whole GPU suite"
timeout 900 python -m pytest tests/ -x -q -m gpu --timeout 600 2>&1 | tail -4 | tee gpurun_out/r2_pytest_gpu.log
echo "== GEMM: tile-width score (BERT shapes now take 128x256 tiles)"
timeout 200 python tools/gemm_quick_perf.py 2>&1 | tail -10 | tee gpurun_out/r2_gemm_perf.log
echo "== GEMM with the multicast-pair kernel (experimental; under a timeout: a protocol bug would hang)"
DLCFN_GEMM_MULTICAST=1 timeout 120 python tools/gemm_quick_perf.py 2>&1 | tail -10 | tee gpurun_out/r2_gemm_perf_multicast.log
echo "== BERT-large, eager LayerNorm vs fused"
DLCFN_FUSED_LN=0 timeout 300 python bench.py --model bert-large --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/r2_bert_eager_ln.json
DLCFN_FUSED_LN=1 timeout 300 python bench.py --model bert-large --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/r2_bert_fused_ln.json
echo "== ResNet-50 default bench"
timeout 400 python bench.py 2>&1 | tail -1 | tee gpurun_out/r2_bench_1gpu.json
echo "== kernel-time breakdown of the step (torch.profiler)"
timeout 300 python tools/profile_step.py 256 > gpurun_out/r2_profile_step.log 2>&1; head -40 gpurun_out/profile_resnet50.txt | cut -c1-150
