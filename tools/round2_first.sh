#!/bin/bash
# First GPU call of the next round (1 GPU, ~3 min): validate what round 1 wrote blind, then re-measure.
#   gpurun --timeout 900 -- 'bash tools/round2_first.sh'
mkdir -p gpurun_out
echo "== experimental kernels (fused dropout+add+LayerNorm)"
DLCFN_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_experimental.py -q -m gpu --timeout 200 -k "layernorm" 2>&1 | tail -15 | tee gpurun_out/r2_experimental_ln.log
echo "== experimental kernels (implicit-GEMM convolution through im2col TMA): fprop, dgrad, wgrad, module - one process each,"
echo "   a wrong assumption about the im2col load would hang on the mbarrier, hence the timeouts"
for k in fprop dgrad wgrad convkxk_module; do
  DLCFN_EXPERIMENTAL=1 timeout 150 python -m pytest tests/test_gpu_experimental.py -q -m gpu --timeout 100 -k "$k" 2>&1 | tail -12 | tee gpurun_out/r2_experimental_conv_$k.log
done
echo "== whole GPU suite"
timeout 900 python -m pytest tests/ -x -q -m gpu --timeout 600 2>&1 | tail -4 | tee gpurun_out/r2_pytest_gpu.log
echo "== GEMM: tile-width score (BERT shapes now take 128x256 tiles)"
timeout 200 python tools/gemm_quick_perf.py 2>&1 | tail -10 | tee gpurun_out/r2_gemm_perf.log
echo "== GEMM on the CTA-pair kernel (tcgen05.mma.cta_group::2; experimental, under timeouts)"
DLCFN_GEMM_2CTA=1 DLCFN_EXPERIMENTAL=1 timeout 200 python -m pytest tests/test_gpu_experimental.py -q -m gpu --timeout 150 -k "pair_variants" 2>&1 | tail -8 | tee gpurun_out/r2_gemm_2cta_tests.log
DLCFN_GEMM_2CTA=1 timeout 120 python tools/gemm_quick_perf.py 2>&1 | tail -10 | tee gpurun_out/r2_gemm_perf_2cta.log
DLCFN_GEMM_MULTICAST=1 DLCFN_EXPERIMENTAL=1 timeout 200 python -m pytest tests/test_gpu_experimental.py -q -m gpu --timeout 150 -k "pair_variants" 2>&1 | tail -8 | tee gpurun_out/r2_gemm_multicast_tests.log
echo "== GEMM with the multicast-pair kernel (experimental; under a timeout: a protocol bug would hang)"
DLCFN_GEMM_MULTICAST=1 timeout 120 python tools/gemm_quick_perf.py 2>&1 | tail -10 | tee gpurun_out/r2_gemm_perf_multicast.log
echo "== BERT-large, eager LayerNorm vs fused"
DLCFN_FUSED_LN=0 timeout 300 python bench.py --model bert-large --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/r2_bert_eager_ln.json
DLCFN_FUSED_LN=1 timeout 300 python bench.py --model bert-large --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/r2_bert_fused_ln.json
echo "== ResNet-50 default bench"
timeout 400 python bench.py 2>&1 | tail -1 | tee gpurun_out/r2_bench_1gpu.json
echo "== ResNet-50 with the 3x3 convolutions on the implicit-GEMM kernel (only meaningful if the conv tests above passed)"
DLCFN_CONVKXK=tcgen05 timeout 400 python bench.py 2>&1 | tail -1 | tee gpurun_out/r2_bench_1gpu_convkxk.json
DLCFN_CONVKXK=tcgen05 DLCFN_CONV1X1=tcgen05 timeout 400 python bench.py 2>&1 | tail -1 | tee gpurun_out/r2_bench_1gpu_allconv_tcgen05.json
echo "== (multi-GPU calls only) hierarchical multi-node mode with virtual nodes: gpurun --gpus 2|4 -- 'DLCFN_EXPERIMENTAL=1 python -m pytest tests/test_gpu_experimental.py -m gpu -k hier_mode -q'"
echo "== ROIAlign / NMS kernels (Mask R-CNN head, experimental)"
DLCFN_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_experimental.py -q -m gpu --timeout 200 -k "roi_align or nms_kernel" 2>&1 | tail -6 | tee gpurun_out/r2_roi_tests.log
echo "== the step as one CUDA graph (experimental): numerics test, then the bench at batch 256 and at a launch-bound batch 32"
DLCFN_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_experimental.py -q -m gpu --timeout 200 -k "graphed_step" 2>&1 | tail -6 | tee gpurun_out/r2_cuda_graph_tests.log
timeout 400 python bench.py --cuda-graph --no-e2e 2>&1 | tail -2 | tee gpurun_out/r2_bench_1gpu_graph.json
timeout 300 python bench.py --batch 32 --no-e2e 2>&1 | tail -1 | tee gpurun_out/r2_bench_b32_eager.json
timeout 300 python bench.py --batch 32 --no-e2e --cuda-graph 2>&1 | tail -2 | tee gpurun_out/r2_bench_b32_graph.json
echo "== kernel-time breakdown of the step (torch.profiler)"
timeout 300 python tools/profile_step.py 256 > gpurun_out/r2_profile_step.log 2>&1; head -40 gpurun_out/profile_resnet50.txt | cut -c1-150
