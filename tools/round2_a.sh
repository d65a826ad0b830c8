#!/bin/bash
# Round 2, GPU call A (1 GPU, ~12 min when nothing hangs): re-establish the round-1 baseline, then validate the kernels that
# round 1 wrote without GPU access.  Every experimental group runs in its own process under its own timeout - a protocol
# bug in a new kernel spins on an mbarrier and must not take the rest of the call with it.
#   gpurun --timeout 2400 -- 'bash tools/round2_a.sh'
mkdir -p gpurun_out
run() { # name, timeout, command...
  local name=$1 t=$2; shift 2
  echo "== $name"
  timeout "$t" "$@" > "gpurun_out/r2_$name.log" 2>&1; local rc=$?
  tail -n 6 "gpurun_out/r2_$name.log" | cut -c1-400
  echo "   [$name exit $rc]"
}
export PYTHONUNBUFFERED=1
run pytest_gpu 900 python -m pytest tests/ -x -q -m gpu --timeout 600
run bench_1gpu 400 python bench.py
E="python -m pytest tests/test_gpu_experimental.py -q -m gpu --timeout 80 -x -k"
DLCFN_EXPERIMENTAL=1 run exp_layernorm 200 $E layernorm
for k in conv_fprop conv_dgrad conv_wgrad convkxk_module; do DLCFN_EXPERIMENTAL=1 run exp_$k 120 $E $k; done
DLCFN_EXPERIMENTAL=1 run exp_roi 150 $E "roi_align or nms_kernel"
DLCFN_EXPERIMENTAL=1 run exp_cuda_graph 200 $E graphed_step
DLCFN_EXPERIMENTAL=1 DLCFN_GEMM_2CTA=1 run exp_gemm_2cta 150 $E pair_variants
DLCFN_EXPERIMENTAL=1 DLCFN_GEMM_MULTICAST=1 run exp_gemm_multicast 150 $E pair_variants
run gemm_perf 150 python tools/gemm_quick_perf.py
DLCFN_GEMM_2CTA=1 run gemm_perf_2cta 100 python tools/gemm_quick_perf.py
DLCFN_GEMM_MULTICAST=1 run gemm_perf_multicast 100 python tools/gemm_quick_perf.py
grep -h "passed\|failed\|error" gpurun_out/r2_exp_*.log gpurun_out/r2_pytest_gpu.log | tail -20
