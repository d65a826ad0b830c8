#!/bin/bash
# Round 2, GPU call B (1 GPU): benches of whatever call A validated, then the step profile.  Pass the names of the validated
# features as arguments: conv graph ln 2cta     e.g.  gpurun --timeout 2400 -- 'bash tools/round2_b.sh conv graph'
mkdir -p gpurun_out
has() { [[ " $ARGS " == *" $1 "* ]]; }
ARGS="$*"
run() { local name=$1 t=$2; shift 2; echo "== $name"; timeout "$t" "$@" 2> "gpurun_out/r2_$name.err" | tail -1 | tee "gpurun_out/r2_$name.json" | cut -c1-600; }
export PYTHONUNBUFFERED=1
run bench_1gpu_b 400 python bench.py
if has conv; then
  echo "== per-layer 3x3 probe"; timeout 300 python tools/conv3x3_probe.py 2>&1 | tail -9 | tee gpurun_out/r2_conv3x3_probe.log
  DLCFN_CONVKXK=tcgen05 run bench_convkxk 400 python bench.py --no-e2e
  DLCFN_CONVKXK=tcgen05 DLCFN_CONV1X1=tcgen05 run bench_allconv_tcgen05 400 python bench.py --no-e2e
fi
if has graph; then
  run bench_graph 400 python bench.py --cuda-graph --no-e2e
  run bench_b32_eager 300 python bench.py --batch 32 --no-e2e
  run bench_b32_graph 300 python bench.py --batch 32 --no-e2e --cuda-graph
fi
run bert_eager_ln 300 python bench.py --model bert-large --steps 10 --warmup 3
if has ln; then DLCFN_FUSED_LN=1 run bert_fused_ln 300 python bench.py --model bert-large --steps 10 --warmup 3; fi
if has 2cta; then DLCFN_GEMM_2CTA=1 run bert_2cta 300 python bench.py --model bert-large --steps 10 --warmup 3; fi
echo "== kernel-time breakdown of the step (torch.profiler)"
timeout 300 python tools/profile_step.py 256 > gpurun_out/r2_profile_step.log 2>&1; head -40 gpurun_out/profile_resnet50.txt | cut -c1-150
