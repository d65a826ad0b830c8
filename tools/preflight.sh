#!/bin/bash
# Everything that can be checked without a GPU, in the order the round-end driver will depend on it:
#   build (nvcc cross-compile + g++), static checks, the CPU test tier, the bodies of the GPU tier on the stand-in GPU, the
#   bench.py contract on the host (1 and 2 ranks), the reference arm's line, the flag-protocol model check.
# bash tools/preflight.sh [--quick]      (--quick skips the full CPU tier)
set -u
cd "$(dirname "$0")/.."
fail=0
step() { local name=$1; shift; local t0=$SECONDS; if "$@" > /tmp/preflight_$$.log 2>&1; then echo "ok    $name ($((SECONDS - t0)) s)"; else echo "FAIL  $name"; tail -n 15 /tmp/preflight_$$.log; fail=1; fi; }
step "build (sm_100a + runtime)" python -c "import __graft_entry__ as g; g.build()"
step "ctypes tables vs C signatures" python tools/check_abi.py
step "undefined names in GPU-only code" python tools/lint_names.py
step "flag-protocol model check (W = 2, 3)" python tools/protocol_model.py --world 2 3
if [ "${1:-}" != "--quick" ]; then
  step "CPU test tier" python -m pytest tests -x -q -m "not gpu" -p no:cacheprovider
fi
step "GPU test bodies on the stand-in GPU" env DLCFN_MOCK_GPU_DRYRUN=1 python -m pytest tests/test_gpu_conv.py tests/test_gpu_nn.py tests/test_gpu_layernorm.py tests/test_gpu_y1_graph.py tests/test_gpu_y2_roi.py tests/test_gpu_z1_zoo.py tests/test_gpu_z2_colsum.py -q -m gpu -p no:cacheprovider
step "bench.py --dry-cpu" python bench.py --dry-cpu --steps 3 --warmup 3
step "bench.py --dry-cpu, 2 ranks" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29719 bench.py --gpus 2 --dry-cpu --steps 3 --warmup 3
step "bench.py --impl reference" python bench.py --impl reference --gpus 1 --steps 2 --warmup 1
rm -f /tmp/preflight_$$.log
[ $fail = 0 ] && echo "preflight: all green" || echo "preflight: FAILURES"
exit $fail
