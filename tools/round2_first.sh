#!/bin/bash
# Kept for the references in the docs: the first GPU call of round 2 is tools/round2_a.sh (baseline + validation of the
# kernels written blind), the second tools/round2_b.sh <validated features> (benches + profile).
# Multi-GPU follow-ups:  gpurun --gpus 2 -- 'python -m pytest tests -m gpu -q -k "world2 or multigpu"'
#                        gpurun --gpus 4 -- 'DLCFN_EXPERIMENTAL=1 python -m pytest tests/test_gpu_experimental.py -m gpu -q -k hier_mode'
exec bash "$(dirname "$0")/round2_a.sh"
