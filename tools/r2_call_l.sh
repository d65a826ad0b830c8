#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29557"
echo "== staged"; timeout 300 $T tests/mp/sharded_linear_worker.py 2>&1 | grep -v "^W0\|^\*\*\*\|OMP_NUM" | tail -12
echo "== tma pull"; DLCFN_PULL_GEMM=tma timeout 300 $T tests/mp/sharded_linear_worker.py 2>&1 | grep -v "^W0\|^\*\*\*\|OMP_NUM" | tail -4
