#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== tests"; DLCFN_EXPECT_SYMM_BACKEND=vmm timeout 600 python -m pytest tests/test_gpu_z3_multigpu.py -x -q -m gpu --timeout 300 -k "collectives_world2 or protocol" 2>&1 | tail -6
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29556"
echo "== sweep allreduce"; timeout 300 $T -m deeplearning_cfn_b200.examples.bandwidth_measure --kv-store allreduce --sweep --max-bytes 16777216 --out gpurun_out/sweep2_allreduce_r2.json 2>/dev/null | tail -9
echo "== sweep nccl"; timeout 300 $T -m deeplearning_cfn_b200.examples.bandwidth_measure --kv-store nccl --sweep --max-bytes 16777216 --out gpurun_out/sweep2_nccl_r2.json 2>/dev/null | tail -9
echo "== stock 2gpu"; timeout 300 $T bench.py --impl torch-stock --gpus 2 --steps 30 --no-e2e 2>/dev/null | tail -1 | tee gpurun_out/stock_2gpu.json | cut -c1-300
