#!/bin/bash
# 2-GPU: multi-GPU test tier on the final tree (LL all-reduce, SyncBN over the heap, staged pull GEMM, protocol stress),
# compute-sanitizer racecheck / synccheck at world 2, small-message latency sweep.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== multigpu tests"; DLCFN_EXPECT_SYMM_BACKEND=vmm timeout 1200 python -m pytest tests/test_gpu_z3_multigpu.py tests/test_gpu_trainer.py -x -q -m gpu --timeout 600 2>&1 | tail -5 | tee gpurun_out/pytest_multigpu_r2_final.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29558"
echo "== sweep allreduce (LL / one-shot / two-shot)"; timeout 300 $T -m deeplearning_cfn_b200.examples.bandwidth_measure --kv-store allreduce --sweep --max-bytes 4194304 --out gpurun_out/sweep2_allreduce_r2.json 2>/dev/null | tail -7
echo "== sweep allreduce one-shot barrier form"; DLCFN_ALLREDUCE_ONESHOT=barrier timeout 300 $T -m deeplearning_cfn_b200.examples.bandwidth_measure --kv-store allreduce --sweep --max-bytes 16384 2>/dev/null | tail -3
echo "== sweep nccl"; timeout 300 $T -m deeplearning_cfn_b200.examples.bandwidth_measure --kv-store nccl --sweep --max-bytes 4194304 --out gpurun_out/sweep2_nccl_r2.json 2>/dev/null | tail -7
echo "== sanitizers"; timeout 900 bash tools/sanitize_collectives.sh 2>&1 | grep -v "^W0\|^\*\*\*\|OMP_NUM" | tail -14
