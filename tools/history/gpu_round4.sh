#!/bin/bash
# 2-GPU round: fused collectives numerics, trainer-mode agreement, N=2 bench in each mode
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
echo "=== bn test"; timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k batchnorm --timeout 120 2>&1 | tail -3
echo "=== collectives"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29611 tests/mp/collectives_worker.py 2>&1 | grep -v "^W0\|^\*\*\*\|OMP_NUM" | tail -25 | tee gpurun_out/collectives2.log
echo "=== trainer modes"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29613 tests/mp/trainer_worker.py 2>&1 | grep -v "^W0\|^\*\*\*\|OMP_NUM" | tail -25 | tee gpurun_out/trainer2.log
for mode in allreduce ps_sync nccl nccl_ps; do
echo "=== bench N=2 $mode"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29620 bench.py --gpus 2 --steps 15 --warmup 4 --mode $mode --no-e2e 2>&1 | grep -E "^\{|Error|error" | tail -3 | tee gpurun_out/bench2_$mode.log
done
