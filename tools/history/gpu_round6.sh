#!/bin/bash
# N-GPU round (N = $1): collectives numerics, bandwidth sweeps vs NCCL, ResNet-50 bench in every mode
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1"
F='^W0|^\*\*\*|OMP_NUM|^$'
echo "=== collectives world $N"; timeout 400 $TR --master-port 29611 tests/mp/collectives_worker.py 2>&1 | grep -Ev "$F" | tail -6 | tee gpurun_out/collectives$N.log
for kind in allreduce nccl ps nccl_ps; do
echo "=== sweep $kind"; timeout 300 $TR --master-port 29630 -m deeplearning_cfn_b200.examples.bandwidth_measure --sweep --kv-store $kind --num-batches 5 --out gpurun_out/sweep_${kind}_$N.json 2>&1 | grep -Ev "$F" | tail -12
done
for kind in allreduce nccl ps nccl_ps; do
echo "=== resnet-200 params $kind"; timeout 300 $TR --master-port 29631 -m deeplearning_cfn_b200.examples.bandwidth_measure --network resnet --num-layers 200 --kv-store $kind --num-batches 3 --out gpurun_out/net_${kind}_$N.json 2>&1 | grep -Ev "$F" | tail -4
done
for mode in allreduce ps_sync ps_dedicated nccl nccl_ps; do
echo "=== bench N=$N $mode"; timeout 400 $TR --master-port 29620 bench.py --gpus $N --steps 15 --warmup 4 --mode $mode --no-e2e 2>&1 | grep -E "^\{|Error|error|Traceback" | cut -c1-330 | tail -3 | tee gpurun_out/bench${N}_$mode.log
done
echo "=== bench N=$N allreduce e2e"; timeout 400 $TR --master-port 29621 bench.py --gpus $N --steps 15 --warmup 4 2>&1 | grep -E "^\{" | tail -1 | tee gpurun_out/bench${N}_e2e.log | cut -c1-200
