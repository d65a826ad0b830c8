#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1"
F='^W0|^\*\*\*|OMP_NUM|^$|destroy_process_group'
echo "=== tune"; TUNE_MIB=16,256 TUNE_BLOCKS=32,128 timeout 200 $TR --master-port 29640 tools/allreduce_tune.py 2>&1 | grep -Ev "$F" | tail -20
echo "=== collectives"; timeout 200 $TR --master-port 29611 tests/mp/collectives_worker.py 2>&1 | grep -Ev "$F" | tail -4
echo "=== sweep allreduce nvls"; DLCFN_ALLREDUCE_ALGO=nvls timeout 200 $TR --master-port 29630 -m deeplearning_cfn_b200.examples.bandwidth_measure --sweep --kv-store allreduce --num-batches 5 --out gpurun_out/sweep2_allreduce_nvls_$N.json 2>&1 | grep -Ev "$F" | tail -11
echo "=== sweep allreduce p2p"; DLCFN_ALLREDUCE_ALGO=p2p timeout 200 $TR --master-port 29632 -m deeplearning_cfn_b200.examples.bandwidth_measure --sweep --kv-store allreduce --num-batches 5 --out gpurun_out/sweep2_allreduce_p2p_$N.json 2>&1 | grep -Ev "$F" | tail -11
echo "=== sweep ps nvls"; DLCFN_ALLREDUCE_ALGO=nvls timeout 200 $TR --master-port 29631 -m deeplearning_cfn_b200.examples.bandwidth_measure --sweep --kv-store ps --num-batches 5 --out gpurun_out/sweep2_ps_nvls_$N.json 2>&1 | grep -Ev "$F" | tail -11
