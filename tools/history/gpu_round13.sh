#!/bin/bash
# N-GPU validation of the NVLS paths + tuned grids (keep it short: charged N x)
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1"
F='^W0|^\*\*\*|OMP_NUM|^$|destroy_process_group'
echo "=== tune"; TUNE_MIB=${TUNE_MIB:-16,256} TUNE_BLOCKS=${TUNE_BLOCKS:-32,128} timeout 240 $TR --master-port 29640 tools/allreduce_tune.py 2>&1 | grep -Ev "$F" | tail -40
echo "=== collectives"; timeout 240 $TR --master-port 29611 tests/mp/collectives_worker.py 2>&1 | grep -Ev "$F" | tail -5
echo "=== kv device"; timeout 200 $TR --master-port 29614 tests/mp/kv_device_worker.py 2>&1 | grep -Ev "$F" | tail -4
echo "=== sharded linear"; timeout 200 $TR --master-port 29615 tests/mp/sharded_linear_worker.py 2>&1 | grep -Ev "$F" | tail -7
echo "=== sweep allreduce (auto algo/grid)"; timeout 240 $TR --master-port 29630 -m deeplearning_cfn_b200.examples.bandwidth_measure --sweep --kv-store allreduce --num-batches 5 --out gpurun_out/sweep2_allreduce_$N.json 2>&1 | grep -Ev "$F" | tail -12
echo "=== sweep ps (auto algo/grid)"; timeout 240 $TR --master-port 29631 -m deeplearning_cfn_b200.examples.bandwidth_measure --sweep --kv-store ps --num-batches 5 --out gpurun_out/sweep2_ps_$N.json 2>&1 | grep -Ev "$F" | tail -12
