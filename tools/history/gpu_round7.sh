#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1"
F='^W0|^\*\*\*|OMP_NUM|^$|destroy_process_group'
echo "=== tune"; timeout 300 $TR --master-port 29640 tools/allreduce_tune.py 2>&1 | grep -Ev "$F" | tail -70
echo "=== collectives (auto backend)"; timeout 300 $TR --master-port 29611 tests/mp/collectives_worker.py 2>&1 | grep -Ev "$F" | tail -6
echo "=== trainer modes"; timeout 300 $TR --master-port 29613 tests/mp/trainer_worker.py 2>&1 | grep -Ev "$F|UserWarning|Consider|losses.append" | tail -6
