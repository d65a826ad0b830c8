#!/bin/bash
# compute-sanitizer racecheck + synccheck over the cross-GPU kernels at world 2 (each rank under its own sanitizer instance).
#   gpurun --gpus 2 -- 'bash tools/sanitize_collectives.sh'
mkdir -p gpurun_out
for TOOL in synccheck racecheck; do
  echo "== $TOOL"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 --no-python \
    bash -c "compute-sanitizer --tool $TOOL --error-exitcode 7 python tests/mp/sanitize_worker.py > gpurun_out/sanitize_world2_${TOOL}_rank\$RANK.log 2>&1; echo rank \$RANK exit \$?"
  for r in 0 1; do grep -h "ERROR SUMMARY\|SANITIZE_RUN_OK\|hazard\|Barrier error" gpurun_out/sanitize_world2_${TOOL}_rank$r.log | head -5; done
done
