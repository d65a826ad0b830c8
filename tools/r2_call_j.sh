#!/bin/bash
# 2-GPU call: multi-GPU test tier (collectives on the own VMM backend, trainer modes, kvstore device, bandwidth tool), benches.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi -L | head -3
echo "== multigpu tests (vmm backend expected)"; DLCFN_EXPECT_SYMM_BACKEND=vmm timeout 1200 python -m pytest tests/test_gpu_z3_multigpu.py tests/test_gpu_trainer.py -x -q -m gpu --timeout 600 2>&1 | tail -15
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555"
echo "== bench 2gpu eager"; timeout 600 $T bench.py --gpus 2 --steps 50 --no-e2e --stock 0 2> gpurun_out/bench_j_2_eager.err | tail -1 | tee gpurun_out/bench_j_2_eager.json | cut -c1-330
echo "== bench 2gpu graph"; timeout 600 $T bench.py --gpus 2 --steps 50 --no-e2e --stock 0 --cuda-graph 2> gpurun_out/bench_j_2_graph.err | tail -1 | tee gpurun_out/bench_j_2_graph.json | cut -c1-330
tail -3 gpurun_out/bench_j_2_graph.err
echo "== bench 2gpu nccl"; timeout 600 $T bench.py --gpus 2 --steps 50 --no-e2e --stock 0 --mode nccl 2> gpurun_out/bench_j_2_nccl.err | tail -1 | tee gpurun_out/bench_j_2_nccl.json | cut -c1-330
echo "== bench 2gpu full default"; timeout 900 $T bench.py --gpus 2 --steps 50 2> gpurun_out/bench_j_2_full.err | tail -1 | tee gpurun_out/bench_j_2_full.json | cut -c1-1600
tail -3 gpurun_out/bench_j_2_full.err
