#!/bin/bash
# 8-GPU call: correctness at N = 8, the five data-parallel modes with the final kernels, BERT-large fused vs NCCL, sweeps.
mkdir -p gpurun_out/scale8_r2
O=gpurun_out/scale8_r2
export PYTHONUNBUFFERED=1
N=${1:-8}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== collectives at $N"; DLCFN_EXPECT_SYMM_BACKEND=vmm timeout 400 $T --master-port 29601 tests/mp/collectives_worker.py 2>&1 | grep -v "^W0\|^\*\*\*\|OMP_NUM" | tail -4
echo "== sharded linear at $N"; timeout 300 $T --master-port 29602 tests/mp/sharded_linear_worker.py 2>&1 | grep -v "^W0\|^\*\*\*\|OMP_NUM" | tail -4 | tee $O/sharded_linear_$N.txt
: > $O/bench${N}_all_modes_r2.jsonl
for mode in allreduce nccl ps_sync ps_dedicated nccl_ps; do
  echo "== resnet $mode"; timeout 400 $T --master-port 29603 bench.py --gpus $N --steps 100 --no-e2e --stock 0 --mode $mode 2> $O/bench_$mode.err | tail -1 | tee -a $O/bench${N}_all_modes_r2.jsonl | cut -c1-260
done
echo "== resnet allreduce eager (no graph)"; timeout 400 $T --master-port 29603 bench.py --gpus $N --steps 100 --no-e2e --stock 0 --no-cuda-graph 2> $O/bench_eager.err | tail -1 | tee -a $O/bench${N}_all_modes_r2.jsonl | cut -c1-260
for mode in allreduce nccl; do
  echo "== bert-large $mode"; timeout 400 $T --master-port 29604 bench.py --gpus $N --model bert-large --steps 20 --warmup 3 --mode $mode 2> $O/bert_$mode.err | tail -1 | tee $O/bench_bert_large_${N}gpu_$mode.json | cut -c1-260
done
for kv in allreduce nccl; do
  echo "== sweep $kv"; timeout 300 $T --master-port 29605 -m deeplearning_cfn_b200.examples.bandwidth_measure --kv-store $kv --sweep --max-bytes 268435456 --out $O/sweep${N}_$kv.json 2>/dev/null | tail -10
done
for kv in allreduce ps nccl; do
  echo "== per-key $kv"; timeout 300 $T --master-port 29606 -m deeplearning_cfn_b200.examples.bandwidth_measure --kv-store $kv --num-batches 5 --out $O/net_${kv}_$N.json 2>/dev/null | tail -2
done
echo "== stock $N gpu"; timeout 400 $T --master-port 29607 bench.py --impl torch-stock --gpus $N --steps 40 --no-e2e 2>/dev/null | tail -1 | tee $O/stock_${N}gpu.json | cut -c1-260
