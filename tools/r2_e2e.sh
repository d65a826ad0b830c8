#!/bin/bash
# End-to-end product run on one 8-GPU box (verdict item 9): provision a stack (1 master + 7 workers, one GPU each), read the
# hostfile / environment the agents produced, run the bundled examples through the launchers, run the scaling sweep, keep the
# artefacts.   gpurun --gpus 8 -- 'bash tools/r2_e2e.sh'
set -u
OUT=gpurun_out/e2e_r2
mkdir -p $OUT
export PYTHONUNBUFFERED=1
HOME_DIR=/tmp/dlcfn-e2e
rm -rf $HOME_DIR
run() { local name=$1 t=$2; shift 2; echo "== $name: $*"; timeout "$t" "$@" > "$OUT/$name.log" 2>&1; echo "   [$name exit $?]"; tail -n 4 "$OUT/$name.log" | cut -c1-300; }
run create_stack 300 python -m deeplearning_cfn_b200 create-stack --stack-name e2e --home $HOME_DIR --parameters WorkerCount=7 InstanceType=b200.1gpu
run describe 60 python -m deeplearning_cfn_b200 describe-stacks --stack-name e2e --home $HOME_DIR
cp $HOME_DIR/stacks/e2e/shared/hostfile* $OUT/ 2>/dev/null; find $HOME_DIR -maxdepth 5 -name "workers*" -o -maxdepth 5 -name "hosts*" | head -5 | xargs -I{} cp {} $OUT/ 2>/dev/null
python -m deeplearning_cfn_b200 describe-stacks --stack-name e2e --home $HOME_DIR 2>/dev/null | python -c "import sys,json; d=json.load(sys.stdin); print(json.dumps(d.get('Stacks',[d])[0].get('Env',{}),indent=1))" > $OUT/env.json 2>/dev/null
for ex in resnet50-allreduce resnet50-ps resnet50-nccl mxnet-cifar10-ps tf-cifar10-ps-async; do
  run example_$ex 600 python -m deeplearning_cfn_b200 run-example $ex --stack-name e2e --home $HOME_DIR --gpus 8
done
run benchmark 1500 python -m deeplearning_cfn_b200.examples.benchmark --gpu_count 8 --networks native:resnet-50:256:224 --kv-stores allreduce dist_sync_device nccl --num-batches 40 --out-dir $OUT/benchmark
ls -la $OUT $OUT/benchmark 2>/dev/null | head -40
run delete_stack 120 python -m deeplearning_cfn_b200 delete-stack --stack-name e2e --home $HOME_DIR
