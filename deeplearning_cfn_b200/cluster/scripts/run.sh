#!/bin/bash
# Horovod-style job runner, executed on the master node of a Mask-RCNN stack.
# Counterpart of examples/distributed-tensorflow/run.sh: same environment contract
# (DEEPLEARNING_WORKERS_COUNT, DEEPLEARNING_WORKER_GPU_COUNT, EFS_MOUNT), same hostfile format
# ("host slots=G"), same tuning exports; the job itself is the dlcfn all-reduce trainer.
set -u
DLCFN_PY=${DLCFN_PY:-python}
rm -f ./hostfile

# GPUs per node: the agent writes the real count; keep the reference's default of 8 when it is 0
if [ "${DEEPLEARNING_WORKER_GPU_COUNT:-0}" -eq "0" ]; then
  echo "DEEPLEARNING_WORKER_GPU_COUNT is not set, setting it to default value of 8"
  DEEPLEARNING_WORKER_GPU_COUNT=8
fi

SRC_DIR=${SRC_DIR:-$HOME}
FSX_MOUNT=${FSX_MOUNT:-/fsx}
if [ -e "$FSX_MOUNT/data" ]; then FILE_SYSTEM="fsx"; DATA_DIR=$FSX_MOUNT/data
elif [ -e "$EFS_MOUNT/data" ]; then FILE_SYSTEM="efs"; DATA_DIR=$EFS_MOUNT/data
else FILE_SYSTEM="ebs"; DATA_DIR=$HOME/data; fi
LOG_DIR=$EFS_MOUNT
echo "Data directory: $DATA_DIR"
echo "Log directory: $LOG_DIR"

[ "$DEEPLEARNING_WORKERS_COUNT" -eq 1 ] || [ $((DEEPLEARNING_WORKERS_COUNT % 2)) -eq 0 ] || \
  { echo "DEEPLEARNING_WORKERS_COUNT must be 1 or even"; exit 1; }

for i in $(seq 1 "$DEEPLEARNING_WORKERS_COUNT"); do
  echo "deeplearning-worker$i slots=$DEEPLEARNING_WORKER_GPU_COUNT" >> ./hostfile
  echo "Running setup on deeplearning-worker$i"
  $DLCFN_PY -m deeplearning_cfn_b200 node-exec "deeplearning-worker$i" -- bash -l setup.sh 1>"setup-$i.log" 2>&1 || \
    { echo "setup failed on deeplearning-worker$i"; exit 1; }
done

NUM_PARALLEL=$((DEEPLEARNING_WORKERS_COUNT * DEEPLEARNING_WORKER_GPU_COUNT))
echo "Number of parallel ranks: $NUM_PARALLEL"
BATCH_NORM=${BATCH_NORM:-FreezeBN}     # or SyncBN
DATE=$(date '+%Y-%m-%d-%H-%M-%S')
RUN_ID=mask-rcnn-coco-$NUM_PARALLEL-$FILE_SYSTEM-$DATE
STEPS_PER_EPOCH=$((120000 / NUM_PARALLEL))
echo "Training started: $(date '+%Y-%m-%d-%H-%M-%S')"

HOROVOD_CYCLE_TIME=0.5 \
HOROVOD_FUSION_THRESHOLD=67108864 \
$DLCFN_PY -m deeplearning_cfn_b200.launch.mpirun -np "$NUM_PARALLEL" \
  --hostfile ./hostfile -bind-to none -map-by slot \
  -x NCCL_SOCKET_IFNAME=^docker0,lo -x NCCL_MIN_NRINGS=8 -x NCCL_DEBUG=${NCCL_DEBUG:-WARN} \
  -x LD_LIBRARY_PATH -x PATH -x HOROVOD_CYCLE_TIME -x HOROVOD_FUSION_THRESHOLD \
  --output-filename "$LOG_DIR/$RUN_ID" \
  $DLCFN_PY -m deeplearning_cfn_b200.examples.maskrcnn_train \
  --logdir "$LOG_DIR/$RUN_ID/train_log/maskrcnn" \
  --config MODE_MASK=True MODE_FPN=True DATA.BASEDIR="$DATA_DIR" \
  TRAIN.STEPS_PER_EPOCH="${STEPS_PER_EPOCH_OVERRIDE:-$STEPS_PER_EPOCH}" \
  TRAIN.LR_SCHEDULE='[240000, 320000, 360000]' BACKBONE.NORM="$BATCH_NORM" TRAINER=horovod ${EXTRA_CONFIG:-}
rc=$?
echo "Training finished: $(date '+%Y-%m-%d-%H-%M-%S')"
exit $rc
