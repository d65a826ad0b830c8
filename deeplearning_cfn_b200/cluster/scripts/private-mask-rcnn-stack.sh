#!/bin/bash
# Creates a Mask-RCNN stack inside an existing private network (MyVpcId, PrivateSubnetId) and waits for it (counterpart of examples/distributed-tensorflow/mask-rcnn-stack.sh:
# same shell-variable interface, create-stack, poll every 30 s while CREATE_IN_PROGRESS, describe-stacks).
AWS_REGION=${AWS_REGION:-b200-box-1}
[[ -z "${AWS_REGION}" ]] && echo "AWS_REGION is required" && exit 1
AWS_AZ=${AWS_AZ:-nvswitch-0}
[[ -z "${AWS_AZ}" ]] && echo "AWS_AZ is required" && exit 1
S3_BUCKET=${S3_BUCKET:-}
[[ -z "${S3_BUCKET}" ]] && echo "S3_BUCKET is required" && exit 1
S3_PREFIX=${S3_PREFIX:-mask-rcnn/b200}
[[ -z "${S3_PREFIX}" ]] && echo "S3_PREFIX is required" && exit 1
KEY_NAME=${KEY_NAME:-local}
[[ -z "${KEY_NAME}" ]] && echo "KEY_NAME is required" && exit 1
INSTANCE_TYPE=${INSTANCE_TYPE:-b200.8gpu}
NUM_WORKERS=${NUM_WORKERS:-1}
EFS_ID=${EFS_ID:-}
SSH_LOCATION=${SSH_LOCATION:-127.0.0.0/8}
DATE=$(date +%s)
STACK_NAME=${STACK_NAME:-mask-rcnn-$DATE}
CFN_TEMPLATE=${CFN_TEMPLATE:-private-mask-rcnn.template}
DATA_TAR=${DATA_TAR:-coco-2017.tar}
SOURCE_TAR=${SOURCE_TAR:-tensorpack.tar}
AMI_ID=${AMI_ID:-}
EFS_SERVES=${EFS_SERVES:-false}
MY_VPC_ID=${MY_VPC_ID:-nvswitch-domain-0}
[[ -z "${MY_VPC_ID}" ]] && echo "MY_VPC_ID is required" && exit 1
PRIVATE_SUBNET_ID=${PRIVATE_SUBNET_ID:-127.1.0.0/16}
[[ -z "${PRIVATE_SUBNET_ID}" ]] && echo "PRIVATE_SUBNET_ID is required" && exit 1
POLL_SECS=${POLL_SECS:-30}
DLCFN="${DLCFN_PY:-python} -m deeplearning_cfn_b200"

$DLCFN create-stack --region "$AWS_REGION" --stack-name "$STACK_NAME" --template-body "file://$CFN_TEMPLATE" --no-wait \
  --parameters \
  ParameterKey=ActivateCondaEnv,ParameterValue= \
  ParameterKey=AMIOverride,ParameterValue=$AMI_ID \
  ParameterKey=EFSFileSystemId,ParameterValue=$EFS_ID \
  ParameterKey=EFSMountPoint,ParameterValue=efs \
  ParameterKey=EFSServesData,ParameterValue=$EFS_SERVES \
  ParameterKey=EbsVolumeSize,ParameterValue=200 \
  ParameterKey=ImageType,ParameterValue=B200Torch \
  ParameterKey=InstanceType,ParameterValue=$INSTANCE_TYPE \
  ParameterKey=KeyName,ParameterValue=$KEY_NAME \
  ParameterKey=MyAvailabilityZone,ParameterValue=$AWS_AZ \
  ParameterKey=S3Bucket,ParameterValue=$S3_BUCKET \
  ParameterKey=S3Prefix,ParameterValue=$S3_PREFIX \
  ParameterKey=RunScript,ParameterValue=run.sh \
  ParameterKey=SetupScript,ParameterValue=setup.sh \
  ParameterKey=SSHLocation,ParameterValue=$SSH_LOCATION \
  ParameterKey=TarData,ParameterValue=$DATA_TAR \
  ParameterKey=TarSource,ParameterValue=$SOURCE_TAR \
  ParameterKey=MyVpcId,ParameterValue=$MY_VPC_ID \
  ParameterKey=PrivateSubnetId,ParameterValue=$PRIVATE_SUBNET_ID \
  ParameterKey=WorkerCount,ParameterValue=$NUM_WORKERS || exit 1

progress=$($DLCFN list-stacks --stack-status-filter CREATE_IN_PROGRESS | grep -c "$STACK_NAME")
while [ "$progress" -ne 0 ]; do
  elapsed=$(( $(date +%s) - DATE ))
  echo "Stack $STACK_NAME status: Create in progress: [ $elapsed secs elapsed ]"
  sleep "$POLL_SECS"
  progress=$($DLCFN list-stacks --stack-status-filter CREATE_IN_PROGRESS | grep -c "$STACK_NAME")
done
sleep 1
$DLCFN describe-stacks --stack-name "$STACK_NAME"
