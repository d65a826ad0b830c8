#!/bin/bash
# Fills the staging bucket (counterpart of examples/distributed-tensorflow/prepare-s3-bucket.sh:21-50):
# data tar (synthetic COCO-shaped shards; there is no network to fetch COCO-2017), source tar, run.sh, setup.sh.
S3_BUCKET=${S3_BUCKET:-}
[[ -z "${S3_BUCKET}" ]] && echo "S3_BUCKET is required" && exit 1
S3_PREFIX=${S3_PREFIX:-mask-rcnn/b200}
${DLCFN_PY:-python} -m deeplearning_cfn_b200 prepare-bucket --bucket "$S3_BUCKET" --prefix "$S3_PREFIX" \
  --num-images "${NUM_IMAGES:-64}" --image-size "${IMAGE_SIZE:-128}"
