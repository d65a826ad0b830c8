"""Data / source staging config sets and the bucket stager.

Reference: the Mask-RCNN template stages a data tar and a source tar from S3 on boot
(mask-rcnn-cfn.yaml: marker files `data.txt`/`nodata.txt` on EFS :774-783; `04_data_efs` master copies
data S3->EFS when EFSServesData :1049-1062; `04_data_ebs`/`05_data_ebs` every node copies S3->EBS
otherwise :784-800; source tar + untar :801-827; `setup.sh` :828+; `08_run_script` on the master
:1107-1118), and `prepare-s3-bucket.sh:21-50` fills that bucket (COCO-2017 tar, TensorPack tar, run.sh,
setup.sh).  Here the "bucket" is a directory (there is no network); everything else keeps its shape.
"""
from __future__ import annotations

import os
import shutil
import stat
import tarfile
import time

import numpy as np


def _bucket_path(st: dict, name: str) -> str:
    return os.path.join(st["Bucket"], st.get("Prefix", ""), name)


def run_staging_configsets(lc: dict, node_root: str, efs_mount: str, is_master: bool, log=None) -> None:
    """Executes the staging config sets named in a launch configuration (no-op for the generic template)."""
    st = lc.get("Staging")
    sets = lc.get("ConfigSets", [])
    if not st or not st.get("Bucket"):
        return
    serves = str(st.get("EFSServesData", "true")).lower() == "true"
    ebs = os.path.join(node_root, "home", "data")           # the node's own scratch volume
    os.makedirs(ebs, exist_ok=True)
    note = (lambda i, s, r="": log(i, s, r)) if log else (lambda *a: None)
    if "data-marker" in sets and is_master:
        # data.txt: "the shared store serves the data"; nodata.txt: "stage your own copy"
        open(os.path.join(efs_mount, "data.txt" if serves else "nodata.txt"), "w").write(str(time.time()))
    tar_data = _bucket_path(st, st["TarData"])
    if "data-efs" in sets and is_master and serves and os.path.exists(tar_data):
        if not os.path.exists(os.path.join(efs_mount, "data", ".complete")):
            _untar(tar_data, os.path.join(efs_mount, "data"))
            open(os.path.join(efs_mount, "data", ".complete"), "w").write("ok")
        note("data-efs", "CREATE_COMPLETE", tar_data)
    if "data-ebs" in sets and not serves and os.path.exists(tar_data):
        _untar(tar_data, ebs)
        note("data-ebs", "CREATE_COMPLETE", tar_data)
    tar_src = _bucket_path(st, st["TarSource"])
    if "source-tar" in sets and os.path.exists(tar_src):
        _untar(tar_src, os.path.join(node_root, "home"))
    for key, cs in (("SetupScript", "setup-script"), ("RunScript", "run-script")):
        src = _bucket_path(st, st[key])
        if cs in sets and os.path.exists(src):
            dst = os.path.join(node_root, "home", os.path.basename(src))
            shutil.copyfile(src, dst)
            os.chmod(dst, os.stat(dst).st_mode | stat.S_IXUSR)


def _untar(path: str, dest: str) -> None:
    os.makedirs(dest, exist_ok=True)
    with tarfile.open(path) as t:
        t.extractall(dest, filter="data")


def data_dir(efs_mount: str, node_home: str, fsx_mount: str | None = None) -> str:
    """run.sh's choice of data directory: FSx -> shared store -> node scratch (run.sh:21-35)."""
    if fsx_mount and os.path.isdir(os.path.join(fsx_mount, "data")):
        return os.path.join(fsx_mount, "data")
    if os.path.exists(os.path.join(efs_mount, "data.txt")) and os.path.isdir(os.path.join(efs_mount, "data")):
        return os.path.join(efs_mount, "data")
    return os.path.join(node_home, "data")


def prepare_bucket(bucket: str, prefix: str = "mask-rcnn/b200", num_images: int = 64, image_size: int = 128,
                   num_classes: int = 80, seed: int = 0, source_dir: str | None = None) -> dict:
    """`prepare-s3-bucket.sh` without a network: synthesise a COCO-shaped detection shard
    (images + boxes + labels as .npz), tar it as coco-2017.tar, tar the training source, and upload both
    together with run.sh / setup.sh."""
    root = os.path.join(bucket, prefix)
    os.makedirs(root, exist_ok=True)
    stage = os.path.join(root, ".stage")
    shutil.rmtree(stage, ignore_errors=True)
    os.makedirs(os.path.join(stage, "coco"), exist_ok=True)
    rng = np.random.default_rng(seed)
    shard = 16
    for s in range(0, num_images, shard):
        n = min(shard, num_images - s)
        imgs = rng.integers(0, 256, (n, image_size, image_size, 3), dtype=np.uint8)
        nbox = rng.integers(1, 8, n)
        boxes, labels = [], []
        for k in range(n):
            xy = rng.uniform(0, image_size * 0.6, (nbox[k], 2))
            wh = rng.uniform(8, image_size * 0.4, (nbox[k], 2))
            boxes.append(np.concatenate([xy, np.minimum(xy + wh, image_size - 1)], 1).astype(np.float32))
            labels.append(rng.integers(1, num_classes + 1, nbox[k]).astype(np.int64))
        np.savez(os.path.join(stage, "coco", f"train2017-{s // shard:05d}.npz"), images=imgs,
                 boxes=np.array(boxes, dtype=object), labels=np.array(labels, dtype=object))
    out = {}
    out["data"] = os.path.join(root, "coco-2017.tar")
    with tarfile.open(out["data"], "w") as t:
        t.add(os.path.join(stage, "coco"), arcname="coco")
    src = source_dir or os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out["source"] = os.path.join(root, "tensorpack.tar")
    with tarfile.open(out["source"], "w") as t:
        t.add(src, arcname=os.path.basename(src), filter=lambda ti: None if "__pycache__" in ti.name or ti.name.endswith(".so") else ti)
    scripts = os.path.join(os.path.dirname(os.path.abspath(__file__)), "scripts")
    for name in ("run.sh", "setup.sh"):
        shutil.copyfile(os.path.join(scripts, name), os.path.join(root, name))
        out[name] = os.path.join(root, name)
    shutil.rmtree(stage, ignore_errors=True)
    return out


def package_event_function(out_path: str) -> str:
    """Build the deployable archive of the scaling-event function (reference `dl_cfn_setup_lambda.zip`: the packaged
    copy of cfn-lambda_function/lambda_function.py next to its SAM stub).  The archive holds the handler under the
    reference's module name plus the control-plane modules it imports and the function template; `Handler:
    lambda_function.lambda_handler` of templates/event-function.template resolves inside it."""
    import os
    import zipfile
    here = os.path.dirname(os.path.abspath(__file__))
    os.makedirs(os.path.dirname(os.path.abspath(out_path)) or ".", exist_ok=True)
    with zipfile.ZipFile(out_path, "w", zipfile.ZIP_DEFLATED) as z:
        z.writestr("lambda_function.py",
                   "from deeplearning_cfn_b200.cluster.events import *  # noqa: F401,F403\n"
                   "from deeplearning_cfn_b200.cluster.events import lambda_handler  # noqa: F401\n")
        z.writestr("deeplearning_cfn_b200/__init__.py", "")
        z.writestr("deeplearning_cfn_b200/cluster/__init__.py", "")
        for mod in ("events.py", "controlplane.py", "queue.py", "spec.py"):
            z.write(os.path.join(here, mod), f"deeplearning_cfn_b200/cluster/{mod}")
        z.write(os.path.join(here, "templates", "event-function.template"), "template.json")
    return out_path
