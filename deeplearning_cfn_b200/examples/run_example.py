"""`run_example` - start one of the bundled workloads on a provisioned cluster.

The reference documents three recipes (README.md:127-163, examples/tensorflow, examples/distributed-tensorflow):
MXNet parameter-server training through `launch.py`, TensorFlow parameter-server training through generated
per-host scripts, and TensorFlow+Horovod through `mpirun`.  `run_example(name, ...)` resolves the stack's
hostfile / environment (DEEPLEARNING_WORKERS_PATH, _COUNT, _GPU_COUNT, EFS_MOUNT) and runs the corresponding
recipe; every recipe is also runnable by hand exactly as in the reference README.

    python -m deeplearning_cfn_b200 run-example mxnet-cifar10-ps --stack-name dlcfn
    python -m deeplearning_cfn_b200 run-example resnet50-allreduce --gpus 8
"""
from __future__ import annotations

import argparse
import os
import subprocess
import sys

PY = sys.executable
EXAMPLES = {
    # name: (kind, description)
    "mxnet-cifar10-ps": ("launch", "CIFAR-10 VGG-11, kvstore dist_device_sync through launch (README.md:139)"),
    "resnet50-ps": ("mpirun", "ResNet-50 synthetic, fused parameter-server path (dist_sync_device)"),
    "resnet50-allreduce": ("mpirun", "ResNet-50 synthetic, fused all-reduce (Horovod train_synthetic.sh <np>)"),
    "resnet50-nccl": ("mpirun", "ResNet-50 synthetic, NCCL all-reduce baseline"),
    "bert-large-allreduce": ("bench", "BERT-large pre-training step, fused all-reduce with bucket fusion"),
    "tf-cifar10-ps-async": ("tfps", "CIFAR-10 CNN, asynchronous PS, per-host scripts from generate-trainer"),
    "maskrcnn-horovod": ("runsh", "Mask-RCNN R50-FPN through run.sh (hostfile with slots, mpirun)"),
    "bandwidth": ("mpirun", "push/pull + all-reduce bandwidth sweep 1 KiB - 1 GiB"),
}


def stack_env(stack_name: str | None, home: str | None) -> dict:
    env = dict(os.environ)
    if stack_name:
        from ..cluster.stack import describe_stack
        env.update(describe_stack(stack_name, home).export_env())
    return env


def build_command(name: str, env: dict, gpus: int | None, extra: list[str]) -> tuple[list[str], dict]:
    count = int(env.get("DEEPLEARNING_WORKERS_COUNT", "1"))
    per = int(env.get("DEEPLEARNING_WORKER_GPU_COUNT", "0"))
    hostfile = env.get("DEEPLEARNING_WORKERS_PATH")
    np_ = gpus or max(1, count * max(1, per))
    if name == "mxnet-cifar10-ps":
        cmd = [PY, "-m", "deeplearning_cfn_b200", "launch", "-n", str(count)]
        if hostfile:
            cmd += ["-H", hostfile]
        else:
            cmd += ["--launcher", "local"]
        cmd += [PY, "-m", "deeplearning_cfn_b200.examples.image_classification", "--dataset", "cifar10", "--model", "vgg11",
                "--epochs", "1", "--kvstore", "dist_device_sync"] + extra
        return cmd, env
    if name in ("resnet50-ps", "resnet50-allreduce", "resnet50-nccl"):
        kv = {"resnet50-ps": "dist_sync_device", "resnet50-allreduce": "horovod", "resnet50-nccl": "nccl"}[name]
        train = [PY, "-m", "deeplearning_cfn_b200.examples.train_imagenet", "--benchmark", "1", "--network", "resnet",
                 "--num-layers", "50", "--batch-size", "256", "--kv-store", kv, "--num-batches", "50"] + extra
        return [PY, "-m", "deeplearning_cfn_b200.launch.mpirun", "-np", str(np_)] + train, env
    if name == "bert-large-allreduce":
        root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        bench = [os.path.join(root, "bench.py"), "--model", "bert-large", "--gpus", str(np_)] + extra
        if np_ == 1:
            return [PY] + bench, env
        return [PY, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={np_}", "--master-addr", "127.0.0.1",
                "--master-port", "29577"] + bench, env
    if name == "bandwidth":
        return [PY, "-m", "deeplearning_cfn_b200.launch.mpirun", "-np", str(np_), PY, "-m",
                "deeplearning_cfn_b200.examples.bandwidth_measure", "--sweep"] + extra, env
    if name == "tf-cifar10-ps-async":
        if not hostfile:
            raise SystemExit("tf-cifar10-ps-async needs a stack (hostfile); create one with create-stack")
        shared = env.get("EFS_MOUNT", "/tmp")
        sdir, ldir = os.path.join(shared, "trainer_scripts"), os.path.join(shared, "logs")
        gen = [PY, "-m", "deeplearning_cfn_b200", "generate-trainer", "--trainer_script_dir", sdir, "--log_dir", ldir,
               "--workers_file_path", hostfile, "--worker_count", str(count), "--worker_gpu_count", str(max(1, per)),
               "--training_script", os.path.join(os.path.dirname(os.path.abspath(__file__)), "cifar10_multi_machine_train.py"),
               "--train_dir", os.path.join(shared, "cifar10_train")] + extra
        hosts = [l.strip() for l in open(hostfile) if l.strip()]
        runs = " ; ".join(f"{PY} -m deeplearning_cfn_b200 node-exec {h} -- bash {os.path.join(sdir, h + '.sh')}" for h in hosts)
        return ["bash", "-c", " ".join(gen) + " && " + runs], env
    if name == "maskrcnn-horovod":
        home = os.path.join(env.get("DLCFN_STACK_DIR", ""), "")
        script = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "cluster", "scripts", "run.sh")
        return ["bash", script] + extra, dict(env, DLCFN_PY=PY, HOME=env.get("HOME", home))
    raise SystemExit(f"unknown example {name!r}; choose from {sorted(EXAMPLES)}")


def run_example(name: str, stack_name: str | None = None, home: str | None = None, gpus: int | None = None,
                extra: list[str] | None = None, dry_run: bool = False) -> int:
    env = stack_env(stack_name, home)
    cmd, env = build_command(name, env, gpus, list(extra or []))
    if dry_run:
        print(" ".join(cmd))
        return 0
    return subprocess.call(cmd, env=env)


def main(argv=None) -> int:
    p = argparse.ArgumentParser(description="run a bundled example on a provisioned cluster")
    p.add_argument("name", nargs="?", default=None)
    p.add_argument("--stack-name", default=None)
    p.add_argument("--home", default=None)
    p.add_argument("--gpus", type=int, default=None)
    p.add_argument("--dry-run", action="store_true")
    p.add_argument("--list", action="store_true")
    a, extra = p.parse_known_args(argv)
    if a.list or a.name is None:
        for k, (_, d) in EXAMPLES.items():
            print(f"{k:24s} {d}")
        return 0
    from ..cluster.stack import StackError
    try:
        return run_example(a.name, a.stack_name, a.home, a.gpus, extra, a.dry_run)
    except StackError as e:
        print(f"error: {e} (create it with `python -m deeplearning_cfn_b200 create-stack --stack-name {a.stack_name or 'NAME'} ...`)",
              file=sys.stderr)
        return 1
    except KeyError as e:
        print(f"error: unknown example {e}; `run-example --list` shows the bundled ones", file=sys.stderr)
        return 1


if __name__ == "__main__":
    sys.exit(main())
