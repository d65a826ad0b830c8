"""`launch.py`-compatible job launcher.

Reference CLI (MX/tools/launch.py:57-118): `-n/--num-workers` (required), `-s/--num-servers` (default = n),
`-H/--hostfile`, `--sync-dst-dir`, `--launcher {local,ssh,mpi,sge,yarn}` (default ssh), `--env-server K:V`*,
`--env-worker K:V`*, `--env NAME`*, then the command.  With a hostfile and launcher=ssh the reference goes
through dmlc's ssh submitter; without a hostfile it falls back to local.  Added here: slurm / mesos /
kubernetes renderers and `-s 0` for all-reduce jobs (tracker = AllreduceTracker instead of a PS scheduler).

    python -m deeplearning_cfn_b200 launch -n $DEEPLEARNING_WORKERS_COUNT -H $DEEPLEARNING_WORKERS_PATH \
        python -m deeplearning_cfn_b200.examples.image_classification --dataset cifar10 --model vgg11 \
        --epochs 1 --kvstore dist_device_sync           # README.md:139 of the reference
"""
from __future__ import annotations

import argparse
import logging
import os
import sys

from . import tracker
from .submitters import get_submitter


def _kv_list(items) -> dict:
    out = {}
    for it in items or []:
        if ":" not in it:
            raise SystemExit(f"expected KEY:VALUE, got {it}")
        k, v = it.split(":", 1)
        out[k] = v
    return out


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="Launch a distributed job")
    p.add_argument("-n", "--num-workers", required=True, type=int, help="number of worker nodes to be launched")
    p.add_argument("-s", "--num-servers", type=int, help="number of server nodes to be launched, "
                   "in default it is equal to NUM_WORKERS")
    p.add_argument("-H", "--hostfile", type=str, help="the hostfile of slave machines which will run the job")
    p.add_argument("--sync-dst-dir", type=str, help="if specified, sync the current directory to this dir on all hosts")
    p.add_argument("--launcher", type=str, default="ssh",
                   choices=["local", "ssh", "mpi", "sge", "yarn", "slurm", "mesos", "kubernetes"])
    p.add_argument("--env-server", action="append", default=[], help="KEY:VALUE given to server processes")
    p.add_argument("--env-worker", action="append", default=[], help="KEY:VALUE given to worker processes")
    p.add_argument("--env", action="append", default=[], help="name of an environment variable to forward")
    p.add_argument("--host-ip", type=str, default="auto")
    p.add_argument("--render-dir", type=str, default=None, help="where batch-scheduler submissions are written")
    p.add_argument("--jobname", type=str, default=None, help="job name of the batch-scheduler submissions (default dlcfn)")
    p.add_argument("--kube-image", type=str, default=None, help="container image of the kubernetes Jobs")
    p.add_argument("command", nargs=argparse.REMAINDER, help="command for launching the program")
    return p


def main(argv=None) -> int:
    args = build_parser().parse_args(argv)
    if args.num_servers is None:
        args.num_servers = args.num_workers
    if args.command and args.command[0] == "--":
        args.command = args.command[1:]
    if not args.command:
        raise SystemExit("no command given")
    args.env_server = _kv_list(args.env_server)
    args.env_worker = _kv_list(args.env_worker)
    # one word = a shell command line as in the reference's README (`launch.py -n 2 "python train.py --lr 0.1"`); several
    # words = an argv, quoted so that empty arguments (`--gpus ""`) and arguments with spaces survive the shell
    import shlex
    cmd = args.command[0] if len(args.command) == 1 else shlex.join(args.command)
    kind = args.launcher
    if kind == "ssh" and args.hostfile is None:
        kind = "local"   # reference behaviour: no hostfile -> run everything here (launch.py:98-118)
    if kind in ("ssh", "mpi") and args.hostfile and not os.path.exists(args.hostfile):
        raise SystemExit(f"hostfile {args.hostfile} not found")
    logging.basicConfig(level=logging.INFO, format="%(asctime)s %(levelname)s %(message)s")
    fun = get_submitter(kind, args, cmd)
    host_ip = "127.0.0.1" if kind == "local" else args.host_ip
    # kubernetes: the scheduler is a Job of the cluster behind a Service (submitters.kubernetes_manifests), not a local process
    trk, handle = tracker.submit(args.num_workers, args.num_servers, fun, host_ip=host_ip,
                                 pscmd=None if kind == "kubernetes" else cmd)
    failed = handle.failed() if handle is not None else []
    if isinstance(trk, tracker.PSTracker):
        trk.stop()
    if failed:
        logging.error("processes failed: %s", failed)
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
