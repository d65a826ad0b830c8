"""In-job bootstrap: prepare the environment of a launched task, then exec the real command.

Reference: `DT/launcher.py:16-75` - the wrapper that batch schedulers (yarn/sge/mesos) run instead of the user
command: it unpacks the job archives (`DMLC_JOB_ARCHIVES`), extends `LD_LIBRARY_PATH` / `CLASSPATH`, and execs the
command.  Same role here: unpack archives into the working directory, put the job's source dir on PYTHONPATH,
derive per-node GPU visibility from the stack (DLCFN_STACK_DIR + DMLC_NODE_HOST) and exec."""
from __future__ import annotations

import os
import subprocess
import sys
import tarfile
import zipfile


def unpack(archive: str, dest: str) -> None:
    if archive.endswith(".zip"):
        with zipfile.ZipFile(archive) as z:
            z.extractall(dest)
    elif archive.endswith((".tar", ".tar.gz", ".tgz")):
        with tarfile.open(archive) as t:
            t.extractall(dest, filter="data")


def main(argv=None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        print("usage: bootstrap <command...>")
        return 1
    env = os.environ
    for a in filter(None, env.get("DMLC_JOB_ARCHIVES", "").split(":")):
        if os.path.exists(a):
            unpack(a, os.getcwd())
    src = env.get("DMLC_JOB_SOURCE_DIR", os.getcwd())
    env["PYTHONPATH"] = os.pathsep.join([src] + [p for p in env.get("PYTHONPATH", "").split(os.pathsep) if p])
    host = env.get("DMLC_NODE_HOST")
    if host and "CUDA_VISIBLE_DEVICES" not in env and env.get("DLCFN_STACK_DIR"):
        from .hostfile import node_env_for
        for k, v in node_env_for(host).items():
            env.setdefault(k, v)
    if env.get("DMLC_JOB_CLUSTER") in ("sge", "yarn", "mesos", "slurm") and "DMLC_TASK_ID" not in env:
        env["DMLC_TASK_ID"] = env.get("SGE_TASK_ID", env.get("SLURM_PROCID", "0"))
    return subprocess.call(argv, env=env)


if __name__ == "__main__":
    sys.exit(main())
