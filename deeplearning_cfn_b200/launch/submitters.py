"""Launcher back-ends: given (nworker, nserver, envs) start every process of a job.

Reference: `DT/local.py` (all processes local, workers FIRST, per-process retry DMLC_NUM_ATTEMPT :12-66 -
how CI fakes a cluster), `DT/ssh.py` (round-robin over the hostfile, servers FIRST, env exported inline,
optional rsync of the working dir :14-92), `DT/mpi.py` (two mpirun invocations :12-77), and the scheduler
front-ends `DT/{sge,slurm,yarn,mesos,kubernetes}.py`.  The scheduler front-ends here render the submission
(script / manifest / argv) and run it only if the scheduler's CLI exists on the box.
"""
from __future__ import annotations

import json
import os
import shutil
import subprocess
import sys
import threading
import time

from .hostfile import Host, parse_hostfile
from .transport import popen_on


class JobHandle:
    def __init__(self):
        self.threads: list[threading.Thread] = []
        self.returncodes: dict = {}
        self.procs: list = []

    def all_done(self) -> bool:
        return all(not t.is_alive() for t in self.threads)

    def wait(self, timeout: float | None = None) -> dict:
        t_end = None if timeout is None else time.time() + timeout
        for t in self.threads:
            t.join(None if t_end is None else max(0.0, t_end - time.time()))
        return self.returncodes

    def failed(self) -> list:
        return [k for k, v in self.returncodes.items() if v not in (0, None)]


def _role_env(envs: dict, role: str, task_id: int, extra: dict | None = None) -> dict:
    e = {k: str(v) for k, v in envs.items()}
    e["DMLC_ROLE"] = role
    e["DMLC_TASK_ID"] = str(task_id)
    if extra:
        e.update({k: str(v) for k, v in extra.items()})
    return e


# --------------------------------------------------------------------------- local
def local_submit(args, cmd: str):
    """Returns fun_submit for tracker.submit: W workers then S servers as local processes."""
    def run_one(handle: JobHandle, key: str, env: dict):
        attempts = int(env.get("DMLC_NUM_ATTEMPT", os.environ.get("DMLC_NUM_ATTEMPT", "0")))
        trial = 0
        full = os.environ.copy()
        full.update(env)
        while True:
            full["DMLC_NUM_ATTEMPT"] = str(trial)
            p = subprocess.Popen(cmd, shell=True, env=full, executable="/bin/bash")
            handle.procs.append(p)
            rc = p.wait()
            if rc == 0 or trial >= attempts:
                handle.returncodes[key] = rc
                return
            trial += 1

    def fun_submit(nworker: int, nserver: int, envs: dict):
        h = JobHandle()
        for i in range(nworker + nserver):
            role = "worker" if i < nworker else "server"
            tid = i if i < nworker else i - nworker
            extra = dict(getattr(args, "env_worker", {}) if role == "worker" else getattr(args, "env_server", {}))
            env = _role_env(envs, role, tid, extra)
            env["DMLC_JOB_CLUSTER"] = "local"
            t = threading.Thread(target=run_one, args=(h, f"{role}{tid}", env), daemon=True)
            t.start()
            h.threads.append(t)
        return h
    return fun_submit


# --------------------------------------------------------------------------- ssh
def sync_dir(local_dir: str, host: Host, remote_dir: str) -> list[str]:
    """rsync argv that mirrors the working directory to a host (DT/ssh.py:14-22)."""
    return ["rsync", "-az", "--rsh", f"ssh -o StrictHostKeyChecking=no -p {host.port}", local_dir.rstrip("/") + "/",
            f"{host.name}:{remote_dir}"]


def ssh_plan(hosts: list[Host], nworker: int, nserver: int) -> list[tuple[str, int, Host]]:
    """(role, task_id, host) per process: servers first, hosts round-robin (DT/ssh.py:77-85)."""
    plan = []
    for i in range(nworker + nserver):
        role = "server" if i < nserver else "worker"
        tid = i if i < nserver else i - nserver
        plan.append((role, tid, hosts[i % len(hosts)]))
    return plan


def ssh_submit(args, cmd: str):
    hosts = parse_hostfile(args.hostfile)
    workdir = args.sync_dst_dir if getattr(args, "sync_dst_dir", None) else os.getcwd()
    pass_env = {k: os.environ[k] for k in getattr(args, "env", []) or [] if k in os.environ}

    def fun_submit(nworker: int, nserver: int, envs: dict):
        if getattr(args, "sync_dst_dir", None) and shutil.which("rsync") and os.environ.get("DLCFN_SSH") == "ssh":
            for hst in hosts:
                subprocess.check_call(sync_dir(os.getcwd(), hst, workdir))
        h = JobHandle()
        for role, tid, hst in ssh_plan(hosts, nworker, nserver):
            extra = dict(getattr(args, "env_worker", {}) if role == "worker" else getattr(args, "env_server", {}))
            env = _role_env(envs, role, tid, extra)
            env["DMLC_NODE_HOST"] = hst.address()
            env.update(pass_env)

            def run(key=f"{role}{tid}", env=env, hst=hst):
                p = popen_on(hst.name, hst.port, env, workdir, cmd)
                h.procs.append(p)
                h.returncodes[key] = p.wait()
            t = threading.Thread(target=run, daemon=True)
            t.start()
            h.threads.append(t)
        return h
    return fun_submit


# --------------------------------------------------------------------------- mpi
def mpi_argv(n: int, envs: dict, role: str, cmd: str, hostfile: str | None, flavor: str = "openmpi") -> list[str]:
    """One mpirun invocation for one role (DT/mpi.py:12-77: OpenMPI passes env with -x, MPICH with -env)."""
    argv = ["mpirun", "-n", str(n)]
    if hostfile:
        argv += ["--hostfile", hostfile]
    e = dict(envs, DMLC_ROLE=role, DMLC_JOB_CLUSTER="mpi")
    for k, v in e.items():
        argv += (["-x", f"{k}={v}"] if flavor == "openmpi" else ["-env", k, str(v)])
    return argv + ["bash", "-c", cmd]


def mpi_submit(args, cmd: str):
    def fun_submit(nworker: int, nserver: int, envs: dict):
        h = JobHandle()
        exe = shutil.which("mpirun")
        for role, n in (("worker", nworker), ("server", nserver)):
            if n == 0:
                continue
            argv = mpi_argv(n, envs, role, cmd, getattr(args, "hostfile", None))
            if exe is None:
                # no MPI runtime in this image: use the dlcfn mpirun-compatible launcher
                argv = [sys.executable, "-m", "deeplearning_cfn_b200.launch.mpirun"] + argv[1:]

            def run(key=role, argv=argv):
                h.returncodes[key] = subprocess.call(argv)
            t = threading.Thread(target=run, daemon=True)
            t.start()
            h.threads.append(t)
        return h
    return fun_submit


# --------------------------------------------------------------------------- batch schedulers (rendered)
def sge_script(nworker: int, nserver: int, envs: dict, cmd: str, queue: str = "default", jobname: str = "dlcfn",
               logdir: str = "log") -> str:
    lines = ["#!/bin/bash", f"#$ -N {jobname}", f"#$ -q {queue}", f"#$ -t 1-{nworker + nserver}", f"#$ -o {logdir}",
             "#$ -j y", "#$ -cwd", "source ~/.bashrc"]
    lines += [f"export {k}={v}" for k, v in envs.items()]
    lines += [f"if [ $SGE_TASK_ID -le {nworker} ]; then export DMLC_ROLE=worker; export DMLC_TASK_ID=$((SGE_TASK_ID-1));",
              f"else export DMLC_ROLE=server; export DMLC_TASK_ID=$((SGE_TASK_ID-{nworker}-1)); fi",
              "export DMLC_JOB_CLUSTER=sge", cmd]
    return "\n".join(lines) + "\n"


def slurm_argv(n: int, nodes: int, envs: dict, role: str, cmd: str) -> list[str]:
    exports = ",".join(["ALL"] + [f"{k}={v}" for k, v in dict(envs, DMLC_ROLE=role, DMLC_JOB_CLUSTER="slurm").items()])
    return ["srun", f"--ntasks={n}", f"--nodes={nodes}", f"--export={exports}", "bash", "-c", cmd]


def kubernetes_manifests(jobname: str, nworker: int, nserver: int, envs: dict, image: str, cmd: str,
                         gpus_per_worker: int = 1) -> list[dict]:
    """Service + scheduler Job (parameter-server jobs), server Job, worker Job.  As in the reference (dmlc_tracker/kubernetes.py
    :107-132) the scheduler lives IN the cluster behind a Service named `<jobname>-scheduler`, and that name is the
    DMLC_PS_ROOT_URI of every pod; the pods of a Job carry the label the Service selects on."""
    sched = f"{jobname}-scheduler"
    if nserver:
        envs = dict(envs, DMLC_PS_ROOT_URI=sched, DMLC_PS_ROOT_PORT=int(envs.get("DMLC_PS_ROOT_PORT") or 9091))

    def job(role: str, n: int) -> dict:
        env = [{"name": k, "value": str(v)} for k, v in dict(envs, DMLC_ROLE=role, DMLC_JOB_CLUSTER="kubernetes").items()]
        c = {"name": f"{jobname}-{role}", "image": image, "command": ["bash", "-c", cmd], "env": env}
        if role == "worker" and gpus_per_worker:
            c["resources"] = {"limits": {"nvidia.com/gpu": gpus_per_worker}}
        return {"apiVersion": "batch/v1", "kind": "Job", "metadata": {"name": f"{jobname}-{role}"},
                "spec": {"parallelism": n, "completions": n,
                         "template": {"metadata": {"labels": {"app": f"{jobname}-{role}"}},
                                      "spec": {"containers": [c], "restartPolicy": "Never"}}}}
    svc = {"apiVersion": "v1", "kind": "Service", "metadata": {"name": sched},
           "spec": {"selector": {"app": sched}, "ports": [{"protocol": "TCP", "port": int(envs.get("DMLC_PS_ROOT_PORT") or 9091)}]}}
    out = [svc, job("scheduler", 1)] if nserver else []
    if nserver:
        out.append(job("server", nserver))
    out.append(job("worker", nworker))
    return out


def yarn_argv(nworker: int, nserver: int, envs: dict, cmd: str, queue: str = "default", jobname: str = "dlcfn",
              worker_cores: int = 1, worker_mem_mb: int = 4096) -> list[str]:
    argv = ["yarn", "jar", "dmlc-yarn.jar", "org.apache.hadoop.yarn.dmlc.Client", "-jobname", jobname, "-queue", queue,
            "-worker_cores", str(worker_cores), "-worker_memory_mb", str(worker_mem_mb)]
    for k, v in dict(envs, DMLC_JOB_CLUSTER="yarn").items():
        argv += ["-env", f"{k}={v}"]
    return argv + ["-nworker", str(nworker), "-nserver", str(nserver), cmd]


def mesos_tasks(nworker: int, nserver: int, envs: dict, cmd: str, cpus: int = 1, mem_mb: int = 4096) -> list[dict]:
    tasks = []
    for role, n in (("server", nserver), ("worker", nworker)):
        for i in range(n):
            env = dict(envs, DMLC_ROLE=role, DMLC_TASK_ID=i, DMLC_JOB_CLUSTER="mesos")
            tasks.append({"name": f"{role}-{i}", "cmd": cmd, "env": {k: str(v) for k, v in env.items()},
                          "resources": {"cpus": cpus, "mem": mem_mb}})
    return tasks


def _rendered_submit(kind: str, args, cmd: str):
    """sge / slurm / yarn / mesos / kubernetes: write the rendered submission next to the job and run the
    scheduler CLI when it is installed."""
    def fun_submit(nworker: int, nserver: int, envs: dict):
        out_dir = getattr(args, "render_dir", None) or os.getcwd()
        os.makedirs(out_dir, exist_ok=True)
        h = JobHandle()
        if kind == "sge":
            path = os.path.join(out_dir, "dlcfn-sge.sh")
            open(path, "w").write(sge_script(nworker, nserver, envs, cmd))
            argv = ["qsub", path]
        elif kind == "slurm":
            argv = slurm_argv(nworker, max(1, nworker), envs, "worker", cmd)
            open(os.path.join(out_dir, "dlcfn-slurm.json"), "w").write(json.dumps(
                {"worker": argv, "server": slurm_argv(nserver, max(1, nserver), envs, "server", cmd)}, indent=1))
        elif kind == "yarn":
            argv = yarn_argv(nworker, nserver, envs, cmd)
            open(os.path.join(out_dir, "dlcfn-yarn.json"), "w").write(json.dumps(argv, indent=1))
        elif kind == "mesos":
            path = os.path.join(out_dir, "dlcfn-mesos.json")
            open(path, "w").write(json.dumps(mesos_tasks(nworker, nserver, envs, cmd), indent=1))
            argv = ["mesos-execute", f"--master={os.environ.get('MESOS_MASTER', '127.0.0.1:5050')}", f"--task_group={path}"]
        else:
            path = os.path.join(out_dir, "dlcfn-k8s.json")
            jobname = getattr(args, "jobname", None) or "dlcfn"
            image = getattr(args, "kube_image", None) or os.environ.get("DLCFN_KUBE_IMAGE", "dlcfn-b200:latest")
            if nserver and os.environ.get("DLCFN_KUBE_SCHEDULER_PORT"):      # port of the scheduler Service (default 9091)
                envs = dict(envs, DMLC_PS_ROOT_PORT=int(os.environ["DLCFN_KUBE_SCHEDULER_PORT"]))
            items = kubernetes_manifests(jobname, nworker, nserver, envs, image, cmd)
            open(path, "w").write(json.dumps({"apiVersion": "v1", "kind": "List", "items": items}, indent=1))
            argv = ["kubectl", "create", "-f", path]
        # run the scheduler CLI(s) when installed (threads, like the other submitters: the tracker serves the rendezvous while
        # the submission is in flight); otherwise the rendered files are the product
        cmds = [argv]
        if kind == "slurm" and nserver > 0:       # slurm: one srun per role
            cmds = [slurm_argv(nserver, max(1, nserver), envs, "server", cmd), argv]
        if shutil.which(cmds[0][0]):
            for i, a in enumerate(cmds):
                def run(key=f"{kind}{i}", a=a):
                    h.returncodes[key] = subprocess.call(a)
                t = threading.Thread(target=run, daemon=True)
                t.start()
                h.threads.append(t)
        else:
            h.returncodes[kind] = None   # rendered only
        return h
    return fun_submit


def get_submitter(kind: str, args, cmd: str):
    if kind == "local":
        return local_submit(args, cmd)
    if kind == "ssh":
        return ssh_submit(args, cmd)
    if kind == "mpi":
        return mpi_submit(args, cmd)
    if kind in ("sge", "slurm", "yarn", "mesos", "kubernetes"):
        return _rendered_submit(kind, args, cmd)
    raise ValueError(f"unknown launcher {kind}")
