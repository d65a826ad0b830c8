"""Per-host launch scripts for parameter-server jobs with one PS per node and one worker per GPU.

Reference: `examples/tensorflow/generate_trainer.py` - from the hostfile it writes `<dir>/<host>.sh`, each
starting 1 PS process on port 2222 with CUDA_VISIBLE_DEVICES='' and G worker processes on ports 2230+i
with CUDA_VISIBLE_DEVICES=i; worker task_index = node*G + i; logs to <log_dir>/{ps,worker}N
(:19-76).  CLI flags are identical (:4-15).  Two behaviours differ on purpose: the interpreter is taken
from the environment (`DLCFN_PY`, default `python`) and `--worker_count` actually limits the number of
nodes used (the reference parses it but ignores it, :64-76).
"""
from __future__ import annotations

import argparse
import os

PS_PORT = 2222
WORKER_PORT0 = 2230


def parse_args(argv=None):
    p = argparse.ArgumentParser(description="Generate per-host scripts for a PS training job")
    p.add_argument("--trainer_script_dir", type=str, required=True,
                   help="where the per-host scripts are written; use the shared store so every node sees them")
    p.add_argument("--log_dir", type=str, default="/tmp/", help="where ps/worker logs go")
    p.add_argument("--workers_file_path", type=str, required=True, help="hostfile written by the bootstrap agent")
    p.add_argument("--worker_count", type=int, required=True, help="number of nodes")
    p.add_argument("--worker_gpu_count", type=int, required=True, help="GPUs per node to use")
    p.add_argument("--training_script", nargs="+", help="training script and its arguments")
    args, unknown = p.parse_known_args(argv)
    args.training_script = " ".join((args.training_script or []) + unknown)
    return args


def worker_hosts(nodes: list[str], gpus: int) -> str:
    return ",".join(f"{n}:{WORKER_PORT0 + i}" for n in nodes for i in range(gpus))


def ps_hosts(nodes: list[str]) -> str:
    return ",".join(f"{n}:{PS_PORT}" for n in nodes)


def host_script(training_script: str, workers: str, ps: str, node_index: int, gpus: int, log_dir: str,
                interpreter: str = "python") -> str:
    log_dir = log_dir.rstrip("/")
    common = f"{interpreter} {training_script} --ps_hosts={ps} --worker_hosts={workers}"
    lines = ["source /etc/profile", "",
             f"CUDA_VISIBLE_DEVICES='' {common} --job_name=ps --task_index={node_index} > {log_dir}/ps{node_index} 2>&1 &", ""]
    for i in range(gpus):
        t = node_index * gpus + i
        lines += [f"CUDA_VISIBLE_DEVICES='{i}' {common} --job_name=worker --task_index={t} > {log_dir}/worker{t} 2>&1 &", ""]
    return "\n".join(lines) + "\n"


def gen_scripts(training_script: str, nodes_file: str, out_dir: str, num_nodes: int, gpus: int, log_dir: str,
                interpreter: str | None = None) -> list[str]:
    with open(nodes_file) as f:
        nodes = [l.strip() for l in f.read().splitlines() if l.strip()]
    if num_nodes:
        nodes = nodes[:num_nodes]
    interpreter = interpreter or os.environ.get("DLCFN_PY", "python")
    os.makedirs(out_dir, exist_ok=True)
    w, p = worker_hosts(nodes, gpus), ps_hosts(nodes)
    paths = []
    for idx, host in enumerate(nodes):
        path = os.path.join(out_dir, host + ".sh")
        with open(path, "w") as f:
            f.write(host_script(training_script, w, p, idx, gpus, log_dir, interpreter))
        paths.append(path)
    return paths


def main(argv=None) -> int:
    a = parse_args(argv)
    os.makedirs(a.log_dir, exist_ok=True)
    gen_scripts(a.training_script, a.workers_file_path, a.trainer_script_dir, a.worker_count, a.worker_gpu_count, a.log_dir)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
