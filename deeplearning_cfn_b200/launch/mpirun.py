"""A small `mpirun`-compatible process launcher for one box (no MPI runtime is installed in the image).

Covers the subset the reference drives (run.sh:70-95, DT/mpi.py:12-77): `-np/-n N`, `--hostfile F` with
`host slots=G` lines, `-x NAME[=VALUE]`, `-env NAME VALUE`, `--output-filename DIR`, and silently accepts the
Open MPI tuning switches run.sh passes (`-bind-to`, `-map-by`, `-mca k v`, ...).  Ranks are mapped by slot:
host order from the hostfile, `slots` consecutive ranks per host.  Every rank gets the torchrun-style
environment (RANK, WORLD_SIZE, LOCAL_RANK, LOCAL_WORLD_SIZE, MASTER_ADDR/PORT) plus OMPI_COMM_WORLD_* so
Horovod-style scripts find what they expect.
"""
from __future__ import annotations

import os
import socket
import subprocess
import sys
import threading

from .hostfile import Host, parse_hostfile
from .transport import popen_on

_TAKES_ONE = {"-bind-to", "--bind-to", "-map-by", "--map-by", "--prefix", "-H", "--host", "-wdir", "--wdir"}
_TAKES_TWO = {"-mca", "--mca", "-gmca"}


def parse(argv: list[str]):
    np_, hostfile, outdir, exports, i = None, None, None, {}, 0
    while i < len(argv):
        a = argv[i]
        if a in ("-np", "-n", "--np"):
            if i + 1 >= len(argv) or not argv[i + 1].isdigit():
                raise SystemExit(f"mpirun: {a} takes a process count, got {argv[i + 1] if i + 1 < len(argv) else 'nothing'!r}")
            np_ = int(argv[i + 1]); i += 2
        elif a in ("--hostfile", "-hostfile", "--machinefile", "-machinefile"):
            hostfile = argv[i + 1]; i += 2
        elif a == "-x":
            kv = argv[i + 1]
            if "=" in kv:
                k, v = kv.split("=", 1)
                exports[k] = v
            elif kv in os.environ:
                exports[kv] = os.environ[kv]
            i += 2
        elif a == "-env":
            exports[argv[i + 1]] = argv[i + 2]; i += 3
        elif a in ("--output-filename", "-output-filename"):
            outdir = argv[i + 1]; i += 2
        elif a in _TAKES_TWO:
            i += 3
        elif a in _TAKES_ONE:
            i += 2
        elif a in ("--allow-run-as-root", "--oversubscribe", "-q", "--tag-output"):
            i += 1
        else:
            break
    return np_, hostfile, outdir, exports, argv[i:]


def rank_map(hosts: list[Host], n: int) -> list[tuple[Host, int, int]]:
    """[(host, local_rank, local_size)] for ranks 0..n-1, filled by slot."""
    out = []
    for h in hosts:
        take = min(h.slots, n - len(out))
        out += [(h, l, take) for l in range(take)]
        if len(out) >= n:
            break
    if len(out) < n:
        raise SystemExit(f"not enough slots: need {n}, hostfile provides {len(out)}")
    return out


def main(argv=None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0] in ("-h", "--help"):
        print("usage: mpirun [-np N] [--hostfile FILE (host slots=G per line)] [-x NAME[=VALUE]]... [--output-filename DIR]\n"
              "              [-bind-to X] [-map-by X] [-mca K V] [--allow-run-as-root] command [args...]\n"
              "starts N ranks (local or over the node transport) with RANK / LOCAL_RANK / WORLD_SIZE / LOCAL_WORLD_SIZE /\n"
              "MASTER_ADDR / MASTER_PORT and OMPI_COMM_WORLD_* set; the first rank that fails terminates the job")
        return 0 if argv else 2
    n, hostfile, outdir, exports, cmd = parse(argv)
    if not cmd:
        raise SystemExit("mpirun: no command")
    hosts = parse_hostfile(hostfile) if hostfile else [Host("127.0.0.1", 22, n or 1)]
    n = n or sum(h.slots for h in hosts)
    ranks = rank_map(hosts, n)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    master = ranks[0][0].address()
    if master.startswith("127.") or master == ranks[0][0].name:
        master = "127.0.0.1"
    rcs = [None] * n
    procs = []

    def run(r: int):
        h, lrank, lsize = ranks[r]
        env = dict(exports, RANK=r, WORLD_SIZE=n, LOCAL_RANK=lrank, LOCAL_WORLD_SIZE=lsize, MASTER_ADDR=master,
                   MASTER_PORT=port, OMPI_COMM_WORLD_RANK=r, OMPI_COMM_WORLD_SIZE=n, OMPI_COMM_WORLD_LOCAL_RANK=lrank,
                   OMPI_COMM_WORLD_LOCAL_SIZE=lsize)
        out = err = None
        if outdir:
            d = os.path.join(outdir, "1", f"rank.{r}")
            os.makedirs(d, exist_ok=True)
            out, err = open(os.path.join(d, "stdout"), "w"), open(os.path.join(d, "stderr"), "w")
        p = popen_on(h.name, h.port, env, os.getcwd(), " ".join(_q(c) for c in cmd), stdout=out, stderr=err)
        procs.append(p)
        rcs[r] = p.wait()

    threads = [threading.Thread(target=run, args=(r,), daemon=True) for r in range(n)]
    for t in threads:
        t.start()
    # like mpirun: the first rank that exits non-zero terminates the job (its peers would otherwise sit in a collective
    # until the transport times out - with NCCL that is many minutes)
    import time
    from .tracker import kill_tree
    aborted = None
    while any(t.is_alive() for t in threads):
        first_bad = next((r for r, rc in enumerate(rcs) if rc not in (None, 0)), None)
        if first_bad is not None and aborted is None:
            aborted = first_bad
            time.sleep(0.5)                      # let ranks that are failing for the same reason report it themselves
            for p in list(procs):
                kill_tree(p)
        time.sleep(0.05)
    bad = [r for r, rc in enumerate(rcs) if rc != 0]
    if bad:
        why = f" (rank {aborted} exited first with code {rcs[aborted]}; the job was terminated)" if aborted is not None else ""
        print(f"mpirun: ranks {bad} exited non-zero: {[rcs[r] for r in bad]}{why}", file=sys.stderr)
        return 1
    return 0


def _q(s: str) -> str:
    import shlex
    return shlex.quote(s)


if __name__ == "__main__":
    sys.exit(main())
