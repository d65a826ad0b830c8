"""How a command reaches a node.

`ssh` is what the reference uses everywhere (DT/ssh.py:77-85, run.sh:48-52, kill-mxnet.py:41-50).  All
nodes of a dlcfn stack live on one box behind loopback aliases, and the image has no sshd, so the
transport is pluggable: DLCFN_SSH=ssh builds real ssh command lines, DLCFN_SSH=local (default when the
target resolves to a loopback address) runs the same shell snippet locally with the node's environment
(its GPU slot as CUDA_VISIBLE_DEVICES, its private root as DLCFN_NODE_ROOT).
"""
from __future__ import annotations

import os
import shlex
import subprocess

from .hostfile import node_env_for, resolve_alias


def ssh_argv(host: str, port: int, remote_cmd: str, user: str | None = None) -> list[str]:
    target = f"{user}@{host}" if user else host
    return ["ssh", "-o", "StrictHostKeyChecking=no", target, "-p", str(port), remote_cmd]


def use_local(host: str) -> bool:
    mode = os.environ.get("DLCFN_SSH", "auto")
    if mode == "local":
        return True
    if mode == "ssh":
        return False
    ip = resolve_alias(host)
    return ip.startswith("127.") or ip in ("localhost", "::1")


def popen_on(host: str, port: int, exports: dict, workdir: str, cmd: str, user: str | None = None,
             stdout=None, stderr=None) -> subprocess.Popen:
    """Run `cmd` "on" `host` with `exports` set, in `workdir`."""
    export_str = " ".join(f"export {k}={shlex.quote(str(v))};" for k, v in exports.items())
    snippet = f"{export_str} cd {shlex.quote(workdir)}; {cmd}"
    if use_local(host):
        env = os.environ.copy()
        env.update(node_env_for(host))
        return subprocess.Popen(["bash", "-c", snippet], env=env, stdout=stdout, stderr=stderr)
    return subprocess.Popen(ssh_argv(resolve_alias(host), port, snippet, user), stdout=stdout, stderr=stderr)


def node_exec(host: str, command: list[str]) -> int:
    if command and command[0] == "--":
        command = command[1:]
    p = popen_on(host, 22, {}, os.getcwd(), " ".join(shlex.quote(c) for c in command))
    return p.wait()
