"""Hostfile parsing and alias resolution.

Formats accepted (all appear in the reference):
  * one alias per line - the agent's `/opt/deeplearning/workers` (dl_cfn_setup.py:94-100)
  * `host[:port]`      - dmlc ssh launcher hostfile (DT/ssh.py:40-54)
  * `host slots=G`     - Open MPI / Horovod hostfile written by run.sh:46-53 (README.md:151-163)
Aliases such as `deeplearning-worker3` resolve through the stack's hosts file (`DLCFN_HOSTS_FILE`,
the box-local stand-in for /etc/hosts) before falling back to the system resolver.
"""
from __future__ import annotations

import os
from dataclasses import dataclass


@dataclass
class Host:
    name: str
    port: int = 22
    slots: int = 1

    def address(self, hosts_file: str | None = None) -> str:
        return resolve_alias(self.name, hosts_file)


class HostfileError(ValueError):
    """A line of a hostfile that cannot be understood (reported as `file:line: what`)."""


def parse_hostfile(path: str) -> list[Host]:
    hosts = []
    with open(path) as f:
        for lineno, raw in enumerate(f, 1):
            line = raw.split("#", 1)[0].strip()
            if not line:
                continue
            parts = line.split()
            name, port, slots = parts[0], 22, 1
            if name.startswith("["):                         # [v6-address]:port
                addr, closing, rest = name[1:].partition("]")
                if not closing or not addr or (rest and not (rest.startswith(":") and rest[1:].isdigit())):
                    raise HostfileError(f"{path}:{lineno}: cannot read host {name!r} (expected [address] or [address]:port)")
                name, port = addr, int(rest[1:]) if rest else 22
            elif name.count(":") == 1:                       # host:port (several colons = a bare IPv6 address)
                host, p = name.split(":")
                if not host or not p.isdigit():
                    raise HostfileError(f"{path}:{lineno}: cannot read host {name!r} (expected host or host:port)")
                name, port = host, int(p)
            for tok in parts[1:]:
                if tok.startswith("slots="):
                    if not tok[6:].isdigit() or int(tok[6:]) < 1:
                        raise HostfileError(f"{path}:{lineno}: slots must be a positive integer, got {tok!r}")
                    slots = int(tok[6:])
                elif tok.startswith("max_slots="):
                    pass
            hosts.append(Host(name, port, slots))
    if not hosts:
        raise HostfileError(f"hostfile {path} is empty")
    return hosts


def write_mpi_hostfile(path: str, names: list[str], slots: int) -> None:
    with open(path, "w") as f:
        for n in names:
            f.write(f"{n} slots={slots}\n")


def load_hosts_map(hosts_file: str | None = None) -> dict:
    hosts_file = hosts_file or os.environ.get("DLCFN_HOSTS_FILE")
    m = {}
    if hosts_file and os.path.exists(hosts_file):
        for line in open(hosts_file):
            parts = line.split("#", 1)[0].split()
            for alias in parts[1:]:
                m[alias] = parts[0]
    return m


def resolve_alias(name: str, hosts_file: str | None = None) -> str:
    return load_hosts_map(hosts_file).get(name, name)


def node_env_for(name: str, stack_dir: str | None = None) -> dict:
    """Per-node environment of a stack node (its GPU slot and private root), looked up by alias."""
    stack_dir = stack_dir or os.environ.get("DLCFN_STACK_DIR")
    if not stack_dir:
        return {}
    from ..cluster.controlplane import ControlPlane
    cp = ControlPlane(stack_dir)
    ip = resolve_alias(name)
    for g in cp.group_names():
        for i in cp.describe_group(g)["Instances"]:
            if i["PrivateIpAddress"] == ip:
                root = cp.node_root(i["InstanceId"])
                env = {"DLCFN_NODE_ROOT": root, "HOME": os.path.join(root, "home")}
                if i.get("Gpus"):
                    env["CUDA_VISIBLE_DEVICES"] = ",".join(str(x) for x in i["Gpus"])
                prof = os.path.join(root, "etc/profile.d/deeplearning.sh")
                if os.path.exists(prof):
                    for line in open(prof):
                        if line.startswith("export ") and "=" in line:
                            k, v = line[7:].strip().split("=", 1)
                            env[k] = v
                return env
    return {}
