"""File-backed control-plane state shared by the provisioner, the scaling-event function and the node
agents of one stack (the roles played by the CloudFormation / AutoScaling / EC2 / IMDS APIs in the
reference: boto calls in dl_cfn_setup.py:209-296, boto3 calls in lambda_function.py:67-169).

Layout of a stack directory:
    stack.json                      status, parameters, outputs, event log
    groups/<Group>.json             min/max/desired, instances, suspended processes
    queues/<Queue>/                 FileQueue spool
    signals/<LogicalId>.json        resource signals (CreationPolicy / WaitCondition)
    nodes/<instance-id>/            the node's private root: etc/hosts, opt/deeplearning/workers, ...
All writers take the stack lock, so the agents may be separate processes.
"""
from __future__ import annotations

import contextlib
import fcntl
import json
import os
import time
import uuid

from .queue import FileQueue


class ControlPlane:
    def __init__(self, stack_dir: str):
        self.dir = os.path.abspath(stack_dir)
        for sub in ("groups", "queues", "signals", "nodes"):
            os.makedirs(os.path.join(self.dir, sub), exist_ok=True)

    # ------------------------------------------------------------------ locking / json io
    @contextlib.contextmanager
    def lock(self):
        fd = os.open(os.path.join(self.dir, ".lock"), os.O_CREAT | os.O_RDWR)
        try:
            fcntl.flock(fd, fcntl.LOCK_EX)
            yield
        finally:
            fcntl.flock(fd, fcntl.LOCK_UN)
            os.close(fd)

    def _read(self, rel: str, default=None):
        p = os.path.join(self.dir, rel)
        if not os.path.exists(p):
            return default
        with open(p) as f:
            return json.load(f)

    def _write(self, rel: str, obj) -> None:
        p = os.path.join(self.dir, rel)
        tmp = p + f".{os.getpid()}.tmp"
        with open(tmp, "w") as f:
            json.dump(obj, f, indent=1)
        os.rename(tmp, p)

    # ------------------------------------------------------------------ stack
    def stack(self) -> dict:
        return self._read("stack.json", {})

    def update_stack(self, **kv) -> dict:
        with self.lock():
            s = self._read("stack.json", {})
            s.update(kv)
            self._write("stack.json", s)
            return s

    def log_event(self, logical_id: str, status: str, reason: str = "") -> None:
        with self.lock():
            s = self._read("stack.json", {})
            s.setdefault("events", []).append({"ts": time.time(), "id": logical_id, "status": status, "reason": reason})
            self._write("stack.json", s)

    # ------------------------------------------------------------------ queues
    def queue(self, name: str) -> FileQueue:
        return FileQueue(os.path.join(self.dir, "queues", name))

    # ------------------------------------------------------------------ scaling groups
    def create_group(self, name: str, min_size: int, max_size: int, desired: int, **extra) -> dict:
        g = {"AutoScalingGroupName": name, "MinSize": int(min_size), "MaxSize": int(max_size),
             "DesiredCapacity": int(desired), "Instances": [], "SuspendedProcesses": []}
        g.update(extra)
        with self.lock():
            self._write(f"groups/{name}.json", g)
        return g

    def describe_group(self, name: str) -> dict | None:
        return self._read(f"groups/{name}.json")

    def group_names(self) -> list[str]:
        return sorted(f[:-5] for f in os.listdir(os.path.join(self.dir, "groups")) if f.endswith(".json"))

    def _mutate_group(self, name: str, fn):
        with self.lock():
            g = self._read(f"groups/{name}.json")
            if g is None:
                raise KeyError(name)
            fn(g)
            self._write(f"groups/{name}.json", g)
            return g

    def add_instance(self, group: str, ip: str, gpus: list[int], lifecycle: str = "Pending") -> dict:
        inst = {"InstanceId": "i-" + uuid.uuid4().hex[:8], "LifecycleState": lifecycle, "HealthStatus": "Healthy",
                "PrivateIpAddress": ip, "State": "pending", "Gpus": gpus, "LaunchTime": time.time()}
        self._mutate_group(group, lambda g: g["Instances"].append(inst))
        os.makedirs(self.node_root(inst["InstanceId"]), exist_ok=True)
        return inst

    def set_instance(self, group: str, instance_id: str, **kv) -> None:
        def f(g):
            for i in g["Instances"]:
                if i["InstanceId"] == instance_id:
                    i.update(kv)
        self._mutate_group(group, f)

    def remove_instance(self, group: str, instance_id: str) -> None:
        self._mutate_group(group, lambda g: g.__setitem__(
            "Instances", [i for i in g["Instances"] if i["InstanceId"] != instance_id]))

    def set_desired_capacity(self, group: str, desired: int) -> None:
        self._mutate_group(group, lambda g: g.__setitem__("DesiredCapacity", int(desired)))

    def suspend_processes(self, group: str, processes: list[str]) -> None:
        self._mutate_group(group, lambda g: g.__setitem__(
            "SuspendedProcesses", sorted(set(g["SuspendedProcesses"]) | set(processes))))

    def describe_instances(self, instance_ids: list[str]) -> list[dict]:
        out = []
        for name in self.group_names():
            for i in self.describe_group(name)["Instances"]:
                if i["InstanceId"] in instance_ids:
                    out.append(dict(i, AutoScalingGroupName=name))
        return out

    # ------------------------------------------------------------------ signals (cfn-signal / signal_resource)
    def signal_resource(self, logical_id: str, unique_id: str, status: str = "SUCCESS", data: str = "") -> None:
        with self.lock():
            s = self._read(f"signals/{logical_id}.json", [])
            if not any(x["id"] == unique_id for x in s):
                s.append({"id": unique_id, "status": status, "data": data, "ts": time.time()})
            self._write(f"signals/{logical_id}.json", s)

    def signals(self, logical_id: str) -> list[dict]:
        return self._read(f"signals/{logical_id}.json", [])

    # ------------------------------------------------------------------ nodes
    def node_root(self, instance_id: str) -> str:
        return os.path.join(self.dir, "nodes", instance_id)
