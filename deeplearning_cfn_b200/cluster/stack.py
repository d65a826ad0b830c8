"""Stack provisioner: `launch_cluster` - turns a template + parameters into a running master/worker
cluster on one box and waits until every node has its hostfile + environment.

Reference flow: CloudFormation creates queues/topic/function/EFS, then the master scaling
group (size 1, CreationPolicy signal), then the worker group (desired = max = WorkerCount, min 0); every
instance boots, runs its config sets and finally the bootstrap agent; the scaling notifications drive the
function (events.lambda_handler); the stack is CREATE_COMPLETE when the master agent signals the wait
handle, otherwise it rolls back after the WaitCondition timeout (deeplearning.template:666-780, 174).

Box mapping: an *instance* is a slot of `InstanceType` GPUs of this box with a loopback address
(127.0.0.x public / 127.1.0.x private subnet); capacity = GPUs present (torch / nvidia-smi) or
`capacity_gpus`; an instance that cannot get its GPUs is a LAUNCH_ERROR -> the function shrinks the
worker group (elastic shrink at provision time, README.md:49).  GPU slots are handed out in the order
of the NVLink topology probe (cluster.discover), so rank order follows the fabric.
"""
from __future__ import annotations

import contextlib
import fcntl
import json
import os
import shutil
import subprocess
import sys
import threading
import time
import uuid
from dataclasses import dataclass, field

from . import agent as agent_mod
from . import events
from .controlplane import ControlPlane
from .discover import probe_topology
from .spec import ClusterSpec, evaluate

DEFAULT_HOME = os.environ.get("DLCFN_HOME", os.path.join(os.path.expanduser("~"), ".dlcfn"))


class StackError(RuntimeError):
    pass


@dataclass
class ClusterHandle:
    stack_name: str
    stack_dir: str
    status: str
    outputs: dict
    master_root: str = ""
    hostfile: str = ""
    hosts_file: str = ""
    env: dict = field(default_factory=dict)
    workers: list = field(default_factory=list)      # [{alias, ip, instance_id, gpus}]
    shared_mount: str = ""

    def export_env(self) -> dict:
        """The DEEPLEARNING_* contract (README.md:89-97) as a dict, plus box-local helpers."""
        e = dict(self.env)
        e["DLCFN_HOSTS_FILE"] = self.hosts_file
        e["DLCFN_STACK_DIR"] = self.stack_dir
        return e


def _parse_profile(path: str) -> dict:
    env = {}
    if os.path.exists(path):
        for line in open(path):
            line = line.strip()
            if line.startswith("export ") and "=" in line:
                k, v = line[7:].split("=", 1)
                env[k] = v
    return env


class Stack:
    def __init__(self, spec: ClusterSpec, home: str | None = None, backend: str = "process",
                 capacity_gpus: int | None = None, poll_interval: float = 0.05, fail_instances: int = 0):
        self.spec = spec
        self.home = home or DEFAULT_HOME
        self.dir = os.path.join(self.home, "stacks", spec.stack_name)
        self.backend = backend
        self.poll = poll_interval
        self.fail_instances = fail_instances       # fault injection: this many worker launches fail
        self.capacity_gpus = capacity_gpus
        self.cp: ControlPlane | None = None
        self._procs: list = []
        self._threads: list = []
        self._agent_rc: dict = {}
        self._next_ip = {"Public": 2, "Private": 2}
        self._free_gpus: list[int] = []

    # ------------------------------------------------------------------ helpers
    def _pseudo(self) -> dict:
        return {"DLCFN::StackId": self.dir, "DLCFN::StackName": self.spec.stack_name, "DLCFN::Region": "b200-box-1",
                "AWS::StackId": self.dir, "AWS::Region": "b200-box-1", "AWS::StackName": self.spec.stack_name}

    def _eval(self, node, attrs=None):
        t = self.spec.template
        return evaluate(node, self.spec.parameters, t.get("Mappings", {}), self._pseudo(), attrs or self._attrs)

    def _alloc_ip(self, subnet: str) -> str:
        n = self._next_ip[subnet]
        self._next_ip[subnet] += 1
        return f"127.{0 if subnet == 'Public' else 1}.{n // 250}.{n % 250 + 1}"

    def _publish(self, topic_env: dict, notification: dict) -> None:
        """SNS -> Lambda: run the function with its declared environment."""
        old = {k: os.environ.get(k) for k in topic_env}
        os.environ.update(topic_env)
        try:
            events.lambda_handler(notification, None)
        finally:
            for k, v in old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v

    # ------------------------------------------------------------------ create
    def create(self) -> ClusterHandle:
        spec = self.spec
        if os.path.exists(os.path.join(self.dir, "stack.json")):
            st = json.load(open(os.path.join(self.dir, "stack.json"))).get("StackStatus")
            if st not in ("DELETE_COMPLETE", None):
                raise StackError(f"stack {spec.stack_name} already exists ({st})")
            shutil.rmtree(self.dir, ignore_errors=True)
        os.makedirs(self.dir, exist_ok=True)
        self.cp = cp = ControlPlane(self.dir)
        cp.update_stack(StackName=spec.stack_name, StackStatus="CREATE_IN_PROGRESS", Parameters=spec.parameters,
                        CreationTime=time.time(), Outputs={}, Template=spec.template.get("Description", ""))
        res = spec.template["Resources"]
        self._attrs = {name: {"Ref": name} for name in res}
        try:
            return self._create_inner(res)
        except Exception as e:
            cp.log_event(spec.stack_name, "CREATE_FAILED", str(e))
            self._rollback()
            raise

    def _create_inner(self, res: dict) -> ClusterHandle:
        spec, cp = self.spec, self.cp
        # --- capacity from the fabric probe
        topo = probe_topology()
        gpus_total = self.capacity_gpus if self.capacity_gpus is not None else topo.num_gpus
        order = topo.rank_order()[:gpus_total] if topo.num_gpus else []
        candidates = list(order) if order else list(range(gpus_total))
        # The box is the region: GPUs held by the other live stacks of this home are not free (two stacks on one box must not
        # be handed the same devices).  What this stack can use is reserved under a home-wide lock BEFORE any instance starts;
        # a stack that finds fewer GPUs than it asked for shrinks like any other capacity shortage.
        want = (1 + spec.worker_count) * spec.gpus_per_instance
        with _home_lock(self.home):
            busy = _gpus_reserved_by_other_stacks(self.home, spec.stack_name)
            self._free_gpus = [g for g in candidates if g not in busy][:want]
            cp.update_stack(Topology=topo.summary(), ReservedGpus=list(self._free_gpus))
        if busy:
            cp.log_event(spec.stack_name, "CAPACITY", f"GPUs {sorted(busy)} are held by other stacks; {len(self._free_gpus)} of {want} reserved")

        # --- shared store (EFS): Retain on delete, reuse when an id is given
        store_id = spec.shared_store_id or ("fs-" + uuid.uuid4().hex[:8])
        store_dir = os.path.join(self.home, "stores", store_id)
        if spec.shared_store_id and not os.path.isdir(store_dir):
            raise StackError(f"shared store {store_id} does not exist")
        os.makedirs(store_dir, exist_ok=True)
        self._attrs["SharedStore"] = {"Ref": store_id, "Path": store_dir}
        cp.log_event("SharedStore", "CREATE_COMPLETE", store_dir)

        # --- queues / topic / function / wait handle
        for name, r in res.items():
            if r["Type"] == "DLCFN::Queue":
                cp.queue(name)
                cp.log_event(name, "CREATE_COMPLETE")
        fn = next((r for r in res.values() if r["Type"] == "DLCFN::Function"), None)
        fn_env = {k: str(v) for k, v in self._eval(fn["Properties"]["Environment"]).items()} if fn else {}
        self._fn_env = fn_env

        # --- master group, gated by its CreationPolicy signal; then workers; then the wait condition
        order_groups = ["MasterAutoScalingGroup", "WorkerAutoScalingGroup"]
        deadline = time.time() + spec.wait_timeout
        for gname in order_groups:
            r = res[gname]
            props = self._eval(r["Properties"])
            cp.create_group(gname, props["MinSize"], props["MaxSize"], props["DesiredCapacity"],
                            Subnet=props.get("Subnet", "Private"))
            cp.log_event(gname, "CREATE_IN_PROGRESS")
            self._launch_group(gname, props)
            if "CreationPolicy" in r:
                need = int(r["CreationPolicy"]["ResourceSignal"]["Count"])
                t_end = time.time() + spec.master_launch_timeout
                while len(cp.signals(gname)) < need:
                    if time.time() > t_end:
                        raise StackError(f"{gname}: no resource signal within {spec.master_launch_timeout}s")
                    time.sleep(self.poll)
            cp.log_event(gname, "CREATE_COMPLETE")

        handle_name = next(n for n, r in res.items() if r["Type"] == "DLCFN::WaitConditionHandle")
        while not any(s["status"] == "SUCCESS" for s in cp.signals(handle_name)):
            failed = [i for i, rc in self._agent_rc.items() if rc not in (None, 0)]
            if failed:
                raise StackError(f"bootstrap agent failed on {failed}")
            if time.time() > deadline:
                raise StackError(f"WaitCondition timed out after {spec.wait_timeout}s")
            time.sleep(self.poll)
        # let the worker agents finish writing their files
        self._join_agents(max(1.0, deadline - time.time()))
        failed = [i for i, rc in self._agent_rc.items() if rc not in (None, 0)]
        if failed:
            raise StackError(f"bootstrap agent failed on {failed}")

        outputs = {k: self._eval(v["Value"]) for k, v in spec.template.get("Outputs", {}).items()}
        used = sorted({g for gname in ("MasterAutoScalingGroup", "WorkerAutoScalingGroup")
                       for inst in (cp.describe_group(gname) or {}).get("Instances", []) for g in inst.get("Gpus", [])})
        cp.update_stack(StackStatus="CREATE_COMPLETE", Outputs=outputs, ReservedGpus=used)    # what a shrink left over is free again
        cp.log_event(spec.stack_name, "CREATE_COMPLETE")
        return self.handle()

    def _launch_group(self, gname: str, props: dict) -> None:
        cp, spec = self.cp, self.spec
        lc = props["LaunchConfiguration"]
        per = spec.gpus_per_instance
        subnet = props.get("Subnet", "Private")
        is_master = gname.startswith("Master")
        desired = int(props["DesiredCapacity"])
        for k in range(desired):
            inject_fail = (not is_master) and k >= desired - self.fail_instances
            if inject_fail or (per > 0 and len(self._free_gpus) < per):
                # InsufficientInstanceCapacity: no instance is created, the function decides what to do
                why = "fault injection" if inject_fail else f"need {per} GPUs, {len(self._free_gpus)} free"
                cp.log_event(gname, "LAUNCH_ERROR", why)
                if is_master:      # nothing can rendezvous without a master: fail now instead of waiting for the launch timeout
                    raise StackError(f"InsufficientInstanceCapacity: the master needs {per} GPUs, {len(self._free_gpus)} are free "
                                     f"on this box (other live stacks hold the rest: list-stacks / delete-stack)")
                self._publish(self._fn_env, events.make_notification(
                    events.LAUNCH_ERROR, gname, "", status_code="Failed",
                    status_message=f"InsufficientInstanceCapacity: {why}"))
                continue
            gpus = [self._free_gpus.pop(0) for _ in range(per)]
            inst = cp.add_instance(gname, self._alloc_ip(subnet), gpus)
            iid = inst["InstanceId"]
            self._boot_instance(gname, inst, lc)
            cp.set_instance(gname, iid, LifecycleState="InService", State="running")
            self._publish(self._fn_env, events.make_notification(events.LAUNCH, gname, iid))

    def _boot_instance(self, gname: str, inst: dict, lc: dict) -> None:
        """UserData + cfn-init config sets (deeplearning.template:490-665): efs-config, download-setup,
        deeplearning-config (= run the agent); the Mask-RCNN template adds data/source staging."""
        cp, spec = self.cp, self.spec
        iid = inst["InstanceId"]
        root = cp.node_root(iid)
        for d in ("etc/profile.d", "opt/deeplearning", "var/log", "home"):
            os.makedirs(os.path.join(root, d), exist_ok=True)
        with open(os.path.join(root, "etc/ssh_config"), "w") as f:
            f.write("Host *\n    ForwardAgent yes\n    StrictHostKeyChecking no\n")
        env = {k: str(v) for k, v in lc["Environment"].items()}
        # efs-config: mount the shared store at /<EFSMountPoint>
        mount = os.path.join(root, spec.shared_mount)
        if not os.path.lexists(mount):
            os.symlink(self._attrs["SharedStore"]["Path"], mount)
        env["EFS_MOUNT"] = mount
        env.update({"DLCFN_NODE_ROOT": root, "DLCFN_INSTANCE_ID": iid,
                    "DLCFN_NODE_GPUS": ",".join(str(g) for g in inst["Gpus"]),
                    "DLCFN_POLL_INTERVAL": str(self.poll)})
        from .staging import run_staging_configsets
        run_staging_configsets(lc, root, mount, is_master=gname.startswith("Master"), log=cp.log_event)
        self._agent_rc[iid] = None
        if self.backend == "thread":
            def run():
                a = agent_mod.Agent(env, poll_interval=self.poll)
                try:
                    a.run()
                    self._agent_rc[iid] = 0
                except Exception as e:
                    self.cp.log_event(iid, "AGENT_FAILED", str(e))
                    self._agent_rc[iid] = 1
            t = threading.Thread(target=run, daemon=True, name=f"agent-{iid}")
            t.start()
            self._threads.append((iid, t))
        else:
            full = dict(os.environ)
            full.update(env)
            full["PYTHONPATH"] = os.pathsep.join(
                [os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))] +
                [p for p in full.get("PYTHONPATH", "").split(os.pathsep) if p])
            p = subprocess.Popen([sys.executable, "-m", "deeplearning_cfn_b200.cluster.agent"], env=full,
                                 stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            self._procs.append((iid, p))

    def _join_agents(self, timeout: float) -> None:
        t_end = time.time() + timeout
        for iid, t in self._threads:
            t.join(max(0.0, t_end - time.time()))
        for iid, p in self._procs:
            try:
                self._agent_rc[iid] = p.wait(max(0.1, t_end - time.time()))
            except subprocess.TimeoutExpired:
                self._agent_rc[iid] = None
        # poll process-backed agents that finished earlier
        for iid, p in self._procs:
            if p.poll() is not None:
                self._agent_rc[iid] = p.returncode

    def _rollback(self) -> None:
        for _, p in self._procs:
            if p.poll() is None:
                p.terminate()
        if self.cp:
            self.cp.update_stack(StackStatus="ROLLBACK_COMPLETE")

    # ------------------------------------------------------------------ describe / delete
    def handle(self) -> ClusterHandle:
        return describe_stack(self.spec.stack_name, self.home)


_LIVE = ("CREATE_IN_PROGRESS", "CREATE_COMPLETE")


@contextlib.contextmanager
def _home_lock(home: str):
    os.makedirs(home, exist_ok=True)
    fd = os.open(os.path.join(home, ".capacity.lock"), os.O_CREAT | os.O_RDWR)
    try:
        fcntl.flock(fd, fcntl.LOCK_EX)
        yield
    finally:
        fcntl.flock(fd, fcntl.LOCK_UN)
        os.close(fd)


def _gpus_reserved_by_other_stacks(home: str, me: str) -> set:
    busy: set = set()
    base = os.path.join(home, "stacks")
    if not os.path.isdir(base):
        return busy
    for name in os.listdir(base):
        if name == me:
            continue
        try:
            with open(os.path.join(base, name, "stack.json")) as f:
                st = json.load(f)
        except (OSError, ValueError):
            continue
        if st.get("StackStatus") in _LIVE:
            busy.update(st.get("ReservedGpus", []))
    return busy


def describe_stack(stack_name: str, home: str | None = None) -> ClusterHandle:
    home = home or DEFAULT_HOME
    sdir = os.path.join(home, "stacks", stack_name)
    if not os.path.exists(os.path.join(sdir, "stack.json")):       # (looking must not create the directory tree of a stack)
        raise StackError(f"no such stack: {stack_name}")
    cp = ControlPlane(sdir)
    st = cp.stack()
    if not st:
        raise StackError(f"no such stack: {stack_name}")
    h = ClusterHandle(stack_name, sdir, st.get("StackStatus", "UNKNOWN"), st.get("Outputs", {}))
    mg = cp.describe_group("MasterAutoScalingGroup")
    if mg and mg["Instances"]:
        root = cp.node_root(mg["Instances"][0]["InstanceId"])
        h.master_root = root
        h.hostfile = os.path.join(root, "opt/deeplearning/workers")
        h.hosts_file = os.path.join(root, "etc/hosts")
        h.env = _parse_profile(os.path.join(root, "etc/profile.d/deeplearning.sh"))
        h.shared_mount = h.env.get("EFS_MOUNT", "")
        ip2inst = {}
        for gname in cp.group_names():
            for i in cp.describe_group(gname)["Instances"]:
                ip2inst[i["PrivateIpAddress"]] = i
        if os.path.exists(h.hosts_file):
            for line in open(h.hosts_file):
                parts = line.split()
                if len(parts) == 2 and parts[1].startswith("deeplearning-worker"):
                    inst = ip2inst.get(parts[0], {})
                    h.workers.append({"alias": parts[1], "ip": parts[0], "instance_id": inst.get("InstanceId"),
                                      "gpus": inst.get("Gpus", [])})
    return h


def list_stacks(home: str | None = None) -> list[dict]:
    home = home or DEFAULT_HOME
    base = os.path.join(home, "stacks")
    out = []
    if os.path.isdir(base):
        for name in sorted(os.listdir(base)):
            st = ControlPlane(os.path.join(base, name)).stack()
            if st:
                out.append({"StackName": name, "StackStatus": st.get("StackStatus"), "CreationTime": st.get("CreationTime")})
    return out


def delete_stack(stack_name: str, home: str | None = None) -> None:
    """Deletes the stack; the shared store has DeletionPolicy Retain (deeplearning.template:456) and
    survives, so a new stack can be created over it and training resumed from its checkpoints."""
    home = home or DEFAULT_HOME
    sdir = os.path.join(home, "stacks", stack_name)
    if not os.path.isdir(sdir):
        raise StackError(f"no such stack: {stack_name}")
    cp = ControlPlane(sdir)
    st = cp.stack()
    keep = {"StackName": stack_name, "StackStatus": "DELETE_COMPLETE", "Outputs": st.get("Outputs", {}),
            "events": st.get("events", [])}
    for sub in ("groups", "queues", "signals", "nodes"):
        shutil.rmtree(os.path.join(sdir, sub), ignore_errors=True)
    with open(os.path.join(sdir, "stack.json"), "w") as f:
        json.dump(keep, f, indent=1)


def launch_cluster(stack_name: str = "dlcfn", template: str = "deeplearning.template", home: str | None = None,
                   backend: str = "process", capacity_gpus: int | None = None, poll_interval: float = 0.05,
                   fail_instances: int = 0, wait_timeout: float | None = None, **parameters) -> ClusterHandle:
    """Public entry point: provision master + workers, run the rendezvous, return hostfile + env."""
    spec = ClusterSpec.from_template(stack_name, template, **parameters)
    if wait_timeout is not None:
        spec.wait_timeout = wait_timeout
        spec.master_launch_timeout = min(spec.master_launch_timeout, wait_timeout)
    return Stack(spec, home, backend, capacity_gpus, poll_interval, fail_instances).create()
