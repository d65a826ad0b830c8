"""Node bootstrap / rank-discovery agent - runs once on every node of a stack.

Reference: `cfn-bootstrap/dl_cfn_setup_v2.py` (call stack in SURVEY §3.1).  Behaviour kept:
  * all configuration arrives through the 12 `AWS_DL_*`/`EFS_MOUNT`/`CFN_PATH` environment variables
    (:393-404); any failure exits 1 (-> stack rollback);
  * MASTER: wait for one `asg-setup` success message per scaling group on the master queue (dedupe by
    group name, delete each message, :122-167), resolve the instances' addresses (:209-280), rank order
    = sorted addresses with the master counted as a worker (:328-340), write the metadata files, then
    broadcast `worker-setup` on the worker queue (:344-355) and signal the wait handle (:285-296);
  * WORKER: poll the worker queue with visibility_timeout=0 and never delete (:169-207), then write the
    SAME files;
  * files: `<root>/etc/hosts` aliases `deeplearning-master`, `deeplearning-worker<i>` (i from 1),
    `<root>/opt/deeplearning/workers` (THE HOSTFILE), `<root>/etc/profile.d/deeplearning.sh` exporting
    DEEPLEARNING_WORKERS_COUNT / _PATH / DEEPLEARNING_WORKER_GPU_COUNT / EFS_MOUNT (:91-115).
Differences by design: `<root>` is the node's private root inside the stack directory (DLCFN_NODE_ROOT;
"/" on a real multi-host deployment); GPU count comes from the node's GPU slot / `nvidia-smi -L`
rather than an instance-type allow-list (which in the reference misses p3dn and yields 0, forcing run.sh
to hard-code 8); the inverted timeout test of :269 is not reproduced.
"""
from __future__ import annotations

import json
import logging
import os
import subprocess
import sys
import time

from .controlplane import ControlPlane

SLEEP_INTERVAL_IN_SECS = float(os.environ.get("DLCFN_POLL_INTERVAL", "30"))
LOGGER = logging.getLogger("dl-cfn-setup")

ENV_KEYS = ("AWS_DL_NODE_TYPE", "AWS_DL_MASTER_QUEUE", "AWS_DL_WORKER_QUEUE", "AWS_DL_WAITCONDITION_TIMEOUT",
            "AWS_DL_MASTERLAUNCH_TIMEOUT", "AWS_DL_STACK_ID", "AWS_DL_WAIT_HANDLE", "AWS_DL_ROLE_NAME",
            "AWS_DL_DEFAULT_USER", "AWS_REGION", "EFS_MOUNT", "CFN_PATH")


class AgentError(RuntimeError):
    pass


def setup_logging(log_loc: str) -> logging.Logger:
    os.makedirs(log_loc, exist_ok=True)
    LOGGER.setLevel(logging.INFO)
    fmt = logging.Formatter("%(asctime)s %(levelname)s: %(filename)s:%(lineno)d %(message)s")
    for h in list(LOGGER.handlers):
        LOGGER.removeHandler(h)
    fh = logging.FileHandler(os.path.join(log_loc, "dl_cfn_setup.log"))
    fh.setFormatter(fmt)
    LOGGER.addHandler(fh)
    if os.environ.get("DLCFN_AGENT_CONSOLE", "0") == "1":
        ch = logging.StreamHandler()
        ch.setFormatter(fmt)
        LOGGER.addHandler(ch)
    return LOGGER


class Agent:
    def __init__(self, env: dict, poll_interval: float | None = None):
        missing = [k for k in ENV_KEYS if k not in env]
        if missing:
            raise AgentError(f"missing environment variables: {missing}")
        self.env = env
        self.node_type = env["AWS_DL_NODE_TYPE"].lower()
        self.cp = ControlPlane(env["AWS_DL_STACK_ID"])
        self.root = env.get("DLCFN_NODE_ROOT", "/")
        self.instance_id = env.get("DLCFN_INSTANCE_ID", "i-unknown")
        self.poll = SLEEP_INTERVAL_IN_SECS if poll_interval is None else poll_interval
        self.host_file = os.path.join(self.root, "etc/hosts")
        self.worker_file = os.path.join(self.root, "opt/deeplearning/workers")
        self.profile = os.path.join(self.root, "etc/profile.d/deeplearning.sh")

    # ------------------------------------------------------------------ instance metadata
    def check_instance_role_availability(self, role_name: str, timeout: float) -> bool:
        """The reference polls IMDS for the instance role (:357-384) because queue access depends on it;
        here the equivalent precondition is that the stack's control plane is reachable."""
        start = time.time()
        while True:
            if os.path.isdir(self.cp.dir) and os.access(self.cp.dir, os.R_OK | os.W_OK):
                LOGGER.info("SUCCESS getting instance role %s", role_name)  # never log credentials (:368-371)
                return True
            if time.time() - start > timeout:
                LOGGER.info("TIMEOUT while checking instance role after %s seconds", timeout)
                return False
            time.sleep(self.poll)

    def get_gpu_count(self) -> int:
        slot = self.env.get("DLCFN_NODE_GPUS")
        if slot is not None:
            return len([g for g in slot.split(",") if g != ""])
        try:
            out = subprocess.check_output(["nvidia-smi", "-L"], stderr=subprocess.DEVNULL).decode()
            return out.count("\n")
        except (OSError, subprocess.CalledProcessError):
            LOGGER.info("Not a GPU node, number of GPUs: 0")
            return 0

    # ------------------------------------------------------------------ master side
    def wait_until_asg_success(self, queue_name: str, timeout: float, expected: int = 2) -> dict:
        q = self.cp.queue(queue_name)
        got: dict = {}
        start = time.time()
        while True:
            for msg in q.receive(max_messages=10, visibility_timeout=60):
                try:
                    c = msg.json()
                    if c is not None and c["event"] == "asg-setup" and c["status"] == "success":
                        if c["asg"] not in got:       # at-least-once delivery: ignore duplicates
                            got[c["asg"]] = c
                        else:
                            LOGGER.info("received duplicate message for %s", c["asg"])
                    q.delete(msg)
                except (TypeError, KeyError, ValueError) as e:
                    LOGGER.error("bad message %r: %s", msg.body, e)
                    continue
            if len(got) == expected:
                break
            if time.time() - start + self.poll > timeout:
                LOGGER.info("timeout while checking asg status after %s seconds", timeout)
                break
            time.sleep(self.poll)
        return got

    def wait_until_instances_active(self, groups: list[str], timeout: float):
        start = time.time()
        master_ids, worker_ids = [], []
        for name in groups:
            g = self.cp.describe_group(name)
            ids = [i["InstanceId"] for i in g["Instances"] if i["HealthStatus"] == "Healthy"]
            (master_ids if "master" in name.lower() else worker_ids).extend(ids)
        pending = set(master_ids + worker_ids)
        masters, workers = {}, {}
        while True:
            for i in self.cp.describe_instances(sorted(pending)):
                if i["State"].lower() == "running":
                    (masters if i["InstanceId"] in master_ids else workers)[i["InstanceId"]] = i["PrivateIpAddress"]
                    pending.discard(i["InstanceId"])
            if not pending:
                break
            if time.time() - start > timeout:   # (the reference's comparison is inverted, :269)
                LOGGER.error("Reached timeout, pending instance ids: %s", sorted(pending))
                break
            time.sleep(self.poll)
        return masters, workers

    def setup_worker_metadata(self, timeout: float):
        start = time.time()
        msgs = self.wait_until_asg_success(self.env["AWS_DL_MASTER_QUEUE"], timeout)
        if len(msgs) != 2:
            raise AgentError(f"did not receive asg success message for all scaling groups, got {sorted(msgs)}")
        master_msg = next(v for k, v in msgs.items() if "master" in k.lower())
        worker_msg = next(v for k, v in msgs.items() if "master" not in k.lower())
        remaining = timeout - (time.time() - start)
        masters, workers = self.wait_until_instances_active([master_msg["asg"], worker_msg["asg"]], remaining)
        if len(masters) != 1:
            raise AgentError(f"expected single master, instead got {masters}")
        master_ip = list(masters.values())[0]
        ips = [master_ip]                      # the master is also a worker
        if not workers:
            LOGGER.info("no worker is launched, using only master instance as worker")
        else:
            ips.extend(workers.values())
            if len(workers) != worker_msg["launched"]:
                LOGGER.error("expected %s running workers, got %s", worker_msg["launched"], workers)
        return master_ip, sorted(ips, key=_ip_key)   # rank order = sorted address

    def send_worker_setup_msg(self, master_ip: str, worker_ips: list[str]) -> None:
        body = {"event": "worker-setup", "master-ip": master_ip, "worker-ips": worker_ips}
        LOGGER.info("sending worker-setup message: %s", json.dumps(body))
        self.cp.queue(self.env["AWS_DL_WORKER_QUEUE"]).send(body)

    def send_cfn_success_signal(self) -> None:
        """cfn-signal --success true <wait handle> (:285-296)."""
        self.cp.signal_resource(self.env["AWS_DL_WAIT_HANDLE"], self.instance_id, "SUCCESS")

    # ------------------------------------------------------------------ worker side
    def wait_for_worker_setup_message(self, timeout: float):
        q = self.cp.queue(self.env["AWS_DL_WORKER_QUEUE"])
        start = time.time()
        while True:
            # visibility_timeout=0 and no delete: every worker consumes the same message
            for msg in q.receive(max_messages=10, visibility_timeout=0):
                try:
                    c = msg.json()
                    if c is not None and c["event"] == "worker-setup":
                        return c["master-ip"], c["worker-ips"]
                except (TypeError, KeyError, ValueError) as e:
                    LOGGER.error("bad message %r: %s", msg.body, e)
            if time.time() - start + self.poll > timeout:
                LOGGER.info("did not receive worker-setup even after %s seconds", timeout)
                return None, None
            time.sleep(self.poll)

    # ------------------------------------------------------------------ both
    def setup_env_variables(self, master_ip: str, worker_ips: list[str]) -> None:
        for p in (self.host_file, self.worker_file, self.profile):
            os.makedirs(os.path.dirname(p), exist_ok=True)
        with open(self.host_file, "a") as hosts, open(self.worker_file, "w+") as w:
            hosts.write(f"{master_ip} deeplearning-master\n")
            for idx, ip in enumerate(worker_ips, start=1):
                hosts.write(f"{ip} deeplearning-worker{idx}\n")
                w.write(f"deeplearning-worker{idx}\n")
        self._chown_to_default_user(self.worker_file)
        gpu_count = self.get_gpu_count()
        with open(self.worker_file) as f:
            n = sum(1 for _ in f)
        with open(self.profile, "a") as f:
            f.write(f"export DEEPLEARNING_WORKERS_COUNT={n}\n")
            f.write(f"export DEEPLEARNING_WORKERS_PATH={self.worker_file}\n")
            f.write(f"export DEEPLEARNING_WORKER_GPU_COUNT={gpu_count}\n")
            f.write(f"export EFS_MOUNT={self.env['EFS_MOUNT']}\n")

    def _chown_to_default_user(self, path: str) -> None:
        """The workers file belongs to the image's login user, who runs the jobs (dl_cfn_setup_v2.py:110-114: chown to
        AWS_DL_DEFAULT_USER).  On a box where that user does not exist, or the agent is not privileged, the file keeps its
        owner - the reference would abort the whole bootstrap on a missing user; a hostfile with the wrong owner is still usable."""
        user = self.env.get("AWS_DL_DEFAULT_USER")
        if not user:
            return
        try:
            import grp
            import pwd
            os.chown(path, pwd.getpwnam(user).pw_uid, grp.getgrnam(user).gr_gid)
            LOGGER.info("chown %s to %s", path, user)
        except (KeyError, PermissionError, OSError) as e:
            LOGGER.info("workers file keeps its owner (default user %r: %s)", user, e)

    def run(self) -> None:
        e = self.env
        setup_timeout = float(e["AWS_DL_WAITCONDITION_TIMEOUT"]) - float(e["AWS_DL_MASTERLAUNCH_TIMEOUT"])
        start = time.time()
        if not self.check_instance_role_availability(e["AWS_DL_ROLE_NAME"], setup_timeout):
            raise AgentError("control plane not reachable")
        setup_timeout -= time.time() - start
        if self.node_type == "master":
            master_ip, ips = self.setup_worker_metadata(setup_timeout)
            self.setup_env_variables(master_ip, ips)
            self.send_worker_setup_msg(master_ip, ips)
            self.send_cfn_success_signal()
        elif self.node_type == "worker":
            master_ip, ips = self.wait_for_worker_setup_message(setup_timeout)
            if master_ip is None or ips is None:
                raise AgentError("FAILED worker metadata setup")
            self.setup_env_variables(master_ip, ips)
        else:
            raise AgentError(f"unknown node type: {e['AWS_DL_NODE_TYPE']}")


def _ip_key(ip: str):
    try:
        return tuple(int(x) for x in ip.split("."))
    except ValueError:
        return (999, ip)


def main(env: dict | None = None) -> int:
    env = dict(os.environ if env is None else env)
    root = env.get("DLCFN_NODE_ROOT", "/")
    setup_logging(os.path.join(root, "var/log"))
    try:
        Agent(env).run()
    except Exception as e:  # any failure -> exit 1 -> stack rollback (:429-431)
        LOGGER.exception(e)
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
