"""Control plane (CPU only): queues, scaling-event function, bootstrap agent, stack provisioner, CLI.
The reference has NO tests for this tier; these pin down the intended behaviour."""
import json
import os
import subprocess
import sys
import threading
import time

import pytest

from deeplearning_cfn_b200.cluster import events
from deeplearning_cfn_b200.cluster.agent import Agent, AgentError, ENV_KEYS
from deeplearning_cfn_b200.cluster.controlplane import ControlPlane
from deeplearning_cfn_b200.cluster.queue import FileQueue, MemoryQueue
from deeplearning_cfn_b200.cluster.spec import ClusterSpec, TemplateError, load_template
from deeplearning_cfn_b200.cluster.stack import StackError, delete_stack, describe_stack, launch_cluster, list_stacks

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(autouse=True)
def fake_gpus(monkeypatch):
    monkeypatch.setenv("DLCFN_FAKE_GPUS", "8")


# ----------------------------------------------------------------------------- queues
@pytest.mark.parametrize("mk", [lambda p: MemoryQueue("q"), lambda p: FileQueue(str(p / "q"))])
def test_queue_sqs_semantics(tmp_path, mk):
    q = mk(tmp_path)
    q.send({"a": 1})
    q.send({"a": 2})
    got = q.receive(10, visibility_timeout=0.2)
    assert [m.json()["a"] for m in got] == [1, 2]
    assert q.receive(10, visibility_timeout=0.2) == []          # hidden while the visibility timeout runs
    time.sleep(0.25)
    again = q.receive(10, visibility_timeout=0.2)
    assert len(again) == 2                                        # not deleted -> delivered again (at-least-once)
    q.delete(again[0])
    time.sleep(0.25)
    assert [m.json()["a"] for m in q.receive(10, 0.0)] == [2]
    # visibility_timeout=0 + no delete: every reader sees the message (the worker-queue broadcast trick)
    assert [m.json()["a"] for m in q.receive(10, 0.0)] == [2]
    assert [m.json()["a"] for m in q.receive(10, 0.0)] == [2]


def test_queue_duplicate_injection(tmp_path):
    q = FileQueue(str(tmp_path / "d"), duplicate_every=2)
    q.send("x")
    q.send("y")
    assert sorted(m.body for m in q.receive(10, 0)) == ["x", "y", "y"]


def _q_producer(path, pid, n):
    from deeplearning_cfn_b200.cluster.queue import FileQueue as FQ
    q = FQ(path)
    for i in range(n):
        q.send({"p": pid, "i": i})


def _q_consumer(path, out, stop_at):
    from deeplearning_cfn_b200.cluster.queue import FileQueue as FQ
    q, got, idle = FQ(path), [], 0
    while time.time() < stop_at and idle < 40:
        msgs = q.receive(5, visibility_timeout=30.0)
        if not msgs:
            idle += 1
            time.sleep(0.02)
            continue
        idle = 0
        for m in msgs:
            got.append([m.json()["p"], m.json()["i"]])
            q.delete(m)
    with open(out, "w") as f:
        json.dump(got, f)


def test_file_queue_under_concurrent_producers_and_consumers(tmp_path):
    """The file-backed queue (the SQS stand-in of the control plane) with 3 producer and 3 consumer PROCESSES: every message
    is delivered, none twice while its visibility timeout runs, and a deleted message is gone."""
    import collections
    import multiprocessing as mp
    path = str(tmp_path / "q")
    FileQueue(path)
    P, N, C = 3, 60, 3
    stop_at = time.time() + 60
    ctx = mp.get_context("spawn")
    outs = [str(tmp_path / f"c{c}.json") for c in range(C)]
    procs = [ctx.Process(target=_q_producer, args=(path, p, N)) for p in range(P)]
    procs += [ctx.Process(target=_q_consumer, args=(path, outs[c], stop_at)) for c in range(C)]
    [p.start() for p in procs]
    [p.join(120) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    seen = collections.Counter()
    for o in outs:
        for p, i in json.load(open(o)):
            seen[(p, i)] += 1
    assert sum(seen.values()) == P * N and set(seen.values()) == {1}, (sum(seen.values()), [k for k, v in seen.items() if v != 1][:5])
    assert len(FileQueue(path)) == 0


# ----------------------------------------------------------------------------- scaling-event function
def _cp_with_groups(tmp_path, desired=2, minsize=0):
    cp = ControlPlane(str(tmp_path / "stack"))
    cp.create_group("s-MasterAutoScalingGroup-X", 1, 1, 1)
    cp.create_group("s-WorkerAutoScalingGroup-Y", minsize, desired, desired)
    os.environ["AWS_DL_STACK_ID"] = cp.dir
    os.environ["AWS_DL_MASTER_SQS_URL"] = "MasterQueue"
    return cp


def test_lambda_launch_complete_and_signal(tmp_path):
    cp = _cp_with_groups(tmp_path)
    inst = cp.add_instance("s-MasterAutoScalingGroup-X", "127.0.0.2", [0], lifecycle="InService")
    events.lambda_handler(events.make_notification(events.LAUNCH, "s-MasterAutoScalingGroup-X", inst["InstanceId"]))
    msgs = cp.queue("MasterQueue").receive(10, 0)
    assert len(msgs) == 1
    body = msgs[0].json()
    assert body == {"min": 1, "desired": 1, "max": 1, "launched": 1, "status": "success",
                    "asg": "s-MasterAutoScalingGroup-X", "event": "asg-setup"}            # SURVEY Appendix D schema
    assert [s["status"] for s in cp.signals("MasterAutoScalingGroup")] == ["SUCCESS"]
    # workers: success only when launched == desired, then ReplaceUnhealthy is suspended
    w1 = cp.add_instance("s-WorkerAutoScalingGroup-Y", "127.1.0.2", [1])
    events.lambda_handler(events.make_notification(events.LAUNCH, "s-WorkerAutoScalingGroup-Y", w1["InstanceId"]))
    assert len(cp.queue("MasterQueue").receive(10, 0)) == 1
    w2 = cp.add_instance("s-WorkerAutoScalingGroup-Y", "127.1.0.3", [2])
    events.lambda_handler(events.make_notification(events.LAUNCH, "s-WorkerAutoScalingGroup-Y", w2["InstanceId"]))
    assert len(cp.queue("MasterQueue").receive(10, 0)) == 2
    assert cp.describe_group("s-WorkerAutoScalingGroup-Y")["SuspendedProcesses"] == ["ReplaceUnhealthy"]


def test_lambda_launch_error_shrinks_to_launched(tmp_path):
    cp = _cp_with_groups(tmp_path, desired=4, minsize=0)
    for k in range(2):
        cp.add_instance("s-WorkerAutoScalingGroup-Y", f"127.1.0.{k + 2}", [k])
    events.lambda_handler(events.make_notification(events.LAUNCH_ERROR, "s-WorkerAutoScalingGroup-Y", "",
                                                   status_code="Failed", status_message="InsufficientInstanceCapacity"))
    g = cp.describe_group("s-WorkerAutoScalingGroup-Y")
    assert g["DesiredCapacity"] == 2 and "ReplaceUnhealthy" in g["SuspendedProcesses"]
    assert cp.queue("MasterQueue").receive(10, 0)[0].json()["launched"] == 2


def test_lambda_launch_error_below_min_does_nothing(tmp_path):
    cp = _cp_with_groups(tmp_path, desired=4, minsize=3)
    cp.add_instance("s-WorkerAutoScalingGroup-Y", "127.1.0.2", [0])
    events.lambda_handler(events.make_notification(events.LAUNCH_ERROR, "s-WorkerAutoScalingGroup-Y", ""))
    assert cp.describe_group("s-WorkerAutoScalingGroup-Y")["DesiredCapacity"] == 4
    assert cp.queue("MasterQueue").receive(10, 0) == []


def test_lambda_other_events_are_harmless(tmp_path):
    _cp_with_groups(tmp_path)
    for ev in (events.TERMINATE, events.TERMINATE_ERROR, events.TEST, "autoscaling:SOMETHING_NEW"):
        events.lambda_handler(events.make_notification(ev, "s-WorkerAutoScalingGroup-Y", "i-1"))


# ----------------------------------------------------------------------------- agent
def _agent_env(cp, node_type, root, gpus="0,1"):
    env = {"AWS_DL_NODE_TYPE": node_type, "AWS_DL_MASTER_QUEUE": "MasterQueue", "AWS_DL_WORKER_QUEUE": "WorkerQueue",
           "AWS_DL_WAITCONDITION_TIMEOUT": "3", "AWS_DL_MASTERLAUNCH_TIMEOUT": "1", "AWS_DL_STACK_ID": cp.dir,
           "AWS_DL_WAIT_HANDLE": "WaitHandle", "AWS_DL_ROLE_NAME": "role", "AWS_DL_DEFAULT_USER": "root",
           "AWS_REGION": "b200-box-1", "EFS_MOUNT": "/myEFSvolume", "CFN_PATH": "/usr/local/bin",
           "DLCFN_NODE_ROOT": str(root), "DLCFN_INSTANCE_ID": "i-test", "DLCFN_NODE_GPUS": gpus}
    assert set(ENV_KEYS) <= set(env)
    return env


def test_agent_requires_full_env_contract(tmp_path):
    cp = ControlPlane(str(tmp_path / "s"))
    env = _agent_env(cp, "Master", tmp_path / "n")
    del env["AWS_DL_WAIT_HANDLE"]
    with pytest.raises(AgentError):
        Agent(env)


def test_master_dedupes_rank_orders_and_broadcasts(tmp_path):
    cp = ControlPlane(str(tmp_path / "s"))
    cp.create_group("x-MasterAutoScalingGroup-1", 1, 1, 1)
    cp.create_group("x-WorkerAutoScalingGroup-1", 0, 2, 2)
    m = cp.add_instance("x-MasterAutoScalingGroup-1", "10.0.0.12", [0, 1], "InService")
    w1 = cp.add_instance("x-WorkerAutoScalingGroup-1", "10.0.1.9", [2, 3], "InService")
    w2 = cp.add_instance("x-WorkerAutoScalingGroup-1", "10.0.1.5", [4, 5], "InService")
    for g, i in (("x-MasterAutoScalingGroup-1", m), ("x-WorkerAutoScalingGroup-1", w1), ("x-WorkerAutoScalingGroup-1", w2)):
        cp.set_instance(g, i["InstanceId"], State="running")
    q = cp.queue("MasterQueue")
    ok = {"min": 1, "desired": 1, "max": 1, "launched": 1, "status": "success", "event": "asg-setup"}
    q.send(dict(ok, asg="x-MasterAutoScalingGroup-1"))
    q.send(dict(ok, asg="x-MasterAutoScalingGroup-1"))           # duplicate delivery (at-least-once)
    q.send("not json at all")
    q.send(dict(ok, asg="x-WorkerAutoScalingGroup-1", launched=2, desired=2, max=2, min=0))
    a = Agent(_agent_env(cp, "Master", tmp_path / "node"), poll_interval=0.01)
    a.run()
    root = tmp_path / "node"
    # master is also a worker; rank order = sorted address (dl_cfn_setup.py:328-340)
    assert (root / "opt/deeplearning/workers").read_text() == "deeplearning-worker1\ndeeplearning-worker2\ndeeplearning-worker3\n"
    assert (root / "etc/hosts").read_text() == ("10.0.0.12 deeplearning-master\n10.0.0.12 deeplearning-worker1\n"
                                                 "10.0.1.5 deeplearning-worker2\n10.0.1.9 deeplearning-worker3\n")
    prof = (root / "etc/profile.d/deeplearning.sh").read_text()
    assert "export DEEPLEARNING_WORKERS_COUNT=3\n" in prof and "export DEEPLEARNING_WORKER_GPU_COUNT=2\n" in prof
    assert f"export DEEPLEARNING_WORKERS_PATH={root}/opt/deeplearning/workers\n" in prof and "export EFS_MOUNT=/myEFSvolume\n" in prof
    assert len(q) <= 1                                             # processed messages were deleted (bad one may remain)
    wmsg = cp.queue("WorkerQueue").receive(10, 0)[0].json()
    assert wmsg == {"event": "worker-setup", "master-ip": "10.0.0.12", "worker-ips": ["10.0.0.12", "10.0.1.5", "10.0.1.9"]}
    assert cp.signals("WaitHandle")[0]["status"] == "SUCCESS"
    # every worker reads the same never-deleted message and writes identical files
    for k in range(2):
        wa = Agent(_agent_env(cp, "Worker", tmp_path / f"w{k}"), poll_interval=0.01)
        wa.run()
        assert (tmp_path / f"w{k}/opt/deeplearning/workers").read_text() == (root / "opt/deeplearning/workers").read_text()
    assert len(cp.queue("WorkerQueue")) == 1


def test_master_times_out_without_asg_messages(tmp_path):
    cp = ControlPlane(str(tmp_path / "s"))
    a = Agent(_agent_env(cp, "Master", tmp_path / "n"), poll_interval=0.05)
    t0 = time.time()
    with pytest.raises(AgentError):
        a.run()
    assert time.time() - t0 < 5


def test_worker_times_out_and_exit_code(tmp_path):
    cp = ControlPlane(str(tmp_path / "s"))
    from deeplearning_cfn_b200.cluster import agent as agent_mod
    env = _agent_env(cp, "Worker", tmp_path / "n")
    env["DLCFN_POLL_INTERVAL"] = "0.05"
    r = subprocess.run([sys.executable, "-m", "deeplearning_cfn_b200.cluster.agent"], env=dict(os.environ, **env, PYTHONPATH=ROOT))
    assert r.returncode == 1                                       # any failure -> exit 1 -> rollback
    assert "did not receive worker-setup" in (tmp_path / "n/var/log/dl_cfn_setup.log").read_text()
    assert agent_mod.main(dict(env, AWS_DL_NODE_TYPE="Gateway")) == 1


# ----------------------------------------------------------------------------- templates
def test_template_parameter_validation():
    with pytest.raises(TemplateError):
        ClusterSpec.from_template("s", WorkerCount=0)
    with pytest.raises(TemplateError):
        ClusterSpec.from_template("s", InstanceType="p3.16xlarge")
    with pytest.raises(TemplateError):
        ClusterSpec.from_template("s", SSHLocation="everywhere")
    with pytest.raises(TemplateError):
        ClusterSpec.from_template("s", EFSFileSystemId="efs-123")
    with pytest.raises(TemplateError):
        ClusterSpec.from_template("s", Bogus="1")
    with pytest.raises(TemplateError):
        ClusterSpec.from_template("9bad")
    s = ClusterSpec.from_template("ok", WorkerCount=3, InstanceType="b200.2gpu")
    assert s.worker_count == 3 and s.gpus_per_instance == 2 and s.wait_timeout == 3300 and s.master_launch_timeout == 600
    m = ClusterSpec.from_template("mr", "mask-rcnn.template", S3Bucket="/tmp/b")
    assert m.wait_timeout == 3600 and m.master_launch_timeout == 1200 and m.extras["EFSServesData"] == "true"
    for name in ("deeplearning.template", "mask-rcnn.template", "private-mask-rcnn.template"):
        t = load_template(name)
        assert {"Parameters", "Mappings", "Resources", "Outputs"} <= set(t)
        assert set(t["Outputs"]) == {"AdminSSHSecurityGroup", "MasterAutoScalingGroup", "WorkerAutoScalingGroup", "MountTargetID"}
    assert "MyVpcId" in load_template("private-mask-rcnn.template")["Parameters"]


def test_validate_template_reports_every_dangling_reference(tmp_path):
    """`validate-template` (aws cloudformation validate-template): the three shipped templates pass and print their parameter
    summary; a template with dangling Ref / GetAtt / FindInMap / DependsOn / an unsupported intrinsic / a Default outside its
    AllowedValues is rejected with ALL problems listed, by the API, by create-stack and by the CLI."""
    import copy
    from deeplearning_cfn_b200.cluster.spec import validate_template
    for name in ("deeplearning.template", "mask-rcnn.template", "private-mask-rcnn.template"):
        summary = validate_template(name)
        keys = [p["ParameterKey"] for p in summary["Parameters"]]
        assert "WorkerCount" in keys and summary["Description"] and summary["Capabilities"] == ["CAPABILITY_NAMED_IAM"]
    t = copy.deepcopy(load_template("deeplearning.template"))
    t["Resources"]["MasterQueue"]["Properties"] = {"Name": {"Ref": "NoSuchParameter"}, "Arn": {"Fn::GetAtt": ["Ghost", "Arn"]},
                                                   "Sub": {"Fn::Sub": "${x}"}}
    t["Resources"]["WorkerQueue"]["DependsOn"] = "AlsoMissing"
    t["Outputs"]["MountTargetID"]["Value"] = {"Fn::FindInMap": ["NoSuchMap", "a", "b"]}
    t["Parameters"]["InstanceType"]["Default"] = "p3.16xlarge"
    t["Surprise"] = {}
    with pytest.raises(TemplateError) as e:
        validate_template(t)
    msg = str(e.value)
    for needle in ("NoSuchParameter", "Ghost", "Fn::Sub is not supported", "AlsoMissing", "NoSuchMap", "InstanceType: its Default",
                   "unknown top-level section 'Surprise'"):
        assert needle in msg, (needle, msg)
    with pytest.raises(TemplateError):
        ClusterSpec.from_template("s", t)                       # create-stack validates before it creates anything
    with pytest.raises(TemplateError):
        validate_template({"Parameters": {}})                   # no Resources
    bad = tmp_path / "bad.template"
    bad.write_text(json.dumps(t))
    base = [sys.executable, "-m", "deeplearning_cfn_b200"]
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run(base + ["validate-template", "--template-body", f"file://{bad}"], env=env, capture_output=True, text=True)
    assert r.returncode == 1 and "Ghost" in r.stderr
    r = subprocess.run(base + ["validate-template"], env=env, capture_output=True, text=True)
    assert r.returncode == 0 and json.loads(r.stdout)["Parameters"][0]["ParameterKey"]
    r = subprocess.run(base + ["--help"], env=env, capture_output=True, text=True)
    assert "run-example" in r.stdout and "validate-template" in r.stdout


# ----------------------------------------------------------------------------- launch_cluster
@pytest.mark.parametrize("backend", ["thread", "process"])
def test_launch_cluster_end_to_end(tmp_path, backend):
    h = launch_cluster("t1", home=str(tmp_path), backend=backend, WorkerCount=3, InstanceType="b200.2gpu")
    assert h.status == "CREATE_COMPLETE"
    assert open(h.hostfile).read().split() == [f"deeplearning-worker{i}" for i in range(1, 5)]
    assert h.env["DEEPLEARNING_WORKERS_COUNT"] == "4" and h.env["DEEPLEARNING_WORKER_GPU_COUNT"] == "2"
    assert h.env["DEEPLEARNING_WORKERS_PATH"] == h.hostfile
    assert os.path.isdir(h.env["EFS_MOUNT"])
    gpus = [tuple(w["gpus"]) for w in h.workers]
    assert sorted(gpus) == [(0, 1), (2, 3), (4, 5), (6, 7)]
    assert h.workers[0]["ip"].startswith("127.0.") and all(w["ip"].startswith("127.1.") for w in h.workers[1:])
    # every node holds identical metadata
    cp = ControlPlane(h.stack_dir)
    for g in cp.group_names():
        for i in cp.describe_group(g)["Instances"]:
            assert open(os.path.join(cp.node_root(i["InstanceId"]), "opt/deeplearning/workers")).read() == open(h.hostfile).read()
    assert set(h.outputs) == {"AdminSSHSecurityGroup", "MasterAutoScalingGroup", "WorkerAutoScalingGroup", "MountTargetID"}
    assert [s["StackName"] for s in list_stacks(str(tmp_path))] == ["t1"]
    with pytest.raises(StackError):
        launch_cluster("t1", home=str(tmp_path), backend=backend)


def test_elastic_shrink_and_store_retention(tmp_path):
    # 8 GPUs, 2 per node -> master + 3 workers fit, 4 more do not: the function shrinks the group (README.md:49)
    h = launch_cluster("big", home=str(tmp_path), backend="thread", WorkerCount=7, InstanceType="b200.2gpu")
    assert h.status == "CREATE_COMPLETE" and h.env["DEEPLEARNING_WORKERS_COUNT"] == "4"
    g = ControlPlane(h.stack_dir).describe_group("WorkerAutoScalingGroup")
    assert g["DesiredCapacity"] == 3 and "ReplaceUnhealthy" in g["SuspendedProcesses"]
    # fault injection: every worker launch fails -> the master alone is the cluster
    # (a different box: stacks of one home share its GPUs, and "big" holds all eight - see the next test)
    h2 = launch_cluster("solo", home=str(tmp_path / "box2"), backend="thread", WorkerCount=2, fail_instances=2, capacity_gpus=1)
    assert h2.env["DEEPLEARNING_WORKERS_COUNT"] == "1"
    # shared store survives deletion and can be re-attached (DeletionPolicy Retain)
    store = h.outputs["MountTargetID"]
    open(os.path.join(h.env["EFS_MOUNT"], "ckpt.bin"), "w").write("weights")
    delete_stack("big", str(tmp_path))
    assert describe_stack("big", str(tmp_path)).status == "DELETE_COMPLETE"
    h3 = launch_cluster("big", home=str(tmp_path), backend="thread", WorkerCount=1, EFSFileSystemId=store)
    assert open(os.path.join(h3.env["EFS_MOUNT"], "ckpt.bin")).read() == "weights"


def test_stacks_of_one_box_do_not_share_gpus(tmp_path):
    """Two live stacks in one home: the second only gets the GPUs the first left (and shrinks to them), a third finds none
    and fails at once with InsufficientInstanceCapacity instead of waiting for the master timeout; deleting a stack frees
    its GPUs."""
    import time
    home = str(tmp_path)
    a = launch_cluster("a", home=home, backend="thread", WorkerCount=2, InstanceType="b200.2gpu")          # 6 of 8 GPUs
    assert a.status == "CREATE_COMPLETE" and a.env["DEEPLEARNING_WORKERS_COUNT"] == "3"
    b = launch_cluster("b", home=home, backend="thread", WorkerCount=2, InstanceType="b200.2gpu")          # 2 left: master only
    assert b.status == "CREATE_COMPLETE" and b.env["DEEPLEARNING_WORKERS_COUNT"] == "1"

    def gpus(name):
        cp = ControlPlane(os.path.join(home, "stacks", name))
        return {g for gn in cp.group_names() for i in cp.describe_group(gn)["Instances"] for g in i["Gpus"]}
    assert len(gpus("a")) == 6 and len(gpus("b")) == 2 and not (gpus("a") & gpus("b"))
    t0 = time.time()
    with pytest.raises(StackError, match="InsufficientInstanceCapacity"):
        launch_cluster("c", home=home, backend="thread", WorkerCount=1, InstanceType="b200.1gpu")
    assert time.time() - t0 < 30
    assert describe_stack("c", home).status == "ROLLBACK_COMPLETE"
    delete_stack("a", home)
    d = launch_cluster("d", home=home, backend="thread", WorkerCount=1, InstanceType="b200.2gpu")
    assert d.status == "CREATE_COMPLETE" and d.env["DEEPLEARNING_WORKERS_COUNT"] == "2" and gpus("d") <= {0, 1, 2, 3, 4, 5, 6, 7} - gpus("b")


def test_rollback_when_master_cannot_launch(tmp_path):
    with pytest.raises(StackError):
        launch_cluster("nogpu", home=str(tmp_path), backend="thread", WorkerCount=1, InstanceType="b200.8gpu",
                       capacity_gpus=4, wait_timeout=1.0)
    assert describe_stack("nogpu", str(tmp_path)).status == "ROLLBACK_COMPLETE"


def test_cpu_instances_need_no_gpus(tmp_path, monkeypatch):
    monkeypatch.setenv("DLCFN_FAKE_GPUS", "0")
    h = launch_cluster("cpu", home=str(tmp_path), backend="thread", WorkerCount=2, InstanceType="cpu.large", ImageType="CPUGloo")
    assert h.env["DEEPLEARNING_WORKERS_COUNT"] == "3" and h.env["DEEPLEARNING_WORKER_GPU_COUNT"] == "0"


def test_mask_rcnn_stack_staging(tmp_path):
    from deeplearning_cfn_b200.cluster.staging import data_dir, prepare_bucket
    bucket = str(tmp_path / "bucket")
    out = prepare_bucket(bucket, num_images=8, image_size=32)
    assert all(os.path.exists(p) for p in out.values())
    h = launch_cluster("mr", "mask-rcnn.template", home=str(tmp_path), backend="thread", WorkerCount=1,
                       InstanceType="b200.4gpu", S3Bucket=bucket, EFSServesData="true")
    efs = h.env["EFS_MOUNT"]
    assert os.path.exists(os.path.join(efs, "data.txt")) and os.path.isdir(os.path.join(efs, "data", "coco"))
    home = os.path.join(h.master_root, "home")
    assert os.access(os.path.join(home, "run.sh"), os.X_OK) and os.path.exists(os.path.join(home, "setup.sh"))
    assert data_dir(efs, home) == os.path.join(efs, "data")
    delete_stack("mr", str(tmp_path))          # it held all 8 GPUs of the box
    h2 = launch_cluster("mr2", "private-mask-rcnn.template", home=str(tmp_path), backend="thread", WorkerCount=1,
                        InstanceType="b200.4gpu", S3Bucket=bucket, EFSServesData="false")
    assert os.path.exists(os.path.join(h2.env["EFS_MOUNT"], "nodata.txt"))
    assert all(w["ip"].startswith("127.1.") for w in h2.workers)     # private variant: master has no public address
    assert os.path.isdir(os.path.join(h2.master_root, "home", "data", "coco"))


def test_cli_create_list_describe_delete(tmp_path):
    env = dict(os.environ, PYTHONPATH=ROOT, DLCFN_FAKE_GPUS="8")
    base = [sys.executable, "-m", "deeplearning_cfn_b200"]
    r = subprocess.run(base + ["create-stack", "--stack-name", "c1", "--home", str(tmp_path), "--parameters",
                               "ParameterKey=WorkerCount,ParameterValue=1", "InstanceType=b200.4gpu"],
                       env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = json.loads(r.stdout)
    assert out["StackStatus"] == "CREATE_COMPLETE" and out["Env"]["DEEPLEARNING_WORKERS_COUNT"] == "2"
    r = subprocess.run(base + ["list-stacks", "--home", str(tmp_path), "--stack-status-filter", "CREATE_COMPLETE"],
                       env=env, capture_output=True, text=True)
    assert "c1" in r.stdout
    r = subprocess.run(base + ["describe-stacks", "--stack-name", "c1", "--home", str(tmp_path)], env=env, capture_output=True, text=True)
    assert json.loads(r.stdout)["Hostfile"].endswith("opt/deeplearning/workers")
    r = subprocess.run(base + ["describe-stack-events", "--stack-name", "c1", "--home", str(tmp_path)], env=env, capture_output=True, text=True)
    evs = json.loads(r.stdout)["StackEvents"]
    assert evs[0]["LogicalResourceId"] == "c1" and evs[0]["ResourceStatus"] == "CREATE_COMPLETE"        # newest first
    assert {"SharedStore", "MasterAutoScalingGroup", "WorkerAutoScalingGroup"} <= {e["LogicalResourceId"] for e in evs}
    assert subprocess.run(base + ["delete-stack", "--stack-name", "c1", "--home", str(tmp_path)], env=env).returncode == 0
    r = subprocess.run(base + ["create-stack", "--stack-name", "c2", "--home", str(tmp_path), "--parameters", "WorkerCount=0"],
                       env=env, capture_output=True, text=True)
    assert r.returncode == 1 and "MinValue" in r.stderr


def test_packaged_event_function_is_self_contained(tmp_path):
    """The scaling-event function ships as an archive (reference dl_cfn_setup_lambda.zip + template.yaml): importable
    on its own under the reference's handler name, without the rest of the framework."""
    import json
    import subprocess
    import sys
    import zipfile
    from deeplearning_cfn_b200.cluster.staging import package_event_function
    out = package_event_function(str(tmp_path / "dl_cfn_setup_lambda.zip"))
    names = set(zipfile.ZipFile(out).namelist())
    assert {"lambda_function.py", "template.json", "deeplearning_cfn_b200/cluster/events.py"} <= names
    tpl = json.loads(zipfile.ZipFile(out).read("template.json"))
    assert tpl["Resources"]["EventFunction"]["Properties"]["Handler"] == "lambda_function.lambda_handler"
    code = ("import sys, json; sys.path.insert(0, %r); import lambda_function as lf; "
            "ev = {'Records': [{'Sns': {'Message': json.dumps({'Event': 'autoscaling:TEST_NOTIFICATION'})}}]}; "
            "m = lf.lambda_handler(ev, None); assert m['Event'].endswith('TEST_NOTIFICATION'); "
            "assert 'torch' not in sys.modules; print('ZIP_OK', lf.__file__)") % out
    r = subprocess.run([sys.executable, "-S", "-c", code], capture_output=True, text=True, cwd=str(tmp_path),
                       env={"PATH": os.environ.get("PATH", ""), "AWS_DL_STACK_ID": str(tmp_path)}, timeout=60)
    assert "ZIP_OK" in r.stdout, r.stdout + r.stderr
