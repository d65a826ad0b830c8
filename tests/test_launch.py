"""Launchers and trackers (CPU only)."""
import os
import subprocess
import sys
import threading

import pytest

from deeplearning_cfn_b200.launch import mpirun, submitters, tf_ps_scripts, tracker
from deeplearning_cfn_b200.launch.hostfile import Host, parse_hostfile, resolve_alias, write_mpi_hostfile
from deeplearning_cfn_b200.launch.launcher import build_parser

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_hostfile_formats(tmp_path):
    p = tmp_path / "h"
    p.write_text("deeplearning-worker1\nnode2:2222\n# comment\nnode3 slots=8\n\n")
    hs = parse_hostfile(str(p))
    assert [(h.name, h.port, h.slots) for h in hs] == [("deeplearning-worker1", 22, 1), ("node2", 2222, 1), ("node3", 22, 8)]
    hosts = tmp_path / "hosts"
    hosts.write_text("127.0.0.3 deeplearning-master\n127.0.0.3 deeplearning-worker1\n127.1.0.3 deeplearning-worker2\n")
    assert resolve_alias("deeplearning-worker2", str(hosts)) == "127.1.0.3"
    assert resolve_alias("unknown-host", str(hosts)) == "unknown-host"
    write_mpi_hostfile(str(tmp_path / "mpi"), ["a", "b"], 4)
    assert (tmp_path / "mpi").read_text() == "a slots=4\nb slots=4\n"
    # IPv6 forms, and lines that cannot be read are reported as file:line instead of a bare int() traceback
    p.write_text("fe80::1 slots=2\n[fe80::2]:2200\n[::1]\n")
    assert [(h.name, h.port, h.slots) for h in parse_hostfile(str(p))] == [("fe80::1", 22, 2), ("fe80::2", 2200, 1), ("::1", 22, 1)]
    from deeplearning_cfn_b200.launch.hostfile import HostfileError
    for bad, needle in (("node1\nnode2:ssh\n", ":2:"), ("node1 slots=many\n", "slots"), ("node1 slots=0\n", "slots"), ("[::1\n", ":1:"),
                        ("# only a comment\n", "empty")):
        p.write_text(bad)
        with pytest.raises(HostfileError, match=needle):
            parse_hostfile(str(p))
    import subprocess
    import sys
    p.write_text("node1:ssh\n")
    r = subprocess.run([sys.executable, "-m", "deeplearning_cfn_b200", "launch", "-n", "1", "-H", str(p), "true"],
                       capture_output=True, text=True, env=dict(os.environ, PYTHONPATH=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    assert r.returncode == 1 and "launch: error:" in r.stderr and "Traceback" not in r.stderr


@pytest.mark.parametrize("n", [1, 2, 3, 7, 8, 16, 31])
def test_link_maps_are_consistent(n):
    tree, parent, ring = tracker.link_maps(n)
    assert sorted(tree) == list(range(n))
    # ring is 0 -> 1 -> ... -> n-1 -> 0 after relabelling
    for r in range(n):
        assert ring[r] == ((r - 1) % n, (r + 1) % n)
    # tree: symmetric, connected, n-1 edges, exactly one root
    edges = {(a, b) for a in tree for b in tree[a]}
    assert all((b, a) in edges for a, b in edges) and len(edges) == 2 * (n - 1)
    assert [r for r in parent if parent[r] == -1] == [0]
    seen, todo = {0}, [0]
    while todo:
        for x in tree[todo.pop()]:
            if x not in seen:
                seen.add(x)
                todo.append(x)
    assert len(seen) == n
    assert all(len(tree[r]) <= 3 for r in tree)


def test_allreduce_tracker_assigns_ranks_and_master():
    n = 5
    trk = tracker.AllreduceTracker("127.0.0.1", n, port=19091, port_end=19999)
    trk.start()
    results = [None] * n

    def worker(i):
        c = tracker.TrackerClient("127.0.0.1", trk.port)
        results[i] = c.start(world=n, jobid=f"job{i}")
        if i == 0:
            c.print("hello from a worker\n")

    ts = [threading.Thread(target=worker, args=(i,)) for i in range(n)]
    [t.start() for t in ts]
    [t.join(20) for t in ts]
    assert sorted(r["rank"] for r in results) == list(range(n))
    assert all(r["world"] == n for r in results)
    assert len({(r["master_addr"], r["master_port"]) for r in results}) == 1
    tree, parent, ring = tracker.link_maps(n)
    for r in results:
        assert r["parent"] == parent[r["rank"]] and sorted(r["neighbors"]) == sorted(tree[r["rank"]])
        assert (r["ring_prev"], r["ring_next"]) == ring[r["rank"]]
    # a restarted worker recovers its old rank from its job id
    again = tracker.TrackerClient("127.0.0.1", trk.port).start(world=n, jobid="job3", cmd="recover")
    assert again["rank"] == [r for i, r in enumerate(results) if i == 3][0]["rank"]
    for r in results:
        tracker.TrackerClient("127.0.0.1", trk.port).shutdown(r["rank"])
    trk.join(10)
    assert not trk.alive() and trk.messages == ["hello from a worker\n"]


def test_role_assignment_order_differs_between_ssh_and_local():
    hosts = [Host("a"), Host("b"), Host("c")]
    plan = submitters.ssh_plan(hosts, nworker=3, nserver=2)
    # ssh: servers first, hosts round-robin (DT/ssh.py:77-85)
    assert [(r, t, h.name) for r, t, h in plan] == [("server", 0, "a"), ("server", 1, "b"), ("worker", 0, "c"),
                                                     ("worker", 1, "a"), ("worker", 2, "b")]


def test_local_launcher_runs_ps_job_with_dmlc_env(tmp_path):
    out = tmp_path / "out"
    out.mkdir()
    script = tmp_path / "job.py"
    script.write_text(
        "import os\n"
        "role=os.environ['DMLC_ROLE']; tid=os.environ.get('DMLC_TASK_ID','s')\n"
        "keys=['DMLC_ROLE','DMLC_NUM_WORKER','DMLC_NUM_SERVER','DMLC_PS_ROOT_URI','DMLC_PS_ROOT_PORT']\n"
        f"open(r'{out}/'+role+tid,'w').write(' '.join(os.environ[k] for k in keys)+' '+os.environ.get('EXTRA',''))\n")
    r = subprocess.run([sys.executable, "-m", "deeplearning_cfn_b200", "launch", "-n", "3", "-s", "2", "--launcher", "local",
                        "--env-worker", "EXTRA:w", "--env-server", "EXTRA:s", sys.executable, str(script)],
                       env=dict(os.environ, PYTHONPATH=ROOT), capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    files = sorted(os.listdir(out))
    assert files == ["schedulers", "server0", "server1", "worker0", "worker1", "worker2"]
    w = (out / "worker1").read_text().split()
    assert w[0] == "worker" and w[1] == "3" and w[2] == "2" and w[3] == "127.0.0.1" and 9091 <= int(w[4]) < 9999 and w[5] == "w"
    assert (out / "server0").read_text().split()[5] == "s"
    assert (out / "schedulers").read_text().split()[0] == "scheduler"


def test_local_launcher_retries_failed_process(tmp_path):
    marker = tmp_path / "attempts"
    script = tmp_path / "flaky.py"
    script.write_text(f"import os,sys\nopen(r'{marker}','a').write(os.environ['DMLC_NUM_ATTEMPT'])\n"
                      "sys.exit(0 if os.environ['DMLC_NUM_ATTEMPT']=='2' else 3)\n")
    r = subprocess.run([sys.executable, "-m", "deeplearning_cfn_b200", "launch", "-n", "1", "-s", "0", "--launcher", "local",
                        sys.executable, str(script)], env=dict(os.environ, PYTHONPATH=ROOT, DMLC_NUM_ATTEMPT="3"),
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert marker.read_text() == "012"


def test_ssh_launcher_over_stack_hostfile(tmp_path, monkeypatch):
    """README.md:139 flow: launch -n $COUNT -H $DEEPLEARNING_WORKERS_PATH <cmd> against a provisioned stack."""
    monkeypatch.setenv("DLCFN_FAKE_GPUS", "4")
    from deeplearning_cfn_b200.cluster.stack import launch_cluster
    h = launch_cluster("job", home=str(tmp_path), backend="thread", WorkerCount=1, InstanceType="b200.2gpu")
    out = tmp_path / "out"
    out.mkdir()
    script = tmp_path / "probe.py"
    script.write_text("import os\n"
                      f"open(r'{out}/'+os.environ['DMLC_ROLE']+os.environ.get('DMLC_TASK_ID',''),'w').write("
                      "os.environ.get('CUDA_VISIBLE_DEVICES','-')+' '+os.environ.get('DMLC_NODE_HOST','-')+' '+os.environ.get('DEEPLEARNING_WORKERS_COUNT','-'))\n")
    env = dict(os.environ, PYTHONPATH=ROOT, **h.export_env())
    r = subprocess.run([sys.executable, "-m", "deeplearning_cfn_b200", "launch", "-n", h.env["DEEPLEARNING_WORKERS_COUNT"],
                        "-H", h.hostfile, sys.executable, str(script)], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    got = {f: (out / f).read_text().split() for f in os.listdir(out) if f.startswith(("worker", "server"))}
    assert sorted(got) == ["server0", "server1", "worker0", "worker1"]
    # processes landed on both nodes, each seeing its own GPU slot and address
    assert {tuple(v[:2]) for v in got.values()} == {("0,1", h.workers[0]["ip"]), ("2,3", h.workers[1]["ip"])}
    assert all(v[2] == "2" for v in got.values())


def test_launcher_cli_defaults():
    a = build_parser().parse_args(["-n", "4", "python", "train.py", "--kvstore", "dist_sync"])
    assert a.num_workers == 4 and a.num_servers is None and a.launcher == "ssh" and a.command[0] == "python"
    with pytest.raises(SystemExit):
        build_parser().parse_args(["python", "x.py"])


def test_rendered_scheduler_submissions():
    envs = {"DMLC_NUM_WORKER": 2, "DMLC_NUM_SERVER": 1, "DMLC_PS_ROOT_URI": "10.0.0.1", "DMLC_PS_ROOT_PORT": 9091}
    sge = submitters.sge_script(2, 1, envs, "python t.py")
    assert "#$ -t 1-3" in sge and "export DMLC_PS_ROOT_PORT=9091" in sge and "DMLC_ROLE=server" in sge
    k8s = submitters.kubernetes_manifests("j", 2, 1, envs, "img", "python t.py", gpus_per_worker=8)
    kinds = [(m["kind"], m["metadata"]["name"]) for m in k8s]
    assert kinds == [("Service", "j-scheduler"), ("Job", "j-scheduler"), ("Job", "j-server"), ("Job", "j-worker")]
    # the scheduler is in the cluster: every pod finds it through the Service, whose selector matches the scheduler pods' label
    assert all({"name": "DMLC_PS_ROOT_URI", "value": "j-scheduler"} in m["spec"]["template"]["spec"]["containers"][0]["env"] for m in k8s[1:])
    assert k8s[0]["spec"]["selector"] == k8s[1]["spec"]["template"]["metadata"]["labels"]
    assert k8s[-1]["spec"]["parallelism"] == 2
    assert k8s[-1]["spec"]["template"]["spec"]["containers"][0]["resources"]["limits"]["nvidia.com/gpu"] == 8
    assert submitters.yarn_argv(2, 1, envs, "cmd")[-5:] == ["-nworker", "2", "-nserver", "1", "cmd"]
    assert len(submitters.mesos_tasks(2, 1, envs, "cmd")) == 3
    s = submitters.slurm_argv(4, 2, envs, "worker", "cmd")
    assert s[:3] == ["srun", "--ntasks=4", "--nodes=2"] and "DMLC_ROLE=worker" in s[3]
    mp = submitters.mpi_argv(2, envs, "server", "cmd", "hf")
    assert mp[:5] == ["mpirun", "-n", "2", "--hostfile", "hf"] and "DMLC_ROLE=server" in mp
    assert "-env" in submitters.mpi_argv(2, envs, "server", "cmd", None, flavor="mpich")


def test_generate_trainer_golden(tmp_path):
    hf = tmp_path / "workers"
    hf.write_text("deeplearning-worker1\ndeeplearning-worker2\n")
    a = tf_ps_scripts.parse_args(["--trainer_script_dir", str(tmp_path / "s"), "--log_dir", str(tmp_path / "log"),
                                  "--workers_file_path", str(hf), "--worker_count", "2", "--worker_gpu_count", "2",
                                  "--training_script", "cifar10_multi_machine_train.py", "--batch_size", "128"])
    assert a.training_script == "cifar10_multi_machine_train.py --batch_size 128"
    paths = tf_ps_scripts.gen_scripts(a.training_script, a.workers_file_path, a.trainer_script_dir, a.worker_count,
                                      a.worker_gpu_count, a.log_dir, interpreter="python")
    assert [os.path.basename(p) for p in paths] == ["deeplearning-worker1.sh", "deeplearning-worker2.sh"]
    ps = "deeplearning-worker1:2222,deeplearning-worker2:2222"
    wk = "deeplearning-worker1:2230,deeplearning-worker1:2231,deeplearning-worker2:2230,deeplearning-worker2:2231"
    log = str(tmp_path / "log")
    want = ("source /etc/profile\n\n"
            f"CUDA_VISIBLE_DEVICES='' python cifar10_multi_machine_train.py --batch_size 128 --ps_hosts={ps} --worker_hosts={wk} "
            f"--job_name=ps --task_index=1 > {log}/ps1 2>&1 &\n\n"
            f"CUDA_VISIBLE_DEVICES='0' python cifar10_multi_machine_train.py --batch_size 128 --ps_hosts={ps} --worker_hosts={wk} "
            f"--job_name=worker --task_index=2 > {log}/worker2 2>&1 &\n\n"
            f"CUDA_VISIBLE_DEVICES='1' python cifar10_multi_machine_train.py --batch_size 128 --ps_hosts={ps} --worker_hosts={wk} "
            f"--job_name=worker --task_index=3 > {log}/worker3 2>&1 &\n\n")
    assert open(paths[1]).read() == want


def test_mpirun_parse_and_rank_map(tmp_path):
    n, hf, outdir, exports, cmd = mpirun.parse(["-np", "6", "--hostfile", "h", "--mca", "plm_rsh_no_tree_spawn", "1",
                                                 "-bind-to", "none", "-map-by", "slot", "-mca", "pml", "ob1",
                                                 "-x", "NCCL_MIN_NRINGS=8", "-x", "PATH", "--output-filename", "logs",
                                                 "python3", "train.py", "--config", "A=1"])
    assert (n, hf, outdir) == (6, "h", "logs") and exports["NCCL_MIN_NRINGS"] == "8" and "PATH" in exports
    assert cmd == ["python3", "train.py", "--config", "A=1"]
    ranks = mpirun.rank_map([Host("a", 22, 4), Host("b", 22, 4)], 6)
    assert [(h.name, l, s) for h, l, s in ranks] == [("a", 0, 4), ("a", 1, 4), ("a", 2, 4), ("a", 3, 4), ("b", 0, 2), ("b", 1, 2)]
    with pytest.raises(SystemExit):
        mpirun.rank_map([Host("a", 22, 2)], 3)
    # real run: 3 local ranks report their environment
    out = tmp_path / "o"
    out.mkdir()
    script = tmp_path / "r.py"
    script.write_text("import os\n"
                      f"open(r'{out}/'+os.environ['RANK'],'w').write(os.environ['WORLD_SIZE']+' '+os.environ['LOCAL_RANK']+' '+os.environ['FOO'])\n")
    hf2 = tmp_path / "hf"
    hf2.write_text("127.0.0.1 slots=3\n")
    rc = mpirun.main(["-np", "3", "--hostfile", str(hf2), "-x", "FOO=bar", sys.executable, str(script)])
    assert rc == 0 and sorted(os.listdir(out)) == ["0", "1", "2"] and (out / "2").read_text() == "3 2 bar"


def test_kill_matches_only_the_named_program(tmp_path):
    import getpass
    import time
    from deeplearning_cfn_b200.launch import kill
    marker = f"dlcfn_kill_target_{os.getpid()}"
    p = subprocess.Popen([sys.executable, "-c", f"import time; time.sleep(60) # {marker}"])
    q = subprocess.Popen([sys.executable, "-c", "import time; time.sleep(60) # innocent"])
    try:
        hf = tmp_path / "h"
        hf.write_text("127.0.0.1\n")
        time.sleep(0.3)
        assert kill.main([str(hf), getpass.getuser(), marker]) == 0
        assert p.wait(10) != 0
        assert q.poll() is None
    finally:
        q.kill()
        if p.poll() is None:
            p.kill()


def test_mpirun_terminates_the_job_when_a_rank_fails(tmp_path):
    """Like mpirun: the first rank that exits non-zero ends the job - a peer blocked in a collective is killed, not waited for."""
    import time
    from conftest import run_job
    script = tmp_path / "r.py"
    script.write_text("import os, sys, time\nif os.environ['RANK'] == '1':\n    sys.exit(3)\ntime.sleep(600)\n")
    t0 = time.time()
    r = run_job([sys.executable, "-m", "deeplearning_cfn_b200.launch.mpirun", "-np", "3", sys.executable, str(script)],
                env=dict(os.environ, PYTHONPATH=ROOT), timeout=120)
    assert r.returncode == 1 and time.time() - t0 < 60
    assert "rank 1 exited first with code 3" in r.stderr


def _write_exe(path, text):
    with open(path, "w") as f:
        f.write(text)
    os.chmod(path, 0o755)


@pytest.mark.parametrize("kind", ["sge", "slurm"])
def test_batch_scheduler_submissions_run_a_job_through_shim_clis(kind, tmp_path):
    """`launch --launcher sge|slurm` end to end: the rendered submission is handed to `qsub` / `srun` - shim commands on PATH
    that do what the real schedulers do with it (an SGE array job: one bash per task id with SGE_TASK_ID set; srun: --ntasks
    copies of the command with the --export environment) - and a real 2-worker / 1-server dist_sync job completes over them.
    (Reference: dmlc_tracker/sge.py, slurm.py submit the same renderings to the real clusters.)"""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    bindir = tmp_path / "bin"
    bindir.mkdir()
    _write_exe(bindir / "qsub", """#!/bin/bash
# minimal Sun Grid Engine: run every task of the array job `#$ -t 1-N` of the submitted script
script="$1"
n=$(grep -oE '^#\\$ -t 1-[0-9]+' "$script" | grep -oE '[0-9]+$')
pids=()
for i in $(seq 1 "$n"); do SGE_TASK_ID=$i bash "$script" & pids+=($!); done
rc=0; for p in "${pids[@]}"; do wait $p || rc=1; done; exit $rc
""")
    _write_exe(bindir / "srun", """#!/bin/bash
# minimal srun: --ntasks=N --nodes=M --export=ALL,K=V,... command...
n=1
while [[ "$1" == --* ]]; do
  case "$1" in
    --ntasks=*) n="${1#--ntasks=}";;
    --export=*) IFS=',' read -ra kv <<< "${1#--export=}"; for e in "${kv[@]}"; do [[ "$e" == *=* ]] && export "$e"; done;;
  esac
  shift
done
pids=()
for i in $(seq 0 $((n-1))); do SLURM_PROCID=$i "$@" & pids+=($!); done
rc=0; for p in "${pids[@]}"; do wait $p || rc=1; done; exit $rc
""")
    env = dict(os.environ, PYTHONPATH=root, PATH=f"{bindir}:{os.environ['PATH']}")
    cmd = [sys.executable, "-m", "deeplearning_cfn_b200", "launch", "-n", "2", "-s", "1", "--launcher", kind, "--render-dir", str(tmp_path),
           sys.executable, os.path.join(root, "tests/mp/dist_kvstore_worker.py"), "dist_async"]
    from conftest import run_job
    r = run_job(cmd, env=env, timeout=240, cwd=str(tmp_path))
    assert r.returncode == 0 and r.stdout.count("ASYNC_OK") == 2, r.stdout[-2000:] + r.stderr[-3000:]
    assert os.path.exists(tmp_path / ("dlcfn-sge.sh" if kind == "sge" else "dlcfn-slurm.json"))


_KUBECTL_SHIM = """#!/usr/bin/env python3
# minimal kubectl: `create -f list.json` starts every Job of the List - `parallelism` pods per Job, each the container's command
# with the container's env - and resolves Service names to this host (cluster DNS); exits when all pods have exited
import json, os, subprocess, sys
assert sys.argv[1:3] == ["create", "-f"], sys.argv
doc = json.load(open(sys.argv[3]))
assert doc["kind"] == "List"
services = {m["metadata"]["name"] for m in doc["items"] if m["kind"] == "Service"}
procs = []
for m in doc["items"]:
    if m["kind"] != "Job":
        continue
    c = m["spec"]["template"]["spec"]["containers"][0]
    env = dict(os.environ)
    for e in c["env"]:
        env[e["name"]] = "127.0.0.1" if e["value"] in services else e["value"]
    for _ in range(m["spec"]["parallelism"]):
        procs.append(subprocess.Popen(c["command"], env=env))
    print("job.batch/" + m["metadata"]["name"] + " created", flush=True)
sys.exit(1 if any(p.wait() for p in procs) else 0)
"""

_MESOS_SHIM = """#!/usr/bin/env python3
# minimal mesos-execute: --master=... --task_group=tasks.json: one process per task with its env, exit status = any failure
import json, os, subprocess, sys
opts = dict(a[2:].split("=", 1) for a in sys.argv[1:] if a.startswith("--") and "=" in a)
tasks = json.load(open(opts["task_group"]))
procs = [subprocess.Popen(["bash", "-c", t["cmd"]], env=dict(os.environ, **t["env"])) for t in tasks]
sys.exit(1 if any(p.wait() for p in procs) else 0)
"""

_YARN_SHIM = """#!/usr/bin/env python3
# minimal `yarn jar dmlc-yarn.jar org.apache.hadoop.yarn.dmlc.Client ...`: what the application master does with the request -
# -nserver server containers and -nworker worker containers, each the command with the -env pairs + DMLC_ROLE / DMLC_TASK_ID
import os, subprocess, sys
a = sys.argv[1:]
assert a[:3] == ["jar", "dmlc-yarn.jar", "org.apache.hadoop.yarn.dmlc.Client"], a
a, env, n = a[3:], dict(os.environ), {}
valued = {"-jobname", "-queue", "-worker_cores", "-worker_memory_mb", "-server_cores", "-server_memory_mb"}
i = 0
while i < len(a) - 1:
    if a[i] == "-env":
        k, v = a[i + 1].split("=", 1)
        env[k] = v
    elif a[i] in ("-nworker", "-nserver"):
        n[a[i][2:]] = int(a[i + 1])
    else:
        assert a[i] in valued, a[i]
    i += 2
procs = [subprocess.Popen(["bash", "-c", a[-1]], env=dict(env, DMLC_ROLE=role, DMLC_TASK_ID=str(t)))
         for role in ("server", "worker") for t in range(n[role])]
sys.exit(1 if any(p.wait() for p in procs) else 0)
"""


@pytest.mark.parametrize("kind", ["yarn", "mesos", "kubernetes"])
def test_yarn_mesos_kubernetes_submissions_run_a_job_through_shim_clis(kind, tmp_path):
    """`launch --launcher yarn|mesos|kubernetes` end to end: the submission goes to `yarn` / `mesos-execute` / `kubectl` - shim
    commands on PATH that do with it what the real scheduler would (containers per role with the requested environment; for
    kubernetes the scheduler itself is one of the Jobs and is found through the Service name) - and a real 2-worker /
    1-server job completes over them.  (Reference: dmlc_tracker/yarn.py, mesos.py, kubernetes.py.)"""
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    bindir = tmp_path / "bin"
    bindir.mkdir()
    shim = {"yarn": ("yarn", _YARN_SHIM), "mesos": ("mesos-execute", _MESOS_SHIM), "kubernetes": ("kubectl", _KUBECTL_SHIM)}[kind]
    _write_exe(bindir / shim[0], shim[1].replace("#!/usr/bin/env python3", "#!" + sys.executable))
    import socket
    with socket.socket() as sk:                      # the scheduler Service's port (the shim cluster shares this host's ports)
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    env = dict(os.environ, PYTHONPATH=root, PATH=f"{bindir}:{os.environ['PATH']}", DLCFN_KUBE_SCHEDULER_PORT=str(port))
    cmd = [sys.executable, "-m", "deeplearning_cfn_b200", "launch", "-n", "2", "-s", "1", "--launcher", kind, "--render-dir", str(tmp_path),
           sys.executable, os.path.join(root, "tests/mp/dist_kvstore_worker.py"), "dist_async"]
    from conftest import run_job
    r = run_job(cmd, env=env, timeout=240, cwd=str(tmp_path))
    assert r.returncode == 0 and r.stdout.count("ASYNC_OK") == 2, r.stdout[-2000:] + r.stderr[-3000:]
    assert os.path.exists(tmp_path / {"yarn": "dlcfn-yarn.json", "mesos": "dlcfn-mesos.json", "kubernetes": "dlcfn-k8s.json"}[kind])


def test_cli_answers_bad_arguments_with_messages_not_tracebacks(tmp_path):
    """Every user-facing command on arguments a user can plausibly get wrong: a message and a non-zero exit, never a Python
    traceback (the list grew out of a sweep over all commands; the three that used to crash are first)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PYTHONPATH=root, DLCFN_HOME=str(tmp_path / "home"), DLCFN_FAKE_GPUS="2")
    cases = ["benchmark --networks bogus", "benchmark --networks native:resnet-50:32", "mpirun -np x true", "mpirun -np",
             "launch -n 2 -s 1", "launch -n 1 -H /nonexistent true", "kill /nonexistent", "generate-trainer --workers_file_path /nonexistent",
             "run-example nosuch", "run-example resnet50-ps --stack-name nosuch", "im2rec /nonexistent /nonexistent", "parse-log /nonexistent",
             "rec2idx /nonexistent", "create-stack --parameters WorkerCount=abc", "create-stack --template-body /nonexistent",
             "describe-stacks --stack-name nosuch", "describe-stack-events --stack-name nosuch", "delete-stack --stack-name nosuch",
             "validate-template --template-body /nonexistent", "node-exec nosuch -- true"]
    for args in cases:
        r = subprocess.run([sys.executable, "-m", "deeplearning_cfn_b200"] + args.split(), env=env, capture_output=True, text=True,
                           stdin=subprocess.DEVNULL, timeout=120, cwd=str(tmp_path))
        assert r.returncode != 0 and "Traceback" not in r.stderr + r.stdout, (args, r.returncode, (r.stderr + r.stdout)[-400:])
