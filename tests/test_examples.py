"""The bundled example recipes run end to end on CPU (small shapes): run_example equivalents of the reference
README flows."""
import json
import os
import subprocess
import sys
import time

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ENV = dict(os.environ, PYTHONPATH=ROOT, DLCFN_FAKE_GPUS="0", CUDA_VISIBLE_DEVICES="")
PY = sys.executable


def run(cmd, env=None, timeout=600):
    from conftest import run_job
    return run_job(cmd, env=dict(ENV, **(env or {})), timeout=timeout)


def test_readme_flow_mxnet_style_ps_over_stack(tmp_path):
    """create-stack -> launch -n $COUNT -H $PATH image_classification --kvstore dist_device_sync (README.md:139)."""
    from deeplearning_cfn_b200.cluster.stack import launch_cluster
    os.environ["DLCFN_FAKE_GPUS"] = "0"
    h = launch_cluster("ex", home=str(tmp_path), backend="thread", WorkerCount=1, InstanceType="cpu.large", ImageType="CPUGloo")
    env = h.export_env()
    r = run([PY, "-m", "deeplearning_cfn_b200", "launch", "-n", env["DEEPLEARNING_WORKERS_COUNT"], "-H",
             env["DEEPLEARNING_WORKERS_PATH"], PY, "-m", "deeplearning_cfn_b200.examples.image_classification",
             "--dataset", "cifar10", "--model", "resnet20", "--epochs", "1", "--kvstore", "dist_device_sync",
             "--batch-size", "16", "--num-samples", "64", "--log-interval", "1"], env=env)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-3000:]
    out = r.stdout + r.stderr
    assert "Node[0] Epoch[0]" in out and "Node[1] Epoch[0]" in out and out.count("validation: accuracy") == 2


@pytest.mark.parametrize("extra", [[], ["--sync_replicas"]], ids=["async", "sync_replicas"])
def test_tf_style_async_ps_from_generated_scripts(tmp_path, extra):
    """generate-trainer -> one script per host -> ps + workers train (asynchronously, or with the SyncReplicas token
    protocol), chief checkpoints weights + moving averages."""
    from deeplearning_cfn_b200.cluster.stack import launch_cluster
    h = launch_cluster("tf", home=str(tmp_path), backend="thread", WorkerCount=1, InstanceType="cpu.large", ImageType="CPUGloo")
    env = h.export_env()
    sdir, ldir, tdir = str(tmp_path / "scripts"), str(tmp_path / "logs"), str(tmp_path / "train")
    r = run([PY, "-m", "deeplearning_cfn_b200", "generate-trainer", "--trainer_script_dir", sdir, "--log_dir", ldir,
             "--workers_file_path", h.hostfile, "--worker_count", "2", "--worker_gpu_count", "1", "--training_script",
             os.path.join(ROOT, "deeplearning_cfn_b200/examples/cifar10_multi_machine_train.py"), "--max_steps", "6", "--batch_size", "8",
             "--train_dir", tdir, "--log_frequency", "2"] + extra, env=env)
    assert r.returncode == 0, r.stderr
    scripts = sorted(os.listdir(sdir))
    assert scripts == ["deeplearning-worker1.sh", "deeplearning-worker2.sh"]
    procs = []
    for s in scripts:   # what the user does by hand in the reference: run each host's script on that host
        text = open(os.path.join(sdir, s)).read().replace("source /etc/profile\n", "")
        open(os.path.join(sdir, s), "w").write(text + "\nwait\n")
        procs.append(subprocess.Popen(["bash", os.path.join(sdir, s)], env=dict(ENV, **env, DLCFN_PY=PY)))
    deadline = time.time() + 300
    for p in procs:
        assert p.wait(max(1, deadline - time.time())) == 0
    logs = {f: open(os.path.join(ldir, f)).read() for f in os.listdir(ldir)}
    assert set(logs) == {"ps0", "ps1", "worker0", "worker1"}
    assert "step 6, loss" in logs["worker0"] and "examples/sec" in logs["worker1"], logs["worker0"][-800:]
    assert os.path.exists(os.path.join(tdir, "model.ckpt"))
    import torch
    sd = torch.load(os.path.join(tdir, "model.ckpt"), map_location="cpu")
    assert sd["global_step"] == 6 and len(sd["moving_average"]) == len(sd["model"])


def test_horovod_style_runner_over_maskrcnn_stack(tmp_path):
    """prepare-bucket -> mask-rcnn stack -> run.sh (hostfile with slots -> dlcfn mpirun -> horovod trainer)."""
    pytest.importorskip("torchvision")
    from deeplearning_cfn_b200.cluster.staging import prepare_bucket
    from deeplearning_cfn_b200.cluster.stack import launch_cluster
    bucket = str(tmp_path / "bucket")
    prepare_bucket(bucket, num_images=8, image_size=64)
    h = launch_cluster("mr", "mask-rcnn.template", home=str(tmp_path), backend="thread", WorkerCount=1,
                       InstanceType="cpu.large", ImageType="CPUGloo", S3Bucket=bucket, EFSServesData="true")
    env = h.export_env()
    home = os.path.join(h.master_root, "home")
    # run from the master's home directory, like the reference
    job_env = dict(ENV)
    job_env.update(env)
    job_env.update(HOME=home, DLCFN_PY=PY, DEEPLEARNING_WORKER_GPU_COUNT="1", STEPS_PER_EPOCH_OVERRIDE="2",
                   EXTRA_CONFIG="PREPROC.MAX_SIZE=64 DATA.NUM_CLASS=5")
    r = subprocess.run(["bash", os.path.join(home, "run.sh")], cwd=home, capture_output=True, text=True, timeout=900,
                       env=job_env)
    assert r.returncode == 0, r.stdout[-2500:] + r.stderr[-2500:]
    assert "Data directory: " + os.path.join(env["EFS_MOUNT"], "data") in r.stdout
    assert "Number of parallel ranks: 2" in r.stdout and "Training finished" in r.stdout
    assert open(os.path.join(home, "hostfile")).read() == "deeplearning-worker1 slots=1\ndeeplearning-worker2 slots=1\n"
    runs = [d for d in os.listdir(env["EFS_MOUNT"]) if d.startswith("mask-rcnn-coco-2-efs-")]
    assert runs and os.path.exists(os.path.join(env["EFS_MOUNT"], runs[0], "train_log/maskrcnn/checkpoint.pt"))


def test_train_imagenet_checkpoint_and_resume(tmp_path):
    prefix = str(tmp_path / "ck" / "r20")
    base = [PY, "-m", "deeplearning_cfn_b200.examples.train_imagenet", "--benchmark", "1", "--network", "resnet",
            "--num-layers", "20", "--image-shape", "3,32,32", "--num-classes", "10", "--batch-size", "8",
            "--num-batches", "3", "--disp-batches", "1", "--model-prefix", prefix, "--lr-step-epochs", "1"]
    r = run(base + ["--num-epochs", "1"])
    assert r.returncode == 0 and os.path.exists(prefix + "-0001.params") and os.path.exists(prefix + "-symbol.json"), r.stderr[-2000:]
    assert "Speed:" in r.stderr + r.stdout
    r = run(base + ["--num-epochs", "2", "--load-epoch", "1"])
    assert r.returncode == 0 and os.path.exists(prefix + "-0002.params"), r.stderr[-2000:]
    assert "Epoch[1]" in r.stderr + r.stdout and "Epoch[0] Batch" not in r.stderr + r.stdout


def test_run_example_cli_lists_and_builds_commands():
    r = run([PY, "-m", "deeplearning_cfn_b200", "run-example", "--list"])
    assert r.returncode == 0 and "mxnet-cifar10-ps" in r.stdout and "maskrcnn-horovod" in r.stdout
    r = run([PY, "-m", "deeplearning_cfn_b200", "run-example", "resnet50-ps", "--gpus", "8", "--dry-run"])
    assert "mpirun -np 8" in r.stdout and "--kv-store dist_sync_device" in r.stdout
    r = run([PY, "-m", "deeplearning_cfn_b200", "benchmark", "--dry-run", "--gpu_count", "8", "--kv-stores", "allreduce"])
    assert r.stdout.count("train_imagenet") == 4 and "-np 8" in r.stdout


def test_profiler_trace_and_aggregate(tmp_path):
    from deeplearning_cfn_b200.utils import profiler
    profiler.set_config(filename=str(tmp_path / "trace.json"))
    profiler.set_state("run")
    for _ in range(3):
        with profiler.scope("KVStoreDistDefaultPush", "kvstore"):
            time.sleep(0.002)
    profiler.marker("epoch-end")
    c = profiler.Counter("bytes_pushed")
    c.increment(10)
    table = profiler.dumps()
    assert "KVStoreDistDefaultPush" in table and " 3 " in table
    fn = profiler.dump()
    ev = json.load(open(fn))["traceEvents"]
    assert sum(1 for e in ev if e["name"] == "KVStoreDistDefaultPush" and e["ph"] == "X") == 3
    assert any(e["ph"] == "i" for e in ev) and any(e["ph"] == "C" for e in ev)
    assert not profiler.is_running()
    # user-defined objects: Domain / Task / Frame / Event / Counter / Marker, pause / resume
    profiler.set_config(filename=str(tmp_path / "user.json"))
    profiler.set_state("run")
    dom = profiler.Domain("training")
    with dom.new_task("epoch"):
        with dom.new_frame("iteration"):
            time.sleep(0.001)
        profiler.pause()
        with profiler.Event("not recorded"):
            pass
        profiler.resume()
        dom.new_counter("samples", 5).increment(3)
        dom.new_marker("checkpoint").mark("global")
    ev = json.load(open(profiler.dump()))["traceEvents"]
    cats = {e["name"]: e.get("cat") for e in ev}
    assert cats["epoch"] == "training:task" and cats["iteration"] == "training:frame" and "not recorded" not in cats
    assert any(e["ph"] == "C" and e["args"].get("training::samples") == 8 for e in ev)
    assert any(e["ph"] == "i" and e["name"] == "checkpoint" and e["s"] == "g" for e in ev)


def test_train_imagenet_from_record_file_two_ranks(tmp_path):
    """Real-data path: .rec/.idx -> native decoder ring, sharded over 2 ranks (gloo), random crop + mirror."""
    import numpy as np
    from deeplearning_cfn_b200.utils import recordio
    rec = recordio.MXIndexedRecordIO(str(tmp_path / "train.idx"), str(tmp_path / "train.rec"), "w")
    rng = np.random.default_rng(1)
    for i in range(48):
        rec.write_idx(i, recordio.pack_img(recordio.IRHeader(0, float(i % 10), i, 0), rng.integers(0, 256, (36, 36, 3), dtype=np.uint8)))
    rec.close()
    r = run([PY, "-m", "deeplearning_cfn_b200.launch.mpirun", "-np", "2", PY, "-m", "deeplearning_cfn_b200.examples.train_imagenet",
             "--network", "resnet", "--num-layers", "20", "--num-classes", "10", "--image-shape", "3,32,32", "--batch-size", "8",
             "--num-epochs", "2", "--kv-store", "horovod", "--data-train", str(tmp_path / "train.rec"), "--rand-crop", "1",
             "--rand-mirror", "1", "--data-nthreads", "2", "--disp-batches", "1"])
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-3000:]
    assert "24 samples in this shard, 3 batches/epoch" in out and "Epoch[1] Train-loss" in out and "Speed:" in out


def test_gluon_style_example_reads_record_files(tmp_path):
    """image_classification --data-dir with train.rec / val.rec (the reference's ImageNet input form) on the native pipeline."""
    import numpy as np
    from deeplearning_cfn_b200.utils import recordio
    rng = np.random.default_rng(2)
    for split, n in (("train", 40), ("val", 16)):
        rec = recordio.MXIndexedRecordIO(str(tmp_path / f"{split}.idx"), str(tmp_path / f"{split}.rec"), "w")
        for i in range(n):
            lab = i % 10
            img = np.clip(rng.integers(0, 40, (36, 36, 3)) + lab * 20, 0, 255).astype(np.uint8)
            rec.write_idx(i, recordio.pack_img(recordio.IRHeader(0, float(lab), i, 0), img))
        rec.close()
    r = run([PY, "-m", "deeplearning_cfn_b200.examples.image_classification", "--dataset", "cifar10", "--model", "resnet20",
             "--data-dir", str(tmp_path), "--epochs", "1", "--batch-size", "8", "--kvstore", "local", "--log-interval", "1"])
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-3000:]
    assert "Epoch[0] Batch [4]" in out and "validation: accuracy" in out


def test_profiler_category_switches_and_mode_env(tmp_path):
    """set_config(profile_all=False, profile_api=False): kvstore / API ranges are dropped, operators and user domains stay;
    MXNET_PROFILER_AUTOSTART=1 with MXNET_PROFILER_MODE=0 starts a symbolic-operators-only profile at import."""
    from deeplearning_cfn_b200.utils import profiler
    profiler.set_config(filename=str(tmp_path / "cat.json"), profile_all=False, profile_api=False, profile_symbolic=True)
    profiler.set_state("run")
    with profiler.scope("KVStoreReduce", "kvstore"):
        pass
    with profiler.scope("conv0", "operator"):
        pass
    with profiler.Domain("mine").new_task("t"):
        pass
    names = {e["name"] for e in json.load(open(profiler.dump()))["traceEvents"]}
    assert "conv0" in names and "t" in names and "KVStoreReduce" not in names
    profiler.set_config(filename=str(tmp_path / "x.json"))          # back to the defaults for later tests
    code = ("from deeplearning_cfn_b200.utils import profiler as p\n"
            "print(p.is_running(), p.category_enabled('operator'), p.category_enabled('kvstore'))\n")
    env = dict(os.environ, MXNET_PROFILER_AUTOSTART="1", MXNET_PROFILER_MODE="0", DLCFN_PROFILER_FILE=str(tmp_path / "auto.json"),
               PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=120)
    assert r.stdout.split() == ["True", "True", "False"], r.stdout + r.stderr


def test_train_imagenet_under_the_parameter_server_launcher(tmp_path):
    """`launch -n 2 -s 2 python train_imagenet.py --benchmark 1 --kv-store dist_sync` (README.md:253-258 of the reference's
    example directory): scheduler and servers become their nodes at package import, the workers take rank / world from the
    PS rendezvous, train in lock step and the job ends by itself."""
    from conftest import run_job
    cmd = [sys.executable, "-m", "deeplearning_cfn_b200", "launch", "-n", "2", "-s", "2", "--launcher", "local", sys.executable, "-m",
           "deeplearning_cfn_b200.examples.train_imagenet", "--network", "resnet", "--num-layers", "20", "--benchmark", "1",
           "--batch-size", "4", "--image-shape", "3,32,32", "--num-classes", "10", "--num-batches", "3", "--num-epochs", "1",
           "--kv-store", "dist_sync", "--disp-batches", "1"]
    r = run_job(cmd, env=dict(os.environ, PYTHONPATH=ROOT, OMP_NUM_THREADS="2"), timeout=420)
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-3000:]
    import re
    losses = sorted(set(re.findall(r"Train-loss=([0-9.]+)", out)))          # regex: the ranks share the stream
    assert "Node[0]" in out and "Node[1]" in out and len(losses) == 1, out[-2000:]      # both ranks, identical loss


def test_train_imagenet_resume_rederives_the_learning_rate_once(tmp_path):
    """--load-epoch N: the rate starts at lr * factor^(steps <= N) and the steps already passed are not applied again
    (fit.py:29-55 of the reference: `begin_epoch` removes them from the scheduler)."""
    from conftest import run_job
    base = [sys.executable, "-m", "deeplearning_cfn_b200.examples.train_imagenet", "--network", "resnet", "--num-layers", "20",
            "--benchmark", "1", "--batch-size", "4", "--image-shape", "3,32,32", "--num-classes", "10", "--num-batches", "2",
            "--gpus", "", "--model-prefix", str(tmp_path / "ck"), "--save-period", "1", "--lr-step-epochs", "1,2,3", "--lr", "0.1"]
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = run_job(base + ["--num-epochs", "2"], env=env, timeout=300)
    assert r.returncode == 0 and os.path.exists(tmp_path / "ck-0002.params"), (r.stdout + r.stderr)[-2000:]
    r = run_job(base + ["--num-epochs", "4", "--load-epoch", "2"], env=env, timeout=300)
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-2000:]
    changes = [l.split("Change learning rate to ")[1].strip() for l in out.splitlines() if "Change learning rate to" in l]
    assert changes == ["1.00000e-04"], changes          # resumed at 1e-3 (two steps passed), one more step at epoch 3
    assert os.path.exists(tmp_path / "ck-0004.params")


def test_image_classification_resume_accepts_periodic_and_best_checkpoints(tmp_path):
    """--resume takes the file written every --save-frequency (arg:/aux: keys) as well as the -best file (plain state dict)."""
    from conftest import run_job
    base = [sys.executable, "-m", "deeplearning_cfn_b200.examples.image_classification", "--dataset", "dummy", "--model", "resnet18_v1",
            "--batch-size", "4", "--gpus", "", "--num-samples", "8", "--prefix", str(tmp_path / "m"), "--save-frequency", "1",
            "--kvstore", "local"]
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = run_job(base + ["--epochs", "1"], env=env, timeout=300)
    assert r.returncode == 0 and os.path.exists(tmp_path / "m-0001.params"), (r.stdout + r.stderr)[-2000:]
    import torch
    from deeplearning_cfn_b200.models import get_model
    from deeplearning_cfn_b200.utils.callback import load_params_file
    raw = torch.load(tmp_path / "m-0001.params", map_location="cpu")
    assert all(k[:4] in ("arg:", "aux:") for k in raw)
    torch.save({k[4:]: v for k, v in raw.items()}, tmp_path / "plain.params")      # what net.save_parameters / -best.params hold
    for f in ("m-0001.params", "plain.params"):
        r = run_job(base + ["--epochs", "2", "--start-epoch", "1", "--resume", str(tmp_path / f)], env=env, timeout=300)
        assert r.returncode == 0, f + (r.stdout + r.stderr)[-2000:]
    res = load_params_file(str(tmp_path / "m-0001.params"), get_model("resnet18_v1", 1000, 224), strict=False)
    assert not res.unexpected_keys and get_model is not None


def test_scaling_sweep_tool_runs_on_a_box_without_gpus(tmp_path):
    """benchmark.py (`--networks mode:net:batch:size --worker_count --gpu_count`): with --gpu_count 0 the sweep is over CPU
    ranks; it writes <net>.csv and the speed-up chart like the reference tool."""
    from conftest import run_job
    r = run_job([sys.executable, "-m", "deeplearning_cfn_b200", "benchmark", "--networks", "native:resnet-20:4:32", "--gpu_count", "0",
                 "--worker_count", "2", "--kv-stores", "local", "--num-batches", "4", "--out-dir", str(tmp_path)],
                env=dict(os.environ, PYTHONPATH=ROOT), timeout=600)
    assert r.returncode == 0, (r.stdout + r.stderr)[-2000:]
    rows = open(tmp_path / "resnet.csv").read().strip().splitlines()
    assert rows[0] == "network,layers,kv_store,gpus,images_per_sec" and len(rows) == 3
    assert all(float(l.split(",")[-1]) > 0 for l in rows[1:]) and (tmp_path / "resnet_speedup.svg").exists()
