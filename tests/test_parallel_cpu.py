"""Data-parallel engine / Horovod-style API / gluon-style Trainer on CPU."""
import os
import subprocess
import sys

import pytest
import torch

import deeplearning_cfn_b200.kv as kv
from deeplearning_cfn_b200.kv.trainer import Trainer

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("naive", ["0", "1"], ids=["overlapped", "naive-serialised"])
def test_gloo_world2_data_parallel_matches_single_process(naive):
    cmd = [sys.executable, "-m", "deeplearning_cfn_b200.launch.mpirun", "-np", "2", sys.executable,
           os.path.join(ROOT, "tests/mp/gloo_trainer_worker.py")]
    from conftest import run_job
    r = run_job(cmd, env=dict(os.environ, PYTHONPATH=ROOT, DLCFN_NAIVE_COLLECTIVES=naive), timeout=300)
    assert r.returncode == 0 and "GLOO_TRAINER_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


def test_vmm_backend_file_descriptor_exchange_world3():
    cmd = [sys.executable, "-m", "deeplearning_cfn_b200.launch.mpirun", "-np", "3", sys.executable,
           os.path.join(ROOT, "tests/mp/fd_exchange_worker.py")]
    from conftest import run_job
    r = run_job(cmd, env=dict(os.environ, PYTHONPATH=ROOT), timeout=120)
    assert r.returncode == 0 and r.stdout.count("FD_EXCHANGE_OK") == 3, r.stdout[-2000:] + r.stderr[-3000:]


def test_vmm_backend_construction_agrees_on_failures_world3():
    """SymmetricHeap._create_vmm_backend over stand-in driver entry points: the success path, and a failure injected on one
    rank in each phase, which every rank must raise (tests/mp/vmm_flow_worker.py)."""
    cmd = [sys.executable, "-m", "deeplearning_cfn_b200.launch.mpirun", "-np", "3", sys.executable,
           os.path.join(ROOT, "tests/mp/vmm_flow_worker.py")]
    from conftest import run_job
    r = run_job(cmd, env=dict(os.environ, PYTHONPATH=ROOT), timeout=180)
    assert r.returncode == 0 and r.stdout.count("VMM_FLOW_OK") == 3, r.stdout[-2000:] + r.stderr[-3000:]


def test_syncbn_statistics_exchange_over_a_heap_matches_torch_distributed_world2():
    cmd = [sys.executable, "-m", "deeplearning_cfn_b200.launch.mpirun", "-np", "2", sys.executable,
           os.path.join(ROOT, "tests/mp/syncbn_heap_worker.py")]
    from conftest import run_job
    r = run_job(cmd, env=dict(os.environ, PYTHONPATH=ROOT), timeout=180)
    assert r.returncode == 0 and r.stdout.count("SYNCBN_HEAP_OK") == 2, r.stdout[-2000:] + r.stderr[-3000:]


def test_hierarchical_multinode_mode_matches_flat_allreduce():
    """2 virtual nodes x 2 ranks on gloo: node-local sum -> per-rail all-reduce of the slices -> node-local broadcast."""
    cmd = [sys.executable, "-m", "deeplearning_cfn_b200.launch.mpirun", "-np", "4", sys.executable,
           os.path.join(ROOT, "tests/mp/hier_trainer_worker.py")]
    from conftest import run_job
    r = run_job(cmd, env=dict(os.environ, PYTHONPATH=ROOT, DLCFN_VIRTUAL_NODE_SIZE="2", OMP_NUM_THREADS="2"), timeout=600)
    assert r.returncode == 0 and "HIER_TRAINER_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


def test_allreduce_job_through_tracker_launcher(tmp_path):
    """launch -n 2 -s 0: no servers -> all-reduce tracker; workers find each other through it."""
    script = tmp_path / "w.py"
    script.write_text(
        "import sys, torch\n"
        f"sys.path.insert(0, r'{ROOT}')\n"
        "from deeplearning_cfn_b200.parallel import horovod as hvd\n"
        "hvd.init()\n"
        "t = hvd.allreduce(torch.ones(3) * (hvd.rank() + 1), average=True)\n"
        "assert hvd.size() == 2 and torch.allclose(t, torch.full((3,), 1.5)), t\n"
        "print('TRACKER_JOB_OK', hvd.rank(), flush=True)\n"
        "hvd.shutdown()\n")
    from conftest import run_job
    r = run_job([sys.executable, "-m", "deeplearning_cfn_b200", "launch", "-n", "2", "-s", "0", "--launcher", "local",
                 sys.executable, str(script)], env=dict(os.environ, PYTHONPATH=ROOT), timeout=180)
    assert r.returncode == 0 and r.stdout.count("TRACKER_JOB_OK") == 2, r.stdout[-2000:] + r.stderr[-3000:]


def _net():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4))


@pytest.mark.parametrize("kvstore,update_on_kv", [("local", False), ("local", True), ("device", None), (None, None)])
def test_trainer_step_matches_manual_sgd(kvstore, update_on_kv):
    net, ref = _net(), _net()
    tr = Trainer(net.named_parameters(), "sgd", {"learning_rate": 0.1, "momentum": 0.9, "wd": 1e-3}, kvstore=kvstore,
                 update_on_kvstore=update_on_kv)
    mom = [torch.zeros_like(p) for p in ref.parameters()]
    x, y = torch.randn(5, 8), torch.randn(5, 4)
    for _ in range(3):
        for m in (net, ref):
            m.zero_grad()
            torch.nn.functional.mse_loss(m(x), y, reduction="sum").backward()
        tr.step(batch_size=5)
        with torch.no_grad():
            for p, m in zip(ref.parameters(), mom):
                g = p.grad / 5
                m.mul_(0.9).add_(-0.1 * 1e-3 * p - 0.1 * g)
                p.add_(m)
    for a, b in zip(net.parameters(), ref.parameters()):
        assert torch.allclose(a, b, atol=1e-6)
    assert tr.learning_rate == 0.1
    tr.set_learning_rate(0.01)
    assert tr.learning_rate == 0.01


def test_trainer_split_allreduce_update_and_states(tmp_path):
    net = _net()
    tr = Trainer(net.named_parameters(), "adam", {"learning_rate": 0.01}, kvstore="local", update_on_kvstore=False)
    x, y = torch.randn(5, 8), torch.randn(5, 4)
    torch.nn.functional.mse_loss(net(x), y).backward()
    tr.allreduce_grads()
    tr.update(batch_size=1)
    f = str(tmp_path / "trainer.states")
    tr.save_states(f)
    tr2 = Trainer(_net().named_parameters(), "adam", {"learning_rate": 0.5}, kvstore="local", update_on_kvstore=False)
    tr2.load_states(f)
    assert tr2.learning_rate == 0.01 and len(tr2._updaters[0].states) == len(tr._updaters[0].states)
    tr3 = Trainer(_net().named_parameters(), "sgd", {"learning_rate": 0.1}, kvstore="local", update_on_kvstore=True)
    with pytest.raises(AssertionError):
        tr3.allreduce_grads()
    with pytest.raises(ValueError):
        Trainer(_net().named_parameters(), kv.optimizer.create("sgd"), {"learning_rate": 0.1})
    stale = Trainer(_net().named_parameters(), "sgd", {"learning_rate": 0.1}, kvstore=None)
    with pytest.raises(UserWarning):
        stale.step(1)
    stale.step(1, ignore_stale_grad=True)


def test_module_fit_and_recordio(tmp_path):
    import numpy as np
    from deeplearning_cfn_b200.kv.module import Module
    from deeplearning_cfn_b200.utils import recordio
    from deeplearning_cfn_b200.utils.callback import Speedometer
    # write a small .rec/.idx, read it back through the prefetching iterator
    rec = recordio.MXIndexedRecordIO(str(tmp_path / "d.idx"), str(tmp_path / "d.rec"), "w")
    rng = np.random.default_rng(0)
    for i in range(40):
        img = np.full((8, 8, 3), (i % 4) * 60, dtype=np.uint8) + rng.integers(0, 10, (8, 8, 3), dtype=np.uint8)
        rec.write_idx(i, recordio.pack_img(recordio.IRHeader(0, i % 4, i, 0), img))
    rec.close()
    r = recordio.MXRecordIO(str(tmp_path / "d.rec"), "r")
    h, img = recordio.unpack_img(r.read())
    assert h.label == 0.0 and img.shape == (8, 8, 3)
    it = recordio.ImageRecordIter(str(tmp_path / "d.rec"), batch_size=8, data_shape=(3, 8, 8), shuffle=True, num_parts=2,
                                  part_index=1, preprocess_threads=2)
    batches = list(it)
    assert len(batches) == 2 and batches[0][0].shape == (8, 8, 8, 3) and batches[0][0].dtype == torch.uint8

    def data():
        it.reset()
        for x, y in it:
            yield x.permute(0, 3, 1, 2).float() / 64.0 - 1.5, y
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Flatten(), torch.nn.Linear(192, 32), torch.nn.ReLU(), torch.nn.Linear(32, 4))
    mod = Module(net)
    mod.fit(data, num_epoch=120, kvstore="local", optimizer="sgd", optimizer_params=(("learning_rate", 0.2),),
            batch_size=8, batch_end_callback=Speedometer(8, 1000))
    assert dict(mod.score(data))["accuracy"] > 0.9
    mod.save_checkpoint(str(tmp_path / "m"), 120, save_optimizer_states=True)
    m2 = Module.load(str(tmp_path / "m"), 120, torch.nn.Sequential(torch.nn.Flatten(), torch.nn.Linear(192, 32), torch.nn.ReLU(), torch.nn.Linear(32, 4)))
    assert abs(dict(m2.score(data))["accuracy"] - dict(mod.score(data))["accuracy"]) < 1e-6


def test_sync_replicas_with_backup_worker():
    """TF SyncReplicasOptimizer semantics on the native PS: aggregate 2 of 3 replicas, drop stale gradients."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PYTHONPATH=root, MXNET_KVSTORE_BIGARRAY_BOUND="40")
    cmd = [sys.executable, "-m", "deeplearning_cfn_b200", "launch", "-n", "3", "-s", "2", "--launcher", "local",
           sys.executable, os.path.join(root, "tests/mp/sync_replicas_worker.py")]
    from conftest import run_job
    r = run_job(cmd, env=env, timeout=240)
    assert r.returncode == 0 and r.stdout.count("SYNCREP_OK") == 3, r.stdout[-2000:] + r.stderr[-3000:]


def test_sync_replicas_accumulator_drops_stale_and_averages():
    from deeplearning_cfn_b200.kv import optimizer as opt
    from deeplearning_cfn_b200.parallel.sync_replicas import SyncReplicasOptimizer
    o = SyncReplicasOptimizer(inner=opt.create("sgd", learning_rate=0.5), replicas_to_aggregate=2, total_num_replicas=3)
    u = opt.get_updater(o)
    w = torch.tensor([1.0, 1.0, 0.0])                       # [weights..., global_step]
    u(0, torch.tensor([2.0, 4.0, 0.0]), w)                  # fresh, 1 of 2
    assert w.tolist() == [1.0, 1.0, 0.0]
    u(0, torch.tensor([4.0, 0.0, 0.0]), w)                  # fresh, 2 of 2 -> mean (3, 2) applied, step 1
    assert torch.allclose(w, torch.tensor([1.0 - 0.5 * 3.0, 1.0 - 0.5 * 2.0, 1.0]))
    u(0, torch.tensor([100.0, 100.0, 0.0]), w)              # stale (local step 0 < global 1): dropped
    assert torch.allclose(w, torch.tensor([-0.5, 0.0, 1.0])) and o.dropped == 1
    u(0, torch.tensor([1.0, 1.0, 1.0]), w)
    u(0, torch.tensor([1.0, 1.0, 1.0]), w)
    assert torch.allclose(w, torch.tensor([-1.0, -0.5, 2.0])) and o.applied == 2


def test_host_storage_pool_recycles_blocks_by_size_class():
    from deeplearning_cfn_b200.utils import storage
    storage.set_strategy("Naive")
    storage.release_all()
    s0 = storage.stats()
    a = storage.host_empty((1000,), torch.float32)            # 4000 B -> one 4 KiB page
    a.fill_(3.0)
    addr = a.data_ptr()
    assert addr % 4096 == 0 and storage.size_class(4000) == 4096
    storage.release(a)
    b = storage.host_empty((1024,), torch.float32)            # same size class: the block comes back
    assert b.data_ptr() == addr
    s1 = storage.stats()
    assert s1["reuses"] == s0["reuses"] + 1 and s1["fresh_allocations"] == s0["fresh_allocations"] + 1
    storage.release(b)
    with pytest.raises(ValueError):
        storage.release(b)
    storage.set_strategy("Round", cutoff_log2=12)
    assert storage.size_class(5000) == 8192 and storage.size_class(70000) == 131072
    storage.set_strategy("Naive")
    assert storage.size_class(5000) == 8192 and storage.size_class(70000) == 73728
    assert storage.release_all() >= 1


def test_shared_memory_tensor_between_processes():
    import subprocess
    import sys
    from deeplearning_cfn_b200.utils import storage
    name = f"dlcfn_test_{os.getpid()}"
    seg = storage.SharedTensor.create(name, (4, 5), torch.float32)
    try:
        seg.tensor.copy_(torch.arange(20, dtype=torch.float32).view(4, 5))
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        code = ("import torch,sys; sys.path.insert(0, %r); from deeplearning_cfn_b200.utils import storage; "
                "s = storage.SharedTensor.attach(%r, (4, 5), torch.float32); assert float(s.tensor.sum()) == 190.0; "
                "s.tensor.mul_(2); s.close(); print('CHILD_OK')") % (root, name)
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
        assert "CHILD_OK" in r.stdout, r.stdout + r.stderr
        assert float(seg.tensor.sum()) == 380.0              # the child's write is visible: same pages
        with pytest.raises(OSError):
            storage.SharedTensor.create(name, (4, 5), torch.float32)     # name is taken while the owner lives
    finally:
        seg.close()
    with pytest.raises(OSError):
        storage.SharedTensor.attach(name, (4, 5), torch.float32)         # owner closed -> unlinked


def test_metrics_match_closed_form():
    from deeplearning_cfn_b200.kv import metric
    p = torch.tensor([[0.1, 0.7, 0.2], [0.8, 0.1, 0.1], [0.2, 0.3, 0.5], [0.3, 0.4, 0.3]])
    y = torch.tensor([1, 2, 2, 0])
    acc = metric.create("acc")
    acc.update([y], [p])
    assert acc.get() == ("accuracy", 0.5)
    top2 = metric.TopKAccuracy(2)
    top2.update([y], [p])
    assert top2.get()[0] == "top_k_accuracy_2" and abs(top2.get()[1] - 0.75) < 1e-9     # rows 0, 2 and 3 (0.3 ties broken by topk)
    ce = metric.create("ce")
    ce.update([y], [p])
    want = -(torch.log(torch.tensor([0.7, 0.1, 0.5, 0.3]))).mean().item()
    assert abs(ce.get()[1] - want) < 1e-5
    ppl = metric.Perplexity(ignore_label=0)
    ppl.update([y], [p])
    want_ppl = float(torch.exp(-(torch.log(torch.tensor([0.7, 0.1, 0.5]))).mean()))
    assert abs(ppl.get()[1] - want_ppl) < 1e-4
    mse, mae, rmse = metric.create("mse"), metric.create("mae"), metric.create("rmse")
    a, b = torch.tensor([1.0, 2.0, 4.0]), torch.tensor([1.0, 1.0, 2.0])
    for m in (mse, mae, rmse):
        m.update([a], [b])
    assert abs(mse.get()[1] - 5.0 / 3) < 1e-6 and abs(mae.get()[1] - 1.0) < 1e-6 and abs(rmse.get()[1] - (5.0 / 3) ** 0.5) < 1e-6
    f1 = metric.create("f1")
    f1.update([torch.tensor([1, 0, 1, 1])], [torch.tensor([[0.2, 0.8], [0.3, 0.7], [0.9, 0.1], [0.4, 0.6]])])
    assert abs(f1.get()[1] - 2 * (2 / 3) * (2 / 3) / (4 / 3)) < 1e-9
    comp = metric.create(["acc", "ce"])
    comp.update([y], [p])
    assert comp.get_name_value()[0] == ("accuracy", 0.5) and comp.get_name_value()[1][0] == "cross-entropy"
    comp.reset()
    assert comp.get()[1][0] != comp.get()[1][0]                      # nan after reset
    custom = metric.np(lambda label, pred: float((pred.argmax(-1) == label).float().mean()), name="myacc")
    custom.update([y], [p])
    assert custom.get() == ("myacc", 0.5)
    with pytest.raises(ValueError):
        metric.create("no-such-metric")


def test_ndarray_iter_last_batch_handling_and_prefetch():
    from deeplearning_cfn_b200.utils.io import NDArrayIter, PrefetchingIter, ResizeIter
    x = torch.arange(10, dtype=torch.float32).view(10, 1)
    y = torch.arange(10)
    it = NDArrayIter(x, y, batch_size=4, last_batch_handle="pad")
    batches = list(it)
    assert [b.pad for b in batches] == [0, 0, 2] and batches[2].data[0].view(-1).tolist() == [8.0, 9.0, 0.0, 1.0]
    assert it.provide_data[0].name == "data" and it.provide_data[0].shape == (4, 1) and it.provide_label[0].name == "softmax_label"
    it = NDArrayIter(x, y, batch_size=4, last_batch_handle="discard")
    assert len(list(it)) == 2
    it = NDArrayIter(x, y, batch_size=4, last_batch_handle="roll_over")
    assert len(list(it)) == 2
    it.reset()
    second = list(it)                                  # 2 carried over + 10 fresh = 3 full batches
    assert len(second) == 3 and second[0].data[0].view(-1).tolist() == [8.0, 9.0, 0.0, 1.0]
    sh = NDArrayIter(x, y, batch_size=5, shuffle=True, seed=3)
    seen = torch.cat([b.label[0] for b in sh])
    assert sorted(seen.tolist()) == list(range(10)) and seen.tolist() != list(range(10))
    rs = ResizeIter(NDArrayIter(x, y, batch_size=4, last_batch_handle="discard"), size=5)
    assert len(list(rs)) == 5
    pf = PrefetchingIter(NDArrayIter(x, y, batch_size=4, last_batch_handle="pad"), depth=2, pin=True)
    got = [b.data[0].clone() for b in pf]
    assert len(got) == 3 and got[0].view(-1).tolist() == [0.0, 1.0, 2.0, 3.0]
    pf.reset()
    xb, yb = next(iter(pf))                            # DataBatch unpacks like a tuple
    assert xb.shape == (4, 1) and yb.tolist() == [0, 1, 2, 3]
    pf.close()


def test_module_fit_with_dataiter_metric_monitor_and_model_helpers(tmp_path, caplog):
    import logging
    from deeplearning_cfn_b200.kv import metric, model
    from deeplearning_cfn_b200.kv.module import Module
    from deeplearning_cfn_b200.utils.callback import LogValidationMetricsCallback, log_train_metric
    from deeplearning_cfn_b200.utils.io import NDArrayIter
    from deeplearning_cfn_b200.utils.monitor import Monitor
    torch.manual_seed(0)
    x = torch.randn(64, 8)
    y = (x[:, 0] > 0).long()
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 2))
    train = NDArrayIter(x, y, batch_size=16, shuffle=True)
    mon = Monitor(interval=2, pattern=".*")
    mod = Module(net)
    with caplog.at_level(logging.INFO):
        mod.fit(train, eval_data=NDArrayIter(x, y, batch_size=16), eval_metric=["acc", metric.CrossEntropy(from_logits=True)],
                num_epoch=8, optimizer="sgd", optimizer_params=(("learning_rate", 0.5),), batch_size=16, monitor=mon,
                batch_end_callback=log_train_metric(2), eval_end_callback=LogValidationMetricsCallback())
    text = caplog.text
    assert "Train-accuracy" in text and "Validation-cross-entropy" in text and "_output" in text
    res = dict(mod.score(NDArrayIter(x, y, batch_size=16), ["acc", metric.TopKAccuracy(2)]))
    assert res["accuracy"] > 0.85 and res["top_k_accuracy_2"] == 1.0
    assert mod.predict(NDArrayIter(x, y, batch_size=24, last_batch_handle="pad")).shape == (64, 2)   # padding removed
    # model.py helpers: store selection, init + update on a 2-"device" local store, checkpoint round trip
    arg = {"w": torch.ones(4), "b": torch.zeros(2)}
    kv0, upd0 = model._create_kvstore("local", 1, arg)
    assert kv0 is None and upd0 is False
    store, upd = model._create_kvstore("local", 2, arg)
    assert store is not None and upd is True
    names = ["w", "b"]
    params = [[arg["w"].clone(), arg["w"].clone()], [arg["b"].clone(), arg["b"].clone()]]
    model._initialize_kvstore(store, params, arg, names, upd)
    store.set_optimizer(kv.optimizer.create("test", rescale_grad=1.0))
    grads = [[torch.ones(4), torch.ones(4) * 2], [None, None]]
    model._update_params_on_kvstore(params, grads, store, names)
    assert torch.equal(params[0][0], torch.full((4,), 4.0)) and torch.equal(params[0][1], params[0][0])
    assert torch.equal(params[1][0], torch.zeros(2))                    # parameter without gradient untouched
    store2, _ = model._create_kvstore("device", 2, arg)
    p2 = [[arg["w"].clone(), arg["w"].clone()]]
    model._initialize_kvstore(store2, p2, arg, ["w"], False)
    upd_fn = kv.optimizer.get_updater(kv.optimizer.create("sgd", learning_rate=0.5))
    model._update_params(p2, [[torch.ones(4), torch.ones(4)]], upd_fn, 2, store2, ["w"])
    assert torch.allclose(p2[0][0], torch.full((4,), 0.0)) and torch.allclose(p2[0][1], p2[0][0])   # 1 - 0.5 * (1 + 1)
    model.save_checkpoint(str(tmp_path / "ck"), 3, net, dict(net.named_parameters()), {})
    sym, a2, aux = model.load_checkpoint(str(tmp_path / "ck"), 3)
    assert sym["class"] == "Sequential" and set(a2) == {n for n, _ in net.named_parameters()} and aux == {}


def test_native_recordio_pipeline_matches_python_reader(tmp_path):
    import struct
    import numpy as np
    from deeplearning_cfn_b200.utils import recordio
    # -- framing: the C++ writer splits payloads at embedded magic words (dmlc continuation records), both readers join them
    w = recordio.NativeRecordIO(str(tmp_path / "f.idx"), str(tmp_path / "f.rec"), "w")
    magic = struct.pack("<I", recordio.MAGIC)
    payloads = [b"hello", b"", b"abcd" + magic + b"efgh" + magic + b"zz", magic + b"tail", b"x" * 1001]
    for i, p in enumerate(payloads):
        w.write(10 + i, p)
    w.close()
    r = recordio.NativeRecordIO(str(tmp_path / "f.idx"), str(tmp_path / "f.rec"), "r")
    assert len(r) == len(payloads) and [r.key(i) for i in range(len(r))] == [10, 11, 12, 13, 14]
    assert [r.read(i) for i in range(len(r))] == payloads
    r.close()
    r2 = recordio.NativeRecordIO(None, str(tmp_path / "f.rec"), "r")          # no index: scanned from the framing
    assert len(r2) == len(payloads) and r2.read(2) == payloads[2]
    r2.close()
    pr = recordio.MXRecordIO(str(tmp_path / "f.rec"), "r")                    # the python reader joins the parts too
    assert [pr.read() for _ in payloads] == payloads and pr.read() is None
    pr.close()
    # -- image batches: python writer -> native iterator, against the python reader
    rec = recordio.MXIndexedRecordIO(str(tmp_path / "d.idx"), str(tmp_path / "d.rec"), "w")
    rng = np.random.default_rng(0)
    imgs = [rng.integers(0, 256, (10, 12, 3), dtype=np.uint8) for _ in range(37)]
    for i, img in enumerate(imgs):
        rec.write_idx(i, recordio.pack_img(recordio.IRHeader(0, float(i % 5), i, 0), img))
    rec.close()
    it = recordio.NativeImageRecordIter(str(tmp_path / "d.rec"), batch_size=8, data_shape=(3, 8, 8), preprocess_threads=3,
                                        prefetch_buffer=2, round_batch=True)
    assert len(it) == 5 and it.num_samples == 37
    seen, pads = [], []
    for x, y in it:
        seen.append((x.clone(), y.clone()))
        pads.append(it.pad)
    assert pads == [0, 0, 0, 0, 3]                                   # 37 = 4 x 8 + 5, wrapped to the epoch start
    for b, (x, y) in enumerate(seen):
        for j in range(8):
            i = (b * 8 + j) % 37
            assert int(y[j]) == i % 5
            assert np.array_equal(x[j].numpy(), imgs[i][1:9, 2:10, :])           # center crop 10x12 -> 8x8
    assert it.errors == 0
    again = [(x.clone(), y.clone()) for x, y in it]                  # second epoch, no shuffle: identical
    assert all(torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) for a, b in zip(seen, again))
    it.close()
    # shards partition the data; shuffling is deterministic per (seed, epoch) and independent of the thread count
    labels = []
    for part in range(2):
        sh = recordio.NativeImageRecordIter(str(tmp_path / "d.rec"), 4, (3, 8, 8), num_parts=2, part_index=part,
                                            round_batch=False)
        labels.append(sum(len(y) for _, y in sh))
        sh.close()
    assert labels == [16, 16]                                        # 19 and 18 samples, partial batches dropped
    a = recordio.NativeImageRecordIter(str(tmp_path / "d.rec"), 8, (3, 8, 8), shuffle=True, seed=7, preprocess_threads=1,
                                       rand_mirror=True, rand_crop=True)
    b = recordio.NativeImageRecordIter(str(tmp_path / "d.rec"), 8, (3, 8, 8), shuffle=True, seed=7, preprocess_threads=4,
                                       rand_mirror=True, rand_crop=True)
    ea, eb = [(x.clone(), y.clone()) for x, y in a], [(x.clone(), y.clone()) for x, y in b]
    assert all(torch.equal(p[0], q[0]) and torch.equal(p[1], q[1]) for p, q in zip(ea, eb))
    assert [int(v) for _, y in ea for v in y][:8] != [i % 5 for i in range(8)]            # really shuffled
    e2 = [(x.clone(), y.clone()) for x, y in a]                      # next epoch: a different order
    assert not all(torch.equal(p[1], q[1]) for p, q in zip(ea, e2))
    a.close()
    b.close()
    # larger target than source: zero padded, image centred; fewer channels requested than stored
    big = recordio.NativeImageRecordIter(str(tmp_path / "d.rec"), 2, (1, 16, 16), round_batch=False)
    x, _ = next(iter(big))
    assert x.shape == (2, 16, 16, 1) and int(x[0, 0, 0, 0]) == 0 and np.array_equal(x[0, 3:13, 2:14, 0].numpy(), imgs[0][:, :, 0])
    big.close()


def test_im2rec_list_and_record_roundtrip(tmp_path):
    import numpy as np
    from PIL import Image
    from deeplearning_cfn_b200.utils import im2rec, recordio
    root = tmp_path / "imgs"
    rng = np.random.default_rng(0)
    for cls in ("cat", "dog", "eel"):
        (root / cls).mkdir(parents=True)
        for i in range(4):
            Image.fromarray(rng.integers(0, 256, (20, 30, 3), dtype=np.uint8)).save(root / cls / f"{i}.png")
    prefix = str(tmp_path / "data")
    assert im2rec.main([prefix, str(root), "--list", "--train-ratio", "0.75"]) == 0
    train = list(im2rec.read_list(prefix + "_train.lst"))
    val = list(im2rec.read_list(prefix + "_val.lst"))
    assert len(train) == 9 and len(val) == 3 and {int(l[0]) for _, _, l in train + val} == {0, 1, 2}
    assert im2rec.main([prefix, str(root), "--resize", "16", "--center-crop"]) == 0
    it = recordio.NativeImageRecordIter(prefix + "_train.rec", 3, (3, 16, 16), round_batch=False)
    seen = 0
    for x, y in it:
        assert x.shape == (3, 16, 16, 3) and set(y.tolist()) <= {0, 1, 2}
        seen += len(y)
    assert seen == 9 and it.errors == 0
    it.close()
    h, img = recordio.unpack_img(recordio.MXIndexedRecordIO(prefix + "_val.idx", prefix + "_val.rec", "r").read())
    assert img.shape == (16, 16, 3)


def test_encoded_image_records_jpeg_png(tmp_path):
    """pack_img(.jpg/.png) writes the reference's encoded payloads; unpack_img detects them by magic number; the native ring
    refuses them with a clear error and image_record_iter falls back to the codec-backed python iterator."""
    import numpy as np
    from PIL import Image
    from deeplearning_cfn_b200.utils import im2rec, recordio
    yy, xx = np.mgrid[0:24, 0:32]
    smooth = np.stack([yy * 8, xx * 6, (yy + xx) * 4], -1).astype(np.uint8)           # JPEG-friendly content
    hdr = recordio.IRHeader(0, 3.0, 7, 0)
    h, img = recordio.unpack_img(recordio.pack_img(hdr, smooth, img_fmt=".png"))
    assert h.label == 3.0 and h.id == 7 and np.array_equal(img, smooth)               # PNG is lossless
    packed = recordio.pack_img(hdr, smooth, quality=95, img_fmt=".jpg")
    assert recordio.image_format(recordio.unpack(packed)[1]) == ".jpg"
    h, img = recordio.unpack_img(packed)
    assert img.shape == smooth.shape and np.abs(img.astype(int) - smooth.astype(int)).mean() < 4.0
    assert len(recordio.pack_img(hdr, smooth, quality=30, img_fmt=".jpg")) < len(packed)
    # a JPEG written by another library (what the reference's im2rec / OpenCV produce) decodes too
    import io
    bio = io.BytesIO()
    Image.fromarray(smooth).save(bio, format="JPEG", quality=90)
    h, img = recordio.unpack_img(recordio.pack(hdr, bio.getvalue()))
    assert img.shape == smooth.shape and np.abs(img.astype(int) - smooth.astype(int)).mean() < 5.0

    root = tmp_path / "imgs"
    for cls in ("a", "b"):
        (root / cls).mkdir(parents=True)
        for i in range(3):
            Image.fromarray(np.roll(smooth, i * 3, 1)).save(root / cls / f"{i}.png")
    prefix = str(tmp_path / "enc")
    assert im2rec.main([prefix, str(root), "--list"]) == 0
    assert im2rec.main([prefix, str(root), "--encoding", ".jpg", "--quality", "90"]) == 0
    with pytest.raises(ValueError, match="encoded"):
        recordio.NativeImageRecordIter(prefix + ".rec", 2, (3, 24, 32))
    it = recordio.image_record_iter(prefix + ".rec", 2, (3, 16, 16), round_batch=False)
    assert isinstance(it, recordio.ImageRecordIter)
    n = 0
    for x, y in it:
        assert x.shape == (2, 16, 16, 3) and x.dtype == torch.uint8 and set(y.tolist()) <= {0, 1}
        n += len(y)
    assert n == 6
    # the center crop matches the decoded image's middle
    it2 = recordio.ImageRecordIter(prefix + ".rec", 1, (3, 16, 16), preprocess_threads=1)
    _, full = recordio.unpack_img(recordio.MXIndexedRecordIO(prefix + ".idx", prefix + ".rec", "r").read_idx(it2.keys[0]))
    x, _ = next(iter(it2))
    assert np.array_equal(x[0].numpy(), full[4:20, 8:24])
    # a finished epoch: iterating again starts the next epoch instead of blocking on an empty queue
    it3 = recordio.ImageRecordIter(prefix + ".rec", 2, (3, 16, 16), shuffle=True, seed=3)
    e1 = [y.tolist() for _, y in it3]
    e2 = [y.tolist() for _, y in it3]
    assert len(e1) == len(e2) == 3 and sorted(sum(e1, [])) == sorted(sum(e2, [])) and len(it3) == 3
    it.close(), it2.close(), it3.close()
    # raw files still take the native ring
    assert im2rec.main([prefix, str(root)]) == 0
    nat = recordio.image_record_iter(prefix + ".rec", 2, (3, 24, 32))
    assert isinstance(nat, recordio.NativeImageRecordIter)
    nat.close()


def test_trainer_handles_parameters_without_gradient():
    """A parameter that takes no part in the step (its hook never fires) is gathered as zeros when the bucket is flushed."""
    from deeplearning_cfn_b200.parallel.trainer import DataParallelTrainer

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.used = torch.nn.Linear(6, 4)
            self.unused = torch.nn.Linear(6, 4)

        def forward(self, x):
            return self.used(x)

    torch.manual_seed(0)
    net = Net()
    tr = DataParallelTrainer(net, mode="local", optimizer="sgd", lr=0.5, wd=0.0, bucket_bytes=64)
    w_unused = net.unused.weight.detach().clone()
    w_used = net.used.weight.detach().clone()
    x, y = torch.randn(8, 6), torch.randint(0, 4, (8,))
    for _ in range(2):
        loss = tr.step(x, y)
    assert torch.isfinite(loss)
    assert torch.equal(net.unused.weight.detach(), w_unused)          # zero gradient, no weight decay: untouched
    assert not torch.equal(net.used.weight.detach(), w_used)
    # and the flat gradient buffer really holds zeros for it (not stale values of an earlier step)
    gv = tr.flat._grad_view[id(net.unused.weight)]
    assert float(gv.abs().sum()) == 0.0


def test_horovod_style_collectives_over_gloo():
    from conftest import run_job
    cmd = [sys.executable, "-m", "deeplearning_cfn_b200.launch.mpirun", "-np", "3", sys.executable,
           os.path.join(ROOT, "tests/mp/hvd_api_worker.py")]
    r = run_job(cmd, env=dict(os.environ, PYTHONPATH=ROOT), timeout=240)
    assert r.returncode == 0 and r.stdout.count("HVD_API_OK") == 3, r.stdout[-2000:] + r.stderr[-3000:]


def test_initializers_by_name_pattern_and_statistics(tmp_path):
    from deeplearning_cfn_b200.kv import initializer as init
    from deeplearning_cfn_b200.kv.module import Module
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Conv2d(16, 32, 3), torch.nn.BatchNorm2d(32), torch.nn.Flatten(), torch.nn.Linear(32, 10))
    init.create("xavier", rnd_type="gaussian", factor_type="in", magnitude=2).init_module(net)
    w = net[0].weight
    assert abs(float(w.std()) - (2.0 / (16 * 9)) ** 0.5) < 0.01 and float(net[0].bias.abs().sum()) == 0.0
    assert torch.equal(net[1].weight, torch.ones(32)) and torch.equal(net[1].running_var, torch.ones(32))   # gamma / moving_var
    assert float(net[1].bias.abs().sum()) == 0.0 and float(net[1].running_mean.abs().sum()) == 0.0
    init.Uniform(0.5)("fc_weight", net[3].weight.data)
    assert float(net[3].weight.abs().max()) <= 0.5 and float(net[3].weight.abs().max()) > 0.3
    o = torch.empty(8, 20)
    init.Orthogonal(scale=1.0)("w", o)
    assert torch.allclose(o @ o.t(), torch.eye(8), atol=1e-4)
    m = init.Mixed([".*bias", ".*"], [init.Constant(0.25), init.Normal(0.02)])
    b, wt = torch.empty(5), torch.empty(5, 5)
    m("layer_bias_custom", b)          # Constant's name rule would zero a *bias* name; this one is not a plain 'bias' suffix
    m("layer_weight", wt)
    assert torch.equal(b, torch.full((5,), 0.25)) and 0.005 < float(wt.std()) < 0.04
    with pytest.raises(ValueError):
        init.create("xavier")("vector_w", torch.empty(7))
    # Load: from a checkpoint written by model.save_checkpoint, with a default for what is missing
    from deeplearning_cfn_b200.kv import model as model_mod
    model_mod.save_checkpoint(str(tmp_path / "ck"), 1, net, {"3.weight": net[3].weight.detach().clone()}, {})
    net2 = torch.nn.Sequential(torch.nn.Conv2d(16, 32, 3), torch.nn.BatchNorm2d(32), torch.nn.Flatten(), torch.nn.Linear(32, 10))
    init.Load(str(tmp_path / "ck-0001.params"), default_init=init.Zero()).init_module(net2)
    assert torch.equal(net2[3].weight, net[3].weight) and float(net2[0].weight.abs().sum()) == 0.0
    mod = Module(torch.nn.Sequential(torch.nn.Linear(4, 4)))
    mod.bind([("data", (2, 4))])
    mod.init_params(initializer="one")
    assert torch.equal(mod.symbol[0].weight.detach(), torch.ones(4, 4)) and float(mod.symbol[0].bias.abs().sum()) == 0.0


def test_module_fine_grained_api_and_params_roundtrip(tmp_path):
    from deeplearning_cfn_b200.kv import metric
    from deeplearning_cfn_b200.kv.module import Module
    from deeplearning_cfn_b200.utils.io import DataBatch
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 8), torch.nn.BatchNorm1d(8), torch.nn.ReLU(), torch.nn.Linear(8, 3))
    mod = Module(net)
    mod.bind([("data", (16, 6))])
    mod.init_params(initializer="xavier")
    mod.init_optimizer("local", "sgd", (("learning_rate", 0.1),))
    x, y = torch.randn(16, 6), torch.randint(0, 3, (16,))
    m = metric.create("acc")
    w0 = net[0].weight.detach().clone()
    mod.forward(DataBatch([x], [y]), is_train=True)
    mod.update_metric(m, [y])
    mod.backward()
    mod.update()
    assert mod.get_outputs()[0].shape == (16, 3) and m.num_inst == 16 and not torch.equal(net[0].weight.detach(), w0)
    mod.forward(DataBatch([x], [y]), is_train=False)
    assert not mod.get_outputs()[0].requires_grad
    arg, aux = mod.get_params()
    assert set(arg) == {"0.weight", "0.bias", "1.weight", "1.bias", "3.weight", "3.bias"} and "1.running_mean" in aux
    mod.save_params(str(tmp_path / "p.params"))
    net2 = torch.nn.Sequential(torch.nn.Linear(6, 8), torch.nn.BatchNorm1d(8), torch.nn.ReLU(), torch.nn.Linear(8, 3))
    mod2 = Module(net2)
    mod2.load_params(str(tmp_path / "p.params"))
    mod2.forward(DataBatch([x], [y]), is_train=False)
    assert torch.allclose(mod2.get_outputs()[0], mod.get_outputs()[0])
    with pytest.raises(ValueError):
        mod2.set_params({"nope": torch.zeros(1)}, allow_missing=True)


def test_graphed_step_falls_back_to_eager_where_a_graph_cannot_be_used():
    """trainer.capture on CPU parameters returns a pass-through that says why and trains exactly like trainer.step."""
    from deeplearning_cfn_b200.parallel.trainer import DataParallelTrainer
    from deeplearning_cfn_b200.parallel.graph import GraphedStep, _flatten, _map

    def make():
        torch.manual_seed(3)
        return torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4))
    x, y = torch.randn(6, 8), torch.randint(0, 4, (6,))
    a = DataParallelTrainer(make(), mode="local", optimizer="sgd_mom", lr=0.1)
    b = DataParallelTrainer(make(), mode="local", optimizer="sgd_mom", lr=0.1)
    step = a.capture(x, y)
    assert isinstance(step, GraphedStep) and not step.graphed and "CUDA" in step.reason and "eager" in repr(step)
    for _ in range(3):
        la, lb = step(x, y), b.step(x, y)
        assert torch.equal(la, lb)
    assert a.step_count == b.step_count == 3
    for pa, pb in zip(a.model.parameters(), b.model.parameters()):
        assert torch.equal(pa, pb)
    # nested inputs (BERT passes tuples) are cloned / flattened structurally
    nested = ((x, x + 1), [y])
    cl = _map(nested, lambda t: t.clone())
    assert isinstance(cl[0], tuple) and isinstance(cl[1], list) and cl[0][1].data_ptr() != nested[0][1].data_ptr()
    assert [t.shape for t in _flatten(nested, [])] == [x.shape, x.shape, y.shape]
    assert GraphedStep._signature_inputs(x, y) != GraphedStep._signature_inputs(x[:3], y[:3])


def test_roi_align_and_nms_semantics_match_torchvision():
    """ops/roi.py: the plain-torch statements of ROIAlign / NMS (what the CUDA kernels implement) against torchvision on
    the host, incl. adaptive sampling, aligned / legacy modes, boxes leaving the map, list-of-boxes input, batched NMS."""
    import torchvision.ops as tvo
    from deeplearning_cfn_b200.ops import roi
    torch.manual_seed(0)
    x = torch.randn(2, 5, 12, 9, dtype=torch.float64)
    rois = torch.tensor([[0, 1.2, 0.7, 6.8, 9.9], [1, -3.0, -2.0, 4.0, 5.0], [1, 3.3, 2.2, 30.0, 40.0], [0, 2.0, 2.0, 2.0, 2.0],
                         [1, 0.0, 0.0, 8.5, 11.5]], dtype=torch.float64)
    for aligned in (False, True):
        for sr in (2, -1):
            for scale in (1.0, 0.5):
                ref = tvo.roi_align(x, rois, (3, 4), scale, sr, aligned)
                got = roi.roi_align_ref(x, rois, (3, 4), scale, sr, aligned)
                assert torch.allclose(got, ref, atol=1e-5), (aligned, sr, scale, (got - ref).abs().max())
    lst = [rois[rois[:, 0] == 0][:, 1:], rois[rois[:, 0] == 1][:, 1:]]
    a = roi.roi_align(x, lst, 3, 1.0, 2, True)                       # CPU tensors: torchvision path, list input
    b = roi.roi_align_ref(x, lst, 3, 1.0, 2, True)
    assert a.shape == (5, 5, 3, 3) and torch.allclose(a, b, atol=1e-5)
    # gradient of the reference statement = torchvision's backward
    xr = x.clone().requires_grad_(True)
    xt = x.clone().requires_grad_(True)
    gy = torch.randn(5, 5, 3, 4, dtype=torch.float64)
    roi.roi_align_ref(xr, rois, (3, 4), 1.0, 2, True).backward(gy)
    tvo.roi_align(xt, rois, (3, 4), 1.0, 2, True).backward(gy)
    assert torch.allclose(xr.grad, xt.grad, atol=1e-5)

    g = torch.Generator().manual_seed(1)
    xy = torch.rand(300, 2, generator=g) * 50
    boxes = torch.cat([xy, xy + 2 + torch.rand(300, 2, generator=g) * 20], 1)
    scores = torch.rand(300, generator=g)
    for thr in (0.3, 0.5, 0.7):
        assert torch.equal(roi.nms_ref(boxes, scores, thr), tvo.nms(boxes, scores, thr))
        assert torch.equal(roi.nms(boxes, scores, thr), tvo.nms(boxes, scores, thr))
    cls = torch.randint(0, 4, (300,), generator=g)
    assert torch.equal(roi.batched_nms(boxes, scores, cls, 0.5), tvo.batched_nms(boxes, scores, cls, 0.5))
    assert roi.nms(boxes[:0], scores[:0], 0.5).numel() == 0 and roi.backend() == "torchvision"


def test_rec2idx_and_parse_log_tools(tmp_path, capsys):
    import numpy as np
    from deeplearning_cfn_b200.__main__ import main as cli
    from deeplearning_cfn_b200.utils import parse_log, recordio
    rec_path = str(tmp_path / "d.rec")
    w = recordio.MXRecordIO(rec_path, "w")
    for i in range(5):
        w.write(recordio.pack_img(recordio.IRHeader(0, float(i), 10 + i, 0), np.full((4, 4, 3), i, dtype=np.uint8)))
    w.close()
    assert cli(["rec2idx", rec_path]) == 0
    r = recordio.MXIndexedRecordIO(str(tmp_path / "d.idx"), rec_path, "r")
    assert sorted(r.keys) == [10, 11, 12, 13, 14]
    h, img = recordio.unpack_img(r.read_idx(13))
    assert h.label == 3.0 and int(img[0, 0, 0]) == 3
    it = recordio.NativeImageRecordIter(rec_path, 5, (3, 4, 4), round_batch=False)     # the rebuilt index feeds the native reader
    x, y = next(iter(it))
    assert sorted(y.tolist()) == [0, 1, 2, 3, 4]
    it.close()

    log = tmp_path / "train.log"
    log.write_text("\n".join([
        "INFO Epoch[0] Batch [20] Speed: 1000.0 samples/sec accuracy=0.1",
        "INFO Epoch[0] Batch [40] Speed: 3000.0 samples/sec accuracy=0.2",
        "INFO Epoch[0] Train-accuracy=0.250000", "INFO Epoch[0] Time cost=12.500", "INFO Epoch[0] Validation-accuracy=0.300000",
        "INFO Epoch[1] Train-accuracy=0.500000", "INFO Epoch[1] Train-top_k_accuracy_5=0.9", "INFO Epoch[1] Time cost=11.000",
        "noise line", "INFO Epoch[1] Validation-accuracy=nan"]))
    cols, rows = parse_log.parse(open(log))
    assert cols == ["epoch", "train-accuracy", "train-top_k_accuracy_5", "val-accuracy", "val-top_k_accuracy_5", "time", "speed"]
    assert rows[0][:2] == [0, 0.25] and rows[0][3] == 0.3 and rows[0][5] == 12.5 and rows[0][6] == 2000.0
    assert rows[1][1] == 0.5 and rows[1][2] == 0.9 and rows[1][6] is None
    assert cli(["parse-log", str(log), "--metric-names", "accuracy"]) == 0
    out = capsys.readouterr().out
    assert "| epoch | train-accuracy | val-accuracy | time | speed |" in out and "| 0 | 0.25 | 0.3 | 12.5 | 2000 |" in out


def test_diagnose_report(capsys):
    from deeplearning_cfn_b200.__main__ import main as cli
    assert cli(["diagnose", "--json"]) == 0
    import json
    info = json.loads(capsys.readouterr().out)
    assert info["package"]["cuda_library"]["present"] and info["package"]["cuda_library"]["loads"] is True
    assert info["package"]["runtime_library"]["loads"] is True and "gloo" in info["torch"]["distributed_backends"]
    assert {"platform", "python", "torch", "gpus", "tools", "environment"} <= set(info)


def test_cpu_thread_budget_is_split_between_colocated_processes(monkeypatch):
    from deeplearning_cfn_b200.utils import cpu
    before = torch.get_num_threads()
    try:
        cores = os.cpu_count() or 1
        assert cpu.share_cpu_threads(1, {}) is None                                        # alone: untouched
        assert cpu.share_cpu_threads(4, {"OMP_NUM_THREADS": "3"}) is None                  # user pinned it
        assert cpu.share_cpu_threads(4, {"DLCFN_KEEP_THREADS": "1"}) is None
        assert cpu.share_cpu_threads(2, {}) == max(1, cores // 2) == torch.get_num_threads()
        assert cpu.share_cpu_threads(8, {"LOCAL_WORLD_SIZE": "4"}) == max(1, cores // 4)   # 2 nodes x 4 ranks
        env = {"DMLC_JOB_CLUSTER": "local", "DMLC_NUM_WORKER": "2", "DMLC_NUM_SERVER": "2"}
        assert cpu.share_cpu_threads(1, env) == max(1, cores // 4)                          # local launcher: workers + servers
    finally:
        torch.set_num_threads(before)


def test_recordio_framing_property_python_and_native_agree(tmp_path):
    """Property test (hypothesis): arbitrary payloads - empty, unaligned lengths, payloads that CONTAIN the record magic at
    aligned and unaligned offsets (the dmlc writer splits those into continuation records) - written by either writer are
    read back identically by either reader."""
    import struct
    from hypothesis import given, settings, strategies as st
    from deeplearning_cfn_b200.utils import recordio
    magic = struct.pack("<I", recordio.MAGIC)
    chunk = st.one_of(st.binary(min_size=0, max_size=40), st.just(magic), st.just(magic * 2), st.just(b"\x00" * 3 + magic))
    payload = st.lists(chunk, min_size=0, max_size=6).map(b"".join)
    counter = [0]

    @settings(max_examples=40, deadline=None)
    @given(st.lists(payload, min_size=1, max_size=6))
    def run(payloads):
        counter[0] += 1
        for writer in ("python", "native"):
            path = str(tmp_path / f"p{counter[0]}_{writer}.rec")
            if writer == "python":
                w = recordio.MXRecordIO(path, "w")
                for b in payloads:
                    w.write(b)
                w.close()
            else:
                w = recordio.NativeRecordIO(None, path, "w")
                for i, b in enumerate(payloads):
                    w.write(i, b)
                w.close()
            r = recordio.MXRecordIO(path, "r")
            got = []
            while True:
                b = r.read()
                if b is None:
                    break
                got.append(b)
            r.close()
            assert got == payloads, (writer, "python reader")
            n = recordio.NativeRecordIO(None, path, "r")
            assert len(n) == len(payloads) and [n.read(i) for i in range(len(n))] == payloads, (writer, "native reader")
            n.close()
    run()


def test_slicing_helpers_partition_their_ranges():
    """Properties of the three slicing rules: executor-group batch slices, inter-node rail slices of a bucket, server shards
    of a bucket - each is an exact partition, rail / shard boundaries fall on 16-byte granules."""
    from types import SimpleNamespace
    from hypothesis import given, settings, strategies as st
    from deeplearning_cfn_b200.kv.module import decide_slices
    from deeplearning_cfn_b200.parallel.trainer import DataParallelTrainer

    @settings(max_examples=200, deadline=None)
    @given(st.integers(1, 512), st.lists(st.integers(1, 5), min_size=1, max_size=8))
    def batch(n, loads):
        try:
            sl = decide_slices(n, loads)
        except ValueError:
            assert n < 2 * len(loads) + sum(loads)        # only tiny batches may fail to give every device a row
            return
        assert sl[0].start == 0 and sl[-1].stop == n and all(a.stop == b.start for a, b in zip(sl, sl[1:]))
        assert all(s.stop > s.start for s in sl)
    batch()

    @settings(max_examples=200, deadline=None)
    @given(st.integers(1, 1 << 16).map(lambda k: k * 64), st.sampled_from([1, 2, 4, 8]), st.sampled_from([2, 4]))
    def rails(numel, node_size, esize):
        fake = SimpleNamespace(node_size=node_size)
        b = SimpleNamespace(numel=numel, group=SimpleNamespace(esize=esize))
        cuts = [DataParallelTrainer._rail_slice(fake, b, l) for l in range(node_size)]
        assert cuts[0][0] == 0 and cuts[-1][1] == numel and all(a[1] == c[0] for a, c in zip(cuts, cuts[1:]))
        assert all((lo * esize) % 16 == 0 for lo, _ in cuts)
        fake2 = SimpleNamespace(rank=0, servers=list(range(node_size)))
        shards = []
        for r in range(node_size):
            fake2.rank = r
            shards.append(DataParallelTrainer._shard(fake2, b))
        assert shards[0][0] == 0 and max(h for _, h in shards) >= numel and all(a[1] == c[0] or c == (numel, numel) or a[1] >= numel
                                                                                for a, c in zip(shards, shards[1:]))
    rails()


def test_native_recordio_reader_and_pipeline_survive_corrupted_files(tmp_path):
    """Truncated, bit-flipped, length-field-smashed and partly random .rec files through rec2idx, the native reader and the
    native image-record pipeline (decoder threads included): every outcome except a crash is acceptable - an error, a
    short read, skipped records.  tools/sanitize_rt.sh / ubsan_rt.sh run this under ASAN / UBSAN (-k recordio)."""
    import random
    import numpy as np
    from deeplearning_cfn_b200.utils import recordio as R
    rng = random.Random(7)
    rec, idx = str(tmp_path / "a.rec"), str(tmp_path / "a.idx")
    w = R.MXIndexedRecordIO(idx, rec, "w")
    for i in range(16):
        img = (np.random.RandomState(i).rand(20, 24, 3) * 255).astype(np.uint8)
        w.write_idx(i, R.pack_img(R.IRHeader(0, float(i % 5), i, 0), img, img_fmt=".raw"))
    w.close()
    good = open(rec, "rb").read()
    survived = 0
    for it in range(48):
        b = bytearray(good)
        k = it % 4
        if k == 0:
            b = b[:rng.randrange(0, len(b))]
        elif k == 1:
            for _ in range(rng.randrange(1, 20)):
                b[rng.randrange(len(b))] ^= 1 << rng.randrange(8)
        elif k == 2:
            off = rng.randrange(0, len(b) - 8)
            b[off:off + 4] = rng.choice([b"\xff\xff\xff\xff", b"\x00\x00\x00\x00", b"\xff\xff\xff\x7f", b"\x00\x00\x00\x80"])
        else:
            off = rng.randrange(0, len(b) - 64)
            b[off:off + 64] = rng.randbytes(64)
        p = str(tmp_path / f"f{it}.rec")
        with open(p, "wb") as f:
            f.write(bytes(b))
        try:
            R.rec2idx(p, p + ".idx")
            r = R.NativeRecordIO(p + ".idx", p, "r")
            for i in range(len(r)):
                try:
                    r.read(i)
                except Exception:                        # noqa: BLE001 - a damaged record may be refused
                    pass
            r.close()
            itr = R.NativeImageRecordIter(p, 4, (3, 16, 16), path_imgidx=p + ".idx", preprocess_threads=2, shuffle=False)
            for bi, _batch in enumerate(itr):
                if bi > 6:
                    break
            itr.close()
            survived += 1
        except Exception:                                # noqa: BLE001 - refusing the file is fine; crashing is not
            pass
    assert survived > 0


def test_host_storage_pool_under_concurrent_random_alloc_and_release():
    """Four threads allocate, touch and release blocks of awkward sizes at random from the pooled host storage
    (csrc/rt/storage.cpp), in both size-class strategies: nothing is lost or handed out twice (first / last byte written by
    the owner stay intact), and everything returns to the pool.  Clean under ASAN / UBSAN / TSAN (profiles/r2/sanitizers_r2.md)."""
    import random
    import threading
    from deeplearning_cfn_b200.utils import storage
    errors = []
    try:
        for kind, cut in (("Naive", 0), ("Round", 12)):
            storage.set_strategy(kind, cut)

            def worker(seed):
                r = random.Random(seed)
                live = []
                try:
                    for _ in range(250):
                        if live and r.random() < 0.45:
                            t, tag = live.pop(r.randrange(len(live)))
                            if int(t[0]) != tag or int(t[-1]) != tag:
                                errors.append(f"block of thread {seed} was overwritten")
                            storage.release(t)
                        else:
                            n = r.choice([1, 7, 64, 4095, 4096, 4097, 65536, r.randrange(1, 200000)])
                            t = storage.host_empty((n,), torch.uint8)
                            tag = (seed * 37 + len(live)) % 251
                            t[0] = tag
                            t[-1] = tag
                            live.append((t, tag))
                    for t, tag in live:
                        if int(t[0]) != tag or int(t[-1]) != tag:
                            errors.append(f"block of thread {seed} was overwritten")
                        storage.release(t)
                except Exception as e:               # noqa: BLE001
                    errors.append(repr(e))
            ts = [threading.Thread(target=worker, args=(s,)) for s in range(4)]
            [t.start() for t in ts]
            [t.join() for t in ts]
            assert not errors, errors[:3]
            assert storage.stats()["live_blocks"] == 0
            storage.release_all()
    finally:
        storage.set_strategy("Naive", 0)
