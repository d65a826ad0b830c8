"""Data-parallel engine / Horovod-style API / gluon-style Trainer on CPU."""
import os
import subprocess
import sys

import pytest
import torch

import deeplearning_cfn_b200.kv as kv
from deeplearning_cfn_b200.kv.trainer import Trainer

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_gloo_world2_data_parallel_matches_single_process():
    cmd = [sys.executable, "-m", "deeplearning_cfn_b200.launch.mpirun", "-np", "2", sys.executable,
           os.path.join(ROOT, "tests/mp/gloo_trainer_worker.py")]
    r = subprocess.run(cmd, env=dict(os.environ, PYTHONPATH=ROOT), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "GLOO_TRAINER_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


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
    r = subprocess.run([sys.executable, "-m", "deeplearning_cfn_b200", "launch", "-n", "2", "-s", "0", "--launcher", "local",
                        sys.executable, str(script)], env=dict(os.environ, PYTHONPATH=ROOT), capture_output=True, text=True,
                       timeout=180)
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
    assert mod.score(data) > 0.9
    mod.save_checkpoint(str(tmp_path / "m"), 120, save_optimizer_states=True)
    m2 = Module.load(str(tmp_path / "m"), 120, torch.nn.Sequential(torch.nn.Flatten(), torch.nn.Linear(192, 32), torch.nn.ReLU(), torch.nn.Linear(32, 4)))
    assert abs(m2.score(data) - mod.score(data)) < 1e-6


def test_sync_replicas_with_backup_worker():
    """TF SyncReplicasOptimizer semantics on the native PS: aggregate 2 of 3 replicas, drop stale gradients."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PYTHONPATH=root, MXNET_KVSTORE_BIGARRAY_BOUND="40")
    cmd = [sys.executable, "-m", "deeplearning_cfn_b200", "launch", "-n", "3", "-s", "2", "--launcher", "local",
           sys.executable, os.path.join(root, "tests/mp/sync_replicas_worker.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
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
