"""KVStore semantics (CPU): local store with fake devices, distributed store as N local processes over the
native transport, python-updater path, engine and fabric planner.  Mirrors the reference's tiers
(tests/python/unittest/test_kvstore.py, tests/nightly/dist_*_kvstore.py, tests/cpp/kvstore/gpu_topology_test.cc)."""
import os
import subprocess
import sys
import threading
import time

import pytest
import torch

import deeplearning_cfn_b200.kv as kv
from deeplearning_cfn_b200.utils import rt

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHAPE = (4, 4)


def _launch(nw, ns, script_args, env=None, timeout=180):
    e = dict(os.environ, PYTHONPATH=ROOT)
    e.update(env or {})
    cmd = [sys.executable, "-m", "deeplearning_cfn_b200", "launch", "-n", str(nw), "-s", str(ns), "--launcher", "local",
           sys.executable] + script_args
    from conftest import run_job
    return run_job(cmd, env=e, timeout=timeout)


# ----------------------------------------------------------------------------- local
@pytest.mark.parametrize("name", ["local", "device"])
def test_single_kv_pair_and_list(name):
    s = kv.create(name)
    assert s.type == name and s.rank == 0 and s.num_workers == 1
    s.init(3, torch.ones(SHAPE))
    s.push(3, torch.ones(SHAPE) * 4)
    out = torch.zeros(SHAPE)
    s.pull(3, out=out)
    assert torch.all(out == 4)                       # no updater: the pushed value replaces the stored one
    keys = [5, 7, 9]
    s.init(keys, [torch.ones(SHAPE)] * 3)
    s.push(keys, [torch.ones(SHAPE) * 2] * 3)
    outs = [torch.zeros(SHAPE) for _ in keys]
    s.pull(keys, out=outs)
    assert all(torch.all(o == 2) for o in outs)


def test_aggregator_over_fake_devices_and_updater():
    s = kv.create("local")
    ndev = 4
    s.init("w", torch.ones(SHAPE))
    s.push("w", [torch.ones(SHAPE)] * ndev)          # list of per-device values is summed
    outs = [torch.zeros(SHAPE) for _ in range(ndev)]
    s.pull("w", out=outs)
    assert all(torch.all(o == ndev) for o in outs)
    s._set_updater(lambda key, g, w: w.add_(g * 2))
    s.push("w", [torch.ones(SHAPE)] * ndev)
    s.pull("w", out=outs)
    assert all(torch.all(o == ndev + 2 * ndev) for o in outs)


def test_invalid_operations_raise():
    s = kv.create("local")
    s.init(1, torch.ones(2))
    with pytest.raises(kv.KVStoreError):
        s.init(1, torch.ones(2))
    with pytest.raises(kv.KVStoreError):
        s.push(2, torch.ones(2))
    with pytest.raises(kv.KVStoreError):
        s.pull(2, out=torch.ones(2))
    with pytest.raises(kv.KVStoreError):
        s.init("name", torch.ones(2))                # mixing int and str keys
    with pytest.raises(kv.KVStoreError):
        kv.create("bogus")
    with pytest.raises(kv.KVStoreError):
        kv.create("nccl").set_gradient_compression({"type": "2bit"})
    with pytest.raises(kv.KVStoreError):
        s.set_gradient_compression({"type": "1bit"})


def test_row_sparse_pull_and_optimizer_states(tmp_path):
    s = kv.create("local")
    s.init(0, torch.arange(20.).view(5, 4))
    out = torch.full((5, 4), -1.0)
    s.row_sparse_pull(0, out=out, row_ids=torch.tensor([3, 1, 3]))
    assert torch.equal(out[1], torch.arange(4., 8.)) and torch.equal(out[3], torch.arange(12., 16.))
    assert float(out[0].abs().sum()) == 0 and float(out[4].abs().sum()) == 0
    s.set_optimizer(kv.optimizer.create("sgd", learning_rate=0.1, momentum=0.9))
    s.push(0, torch.ones(5, 4))
    f = str(tmp_path / "states")
    s.save_optimizer_states(f, dump_optimizer=True)
    s2 = kv.create("local")
    s2.set_optimizer(kv.optimizer.create("sgd", learning_rate=0.1, momentum=0.9))
    s2.load_optimizer_states(f)
    assert torch.allclose(s2._updater.states[0], s._updater.states[0])


def test_local_2bit_compression_matches_quantisation_rule():
    s = kv.create("device")
    s.set_gradient_compression({"type": "2bit", "threshold": 0.5})
    s.init(0, torch.zeros(32))
    g = torch.linspace(-1, 1, 32)
    s.push(0, [g, g])
    out = torch.zeros(32)
    s.pull(0, out=out)
    q = torch.where(g >= 0.5, 0.5, torch.where(g <= -0.5, -0.5, 0.0))
    assert torch.allclose(out, 2 * q)


# ----------------------------------------------------------------------------- optimizers
@pytest.mark.parametrize("name", ["sgd", "nag", "signum", "adam", "adagrad", "rmsprop", "adadelta", "ftrl", "adamax",
                                  "nadam", "ftml", "dcasgd", "sgld", "lbsgd", "ccsgd", "test"])
def test_every_optimizer_steps_and_decreases_quadratic(name):
    torch.manual_seed(0)
    o = kv.optimizer.create(name, learning_rate=0.05) if name != "test" else kv.optimizer.create(name, rescale_grad=-0.05)
    if hasattr(o, "momentum") and name in ("sgd", "nag", "dcasgd", "lbsgd", "ccsgd"):
        o.momentum = 0.9
    u = kv.optimizer.get_updater(o)
    w = torch.full((16,), 3.0)
    f0 = float((w ** 2).sum())
    for _ in range(200):
        u(0, 2 * w, w)
    assert torch.isfinite(w).all()
    if name != "sgld":
        assert float((w ** 2).sum()) < f0


def test_sgd_matches_reference_formula_and_multi_precision():
    o = kv.optimizer.create("sgd", learning_rate=0.1, momentum=0.9, wd=0.01, rescale_grad=0.5, clip_gradient=1.0,
                            multi_precision=True)
    u = kv.optimizer.get_updater(o)
    w = torch.ones(8, dtype=torch.bfloat16)
    w32, m = torch.ones(8), torch.zeros(8)
    for i in range(4):
        g = torch.full((8,), 3.0 + i)
        u(0, g.bfloat16(), w)
        gg = (g * 0.5).clamp(-1, 1)
        m = 0.9 * m - 0.1 * 0.01 * w32 - 0.1 * gg
        w32 = w32 + m
    master, _ = u.states[0]
    assert torch.allclose(master, w32, atol=1e-6) and torch.allclose(w.float(), w32.bfloat16().float())
    assert "kind=sgd_mom" in o.native_config()


def test_lr_schedulers():
    f = kv.optimizer.FactorScheduler(step=10, factor=0.5, base_lr=1.0)
    assert f(5) == 1.0 and f(10) == 0.5 and f(25) == 0.25
    m = kv.optimizer.MultiFactorScheduler(step=[5, 10], factor=0.1, base_lr=1.0)
    assert m(5) == 1.0 and abs(m(6) - 0.1) < 1e-12 and abs(m(11) - 0.01) < 1e-12
    p = kv.optimizer.PolyScheduler(max_update=100, base_lr=1.0, pwr=2)
    assert p(0) == 1.0 and p(100) == 0.0 and abs(p(50) - 0.25) < 1e-12
    with pytest.raises(ValueError):
        kv.optimizer.MultiFactorScheduler(step=[5, 5])


# ----------------------------------------------------------------------------- distributed (local processes)
def test_dist_sync_world2_two_servers():
    r = _launch(2, 2, [os.path.join(ROOT, "tests/mp/dist_kvstore_worker.py"), "dist_sync"])
    assert r.returncode == 0 and r.stdout.count("SYNC_OK") == 2, r.stdout[-2000:] + r.stderr[-3000:]


def test_dist_sync_seven_workers_seven_servers():
    # the reference CI shape: launch.py -n 7 --launcher local (ci/docker/runtime_functions.sh:916-931)
    r = _launch(7, 7, [os.path.join(ROOT, "tests/mp/dist_kvstore_worker.py"), "dist_device_sync"], timeout=300)
    assert r.returncode == 0 and r.stdout.count("SYNC_OK") == 7, r.stdout[-2000:] + r.stderr[-3000:]


@pytest.mark.parametrize("ns", [2, 3])
def test_dist_store_regressions_big_keys_dtype_and_staging(ns):
    # advisor round 1: 2-bit shards not on word boundaries, row-sparse rows across the split, bf16 grad on an fp32 key,
    # staging buffer reused under a queued push
    r = _launch(2, ns, [os.path.join(ROOT, "tests/mp/dist_regress_worker.py"), "store"])
    assert r.returncode == 0 and r.stdout.count("REGRESS_OK") == 2, r.stdout[-2000:] + r.stderr[-3000:]


def test_dist_trainer_rescale_grad_and_lr_reach_the_servers():
    r = _launch(2, 1, [os.path.join(ROOT, "tests/mp/dist_regress_worker.py"), "trainer"])
    assert r.returncode == 0 and r.stdout.count("TRAINER_OK") == 2, r.stdout[-2000:] + r.stderr[-3000:]


def test_c_level_store_api_local_and_dist(tmp_path):
    """A C program on the handle-based store API of libdlcfn_rt.so (MXKVStore* counterpart): local store with an updater, then
    the same binary as scheduler / server / 2 workers of a dist_sync job under the local launcher."""
    import shutil
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    libdir = os.path.join(ROOT, "deeplearning_cfn_b200", "_lib")
    exe = str(tmp_path / "kvcapi")
    subprocess.run([cc, "-O1", "-o", exe, os.path.join(ROOT, "tests/c/kvstore_capi_test.c"), "-L", libdir, "-ldlcfn_rt",
                    f"-Wl,-rpath,{libdir}"], check=True)
    r = subprocess.run([exe, "local"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "CAPI_OK" in r.stdout, r.stdout + r.stderr
    e = dict(os.environ, PYTHONPATH=ROOT)
    cmd = [sys.executable, "-m", "deeplearning_cfn_b200", "launch", "-n", "2", "-s", "1", "--launcher", "local", exe, "dist_sync"]
    from conftest import run_job
    r = run_job(cmd, env=e, timeout=120)
    assert r.returncode == 0 and r.stdout.count("CAPI_OK") == 2, r.stdout[-2000:] + r.stderr[-3000:]


def test_dist_async():
    r = _launch(3, 2, [os.path.join(ROOT, "tests/mp/dist_kvstore_worker.py"), "dist_async"])
    assert r.returncode == 0 and r.stdout.count("ASYNC_OK") == 3, r.stdout[-2000:] + r.stderr[-3000:]


def test_native_server_optimizers_match_python_definitions():
    r = _launch(2, 2, [os.path.join(ROOT, "tests/mp/dist_kvstore_worker.py"), "opt_parity"])
    assert r.returncode == 0 and r.stdout.count("OPT_OK") == 2, r.stdout[-2000:] + r.stderr[-3000:]


@pytest.mark.parametrize("name,kw", [("nag", dict(momentum=0.9, wd=0.01)), ("signum", dict(momentum=0.9, wd_lh=0.01)),
                                     ("signum", dict(momentum=0.0)), ("adagrad", dict(wd=0.01)),
                                     ("rmsprop", dict(wd=0.01, clip_weights=2.0)), ("adadelta", dict(wd=0.001)),
                                     ("ftrl", dict(lamda1=0.01, wd=0.01)), ("adamax", dict(wd=0.01)),
                                     ("sgd", dict(momentum=0.9, wd=0.01)), ("adam", dict(wd=0.01))])
def test_flat_reference_optimizer_matches_python_definitions(name, kw):
    """ops.reference.fused_optimizer is what the CUDA kernels are tested against; pin it to kv/optimizer.py."""
    from deeplearning_cfn_b200.ops import reference as ref_ops
    torch.manual_seed(3)
    o = kv.optimizer.create(name, learning_rate=0.05, rescale_grad=0.5, clip_gradient=1.5, **kw)
    cfg = dict(x.split("=") for x in o.native_config().split(";"))
    kind = cfg.pop("kind")
    hp = {k: float(v) for k, v in cfg.items()}
    u = kv.optimizer.get_updater(o)
    w_py = torch.randn(50)
    master = w_py.clone()
    s1, s2 = torch.zeros(50), torch.zeros(50)
    for t in range(1, 5):
        g = torch.randn(50) * 2
        u(0, g, w_py)
        lr = hp["lr"]
        if kind == "adam":
            lr = lr * (1 - hp["beta2"] ** t) ** 0.5 / (1 - hp["beta1"] ** t)
        elif kind == "adamax":
            lr = lr / (1 - hp["beta1"] ** t)
        ref_ops.fused_optimizer(kind, g, master, s1, s2, None, **{**hp, "lr": lr})
        assert torch.allclose(master, w_py, rtol=1e-5, atol=1e-6), (name, t)


def test_dist_sync_python_updater_on_server():
    r = _launch(2, 1, [os.path.join(ROOT, "tests/mp/dist_kvstore_worker.py"), "dist_sync"], env={"DLCFN_PS_PY_UPDATER": "1"})
    assert r.returncode == 0 and r.stdout.count("SYNC_OK") == 2, r.stdout[-2000:] + r.stderr[-3000:]


def test_dist_sync_survives_dropped_messages_with_resend():
    # fault injection: 10 % of received data messages are dropped; at-least-once delivery repairs it
    r = _launch(2, 2, [os.path.join(ROOT, "tests/mp/dist_kvstore_worker.py"), "dist_sync"],
                env={"PS_DROP_MSG": "10", "PS_RESEND": "1", "PS_RESEND_TIMEOUT": "200"}, timeout=300)
    assert r.returncode == 0 and r.stdout.count("SYNC_OK") == 2, r.stdout[-2000:] + r.stderr[-3000:]


# ----------------------------------------------------------------------------- engine
def test_engine_orders_writers_and_readers_and_priorities():
    e = rt.Engine(nthreads=4)
    v = e.new_var()
    log = []
    for i in range(20):
        e.push(lambda i=i: (time.sleep(0.001), log.append(i)), writes=(v,))
    e.wait_for_var(v)
    assert log == list(range(20))                                  # writers of one var run in program order
    a, b = e.new_var(), e.new_var()
    seen = []
    gate = threading.Event()
    e.push(lambda: gate.wait(5), writes=(a,))
    for i in range(6):
        e.push(lambda i=i: seen.append(("r", i)), reads=(a,), writes=())
    e.push(lambda: seen.append(("w", 0)), writes=(a,))
    e.push(lambda: seen.append(("other", 0)), writes=(b,))
    time.sleep(0.05)
    assert seen == [("other", 0)]                                   # independent var is not blocked
    gate.set()
    e.wait_all()
    assert seen[-1] == ("w", 0) and sorted(seen[1:7]) == [("r", i) for i in range(6)]   # readers between writers
    # priority: with one worker thread busy, higher priority ops overtake lower ones
    e1 = rt.Engine(nthreads=1)
    g2, order = threading.Event(), []
    e1.push(lambda: g2.wait(5), writes=(e1.new_var(),))
    for p in (0, -5, 10, 3):
        e1.push(lambda p=p: order.append(p), writes=(e1.new_var(),), priority=p)
    g2.set()
    e1.wait_all()
    assert order == [10, 3, 0, -5]
    e.close()
    e1.close()


def test_engine_propagates_exceptions_and_naive_mode():
    e = rt.Engine(nthreads=2)
    v = e.new_var()
    e.push(lambda: (_ for _ in ()).throw(ValueError("boom")), writes=(v,))
    with pytest.raises(ValueError):
        e.wait_for_var(v)
    e.close()
    n = rt.Engine(naive=True)
    out = []
    n.push(lambda: out.append(1), writes=(n.new_var(),))
    assert out == [1]                                               # NaiveEngine: synchronous
    n.close()


# ----------------------------------------------------------------------------- fabric planner (GPU-free)
def _dgx1():
    """Hybrid cube mesh of 8 GPUs: NVLink inside and between two quads, PCIe (weight 1) otherwise."""
    n = 8
    w = [[1 if i != j else 0 for j in range(n)] for i in range(n)]
    nv = [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3), (4, 5), (4, 6), (4, 7), (5, 6), (5, 7), (6, 7),
          (0, 4), (1, 5), (2, 6), (3, 7)]
    for a, b in nv:
        w[a][b] = w[b][a] = 3
    return w


def _check_tree(parents, root, n):
    assert parents[root] == -1 and all(p >= 0 for i, p in enumerate(parents) if i != root)
    for i in range(n):                                               # every node reaches the root: spanning tree
        cur, hops = i, 0
        while cur != root:
            cur = parents[cur]
            hops += 1
            assert hops <= n
    depth = max(len(_path(parents, i, root)) for i in range(n))
    assert depth <= 4                                                # ~log2(8) + 1 levels


def _path(parents, i, root):
    p = []
    while i != root:
        p.append(i)
        i = parents[i]
    return p


def test_planner_backtracking_mode_uses_only_existing_links(monkeypatch):
    """MXNET_KVSTORE_TREE_BACKTRACK=1 (gpu_topology.h:789-957): every edge of every tree is a direct link, every tree spans all
    GPUs with the doubling depth, on the DGX-1 hybrid cube mesh and on a sparse ring-like fabric where KL may pair unlinked GPUs."""
    monkeypatch.setenv("MXNET_KVSTORE_TREE_BACKTRACK", "1")
    ring = [[0] * 8 for _ in range(8)]
    for i in range(8):
        for d in (1, 2):
            ring[i][(i + d) % 8] = ring[(i + d) % 8][i] = 2 if d == 1 else 1
    for link in (_dgx1(), ring):
        n = len(link)
        trees = rt.plan_trees(link)
        for root, parents in enumerate(trees):
            _check_tree(parents, root, n)
            for i, p in enumerate(parents):
                if p >= 0:
                    assert link[i][p] > 0, f"tree {root}: edge {i}->{p} is not a direct link"
            assert max(len(_path(parents, i, root)) for i in range(n)) - 1 <= 3        # log2(8) levels of doubling
    monkeypatch.setenv("MXNET_KVSTORE_TREE_BACKTRACK", "0")
    with pytest.raises(ValueError):
        monkeypatch.setenv("MXNET_KVSTORE_TREE_BACKTRACK", "1")
        rt.plan_trees([[0, 1, 0], [1, 0, 0], [0, 0, 0]])                               # GPU 2 is isolated


def test_planner_uniform_nvswitch_and_dgx1():
    uni = [[0 if i == j else 18 for j in range(8)] for i in range(8)]
    assert rt.topo_is_uniform(uni) and not rt.topo_is_uniform(_dgx1())
    for link in (uni, _dgx1()):
        trees = rt.plan_trees(link)
        for r in range(8):
            _check_tree(trees[r], r, 8)
    # on the cube mesh the planner prefers NVLink edges
    w = _dgx1()
    trees = rt.plan_trees(w, penalty=1.0)
    nv_edges = sum(1 for r in range(8) for i, p in enumerate(trees[r]) if p >= 0 and w[i][p] == 3)
    assert nv_edges >= 0.85 * 8 * 7
    ring = rt.plan_ring(w, 0)
    assert sorted(ring) == list(range(8)) and all(w[a][b] == 3 for a, b in zip(ring, ring[1:]))
    rnd = [[0 if i == j else ((i * 7 + j * 3) % 5 + 1) for j in range(6)] for i in range(6)]
    rnd = [[max(rnd[i][j], rnd[j][i]) if i != j else 0 for j in range(6)] for i in range(6)]
    for r, t in enumerate(rt.plan_trees(rnd)):
        _check_tree(t, r, 6)


# ----------------------------------------------------------------------------- row-sparse (lazy) updates
@pytest.mark.parametrize("name,kw", [("sgd", dict(momentum=0.9)), ("adam", {}), ("adagrad", {}), ("ftrl", {}),
                                     ("groupadagrad", {})])
def test_row_sparse_update_touches_only_listed_rows_and_matches_dense_rule(name, kw):
    torch.manual_seed(5)
    o = kv.optimizer.create(name, learning_rate=0.1, **kw)
    u = kv.optimizer.get_updater(o)
    w = torch.randn(12, 6)
    w0 = w.clone()
    ids = torch.tensor([3, 7, 3, 11])                       # row 3 twice: gradients are summed first
    g = torch.randn(4, 6)
    u.update_row_sparse(0, ids, g, w)
    untouched = [r for r in range(12) if r not in (3, 7, 11)]
    assert torch.equal(w[untouched], w0[untouched])
    # dense rule on the touched rows only
    o2 = kv.optimizer.create(name, learning_rate=0.1, **kw)
    sub = w0[[3, 7, 11]].clone()
    gd = torch.stack([g[0] + g[2], g[1], g[3]])
    st = o2.create_state(0, sub)
    o2.update(0, sub, gd, st)
    assert torch.allclose(w[[3, 7, 11]], sub, atol=1e-6)
    assert not torch.equal(w[[3, 7, 11]], w0[[3, 7, 11]])


def test_device_kvstore_tree_reduce_matches_flat_sum(monkeypatch):
    """CommDeviceTree: topology trees from a (DGX-1 style) link matrix; small keys reduce on one tree, big keys are
    row-sliced over all trees (reduce-scatter + gather).  Fake devices = CPU tensors, as in the reference unit tests."""
    monkeypatch.setenv("MXNET_KVSTORE_USETREE", "1")
    monkeypatch.setenv("MXNET_KVSTORE_TREE_ARRAY_BOUND", "64")
    s = kv.create("device")
    n = 8
    dgx1 = [[0.0] * n for _ in range(n)]
    for a, b, w in [(0, 1, 1), (0, 2, 1), (0, 3, 2), (1, 2, 2), (1, 3, 1), (2, 3, 1), (4, 5, 1), (4, 6, 1), (4, 7, 2), (5, 6, 2),
                    (5, 7, 1), (6, 7, 1), (0, 4, 2), (1, 5, 2), (2, 6, 1), (3, 7, 1)]:
        dgx1[a][b] = dgx1[b][a] = float(w)
    s.link_matrix = dgx1
    torch.manual_seed(0)
    for key, shape in [(3, (4, 4)), (5, (40, 8))]:          # 16 elements: one tree; 320 elements: sliced over 8 trees
        s.init(key, torch.zeros(shape))
        grads = [torch.randn(shape) for _ in range(n)]
        s.push(key, grads)
        out = torch.zeros(shape)
        s.pull(key, out=out)
        assert torch.allclose(out, torch.stack(grads).sum(0), atol=1e-5)
    trees = s._plan(n)
    for r, parent in enumerate(trees):                      # every tree spans all devices and only uses real links
        assert parent[r] == -1 and sorted(i for i, p in enumerate(parent) if p >= 0) == [i for i in range(n) if i != r]
        assert all(dgx1[i][p] > 0 for i, p in enumerate(parent) if p >= 0)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("n,k", [(7, 1), (4096 * 3 + 17, 5), (100_000, 9)])
def test_native_host_reducer_matches_sum(dtype, n, k):
    """CommCPU::ReduceSumCPU equivalent: 4 inputs per pass, chunked over a thread team above the big-array bound."""
    torch.manual_seed(n + k)
    xs = [torch.randn(n).to(dtype) for _ in range(k)]
    for nthreads, bound in ((1, 10 ** 9), (4, 1000)):
        out = rt.reduce_sum_cpu(xs, nthreads=nthreads, bigarray_bound=bound)
        if dtype == torch.float32:
            want = torch.stack(xs).sum(0)
            assert torch.allclose(out, want, rtol=1e-5, atol=1e-5)
        else:
            want = torch.stack([x.float() for x in xs]).sum(0)
            assert torch.allclose(out.float(), want, rtol=2e-2, atol=2e-2)
    inplace = xs[0].clone()
    rt.reduce_sum_cpu([inplace] + xs[1:], out=inplace, nthreads=4, bigarray_bound=1000)      # dst aliases srcs[0]
    assert torch.allclose(inplace.float(), torch.stack([x.float() for x in xs]).sum(0), rtol=2e-2, atol=2e-2)


def test_remote_server_profiling_writes_rank_prefixed_traces(tmp_path):
    import json
    r = _launch(2, 2, [os.path.join(ROOT, "tests/mp/server_profile_worker.py"), str(tmp_path)],
                env={"MXNET_KVSTORE_BIGARRAY_BOUND": "1000"})
    assert r.returncode == 0 and r.stdout.count("PROFILED") == 2, r.stdout[-2000:] + r.stderr[-3000:]
    files = sorted(f for f in os.listdir(tmp_path) if f.endswith("server_profile.json"))
    assert files == ["rank0_server_profile.json", "rank1_server_profile.json"]
    seen = set()
    for f in files:
        ev = json.load(open(tmp_path / f))["traceEvents"]
        names = {e["name"] for e in ev}
        assert "KVStoreDistServerPush" in names and "KVStoreDistServerPull" in names
        assert all(e["ph"] == "X" and e["dur"] >= 0 and e["cat"] == "kvstore_server" for e in ev)
        seen |= {e["args"]["key"] for e in ev}
    assert len(seen) >= 2          # the small key on one server, the big key's shards on both


def test_dead_worker_is_replaced_and_the_sync_round_completes():
    r = _launch(2, 1, [os.path.join(ROOT, "tests/mp/recovery_worker.py")],
                env={"DMLC_NUM_ATTEMPT": "1", "PS_HEARTBEAT_INTERVAL": "1", "PS_HEARTBEAT_TIMEOUT": "3"}, timeout=120)
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-3000:]
    assert "worker 1 crashing" in out and "worker 0 attempt 0 RECOVERY_OK" in out and "worker 1 attempt 1 RECOVERY_OK" in out


@pytest.mark.parametrize("update_on_kvstore_store", ["local", "device"])
def test_module_with_several_contexts_matches_single_device(update_on_kvstore_store):
    """Module(context=[...]) = the reference's DataParallelExecutorGroup, faked like the reference's own unit tests with
    several CPU contexts: batch sliced by work_load_list, per-device gradients reduced through the kvstore, every replica
    updated identically - and the trajectory equals the single-context module on the whole batch."""
    import torch
    from deeplearning_cfn_b200.kv.module import Module, decide_slices

    assert decide_slices(10, [1, 1, 1]) == [slice(0, 3), slice(3, 6), slice(6, 10)]
    assert decide_slices(8, [3, 1]) == [slice(0, 6), slice(6, 8)]
    with pytest.raises(ValueError):
        decide_slices(2, [1, 1, 1, 1])

    def net():
        torch.manual_seed(0)
        return torch.nn.Sequential(torch.nn.Linear(6, 12), torch.nn.Tanh(), torch.nn.Linear(12, 3))
    g = torch.Generator().manual_seed(1)
    batches = [(torch.randn(10, 6, generator=g), torch.randint(0, 3, (10,), generator=g)) for _ in range(4)]
    opt = (("learning_rate", 0.1), ("momentum", 0.9))
    single = Module(net(), context="cpu")
    single.fit(batches, num_epoch=2, kvstore="local", optimizer="sgd", optimizer_params=opt, batch_size=10)
    multi = Module(net(), context=["cpu", "cpu", "cpu"], work_load_list=[2, 1, 1])
    assert multi.multi_device and len(multi._replicas) == 3
    multi.fit(batches, num_epoch=2, kvstore=update_on_kvstore_store, optimizer="sgd", optimizer_params=opt, batch_size=10)
    for a, b in zip(single.symbol.parameters(), multi.symbol.parameters()):
        assert torch.allclose(a, b, atol=1e-5), float((a - b).abs().max())
    for rp in multi._rparams[1:]:                        # replicas stay bit-identical to each other
        for a, b in zip(multi._rparams[0], rp):
            assert torch.equal(a, b)
    # merged outputs cover the whole batch; per-device outputs on request
    multi.forward(batches[0], is_train=False)
    assert multi.get_outputs()[0].shape == (10, 3)
    parts = multi.get_outputs(merge_multi_context=False)[0]
    assert [p.shape[0] for p in parts] == [5, 2, 3]
    assert torch.allclose(torch.cat(parts), multi.get_outputs()[0])
    acc = multi.score(batches, "acc")[0][1]
    assert 0.0 <= acc <= 1.0
    # no kvstore named: several devices still aggregate
    m2 = Module(net(), context=["cpu", "cpu"])
    m2.bind([("data", (10, 6))])
    m2.init_params()
    m2.init_optimizer(kvstore=None, optimizer="sgd", optimizer_params=opt)
    assert m2._kv is not None and m2._kv.type == "local"
    m2.forward_backward(batches[0])
    m2.update()
    for a, b in zip(*m2._rparams):
        assert torch.equal(a, b)


def test_package_import_turns_servers_into_nodes_and_store_closes_at_exit(tmp_path):
    """The reference contract of `import mxnet` + the KVStoreDist destructor: the user script imports only the top-level
    package before doing unrelated work, never calls close(), keeps the store in a module global - servers and scheduler
    still become their nodes at the import, and the job still ends (closing barrier + server stop run at interpreter exit)."""
    script = tmp_path / "user_script.py"
    script.write_text(
        "import sys\n"
        f"sys.path.insert(0, r'{ROOT}')\n"
        "import deeplearning_cfn_b200                      # servers / scheduler never get past this line\n"
        "import os, torch\n"
        "assert os.environ['DMLC_ROLE'] == 'worker', os.environ['DMLC_ROLE']\n"
        "import deeplearning_cfn_b200.kv as kv\n"
        "store = kv.create('dist_sync')\n"
        "store.init(1, torch.zeros(4)); store.push(1, torch.ones(4) * (store.rank + 1))\n"
        "out = torch.zeros(4); store.pull(1, out=out)\n"
        "print('USER_SCRIPT_DONE', store.rank, out.tolist(), flush=True)\n")
    r = _launch(2, 2, [str(script)])
    assert r.returncode == 0 and r.stdout.count("USER_SCRIPT_DONE") == 2, r.stdout[-2000:] + r.stderr[-3000:]
    assert "[3.0, 3.0, 3.0, 3.0]" in r.stdout


def test_dist_store_outside_a_launcher_fails_with_a_clear_error(monkeypatch):
    for v in ("DMLC_PS_ROOT_URI", "DMLC_PS_ROOT_PORT", "DMLC_NUM_WORKER", "DMLC_NUM_SERVER"):
        monkeypatch.delenv(v, raising=False)
    with pytest.raises(kv.KVStoreError, match="launcher"):
        kv.create("dist_sync")


def test_job_ends_when_a_worker_crashes(tmp_path):
    """A synchronous job cannot continue without one of its workers: the failing worker leaves with its error (it does not
    enter the closing barrier), the launcher kills the remaining process trees, ends the scheduler and reports failure -
    instead of the survivors waiting in a push forever."""
    import time
    script = tmp_path / "crash.py"
    script.write_text(
        "import sys, torch\n"
        f"sys.path.insert(0, r'{ROOT}')\n"
        "import deeplearning_cfn_b200.kv as kv\n"
        "s = kv.create('dist_sync')\n"
        "s.init(1, torch.zeros(4))\n"
        "if s.rank == 1:\n"
        "    raise RuntimeError('worker 1 dies')\n"
        "s.push(1, torch.ones(4)); out = torch.zeros(4); s.pull(1, out=out)\n"
        "print('SURVIVOR_DONE', flush=True)\n")
    t0 = time.time()
    r = _launch(2, 1, [str(script)], timeout=120)
    took = time.time() - t0
    assert r.returncode != 0 and took < 90, (r.returncode, took, r.stdout[-1000:] + r.stderr[-2000:])
    assert "worker 1 dies" in r.stderr and "stopping the job" in r.stderr and "SURVIVOR_DONE" not in r.stdout
    import psutil
    left = [p for p in psutil.process_iter(["cmdline"]) if str(script) in " ".join(p.info["cmdline"] or [])]
    assert not left, [p.pid for p in left]


def test_scheduler_survives_garbage_connections(tmp_path):
    """Fuzz the scheduler's port (random bytes, a frame header announcing 2^40 bytes, connect-and-close, truncated frames)
    before the job's nodes arrive: it must drop those connections and still run the job."""
    import random
    import signal
    import socket
    import struct
    import subprocess
    import time
    port = 9300 + os.getpid() % 200
    env = dict(os.environ, PYTHONPATH=ROOT, DMLC_PS_ROOT_URI="127.0.0.1", DMLC_PS_ROOT_PORT=str(port), DMLC_NUM_WORKER="1",
               DMLC_NUM_SERVER="1")
    procs = []

    def spawn(role, cmd):
        p = subprocess.Popen(cmd, env=dict(env, DMLC_ROLE=role), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                             start_new_session=True)
        procs.append(p)
        return p
    try:
        sched = spawn("scheduler", [sys.executable, "-c", "import deeplearning_cfn_b200.kv"])
        for _ in range(150):
            try:
                socket.create_connection(("127.0.0.1", port), timeout=1).close()
                break
            except OSError:
                time.sleep(0.2)
        rng = random.Random(0)
        for i in range(40):
            s = socket.create_connection(("127.0.0.1", port), timeout=2)
            body = rng.randbytes(rng.randint(0, 300))
            data = [rng.randbytes(rng.randint(1, 200)), struct.pack("<IQ", 0xDEADBEEF, 1 << 40) + rng.randbytes(50), b"",
                    rng.randbytes(12) + rng.randbytes(rng.randint(0, 5000)),
                    struct.pack("<IQ", 0xD1CF2026, len(body)) + body,                     # well-framed, undecodable message
                    struct.pack("<IQ", 0xD1CF2026, 1000) + body[:10]][i % 6]              # frame cut short
            s.sendall(data)
            s.close()
        time.sleep(0.3)
        assert sched.poll() is None, sched.communicate()[0][-2000:]
        server = spawn("server", [sys.executable, "-c", "import deeplearning_cfn_b200.kv"])
        worker = spawn("worker", [sys.executable, os.path.join(ROOT, "tests/mp/dist_kvstore_worker.py"), "dist_sync"])
        out, _ = worker.communicate(timeout=120)
        assert worker.returncode == 0 and "SYNC_OK" in out, out[-2000:]
        assert server.wait(timeout=60) == 0 and sched.wait(timeout=60) == 0
    finally:
        for p in procs:
            if p.poll() is None:
                os.killpg(p.pid, signal.SIGKILL)


def test_dead_server_is_reported_through_heartbeats(tmp_path):
    """Failure detection (reference: PS_HEARTBEAT_INTERVAL / PS_HEARTBEAT_TIMEOUT, `kv.get_num_dead_node`): SIGKILL one of two
    servers after the rendezvous; a worker asking the scheduler sees exactly one dead node in the server group."""
    import signal
    import subprocess
    import time
    port = 9500 + os.getpid() % 150
    env = dict(os.environ, PYTHONPATH=ROOT, DMLC_PS_ROOT_URI="127.0.0.1", DMLC_PS_ROOT_PORT=str(port), DMLC_NUM_WORKER="1",
               DMLC_NUM_SERVER="2", PS_HEARTBEAT_INTERVAL="1", PS_HEARTBEAT_TIMEOUT="3")
    script = tmp_path / "w.py"
    script.write_text(
        "import os, sys, time, torch\n"
        f"sys.path.insert(0, r'{ROOT}')\n"
        "import deeplearning_cfn_b200.kv as kv\n"
        "s = kv.create('dist_async')\n"
        "s.init(1, torch.zeros(4))\n"
        "assert s.get_num_dead_node(2, 3) == 0\n"
        "print('WORKER_READY', flush=True)\n"
        "t0, seen = time.time(), 0\n"
        "while time.time() - t0 < 40 and not seen:\n"
        "    seen = s.get_num_dead_node(2, 3)\n"
        "    time.sleep(0.5)\n"
        "print('DEAD_SERVERS', seen, flush=True)\n"
        "os._exit(0)\n")                      # no closing barrier: one server is gone

    def spawn(role, cmd):
        return subprocess.Popen(cmd, env=dict(env, DMLC_ROLE=role), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                                start_new_session=True)
    procs = [spawn("scheduler", [sys.executable, "-c", "import deeplearning_cfn_b200.kv"])]
    procs += [spawn("server", [sys.executable, "-c", "import deeplearning_cfn_b200.kv"]) for _ in range(2)]
    worker = spawn("worker", [sys.executable, str(script)])
    procs.append(worker)
    try:
        line = worker.stdout.readline()
        assert "WORKER_READY" in line, line + worker.stdout.read()[-2000:]
        os.killpg(procs[2].pid, signal.SIGKILL)
        out, _ = worker.communicate(timeout=90)
        assert "DEAD_SERVERS 1" in out, out[-2000:]
    finally:
        for p in procs:
            if p.poll() is None:
                os.killpg(p.pid, signal.SIGKILL)


def test_job_ends_when_a_server_is_killed(tmp_path):
    """SIGKILL one server of a running dist_sync job: the launcher notices the exit code, kills the remaining process trees and
    returns non-zero within seconds; nothing of the job is left behind."""
    import signal
    import subprocess
    import time
    import psutil
    script = tmp_path / "loop.py"
    script.write_text(
        "import sys, time, torch\n"
        f"sys.path.insert(0, r'{ROOT}')\n"
        "import deeplearning_cfn_b200.kv as kv\n"
        "s = kv.create('dist_sync')\n"
        "s.init(1, torch.zeros(4))\n"
        "print('WORKER_READY', flush=True)\n"
        "for _ in range(400):\n"
        "    s.push(1, torch.ones(4)); out = torch.zeros(4); s.pull(1, out=out); time.sleep(0.2)\n")
    p = subprocess.Popen([sys.executable, "-m", "deeplearning_cfn_b200", "launch", "-n", "2", "-s", "2", "--launcher", "local",
                          sys.executable, str(script)], env=dict(os.environ, PYTHONPATH=ROOT), stdout=subprocess.PIPE,
                         stderr=subprocess.STDOUT, text=True, start_new_session=True)
    try:
        servers, deadline = [], time.time() + 60
        while time.time() < deadline and len(servers) < 2:
            time.sleep(1.0)
            servers = []
            for c in psutil.Process(p.pid).children(recursive=True):
                try:
                    if c.environ().get("DMLC_ROLE") == "server" and "python" in c.name():
                        servers.append(c)
                except (psutil.Error, OSError):
                    pass
        assert len(servers) == 2
        time.sleep(3.0)                              # let the job get going
        servers[0].kill()
        out, _ = p.communicate(timeout=60)
        assert p.returncode != 0 and "stopping the job" in out, out[-2000:]
        time.sleep(1.0)
        left = [c.pid for c in psutil.process_iter(["cmdline", "status"])
                if str(script) in " ".join(c.info["cmdline"] or []) and c.info["status"] != psutil.STATUS_ZOMBIE]
        assert not left, left
    finally:
        if p.poll() is None:
            os.killpg(p.pid, signal.SIGKILL)


def test_module_fit_with_dist_sync_store_under_the_launcher(tmp_path):
    """The symbolic recipe (`Module.fit(kvstore='dist_sync')`, update on the servers, push / pull with priority -index) as a
    2-worker job: different data per rank, identical weights afterwards, the job ends without an explicit close()."""
    script = tmp_path / "fit.py"
    script.write_text(
        "import sys, torch\n"
        f"sys.path.insert(0, r'{ROOT}')\n"
        "import deeplearning_cfn_b200.kv as kv\n"
        "from deeplearning_cfn_b200.kv.module import Module\n"
        "torch.manual_seed(0)\n"
        "net = torch.nn.Sequential(torch.nn.Linear(6, 12), torch.nn.ReLU(), torch.nn.Linear(12, 3))\n"
        "store = kv.create('dist_sync')\n"
        "g = torch.Generator().manual_seed(1 + store.rank)\n"
        "batches = [(torch.randn(8, 6, generator=g), torch.randint(0, 3, (8,), generator=g)) for _ in range(5)]\n"
        "m = Module(net, context='cpu')\n"
        "m.fit(batches, num_epoch=2, kvstore=store, optimizer='sgd', optimizer_params=(('learning_rate', 0.1), ('momentum', 0.9)),\n"
        "      batch_size=8)\n"
        "w = torch.cat([p.detach().reshape(-1) for p in net.parameters()])\n"
        "print(f'MODULE_DIST_OK {float(w.sum()):.6f} {float(w.abs().sum()):.6f}', flush=True)\n")
    r = _launch(2, 1, [str(script)], timeout=240)
    import re
    found = re.findall(r"MODULE_DIST_OK (-?[0-9.]+) (-?[0-9.]+)", r.stdout)      # the two ranks share stdout: lines may interleave
    assert r.returncode == 0 and len(found) == 2 and found[0] == found[1], r.stdout[-2000:] + r.stderr[-2000:]


def test_nccl_store_grouped_push_packs_16_keys_and_matches_per_key_pushes():
    """The packing / unpacking of the NCCL store's grouped multi-key push (kvstore_nccl.h:221-307: <= 16 keys per group) with a
    host stand-in for the one NCCL reduce per group: same stored values as key-by-key pushes, ceil(K / 16) reductions, mixed
    dtype groups fall back to per-key pushes."""
    from deeplearning_cfn_b200.kv.kvstore import KVStoreNCCL, KVStoreLocal
    torch.manual_seed(0)
    shapes = [(3, 5), (7,), (2, 2, 2)] * 7                                   # 21 keys -> groups of 16 + 5
    keys = list(range(len(shapes)))
    vals = [[torch.randn(s) for _ in range(3)] for s in shapes]              # 3 "devices"
    ref = KVStoreLocal("local")
    got = KVStoreNCCL("nccl")
    for st in (ref, got):
        st.init(keys, [torch.zeros(s) for s in shapes])
        st._set_updater(lambda k, g, w: w.add_(g * (1 + k)))
    calls = []

    def reduce_packed(packed):
        calls.append(packed[0].numel())
        return torch.stack(packed).sum(0)

    ref.push(keys, vals)
    got._push_grouped(keys, vals, 0, reduce_packed)
    assert len(calls) == 2 and calls[0] == sum(torch.Size(s).numel() for s in shapes[:16])
    for k, s in zip(keys, shapes):
        a, b = torch.zeros(s), torch.zeros(s)
        ref.pull(k, out=a)
        got.pull(k, out=b)
        assert torch.allclose(a, b, atol=1e-6), k
    # a group with two dtypes is pushed key by key (host tensors: the plain reduce), nothing goes through the packed path
    calls.clear()
    mixed = [[torch.ones(4), torch.ones(4)], [torch.ones(4, dtype=torch.float64), torch.ones(4, dtype=torch.float64)]]
    st = KVStoreNCCL("nccl")
    st.init([0, 1], [torch.zeros(4), torch.zeros(4, dtype=torch.float64)])
    st._push_grouped([0, 1], mixed, 0, reduce_packed)
    out = torch.zeros(4)
    st.pull(0, out=out)
    assert calls == [] and torch.equal(out, torch.full((4,), 2.0))


def test_message_decoder_survives_structured_fuzzing():
    """The frame decoder of the native transport (csrc/rt/van.cpp: Unpack, the first code that touches bytes from the network)
    under structured fuzzing through its test hook: valid messages decode to their entry counts; 20 000 mutations - truncation,
    bit flips, every length / count field replaced by boundary values (0, n - 1, 2^31, 2^32, 2^63, 2^64 - 1) - decode or are
    rejected without a crash.  tools/sanitize_rt.sh / ubsan_rt.sh run this under ASAN / UBSAN (the pointer overflow UBSAN found
    in round 2 was in exactly these length checks)."""
    import random
    import struct
    L = rt.lib()
    rng = random.Random(1234)

    def pack(nodes, data, body):
        out = struct.pack("<7i", 3, 9, 0, 0, 1, 0, rng.randrange(1 << 20)) + struct.pack("<ii", rng.randrange(8), 7)
        out += struct.pack("<Q", rng.randrange(1 << 40)) + struct.pack("<Q", len(body)) + body
        out += struct.pack("<i", len(nodes))
        for host in nodes:
            out += struct.pack("<4i", 1, 8, 9000, 0) + struct.pack("<Q", len(host)) + host
        out += struct.pack("<i", len(data))
        for d in data:
            out += struct.pack("<Q", len(d)) + d
        return out

    seeds = []
    for _ in range(50):
        nodes = [rng.randbytes(rng.randrange(0, 20)) for _ in range(rng.randrange(0, 4))]
        data = [rng.randbytes(rng.randrange(0, 300)) for _ in range(rng.randrange(0, 4))]
        p = pack(nodes, data, rng.randbytes(rng.randrange(0, 64)))
        assert L.dlcfn_rt_debug_unpack(p, len(p)) == 1 + len(nodes) + len(data)
        seeds.append(p)
    extremes = [0, 1, 0x7FFFFFFF, 0x80000000, 0xFFFFFFFF, 1 << 32, (1 << 63) - 1, 1 << 63, (1 << 64) - 1]
    rejected = 0
    for it in range(20000):
        b = bytearray(rng.choice(seeds))
        kind = it % 5
        if kind == 0:
            b = b[:rng.randrange(0, len(b) + 1)]
        elif kind == 1:
            for _ in range(rng.randrange(1, 6)):
                b[rng.randrange(len(b))] ^= 1 << rng.randrange(8)
        elif kind == 2:                                    # an 8-byte field somewhere becomes a boundary value
            off = rng.randrange(0, max(1, len(b) - 8))
            b[off:off + 8] = struct.pack("<Q", rng.choice(extremes + [len(b) - off, len(b) - off - 7, len(b)]))
        elif kind == 3:                                    # a 4-byte field (counts) becomes a boundary value
            off = rng.randrange(0, max(1, len(b) - 4))
            b[off:off + 4] = struct.pack("<I", rng.choice([0, 1, 0x7FFFFFFF, 0x80000000, 0xFFFFFFFF, 1000]))
        else:
            b = bytearray(rng.randbytes(rng.randrange(0, 400)))
        rc = L.dlcfn_rt_debug_unpack(bytes(b), len(b))
        assert rc >= 0
        rejected += rc == 0
    assert 1000 < rejected < 20000          # both outcomes occur: the fuzzer neither only breaks nor only preserves messages


def test_planners_return_valid_trees_or_refuse_on_arbitrary_link_matrices(monkeypatch):
    """KL and backtracking tree planners and the ring planner (csrc/rt/topo.cpp) on 600 random link matrices - sparse,
    asymmetric, real-valued, with NaN / huge / negative entries: the answer is a well-formed set of trees (root marked, parents
    in range) / a permutation, or ValueError; nothing else.  Clean under ASAN / UBSAN (tools/*_rt.sh -k planner)."""
    import random
    rng = random.Random(3)
    ok = refused = 0
    for it in range(600):
        n = rng.randrange(1, 9)
        kind = it % 4
        link = [[0.0] * n for _ in range(n)]
        for i in range(n):
            for j in range(n):
                if i == j:
                    continue
                if kind == 0:
                    link[i][j] = rng.choice([0, 0, 1, 2, 18])
                elif kind == 1:
                    link[i][j] = link[j][i] = rng.choice([0, 1, 4])
                elif kind == 2:
                    link[i][j] = rng.random() * 20
                else:
                    link[i][j] = rng.choice([0, 1e30, -1, float("nan"), 3])
        for bt in ("0", "1"):
            monkeypatch.setenv("MXNET_KVSTORE_TREE_BACKTRACK", bt)
            try:
                trees = rt.plan_trees(link)
            except ValueError:
                refused += 1
                continue
            assert len(trees) == n and all(len(t) == n for t in trees)
            for r, t in enumerate(trees):
                assert t[r] == -1 and all(-1 <= p < n for p in t)
            ok += 1
        try:
            assert sorted(rt.plan_ring(link, rng.randrange(n))) == list(range(n))
        except ValueError:
            pass
    assert ok > 300 and refused > 30


def test_engine_random_dependency_graphs_keep_per_variable_order():
    """Random read / write sets, priorities and thread counts through the native dependency engine (csrc/rt/engine.cpp): every
    reader observes the last writer pushed before it, writers of a variable run in push order, wait_for_var may be called at any
    time.  (30 such graphs were also run under ThreadSanitizer: profiles/r2/sanitizers_r2.md.)"""
    import random
    import threading
    rng = random.Random(5)
    for trial in range(10):
        eng = rt.Engine(rng.choice([1, 2, 4, 8]))
        nv = rng.randrange(1, 8)
        vars_ = [eng.new_var() for _ in range(nv)]
        state = {v: 0 for v in vars_}                  # index of the last writer that ran
        pushed_last = {v: 0 for v in vars_}            # index of the last writer pushed so far
        violations, lock = [], threading.Lock()
        for op in range(1, rng.randrange(50, 300) + 1):
            touched = rng.sample(vars_, rng.randrange(0, min(3, nv) + 1))
            nw = rng.randrange(0, len(touched) + 1)
            writes, reads = touched[:nw], touched[nw:]
            want = {v: pushed_last[v] for v in touched}
            for v in writes:
                pushed_last[v] = op

            def fn(op=op, reads=reads, writes=writes, want=want):
                for v in reads + writes:
                    if state[v] != want[v]:
                        with lock:
                            violations.append((op, v, state[v], want[v]))
                for v in writes:
                    state[v] = op
            eng.push(fn, reads, writes, priority=rng.randrange(-5, 6))
            if rng.random() < 0.02:
                eng.wait_for_var(rng.choice(vars_))
        eng.wait_all()
        assert not violations, (trial, violations[:5])
        assert state == pushed_last
        eng.close()
