"""Worker body of the distributed KVStore tests; started by `launch -n W -s S --launcher local`.
Pattern of the reference nightly (MX/tests/nightly/dist_sync_kvstore.py:101-112): every worker pushes
ones*(rank+1) under the `Test` optimizer, so the value after round i is known in closed form."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import deeplearning_cfn_b200.kv as kv  # noqa: E402  (servers/scheduler never return from this import)

mode = sys.argv[1] if len(sys.argv) > 1 else "dist_sync"
BIG = 1200
os.environ.setdefault("MXNET_KVSTORE_BIGARRAY_BOUND", "1000")


def check(t, want, what):
    if not torch.allclose(t.float(), torch.full_like(t.float(), float(want)), rtol=1e-3, atol=1e-3):
        raise AssertionError(f"{what}: got {t.reshape(-1)[:4].tolist()} want {want}")


def optimizer_parity():
    """Every optimizer that has a native server implementation (csrc/rt/kv.cpp:ApplyUpdate, same numbering as
    csrc/cuda/optimizer.cuh) against the python definition in kv/optimizer.py, 2 workers x 1+ servers."""
    store = kv.create("dist_sync")
    r, n = store.rank, store.num_workers
    cases = [("sgd", dict(momentum=0.9, wd=0.01)), ("nag", dict(momentum=0.9, wd=0.01)),
             ("signum", dict(momentum=0.9, wd_lh=0.01)), ("signum", dict(momentum=0.0)),
             ("adam", dict(wd=0.01)), ("adagrad", dict(wd=0.01)), ("rmsprop", dict(wd=0.01, clip_weights=2.0)),
             ("adadelta", dict(wd=0.001)), ("ftrl", dict(lamda1=0.01, wd=0.01)), ("adamax", dict(wd=0.01))]
    torch.manual_seed(0)
    for ci, (name, kw) in enumerate(cases):
        key = 100 + ci
        w0 = torch.randn(37)
        store.init(key, w0)
        o = kv.optimizer.create(name, learning_rate=0.05, rescale_grad=0.5, clip_gradient=1.5, **kw)
        assert o.native_config() is not None, name
        store.set_optimizer(o)
        local = kv.optimizer.get_updater(kv.optimizer.create(name, learning_rate=0.05, rescale_grad=0.5,
                                                             clip_gradient=1.5, **kw))
        wl = w0.clone()
        for i in range(4):
            gs = [torch.randn(37) * (w + 1) for w in range(n)]      # same stream on every worker
            store.push(key, gs[r])
            out = torch.zeros(37)
            store.pull(key, out=out)
            local(key, sum(gs), wl)
            if not torch.allclose(out, wl, rtol=2e-4, atol=2e-5):
                raise AssertionError(f"{name} round {i}: server {out[:4].tolist()} python {wl[:4].tolist()}")
    store._barrier()
    print(f"worker {r} OPT_OK", flush=True)
    store.close()


def main():
    if mode == "opt_parity":
        return optimizer_parity()
    store = kv.create(mode)
    r, n = store.rank, store.num_workers
    assert store.type == mode and 0 <= r < n
    rate = 2.0
    shapes = {3: (4, 5), 5: (7,), "big": (BIG, 3)}                       # "big" crosses BIGARRAY_BOUND -> split over servers
    int_keys = [3, 5]
    if mode.endswith("_async"):
        store.init(3, torch.ones(shapes[3]))
        store.set_optimizer(kv.optimizer.create("test", rescale_grad=rate))
        for i in range(5):
            store.push(3, torch.ones(shapes[3]) * (r + 1))
        store._barrier()
        out = torch.zeros(shapes[3])
        store.pull(3, out=out)
        check(out, 1 + rate * 5 * n * (n + 1) / 2, "async total")       # every push applied exactly once
        print(f"worker {r} ASYNC_OK", flush=True)
        store.close()
        return
    # ---- sync: dense, list keys, big key
    store.init(int_keys, [torch.ones(shapes[k]) for k in int_keys])
    store.init(99, torch.ones(shapes["big"]))
    try:
        store.init(3, torch.ones(shapes[3]))
        raise SystemExit("duplicate init must fail")
    except kv.KVStoreError:
        pass
    try:
        store.push(1234, torch.ones(2))
        raise SystemExit("push to an uninitialised key must fail")
    except kv.KVStoreError:
        pass
    # no optimizer yet: the server stores the SUM over workers (kvstore_dist_server.h:358-361)
    store.push(5, torch.ones(shapes[5]) * (r + 1))
    o = torch.zeros(shapes[5])
    store.pull(5, out=o)
    check(o, n * (n + 1) / 2, "sum without updater")
    store.set_optimizer(kv.optimizer.create("test", rescale_grad=rate))
    base5 = n * (n + 1) / 2
    for i in range(3):
        store.push(3, [torch.ones(shapes[3]) * (r + 1)] * 2)             # two "devices" per worker: local reduce first
        store.push(99, torch.ones(shapes["big"]) * (r + 1), priority=-1)
        store.push(5, torch.ones(shapes[5]) * (r + 1))
        o3, ob, o5 = torch.zeros(shapes[3]), torch.zeros(shapes["big"]), torch.zeros(shapes[5])
        store.pull([3, 99], out=[o3, ob])
        store.pull(5, out=[o5, torch.zeros(shapes[5])])
        check(o3, 1 + rate * (i + 1) * 2 * n * (n + 1) / 2, f"round {i} key 3")
        check(ob, 1 + rate * (i + 1) * n * (n + 1) / 2, f"round {i} big key")
        check(o5, base5 + rate * (i + 1) * n * (n + 1) / 2, f"round {i} key 5")
    # ---- multi-precision wire: bf16 values, fp32 master on the server
    store.init(7, torch.ones(9, dtype=torch.bfloat16))
    for i in range(2):
        store.push(7, torch.ones(9, dtype=torch.bfloat16) * (r + 1))
        o7 = torch.zeros(9, dtype=torch.bfloat16)
        store.pull(7, out=o7)
        check(o7, 1 + rate * (i + 1) * n * (n + 1) / 2, f"bf16 round {i}")
    # ---- row-sparse push/pull
    store.init(11, torch.zeros(20, 4))
    rows = torch.tensor([1, 7, 19]) if r % 2 == 0 else torch.tensor([7, 8])
    store.push_row_sparse(11, torch.ones(len(rows), 4) * (r + 1), rows)
    out = torch.full((20, 4), -1.0)
    store.row_sparse_pull(11, out=out, row_ids=torch.tensor([7, 19, 7, 2]))
    want7 = rate * sum(w + 1 for w in range(n))
    want19 = rate * sum(w + 1 for w in range(n) if w % 2 == 0)
    check(out[7], want7, "rsp row 7")
    check(out[19], want19, "rsp row 19")
    check(out[2], 0.0, "rsp untouched row")
    assert float(out[1].abs().sum()) == 0.0, "rows that were not requested stay zero"
    # ---- 2-bit compression: expected value recomputed from the quantisation rule
    store.set_gradient_compression({"type": "2bit", "threshold": 0.5})
    store.init(13, torch.zeros(40))
    g = torch.full((40,), 0.3) * (r + 1)
    store.push(13, g)                                                    # residual carries what was not sent
    o13 = torch.zeros(40)
    store.pull(13, out=o13)
    sent = sum(0.5 if 0.3 * (w + 1) >= 0.5 else 0.0 for w in range(n))
    check(o13, rate * sent, "2bit first push")
    store.push(13, g)
    store.pull(13, out=o13)
    sent2 = 0.0
    for w in range(n):
        res = 0.3 * (w + 1)
        res = res - 0.5 if res >= 0.5 else res
        res += 0.3 * (w + 1)
        sent2 += 0.5 if res >= 0.5 else 0.0
    check(o13, rate * (sent + sent2), "2bit second push (error feedback)")
    assert store.get_num_dead_node(2 | 4, 60) == 0
    try:
        store.save_optimizer_states("/tmp/never")
        raise SystemExit("saving optimizer states must be refused in distributed mode")
    except kv.KVStoreError:
        pass
    store._barrier()
    print(f"worker {r} SYNC_OK", flush=True)
    store.close()


if __name__ == "__main__":
    main()
