"""KVStore: the key-value parameter store API of the reference, on B200.

API parity target: `mx.kv.create(name)` and the `KVStore` methods of MX/python/mxnet/kvstore.py:97-677
(init / push / pull / row_sparse_pull / set_gradient_compression / set_optimizer / save_- and
load_optimizer_states / type / rank / num_workers / _barrier / _send_command_to_servers); behavioural
contract distilled in SURVEY Appendix B.  Type-string rules follow kvstore.cc:40-77: contains `dist` ->
distributed, `device` -> reduce on GPU, `nccl` -> NCCL baseline, `_async` -> no per-key barrier.

Back-ends
  local / device         KVStoreLocal: one process, values may be lists of per-device tensors.  `local` reduces
                         on the host (reference CommCPU, comm.h:103-449), `device` reduces on a GPU with the
                         dlcfn elementwise-sum kernel after P2P copies (reference CommDevice, comm.h:451-800).
  nccl                   KVStoreNCCL: single-process multi-GPU ncclReduce + ncclBcast - the baseline shape of
                         kvstore_nccl.h:221-405 (never the product path).
  dist_sync / dist_async / dist_device_sync (CPU wire)
                         KVStoreDist over the native transport in csrc/rt (ps-lite replacement): dedicated
                         server processes, key sharding, sync aggregation or async apply on the server.
  dist_device_sync (GPU) KVStoreDistDevice: when torch.distributed is up on CUDA, keys live in the symmetric
                         peer-HBM heap and push -> server reduce -> optimizer -> pull is ONE fused kernel per key
                         (csrc/cuda/collectives.cu:kps_sync); servers are colocated GPU shards.
"""
from __future__ import annotations

import ctypes
import math
import os
import sys
import pickle

import torch

from ..ops import native
from ..ops import reference as ref_ops
from ..utils import rt
from . import optimizer as opt

_DT = {torch.float32: rt.DT_F32, torch.bfloat16: rt.DT_BF16, torch.float16: rt.DT_F16}


class KVStoreError(RuntimeError):
    pass


def _as_list(x):
    return list(x) if isinstance(x, (list, tuple)) else [x]


def _key_value_pairs(key, value):
    """Normalise (key, value) to parallel lists: keys [k...], values [[per-device tensors]...]."""
    if isinstance(key, (list, tuple)):
        keys = list(key)
        vals = list(value)
        if len(vals) != len(keys):
            # flat list of per-device values grouped by key
            if len(vals) % len(keys) != 0:
                raise KVStoreError("number of values is not a multiple of the number of keys")
            n = len(vals) // len(keys)
            vals = [vals[i * n:(i + 1) * n] for i in range(len(keys))]
        return keys, [_as_list(v) for v in vals]
    return [key], [_as_list(value)]


class KVStore:
    def __init__(self, name: str):
        self._type = name
        self._updater = None
        self._optimizer = None
        self._str_keys: dict = {}
        self._key_kind = None
        self._compression = None
        self._residuals: dict = {}

    # ------------------------------------------------------------------ identity
    @property
    def type(self) -> str:
        return self._type

    @property
    def rank(self) -> int:
        return 0

    @property
    def num_workers(self) -> int:
        return 1

    # ------------------------------------------------------------------ keys
    def _ikey(self, key, create: bool = False) -> int:
        kind = "str" if isinstance(key, str) else "int"
        if self._key_kind is None:
            self._key_kind = kind
        elif self._key_kind != kind:
            raise KVStoreError("inconsistent key types: do not mix int and str keys (kvstore_local.h:93-166)")
        if kind == "int":
            return int(key)
        if key not in self._str_keys:
            if not create:
                raise KVStoreError(f"key {key!r} has not been initialized")
            self._str_keys[key] = len(self._str_keys)
        return self._str_keys[key]

    # ------------------------------------------------------------------ API
    def init(self, key, value) -> None:
        raise NotImplementedError

    def push(self, key, value, priority: int = 0) -> None:
        raise NotImplementedError

    def pull(self, key, out=None, priority: int = 0, ignore_sparse: bool = True) -> None:
        raise NotImplementedError

    def row_sparse_pull(self, key, out=None, priority: int = 0, row_ids=None) -> None:
        raise NotImplementedError

    def set_gradient_compression(self, compression_params: dict) -> None:
        """{'type': '2bit', 'threshold': 0.5} (kvstore.py:397-457; gradient_compression.cc)."""
        t = compression_params.get("type", "none")
        if t not in ("2bit", "none"):
            raise KVStoreError(f"unknown gradient compression type {t!r}")
        if "nccl" in self._type and t != "none":
            raise KVStoreError("gradient compression is not supported by the nccl kvstore (kvstore_nccl.h:181-189)")
        thr = float(compression_params.get("threshold", 0.5))
        if thr <= 0:
            raise KVStoreError("threshold must be greater than 0")
        self._compression = None if t == "none" else {"type": t, "threshold": thr}

    def set_optimizer(self, optimizer) -> None:
        self._optimizer = opt.create(optimizer) if not isinstance(optimizer, opt.Optimizer) else optimizer
        self._set_updater(opt.get_updater(self._optimizer))

    def _set_updater(self, updater) -> None:
        self._updater = updater

    def save_optimizer_states(self, fname: str, dump_optimizer: bool = False) -> None:
        if self._updater is None:
            raise KVStoreError("Cannot save states for distributed training")  # kvstore.py:550
        with open(fname, "wb") as f:
            f.write(self._updater.get_states(dump_optimizer))

    def load_optimizer_states(self, fname: str) -> None:
        if self._updater is None:
            raise KVStoreError("Cannot load states for distributed training")
        with open(fname, "rb") as f:
            self._updater.set_states(f.read())

    def _barrier(self) -> None:
        pass

    def _send_command_to_servers(self, head: int, body: str) -> None:
        pass

    def get_num_dead_node(self, node_id: int, timeout: int = 60) -> int:
        return 0

    # ------------------------------------------------------------------ helpers shared by back-ends
    def _compress_roundtrip(self, ikey: int, dev_idx: int, grad: torch.Tensor) -> torch.Tensor:
        """quantize (with error feedback) -> dequantize one device's gradient (comm.h:552-596)."""
        thr = self._compression["threshold"]
        rkey = (ikey, dev_idx)
        if rkey not in self._residuals:
            self._residuals[rkey] = torch.zeros(grad.numel(), dtype=torch.float32, device=grad.device)
        res = self._residuals[rkey]
        flat = grad.reshape(-1)
        if grad.is_cuda:
            words = torch.empty((flat.numel() + 15) // 16, dtype=torch.int32, device=grad.device)
            g = flat if flat.dtype in (torch.float32, torch.bfloat16) else flat.float()
            native.quantize_2bit(g.contiguous(), res, words, thr)
            out = torch.empty(flat.numel(), dtype=torch.float32, device=grad.device)
            native.dequantize_2bit(words, out, thr)
        else:
            words = ref_ops.quantize_2bit(flat, res, thr)
            out = ref_ops.dequantize_2bit(words, flat.numel(), thr)
        return out.view(grad.shape).to(grad.dtype)


# =============================================================================== local / device
class KVStoreLocal(KVStore):
    """`local` (host reduce, CommCPU) / `device` (GPU reduce, CommDevice; with MXNET_KVSTORE_USETREE=1 the
    topology-aware tree reduce of CommDeviceTree).  Runtime knobs honoured, names as in the reference (env_var.md):
    MXNET_KVSTORE_USETREE, MXNET_KVSTORE_TREE_ARRAY_BOUND (1e7), MXNET_KVSTORE_TREE_LINK_USAGE_PENALTY (0.7),
    MXNET_KVSTORE_LOGTREE, MXNET_ENABLE_GPU_P2P (1; 0 stages device-to-device copies through the host),
    MXNET_GPU_COPY_NTHREADS (2; number of copy streams used to stage peer gradients), MXNET_KVSTORE_SERIAL_PUSH
    (0; 1 synchronises after every push - the reference's benchmarking switch, comm.h:109)."""

    def __init__(self, name: str = "local"):
        super().__init__(name)
        self._store: dict = {}
        self._on_device = "device" in name
        self._use_tree = self._on_device and os.environ.get("MXNET_KVSTORE_USETREE", "0") == "1"
        self._tree_bound = int(float(os.environ.get("MXNET_KVSTORE_TREE_ARRAY_BOUND", "10000000")))
        self._p2p = os.environ.get("MXNET_ENABLE_GPU_P2P", "1") != "0"
        self._serial_push = os.environ.get("MXNET_KVSTORE_SERIAL_PUSH", "0") == "1"
        self._n_copy_streams = max(1, int(os.environ.get("MXNET_GPU_COPY_NTHREADS", "2")))
        self._copy_streams: dict = {}
        self._trees: dict = {}
        self.link_matrix = None        # optional [n][n] link-strength matrix; default: probe (CUDA) / uniform

    # -- device-to-device staging (reference C4: cudaMemcpyPeerAsync on the copy workers' streams)
    def _stage(self, v: torch.Tensor, dev, slot: int) -> torch.Tensor:
        v = v.detach()
        if v.device == dev:
            return v.contiguous()
        if not self._p2p and v.is_cuda:
            return v.to("cpu").to(dev).contiguous()
        if v.is_cuda and getattr(dev, "type", None) == "cuda":
            key = (dev.index, slot % self._n_copy_streams)
            st = self._copy_streams.get(key)
            if st is None:
                st = self._copy_streams[key] = torch.cuda.Stream(device=dev)
            st.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(st):
                out = v.to(dev, non_blocking=True).contiguous()
            torch.cuda.current_stream(dev).wait_stream(st)
            return out
        return v.to(dev).contiguous()

    # -- CommDeviceTree (comm_tree.h:181-259, gpu_topology.h): one reduction tree per root from the link matrix
    def _plan(self, n: int) -> list:
        if n not in self._trees:
            link = self.link_matrix
            if link is None:
                link = [[0.0 if i == j else 1.0 for j in range(n)] for i in range(n)]    # NVSwitch: uniform
            self._trees[n] = rt.plan_trees(link)
        return self._trees[n]

    @staticmethod
    def _reduce_on_tree(parts: list, parent: list, root: int) -> torch.Tensor:
        n = len(parts)

        def depth(i):
            d = 0
            while parent[i] >= 0:
                i = parent[i]
                d += 1
            return d

        acc = [p.detach() for p in parts]
        for node in sorted(range(n), key=depth, reverse=True):      # leaves first
            p = parent[node]
            if p < 0:
                continue
            acc[p] = acc[p] + acc[node].to(acc[p].device)           # one hop along a tree edge
        return acc[root]

    def _reduce_tree(self, ikey: int, vals: list) -> torch.Tensor:
        n = len(vals)
        trees = self._plan(n)
        v0 = vals[0]
        if v0.numel() > self._tree_bound and v0.dim() >= 1 and v0.shape[0] >= n:
            # big array: row-slice into n pieces, reduce-scatter piece j on the tree rooted at j, then gather
            pieces = [torch.chunk(v.detach(), n, dim=0) for v in vals]
            merged = [self._reduce_on_tree([pieces[d][j] for d in range(n)], trees[j], j) for j in range(len(pieces[0]))]
            dev = v0.device
            return torch.cat([m.to(dev) for m in merged], dim=0)
        root = ikey % n
        return self._reduce_on_tree(list(vals), trees[root], root)

    def _reduce(self, ikey: int, vals: list) -> torch.Tensor:
        if self._compression is not None:
            vals = [self._compress_roundtrip(ikey, i, v) for i, v in enumerate(vals)]
        if len(vals) == 1:
            return vals[0].detach().clone()
        if self._use_tree:
            return self._reduce_tree(ikey, vals)
        if self._on_device and vals[0].is_cuda:
            # merge on the device that holds the stored value (reference: least-loaded GPU, comm.h:678-712)
            dev = self._store[ikey].device if ikey in self._store and self._store[ikey].is_cuda else vals[0].device
            with torch.cuda.device(dev):
                staged = [self._stage(v, dev, i) for i, v in enumerate(vals)]
                out = torch.empty_like(staged[0])
                if out.dtype in (torch.float32, torch.bfloat16):
                    native.elementwise_sum([s.reshape(-1) for s in staged], out.reshape(-1))
                else:
                    out = torch.stack(staged).sum(0)
            return out
        # host reduce (CommCPU): copy every device gradient to the host, sum there with the native threaded reducer
        host = [v.detach().to("cpu").contiguous() for v in vals]
        if host[0].dtype in (torch.float32, torch.bfloat16) and all(h.dtype == host[0].dtype for h in host):
            return rt.reduce_sum_cpu([h.reshape(-1) for h in host]).view(vals[0].shape)
        acc = host[0].to(torch.float32).clone()
        for v in host[1:]:
            acc += v.to(torch.float32)
        return acc.to(vals[0].dtype)

    def init(self, key, value) -> None:
        keys, vals = _key_value_pairs(key, value)
        for k, v in zip(keys, vals):
            ik = self._ikey(k, create=True)
            if ik in self._store:
                raise KVStoreError(f"duplicate init of key {k}")
            t = v[0].detach().clone()
            self._store[ik] = t if (self._on_device or not t.is_cuda) else t.cpu()

    def push(self, key, value, priority: int = 0) -> None:
        keys, vals = _key_value_pairs(key, value)
        for k, v in zip(keys, vals):
            ik = self._ikey(k)
            if ik not in self._store:
                raise KVStoreError(f"key {k} has not been initialized")
            self._apply_merged(ik, k, self._reduce(ik, v))

    def _apply_merged(self, ik: int, k, merged: torch.Tensor) -> None:
        """The store's half of a push: the reduced value meets the updater (or replaces the stored value)."""
        if ik not in self._store:
            raise KVStoreError(f"key {k} has not been initialized")
        stored = self._store[ik]
        merged = merged.to(stored.device)
        if self._updater is not None:
            self._updater(ik if not isinstance(k, str) else k, merged, stored)
        else:
            stored.copy_(merged.to(stored.dtype))
        if self._serial_push and stored.is_cuda:
            torch.cuda.synchronize(stored.device)

    def pull(self, key, out=None, priority: int = 0, ignore_sparse: bool = True) -> None:
        if out is None:
            raise KVStoreError("pull needs `out`")
        keys, outs = _key_value_pairs(key, out)
        for k, o in zip(keys, outs):
            ik = self._ikey(k)
            if ik not in self._store:
                raise KVStoreError(f"key {k} has not been initialized")
            for t in o:
                t.copy_(self._store[ik].to(t.device, t.dtype), non_blocking=True)

    def row_sparse_pull(self, key, out=None, priority: int = 0, row_ids=None) -> None:
        """Pull only the rows named by row_ids (deduplicated + sorted, kvstore_local.h:239-261,395)."""
        if out is None or row_ids is None:
            raise KVStoreError("row_sparse_pull needs `out` and `row_ids`")
        keys, outs = _key_value_pairs(key, out)
        rids = _as_list(row_ids)
        for k, o in zip(keys, outs):
            ik = self._ikey(k)
            stored = self._store[ik]
            for j, t in enumerate(o):
                rid = rids[j if len(rids) > 1 else 0].to(torch.int64)
                if rid.is_cuda:
                    uniq = native.unique_rows(rid.reshape(-1), stored.shape[0])
                else:
                    uniq = torch.unique(rid.reshape(-1))
                t.zero_()
                t[uniq.to(t.device)] = stored[uniq.to(stored.device)].to(t.device, t.dtype)


# =============================================================================== nccl (baseline)
class KVStoreNCCL(KVStoreLocal):
    """Single-process multi-GPU reduce-to-root + broadcast through NCCL (kvstore_nccl.h:221-405).  Like the reference, a
    multi-key push / pull is issued in groups of up to 16 keys (kvstore_nccl.h:221-307: one ncclGroupStart / End per group):
    the values of a group that share a dtype are packed per device into one buffer, reduced / broadcast with ONE NCCL call and
    unpacked - one launch and one ring traversal instead of sixteen."""

    GROUP = 16

    def __init__(self, name: str = "nccl"):
        super().__init__(name)
        self._on_device = True

    def push(self, key, value, priority: int = 0) -> None:
        keys, vals = _key_value_pairs(key, value)
        groupable = len(keys) > 1 and all(len(v) > 1 and all(t.is_cuda for t in v) for v in vals) and self._compression is None
        if not groupable:
            return super().push(key, value, priority)
        import torch.cuda.nccl as nccl

        def reduce_packed(packed):
            out = torch.empty_like(packed[0])
            nccl.reduce(packed, output=out, root=0)
            return out

        self._push_grouped(keys, vals, priority, reduce_packed)

    def _push_grouped(self, keys, vals, priority, reduce_packed) -> None:
        """Groups of <= GROUP keys; a group whose values agree in dtype / device list travels as ONE packed buffer per device
        through `reduce_packed` (list of per-device buffers -> their sum on the first device) and is unpacked key by key into
        the updater, in key order - the order a per-key push would have used."""
        for g0 in range(0, len(keys), self.GROUP):
            gk, gv = keys[g0:g0 + self.GROUP], vals[g0:g0 + self.GROUP]
            same = len({(v[0].dtype, len(v), tuple(t.device for t in v)) for v in gv}) == 1
            if not same:
                for k, v in zip(gk, gv):
                    KVStoreLocal.push(self, k, v, priority)
                continue
            ndev = len(gv[0])
            packed = [torch.cat([v[d].detach().reshape(-1) for v in gv]) for d in range(ndev)]
            out = reduce_packed(packed)
            pos = 0
            for k, v in zip(gk, gv):
                n = v[0].numel()
                merged = out[pos:pos + n].view(v[0].shape)
                pos += n
                self._apply_merged(self._ikey(k), k, merged)

    def _reduce(self, ikey: int, vals: list) -> torch.Tensor:
        if len(vals) == 1 or not vals[0].is_cuda:
            return super()._reduce(ikey, vals)
        import torch.cuda.nccl as nccl
        ins = [v.detach().contiguous() for v in vals]
        out = torch.empty_like(ins[0])
        nccl.reduce(ins, output=out, root=0)
        return out

    def pull(self, key, out=None, priority: int = 0, ignore_sparse: bool = True) -> None:
        keys, outs = _key_value_pairs(key, out)
        for k, o in zip(keys, outs):
            ik = self._ikey(k)
            if ik not in self._store:
                raise KVStoreError(f"key {k} has not been initialized")
            if len(o) > 1 and all(t.is_cuda for t in o) and o[0].device == self._store[ik].device:
                import torch.cuda.nccl as nccl
                o[0].copy_(self._store[ik])
                nccl.broadcast(list(o), root=0)
            else:
                for t in o:
                    t.copy_(self._store[ik].to(t.device, t.dtype))

    def row_sparse_pull(self, *a, **kw):
        raise KVStoreError("the nccl kvstore is dense only (kvstore_nccl.h:181-189)")


# =============================================================================== dist (CPU wire)
class KVStoreDist(KVStoreLocal):
    """Worker side of dist_sync / dist_async over the native transport (reference KVStoreDist)."""

    def __init__(self, name: str):
        super().__init__(name)
        missing = [v for v in ("DMLC_PS_ROOT_URI", "DMLC_PS_ROOT_PORT", "DMLC_NUM_WORKER", "DMLC_NUM_SERVER") if v not in os.environ]
        if missing:
            # the reference aborts in ps::Postoffice with a CHECK on the same variables; without them the node would wait
            # forever for a scheduler on the default address
            raise KVStoreError(f"a `{name}` kvstore is a node of a parameter-server job: start the script through the launcher "
                               f"(`python -m deeplearning_cfn_b200 launch -n W -s S ...`), which sets {', '.join(missing)}")
        if not torch.cuda.is_available():
            from ..utils.cpu import share_cpu_threads
            share_cpu_threads()                      # co-located CPU workers / servers split the cores
        self._L = rt.lib()
        if self._L.dlcfn_ps_start() != 0:
            raise KVStoreError("cannot join the parameter-server job (is the scheduler up? DMLC_PS_ROOT_URI/PORT)")
        self._sync = "_async" not in name
        self._shapes: dict = {}
        self._bufs: dict = {}
        self._pending: dict = {}
        self._engine = rt.Engine(nthreads=int(os.environ.get("MXNET_KVSTORE_REDUCTION_NTHREADS", "4")))
        self._vars: dict = {}
        if not self._L.dlcfn_ps_is_recovery():
            self._L.dlcfn_ps_barrier(rt.GROUP_ALL)          # kvstore_dist.h:52-56
        if self.rank == 0 and self._sync:
            self._send_command_to_servers(rt.CMD_SYNC_MODE, "")   # kvstore.cc:55-58
        self._closed = False
        # The reference's destructor (kvstore_dist.h:62-76) runs when the process ends even if the script never drops the
        # store.  `__del__` is not reliable for objects that are still referenced at interpreter shutdown (a module-level
        # `kv = create(...)`), and a worker that skips the closing barrier leaves servers and peers waiting - so the
        # protocol is also run from an atexit hook that holds only a weak reference.
        import atexit
        import weakref
        ref = weakref.ref(self)

        def _close_at_exit(ref=ref):
            store = ref()
            # an unhandled exception is ending this worker: do not enter the closing barrier (its peers are not there -
            # they are blocked on this worker's missing push), just leave with the error code so the launcher ends the job
            if getattr(sys, "last_exc", None) is not None or getattr(sys, "last_type", None) is not None:
                return
            if store is not None:
                try:
                    store.close()
                except Exception:       # noqa: BLE001 - nothing useful can be done this late
                    pass
        atexit.register(_close_at_exit)

    @property
    def rank(self) -> int:
        return self._L.dlcfn_ps_rank()

    @property
    def num_workers(self) -> int:
        return self._L.dlcfn_ps_num_workers()

    @property
    def num_servers(self) -> int:
        return self._L.dlcfn_ps_num_servers()

    def _barrier(self) -> None:
        self._engine.wait_all()
        self._L.dlcfn_ps_barrier(rt.GROUP_WORKER)            # workers only (kvstore_dist.h:105-107)

    def _send_command_to_servers(self, head: int, body) -> None:
        b = body if isinstance(body, bytes) else str(body).encode()
        self._L.dlcfn_kv_send_command(int(head), b, len(b))

    def get_num_dead_node(self, node_id: int, timeout: int = 60) -> int:
        return self._L.dlcfn_ps_num_dead(node_id, timeout)

    def set_gradient_compression(self, compression_params: dict) -> None:
        super().set_gradient_compression(compression_params)
        if self._compression and self.rank == 0:
            self._send_command_to_servers(rt.CMD_SET_GC, f"2bit:{self._compression['threshold']}")
        self._barrier()

    def set_optimizer(self, optimizer) -> None:
        """Ships the optimizer to the servers (kvstore.py:480-497): natively understood optimizers as a config
        string for the C++ server, anything else pickled for the server-side python updater."""
        o = opt.create(optimizer) if not isinstance(optimizer, opt.Optimizer) else optimizer
        self._optimizer = o
        if self.rank == 0:
            cfg = o.native_config() if os.environ.get("DLCFN_PS_PY_UPDATER", "0") != "1" else None
            if cfg is not None:
                self._send_command_to_servers(rt.CMD_SET_OPTIMIZER, cfg)
            else:
                self._send_command_to_servers(rt.CMD_CONTROLLER, pickle.dumps(o, 0))
            if o.multi_precision:
                self._send_command_to_servers(rt.CMD_SET_MULTI_PRECISION, "")
        self._barrier()

    def _var(self, ik: int) -> int:
        if ik not in self._vars:
            self._vars[ik] = self._engine.new_var()
        return self._vars[ik]

    def _host(self, ik: int, like: torch.Tensor) -> torch.Tensor:
        """Pinned-host staging buffer per key (reference comm_buf_, kvstore_dist.h:335-342)."""
        b = self._bufs.get(ik)
        if b is None or b.numel() != like.numel() or b.dtype != like.dtype:
            b = torch.empty(like.numel(), dtype=like.dtype)
            if torch.cuda.is_available():
                b = b.pin_memory()
            self._bufs[ik] = b
        return b

    def init(self, key, value) -> None:
        keys, vals = _key_value_pairs(key, value)
        for k, v in zip(keys, vals):
            ik = self._ikey(k, create=True)
            if ik in self._shapes:
                raise KVStoreError(f"duplicate init of key {k}")
            t = v[0].detach()
            if t.dtype not in _DT:
                raise KVStoreError(f"unsupported dtype {t.dtype}")
            self._shapes[ik] = (tuple(t.shape), t.dtype)
            # shard boundaries of a big key: whole 16-element words of the 2-bit wire format and, for >1-D keys, whole rows
            # (row-sparse requests address rows by shard-local id) - same split for every request type of the key
            row_len = int(torch.tensor(t.shape[1:]).prod()) if t.dim() > 1 else 1
            self._L.dlcfn_kv_set_key_unit(ik, math.lcm(16, max(row_len, 1)))
            if self.rank == 0:   # only rank 0's value initialises the servers (kvstore_dist.h:196-208)
                h = self._host(ik, t)
                h.copy_(t.reshape(-1))
                ts = self._L.dlcfn_kv_push(ik, h.data_ptr(), h.numel(), _DT[t.dtype])
                self._L.dlcfn_kv_wait(ts)
        if not self._L.dlcfn_ps_is_recovery():
            self._L.dlcfn_ps_barrier(rt.GROUP_WORKER)

    def push(self, key, value, priority: int = 0) -> None:
        keys, vals = _key_value_pairs(key, value)
        for k, v in zip(keys, vals):
            ik = self._ikey(k)
            if ik not in self._shapes:
                raise KVStoreError(f"key {k} has not been initialized")
            shape, dtype = self._shapes[ik]
            if tuple(v[0].shape) != shape:
                raise KVStoreError(f"shape mismatch for key {k}: {tuple(v[0].shape)} vs {shape}")
            comp = self._compression
            if comp is not None:
                self._compression = None          # compress AFTER the local reduce, once (kvstore_dist.h:378-410)
            merged = self._reduce(ik, v)          # local multi-device reduce first (comm_->Reduce)
            self._compression = comp
            if merged.dtype != dtype:             # the wire carries the key's init dtype (a bf16 gradient of an fp32 key
                merged = merged.to(dtype)         # would otherwise be reinterpreted byte-wise by the native push)
            h = self._host(ik, merged)
            n = merged.numel()

            if comp is not None:
                thr = comp["threshold"]
                rk = (ik, "wire")
                if rk not in self._residuals:
                    self._residuals[rk] = torch.zeros(n, dtype=torch.float32)
                words = ref_ops.quantize_2bit(merged.detach().float().cpu().reshape(-1), self._residuals[rk], thr).contiguous()

                def do_push(ik=ik, words=words, n=n):
                    ts = self._L.dlcfn_kv_push_2bit(ik, words.data_ptr(), n)
                    self._L.dlcfn_kv_wait(ts)
            else:
                # the staging buffer is borrowed by a queued push of the same key until that push has been sent: wait for it
                # before overwriting (the reference copies into comm_buf inside the engine op, kvstore_dist.h:335-342)
                self._engine.wait_for_var(self._var(ik))
                h.copy_(merged.reshape(-1))

                def do_push(ik=ik, h=h, dtype=dtype):
                    ts = self._L.dlcfn_kv_push(ik, h.data_ptr(), h.numel(), _DT[dtype])
                    self._L.dlcfn_kv_wait(ts)        # in sync mode the reply IS the per-key barrier
            # asynchronous, ordered against later pulls of the same key through its variable
            self._engine.push(do_push, reads=(), writes=(self._var(ik),), priority=priority)

    def pull(self, key, out=None, priority: int = 0, ignore_sparse: bool = True) -> None:
        if out is None:
            raise KVStoreError("pull needs `out`")
        keys, outs = _key_value_pairs(key, out)
        done = []
        for k, o in zip(keys, outs):
            ik = self._ikey(k)
            if ik not in self._shapes:
                raise KVStoreError(f"key {k} has not been initialized")
            shape, dtype = self._shapes[ik]
            recv = torch.empty(int(torch.tensor(shape).prod()) if shape else 1, dtype=dtype)

            def do_pull(ik=ik, recv=recv, dtype=dtype):
                ts = self._L.dlcfn_kv_pull(ik, recv.data_ptr(), recv.numel(), _DT[dtype])
                self._L.dlcfn_kv_wait(ts)
            self._engine.push(do_pull, reads=(), writes=(self._var(ik),), priority=priority)
            done.append((ik, recv, shape, o))
        for ik, recv, shape, o in done:
            self._engine.wait_for_var(self._var(ik))
            for t in o:        # comm_->Broadcast to every local device
                t.copy_(recv.view(shape).to(t.device, t.dtype), non_blocking=True)

    def row_sparse_pull(self, key, out=None, priority: int = 0, row_ids=None) -> None:
        if out is None or row_ids is None:
            raise KVStoreError("row_sparse_pull needs `out` and `row_ids`")
        keys, outs = _key_value_pairs(key, out)
        rids = _as_list(row_ids)
        for k, o in zip(keys, outs):
            ik = self._ikey(k)
            shape, dtype = self._shapes[ik]
            row_len = int(torch.tensor(shape[1:]).prod()) if len(shape) > 1 else 1
            total = shape[0] * row_len
            self._engine.wait_for_var(self._var(ik))
            for j, t in enumerate(o):
                uniq = torch.unique(rids[j if len(rids) > 1 else 0].to(torch.int64).cpu().reshape(-1)).contiguous()
                rows = torch.zeros(uniq.numel() * row_len, dtype=dtype)
                ts = self._L.dlcfn_kv_pull_rsp(ik, rows.data_ptr(), uniq.data_ptr(), uniq.numel(), row_len, total, _DT[dtype])
                self._L.dlcfn_kv_wait(ts)
                t.zero_()
                t[uniq.to(t.device)] = rows.view(uniq.numel(), *shape[1:]).to(t.device, t.dtype)

    def push_row_sparse(self, key, rows: torch.Tensor, row_ids: torch.Tensor, priority: int = 0) -> None:
        """Row-sparse gradient push: only `rows` (for `row_ids`) travel (kvstore_dist.h:437-513)."""
        ik = self._ikey(key)
        shape, dtype = self._shapes[ik]
        row_len = int(torch.tensor(shape[1:]).prod()) if len(shape) > 1 else 1
        total = shape[0] * row_len
        r = rows.detach().to("cpu", dtype).contiguous()
        ids = row_ids.detach().to("cpu", torch.int64).contiguous()

        def do_push():
            ts = self._L.dlcfn_kv_push_rsp(ik, r.data_ptr(), ids.data_ptr(), ids.numel(), row_len, total, _DT[dtype])
            self._L.dlcfn_kv_wait(ts)
        self._engine.push(do_push, reads=(), writes=(self._var(ik),), priority=priority)

    def save_optimizer_states(self, fname, dump_optimizer=False):
        raise KVStoreError("Cannot save states for distributed training")   # kvstore.py:550

    def load_optimizer_states(self, fname):
        raise KVStoreError("Cannot load states for distributed training")

    def close(self) -> None:
        """Destructor protocol of KVStoreDist (kvstore_dist.h:62-76): barrier, rank 0 stops the servers."""
        if self._closed:
            return
        self._closed = True
        self._engine.wait_all()
        # every live worker takes part in the closing barrier - a node that rejoined after a failure skipped the START-UP
        # barriers (its peers had passed them long before), but at shutdown its peers are waiting for it here
        self._L.dlcfn_ps_barrier(rt.GROUP_WORKER)
        if self.rank == 0:
            self._send_command_to_servers(rt.CMD_STOP_SERVER, "")
        self._engine.close()
        self._L.dlcfn_ps_finalize(1)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# =============================================================================== dist_device_sync on peer HBM
class KVStoreDistDevice(KVStore):
    """dist_device_sync with one process per GPU: keys live in the symmetric peer-HBM heap, every rank is a
    worker and a server shard, and push+reduce+update+pull is one fused kernel per key."""

    def __init__(self, name: str = "dist_device_sync", heap_bytes: int | None = None):
        super().__init__(name)
        import torch.distributed as dist
        from ..parallel.symm import SymmetricHeap
        if not (dist.is_available() and dist.is_initialized() and torch.cuda.is_available()):
            raise KVStoreError("KVStoreDistDevice needs an initialized torch.distributed job on CUDA")
        self._dist = dist
        self._rank, self._world = dist.get_rank(), dist.get_world_size()
        self.heap = SymmetricHeap(heap_bytes or int(os.environ.get("DLCFN_KV_HEAP_BYTES", str(1 << 30))))
        self._keys: dict = {}
        self._opt_kind, self._hp = "none", {}
        self._bigarray = int(os.environ.get("MXNET_KVSTORE_BIGARRAY_BOUND", "1000000"))

    @property
    def rank(self):
        return self._rank

    @property
    def num_workers(self):
        return self._world

    def _barrier(self):
        self.heap.barrier()
        torch.cuda.current_stream().synchronize()

    def set_optimizer(self, optimizer) -> None:
        o = opt.create(optimizer) if not isinstance(optimizer, opt.Optimizer) else optimizer
        cfg = o.native_config()
        if cfg is None:
            raise KVStoreError(f"{type(o).__name__} has no fused server kernel (csrc/cuda/optimizer.cuh lists the 13 that do)")
        self._optimizer = o
        d = dict(kv.split("=") for kv in cfg.split(";"))
        self._opt_kind = d.pop("kind")
        self._hp = {"lr": float(d.get("lr", 0)), "wd": float(d.get("wd", 0)), "momentum": float(d.get("momentum", 0)),
                    "beta1": float(d.get("beta1", 0.9)), "beta2": float(d.get("beta2", 0.999)),
                    "eps": float(d.get("eps", 1e-8)), "rescale": float(d.get("rescale", 1.0)),
                    "clip": float(d.get("clip", -1.0))}
        for st in self._keys.values():
            self._alloc_state(st)

    def _shard(self, ikey: int, n: int, esize: int) -> tuple[int, int]:
        vec = 16 // esize
        nvec = (n + vec - 1) // vec
        if n < self._bigarray:     # small key: one server (key*9973) % S
            owner = (ikey * 9973) % self._world
            return (0, nvec * vec) if owner == self._rank else (0, 0)
        per = (nvec + self._world - 1) // self._world
        return min(self._rank * per, nvec) * vec, min((self._rank + 1) * per, nvec) * vec

    def _alloc_state(self, st: dict) -> None:
        n = st["hi"] - st["lo"]
        dev = self.heap.device
        need1 = self._opt_kind in native.OPT_USES_S1
        need2 = self._opt_kind in native.OPT_USES_S2
        if n and st.get("master") is None:
            st["master"] = st["w"][st["lo"]:st["hi"]].float().clone()
        st["s1"] = torch.zeros(n, device=dev) if (need1 and n) else None
        st["s2"] = torch.zeros(n, device=dev) if (need2 and n) else None

    def init(self, key, value) -> None:
        keys, vals = _key_value_pairs(key, value)
        for k, v in zip(keys, vals):
            ik = self._ikey(k, create=True)
            if ik in self._keys:
                raise KVStoreError(f"duplicate init of key {k}")
            t = v[0].detach()
            if t.dtype not in (torch.float32, torch.bfloat16):
                raise KVStoreError("dist_device_sync keys must be float32 or bfloat16")
            esz = t.element_size()
            vec = 16 // esz
            npad = (t.numel() + vec - 1) // vec * vec
            goff, g = self.heap.alloc_tensor(npad, t.dtype)
            woff, w = self.heap.alloc_tensor(npad, t.dtype)
            g.zero_()
            w.zero_()
            w[:t.numel()].copy_(t.reshape(-1))
            lo, hi = self._shard(ik, t.numel(), esz)
            st = {"goff": goff, "woff": woff, "g": g, "w": w, "n": t.numel(), "npad": npad, "shape": tuple(t.shape),
                  "dtype": t.dtype, "lo": lo, "hi": hi, "master": None}
            self._keys[ik] = st
            # only rank 0's value counts: broadcast it into every rank's weight region
            self.heap.broadcast_(woff, npad * esz, root=0)
            torch.cuda.current_stream().synchronize()
            self._alloc_state(st)

    def push(self, key, value, priority: int = 0) -> None:
        keys, vals = _key_value_pairs(key, value)
        for k, v in zip(keys, vals):
            ik = self._ikey(k)
            if ik not in self._keys:
                raise KVStoreError(f"key {k} has not been initialized")
            st = self._keys[ik]
            g = st["g"][:st["n"]]
            if len(v) == 1:
                g.copy_(v[0].reshape(-1))
            else:
                native.elementwise_sum([x.reshape(-1).to(g.device, g.dtype).contiguous() for x in v], g)
            if self._compression is not None:
                # 2-bit wire format with error feedback (K9): what this worker contributes is dequantize(quantize(g +
                # residual)); the fused server kernel then sums the workers' contributions as usual.  On NVLink the
                # saved bytes do not matter, the point is the reference's numerical semantics.
                g.copy_(self._compress_roundtrip(ik, 0, g).reshape(-1))
            hp = dict(self._hp) if self._hp else {"lr": 0.0}
            kind = self._opt_kind
            if kind == "adam":
                o = self._optimizer
                t = o._update_count(ik)
                hp["lr"] = hp["lr"] * (1 - hp["beta2"] ** t) ** 0.5 / (1 - hp["beta1"] ** t)
            elif kind == "adamax":
                hp["lr"] = hp["lr"] / (1 - hp["beta1"] ** self._optimizer._update_count(ik))
            self.heap.ps_sync_step(st["goff"], st["woff"], st["npad"], st["lo"], st["hi"], (1 << self._world) - 1,
                                   st["master"], st["s1"], st["s2"], st["dtype"], kind, **hp)

    def pull(self, key, out=None, priority: int = 0, ignore_sparse: bool = True) -> None:
        if out is None:
            raise KVStoreError("pull needs `out`")
        keys, outs = _key_value_pairs(key, out)
        for k, o in zip(keys, outs):
            ik = self._ikey(k)
            if ik not in self._keys:
                raise KVStoreError(f"key {k} has not been initialized")
            st = self._keys[ik]
            for t in o:
                t.copy_(st["w"][:st["n"]].view(st["shape"]).to(t.dtype))

    def row_sparse_pull(self, key, out=None, priority: int = 0, row_ids=None) -> None:
        keys, outs = _key_value_pairs(key, out)
        rids = _as_list(row_ids)
        for k, o in zip(keys, outs):
            st = self._keys[self._ikey(k)]
            full = st["w"][:st["n"]].view(st["shape"])
            for j, t in enumerate(o):
                uniq = native.unique_rows(rids[j if len(rids) > 1 else 0].to(full.device, torch.int64).reshape(-1), full.shape[0])
                t.zero_()
                t[uniq] = full[uniq].to(t.dtype)

    def save_optimizer_states(self, fname, dump_optimizer=False):
        raise KVStoreError("Cannot save states for distributed training")


# =============================================================================== factory
def create(name: str = "local") -> KVStore:
    """kvstore.cc:40-77 substring rules."""
    if not isinstance(name, str):
        raise TypeError("name must be a string")
    n = name.lower()
    if "dist" in n:
        import torch.distributed as dist
        gpu_job = dist.is_available() and dist.is_initialized() and torch.cuda.is_available() and \
            dist.get_backend() == "nccl"
        if "device" in n and gpu_job and "DMLC_PS_ROOT_URI" not in os.environ:
            return KVStoreDistDevice(name)
        return KVStoreDist(name)
    if "nccl" in n:
        return KVStoreNCCL(name)
    if n in ("local", "device") or n.startswith("local") or "device" in n:
        return KVStoreLocal(name)
    raise KVStoreError(f"unknown kvstore type {name!r}")
