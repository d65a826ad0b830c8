"""ctypes bindings for the native CPU runtime `libdlcfn_rt.so` (csrc/rt): parameter-server transport
(ps-lite replacement), dependency engine, fabric planner."""
from __future__ import annotations

import ctypes
import threading
from ctypes import (CFUNCTYPE, POINTER, c_char_p, c_double, c_float, c_int, c_int64, c_longlong, c_size_t,
                    c_void_p)
from pathlib import Path

import os as _os

# DLCFN_RT_LIB points at an alternative build of the runtime (tools/sanitize_rt.sh: address-sanitizer build)
_LIB_PATH = Path(_os.environ.get("DLCFN_RT_LIB") or (Path(__file__).resolve().parent.parent / "_lib" / "libdlcfn_rt.so"))
_lib = None
_lock = threading.Lock()

ROLE_SCHEDULER, ROLE_SERVER, ROLE_WORKER = 0, 1, 2
GROUP_SCHEDULER, GROUP_SERVER, GROUP_WORKER = 1, 2, 4
GROUP_ALL = 7
# command heads (csrc/rt/ps.h CommandType; reference kvstore_dist_server.h:46-49)
CMD_CONTROLLER, CMD_SET_MULTI_PRECISION, CMD_STOP_SERVER, CMD_SYNC_MODE, CMD_SET_GC, CMD_SET_PROFILER, CMD_SET_OPTIMIZER = range(7)
DT_F32, DT_BF16, DT_F16 = 0, 1, 2

ENGINE_FN = CFUNCTYPE(c_int, c_void_p)
UPDATER_FN = CFUNCTYPE(None, c_int, POINTER(c_float), POINTER(c_float), c_size_t, c_void_p)
CONTROLLER_FN = CFUNCTYPE(None, c_int, c_void_p, c_size_t, c_void_p)


def lib_path() -> Path:
    return _LIB_PATH


def lib():
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not _LIB_PATH.exists():
                from .. import _build
                _build.build_rt()
            L = ctypes.CDLL(str(_LIB_PATH))
            _declare(L)
            _lib = L
    return _lib


def _declare(L) -> None:
    L.dlcfn_ps_start.restype = c_int
    L.dlcfn_ps_finalize.argtypes = [c_int]
    for f in ("dlcfn_ps_role", "dlcfn_ps_rank", "dlcfn_ps_node_id", "dlcfn_ps_num_workers", "dlcfn_ps_num_servers",
              "dlcfn_ps_is_recovery"):
        getattr(L, f).restype = c_int
    L.dlcfn_ps_barrier.argtypes = [c_int]
    L.dlcfn_ps_num_dead.argtypes = [c_int, c_int]
    L.dlcfn_ps_num_dead.restype = c_int
    L.dlcfn_ps_bytes_sent.restype = c_longlong
    L.dlcfn_ps_bytes_recv.restype = c_longlong
    L.dlcfn_kv_push.argtypes = [c_int, c_void_p, c_size_t, c_int]
    L.dlcfn_kv_push.restype = c_int
    L.dlcfn_kv_pull.argtypes = [c_int, c_void_p, c_size_t, c_int]
    L.dlcfn_kv_pull.restype = c_int
    L.dlcfn_kv_push_rsp.argtypes = [c_int, c_void_p, c_void_p, c_size_t, c_size_t, c_size_t, c_int]
    L.dlcfn_kv_push_rsp.restype = c_int
    L.dlcfn_kv_pull_rsp.argtypes = [c_int, c_void_p, c_void_p, c_size_t, c_size_t, c_size_t, c_int]
    L.dlcfn_kv_pull_rsp.restype = c_int
    L.dlcfn_kv_push_2bit.argtypes = [c_int, c_void_p, c_size_t]
    L.dlcfn_kv_push_2bit.restype = c_int
    L.dlcfn_kv_wait.argtypes = [c_int]
    L.dlcfn_kv_send_command.argtypes = [c_int, c_char_p, c_size_t]
    L.dlcfn_kv_set_bigarray_bound.argtypes = [c_size_t]
    L.dlcfn_kv_set_key_unit.argtypes = [c_int, c_size_t]
    L.dlcfn_kv_partition.argtypes = [c_int, c_size_t, POINTER(c_int), POINTER(c_size_t), POINTER(c_size_t), c_int]
    L.dlcfn_kv_partition.restype = c_int
    L.dlcfn_kv_server_set_updater.argtypes = [UPDATER_FN, c_void_p]
    L.dlcfn_kv_server_set_controller.argtypes = [CONTROLLER_FN, c_void_p]
    L.dlcfn_engine_create.argtypes = [c_int, c_int]
    L.dlcfn_engine_create.restype = c_void_p
    L.dlcfn_engine_destroy.argtypes = [c_void_p]
    L.dlcfn_engine_new_var.argtypes = [c_void_p]
    L.dlcfn_engine_new_var.restype = c_int
    L.dlcfn_engine_push.argtypes = [c_void_p, ENGINE_FN, c_void_p, POINTER(c_int), c_int, POINTER(c_int), c_int, c_int]
    L.dlcfn_engine_wait_var.argtypes = [c_void_p, c_int]
    L.dlcfn_engine_wait_var.restype = c_int
    L.dlcfn_engine_wait_all.argtypes = [c_void_p]
    L.dlcfn_topo_is_uniform.argtypes = [POINTER(c_double), c_int]
    L.dlcfn_topo_is_uniform.restype = c_int
    L.dlcfn_topo_plan_trees.argtypes = [POINTER(c_double), c_int, c_double, POINTER(c_int)]
    L.dlcfn_topo_plan_trees.restype = c_int
    # handle-based C-level store API (csrc/rt/kvstore_capi.cpp; for C / C++ programs - declared here for the ABI check and tests)
    from ctypes import c_char_p as _ccp, c_void_p as _cvp
    L.DLCFNKVGetLastError.argtypes = []
    L.DLCFNKVGetLastError.restype = _ccp
    L.DLCFNKVStoreCreate.argtypes = [_ccp, POINTER(_cvp)]
    L.DLCFNKVStoreCreate.restype = c_int
    L.DLCFNKVStoreFree.argtypes = [_cvp]
    L.DLCFNKVStoreFree.restype = c_int
    L.dlcfn_topo_plan_trees_backtrack.argtypes = [POINTER(c_double), c_int, c_double, POINTER(c_int)]
    L.dlcfn_topo_plan_trees_backtrack.restype = c_int
    L.dlcfn_topo_plan_ring.argtypes = [POINTER(c_double), c_int, c_int, POINTER(c_int)]
    L.dlcfn_topo_plan_ring.restype = c_int
    from ctypes import c_char_p as _bytes_p, c_size_t as _sz
    L.dlcfn_rt_debug_unpack.argtypes = [_bytes_p, _sz]         # test hook: decode + re-encode one frame payload (van.cpp)
    L.dlcfn_rt_debug_unpack.restype = c_int


# --------------------------------------------------------------------------- fabric planner
def _mat(link):
    n = len(link)
    arr = (c_double * (n * n))(*[float(link[i][j]) for i in range(n) for j in range(n)])
    return n, arr


def reduce_sum_cpu(srcs: list, out=None, nthreads: int | None = None, bigarray_bound: int | None = None):
    """out = sum(srcs) on the host with the native reducer (csrc/rt/reduce.cpp; reference CommCPU::ReduceSumCPU):
    4 inputs per pass, arrays >= MXNET_KVSTORE_BIGARRAY_BOUND elements split into 4096-element chunks over
    MXNET_KVSTORE_REDUCTION_NTHREADS threads.  fp32 / bf16 contiguous CPU tensors of equal size."""
    import os
    import torch
    L = lib()
    if not getattr(L, "_reduce_declared", False):
        L.dlcfn_reduce_sum_cpu.argtypes = [c_void_p, c_int, c_void_p, c_size_t, c_int, c_int, c_size_t]
        L.dlcfn_reduce_sum_cpu.restype = c_int
        L._reduce_declared = True
    t0 = srcs[0]
    code = {torch.float32: 0, torch.bfloat16: 1}.get(t0.dtype)
    if code is None:
        raise ValueError(f"reduce_sum_cpu supports float32 and bfloat16, got {t0.dtype}")
    srcs = [s.detach().contiguous() for s in srcs]
    assert all((not s.is_cuda) and s.dtype == t0.dtype and s.numel() == t0.numel() for s in srcs)
    if out is None:
        out = torch.empty_like(srcs[0])
    assert out.is_contiguous() and out.numel() == t0.numel() and out.dtype == t0.dtype
    nthreads = nthreads if nthreads is not None else int(os.environ.get("MXNET_KVSTORE_REDUCTION_NTHREADS", "4"))
    bound = bigarray_bound if bigarray_bound is not None else int(os.environ.get("MXNET_KVSTORE_BIGARRAY_BOUND", "1000000"))
    arr = (c_void_p * len(srcs))(*[s.data_ptr() for s in srcs])
    if L.dlcfn_reduce_sum_cpu(ctypes.cast(arr, c_void_p), len(srcs), out.data_ptr(), t0.numel(), code, nthreads, bound) != 0:
        raise ValueError("reduce_sum_cpu: bad arguments")
    return out


def topo_is_uniform(link) -> bool:
    n, arr = _mat(link)
    return bool(lib().dlcfn_topo_is_uniform(arr, n))


def plan_trees(link, penalty: float | None = None) -> list[list[int]]:
    """One reduction tree per root: result[r][i] = parent of GPU i in the tree rooted at r (-1 = root)."""
    import os
    if penalty is None:
        penalty = float(os.environ.get("MXNET_KVSTORE_TREE_LINK_USAGE_PENALTY", "0.7"))
    n, arr = _mat(link)
    out = (c_int * (n * n))()
    if os.environ.get("MXNET_KVSTORE_TREE_BACKTRACK", "0") == "1":     # gpu_topology.h:789-957: exhaustive search, existing links only
        rc = lib().dlcfn_topo_plan_trees_backtrack(arr, n, penalty, out)
        if rc == -2:
            raise ValueError("the link matrix has no spanning tree over direct links for some root")
        if rc != 0:
            raise ValueError("bad link matrix")
    elif lib().dlcfn_topo_plan_trees(arr, n, penalty, out) != 0:
        raise ValueError("bad link matrix")
    trees = [[out[r * n + i] for i in range(n)] for r in range(n)]
    if os.environ.get("MXNET_KVSTORE_LOGTREE", "0") == "1":      # gpu_topology.h:46
        print("link matrix:")
        for r in range(n):
            print("  " + " ".join(f"{arr[r * n + c]:g}" for c in range(n)))
        for r, t in enumerate(trees):
            print(f"tree rooted at {r}: parent = {t}")
    return trees


def plan_ring(link, start: int = 0) -> list[int]:
    n, arr = _mat(link)
    out = (c_int * n)()
    if lib().dlcfn_topo_plan_ring(arr, n, start, out) != 0:
        raise ValueError("bad ring request")
    return list(out)


# --------------------------------------------------------------------------- dependency engine
class Engine:
    """Versioned-variable dependency engine (reference E1).  `push(fn, reads, writes, priority)` runs fn on a
    worker thread once its dependencies are met; larger priority runs earlier."""

    def __init__(self, nthreads: int = 0, naive: bool = False):
        self._L = lib()
        self._h = self._L.dlcfn_engine_create(nthreads, int(naive))
        self._keep: dict[int, object] = {}
        self._keep_lock = threading.Lock()
        self._next = 0
        self.errors: list[BaseException] = []

    def new_var(self) -> int:
        return self._L.dlcfn_engine_new_var(self._h)

    def push(self, fn, reads=(), writes=(), priority: int = 0) -> None:
        with self._keep_lock:
            token = self._next
            self._next += 1

        def tramp(_user):
            try:
                fn()
                rc = 0
            except BaseException as e:  # propagate through the written vars (threaded_engine.h:179-180)
                self.errors.append(e)
                rc = 1
            with self._keep_lock:
                self._keep.pop(token, None)
            return rc

        cb = ENGINE_FN(tramp)
        with self._keep_lock:
            self._keep[token] = cb
        r = (c_int * len(reads))(*reads)
        w = (c_int * len(writes))(*writes)
        self._L.dlcfn_engine_push(self._h, cb, None, r, len(reads), w, len(writes), priority)

    def wait_for_var(self, var: int) -> None:
        if self._L.dlcfn_engine_wait_var(self._h, var) != 0 and self.errors:
            raise self.errors.pop(0)

    def wait_all(self) -> None:
        self._L.dlcfn_engine_wait_all(self._h)
        if self.errors:
            raise self.errors.pop(0)

    def close(self) -> None:
        if self._h:
            self._L.dlcfn_engine_wait_all(self._h)
            self._L.dlcfn_engine_destroy(self._h)
            self._h = None
