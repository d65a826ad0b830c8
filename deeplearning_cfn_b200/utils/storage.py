"""Storage managers (reference E3, MX/src/storage/): pooled pinned-host memory, pooled device memory, and named
shared-memory segments, all backed by the native pools in csrc/rt/storage.cpp and csrc/cuda/storage.cu.

    buf = storage.pinned_empty((256, 224, 224, 3), torch.uint8)   # page-locked, recycled by size class
    storage.release(buf)
    seg = storage.SharedTensor.create("batch0", (256, 3, 224, 224), torch.float32)   # producer process
    view = storage.SharedTensor.attach("batch0", (256, 3, 224, 224), torch.float32)  # consumer process

Env (same names as the reference): MXNET_GPU_MEM_POOL_TYPE=Naive|Round, MXNET_GPU_MEM_POOL_ROUND_LINEAR_CUTOFF,
MXNET_GPU_MEM_POOL_RESERVE.
"""
from __future__ import annotations

import ctypes
import math
from ctypes import POINTER, c_int, c_size_t, c_uint64, c_ulonglong, c_void_p

import torch

from . import rt

_DECL = False
_pinned: set[int] = set()                 # base addresses registered with cudaHostRegister
_live: dict[int, tuple[int, int]] = {}    # tensor data_ptr -> (base, nbytes)


def _L():
    global _DECL
    L = rt.lib()
    if not _DECL:
        L.dlcfn_storage_alloc.argtypes = [c_size_t]
        L.dlcfn_storage_alloc.restype = c_void_p
        L.dlcfn_storage_free.argtypes = [c_void_p]
        L.dlcfn_storage_free.restype = c_int
        L.dlcfn_storage_size_class.argtypes = [c_size_t]
        L.dlcfn_storage_size_class.restype = c_size_t
        L.dlcfn_storage_set_strategy.argtypes = [c_int, c_int]
        L.dlcfn_storage_stats.argtypes = [POINTER(c_uint64)]
        L.dlcfn_storage_list_pooled.argtypes = [POINTER(c_void_p), c_size_t]
        L.dlcfn_storage_list_pooled.restype = c_size_t
        L.dlcfn_storage_release_all.restype = c_size_t
        L.dlcfn_shm_open.argtypes = [ctypes.c_char_p, c_size_t, c_int, POINTER(c_size_t)]
        L.dlcfn_shm_open.restype = c_void_p
        L.dlcfn_shm_close.argtypes = [c_void_p]
        L.dlcfn_shm_close.restype = c_int
        _DECL = True
    return L


def _tensor_over(addr: int, nbytes: int, shape, dtype: torch.dtype) -> torch.Tensor:
    buf = (ctypes.c_uint8 * nbytes).from_address(addr)
    return torch.frombuffer(buf, dtype=dtype, count=math.prod(shape)).view(shape)


def set_strategy(kind: str = "Naive", cutoff_log2: int = 0) -> None:
    _L().dlcfn_storage_set_strategy(1 if kind.lower() == "round" else 0, cutoff_log2)


def size_class(nbytes: int) -> int:
    return int(_L().dlcfn_storage_size_class(nbytes))


def host_empty(shape, dtype: torch.dtype = torch.uint8, pin: bool = False) -> torch.Tensor:
    """A tensor over a pooled, page-aligned host block.  pin=True page-locks the block (once per block: a recycled block
    stays registered) so that H2D copies from it are asynchronous DMA."""
    shape = tuple(shape) if not isinstance(shape, int) else (shape,)
    nbytes = max(1, math.prod(shape) * torch.empty((), dtype=dtype).element_size())
    L = _L()
    base = L.dlcfn_storage_alloc(nbytes)
    if not base:
        raise MemoryError(f"host pool could not allocate {nbytes} bytes")
    if pin and torch.cuda.is_available() and base not in _pinned:
        rc = torch.cuda.cudart().cudaHostRegister(base, size_class(nbytes), 0)
        if int(rc) != 0:
            L.dlcfn_storage_free(base)
            raise RuntimeError(f"cudaHostRegister failed with {rc}")
        _pinned.add(base)
    t = _tensor_over(base, nbytes, shape, dtype)
    _live[t.data_ptr()] = (base, nbytes)
    return t


def pinned_empty(shape, dtype: torch.dtype = torch.uint8) -> torch.Tensor:
    return host_empty(shape, dtype, pin=True)


def release(t: torch.Tensor) -> None:
    """Hand the block behind `t` back to the pool (the tensor must not be used afterwards)."""
    ent = _live.pop(t.data_ptr(), None)
    if ent is None:
        raise ValueError("tensor does not come from utils.storage")
    if _L().dlcfn_storage_free(ent[0]) != 0:
        raise RuntimeError("double release")


def stats() -> dict:
    a = (c_uint64 * 5)()
    _L().dlcfn_storage_stats(a)
    return {"bytes_live": a[0], "bytes_pooled": a[1], "fresh_allocations": a[2], "reuses": a[3], "live_blocks": a[4]}


def release_all() -> int:
    """Free every pooled (not live) host block, unpinning first.  Returns the number of blocks released."""
    L = _L()
    n = L.dlcfn_storage_list_pooled(None, 0)
    if n and _pinned and torch.cuda.is_available():
        arr = (c_void_p * n)()
        L.dlcfn_storage_list_pooled(arr, n)
        for p in arr:
            if p in _pinned:
                torch.cuda.cudart().cudaHostUnregister(p)
                _pinned.discard(p)
    return int(L.dlcfn_storage_release_all())


class SharedTensor:
    """A tensor in a named POSIX shared-memory segment (CPUSharedStorageManager).  The creator owns the name: closing it
    unlinks the segment; attachers keep their mapping until they close."""

    def __init__(self, name: str, addr: int, nbytes: int, shape, dtype, owner: bool):
        self.name, self._addr, self.nbytes, self.owner = name, addr, nbytes, owner
        self.tensor = _tensor_over(addr, nbytes, tuple(shape), dtype)

    @staticmethod
    def _open(name, shape, dtype, create):
        nbytes = max(1, math.prod(shape) * torch.empty((), dtype=dtype).element_size())
        out = c_size_t(0)
        key = ("/" + name.lstrip("/")).encode()
        addr = _L().dlcfn_shm_open(key, nbytes, 1 if create else 0, ctypes.byref(out))
        if not addr:
            err = ctypes.get_errno()
            raise OSError(err, f"shared segment {name!r} could not be {'created' if create else 'attached'}")
        return SharedTensor(name, addr, nbytes, shape, dtype, create)

    @classmethod
    def create(cls, name: str, shape, dtype: torch.dtype = torch.float32) -> "SharedTensor":
        return cls._open(name, tuple(shape), dtype, True)

    @classmethod
    def attach(cls, name: str, shape, dtype: torch.dtype = torch.float32) -> "SharedTensor":
        return cls._open(name, tuple(shape), dtype, False)

    def close(self) -> None:
        if self._addr:
            self.tensor = None
            _L().dlcfn_shm_close(self._addr)
            self._addr = 0

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


# ------------------------------------------------------------------------------------------------- device pool
class DevicePool:
    """Pooled cudaMalloc blocks for the runtime's own recurring device buffers (csrc/cuda/storage.cu)."""

    def __init__(self):
        from ..ops import native
        self._n = native
        L = native.lib()
        if not getattr(L, "_pool_declared2", False):
            L.dlcfn_pool_alloc.argtypes = [c_size_t, POINTER(c_void_p)]
            L.dlcfn_pool_alloc.restype = c_int
            L.dlcfn_pool_free.argtypes = [c_void_p]
            L.dlcfn_pool_free.restype = c_int
            L.dlcfn_pool_set_strategy.argtypes = [c_int, c_int, c_int]
            L.dlcfn_pool_stats.argtypes = [POINTER(c_ulonglong)]
            L._pool_declared2 = True
        self.L = L

    def alloc(self, nbytes: int) -> int:
        p = c_void_p(0)
        self._n.check(self.L.dlcfn_pool_alloc(nbytes, ctypes.byref(p)), "pool_alloc")
        return int(p.value)

    def free(self, ptr: int) -> None:
        self._n.check(self.L.dlcfn_pool_free(ptr), "pool_free")

    def set_strategy(self, kind: str = "Naive", cutoff_log2: int = 0, reserve_pct: int = -1) -> None:
        self.L.dlcfn_pool_set_strategy(1 if kind.lower() == "round" else 0, cutoff_log2, reserve_pct)

    def release_all(self) -> None:
        self.L.dlcfn_pool_release_all()

    def stats(self) -> dict:
        a = (c_ulonglong * 5)()
        self.L.dlcfn_pool_stats(a)
        return {"bytes_live": a[0], "bytes_pooled": a[1], "fresh_allocations": a[2], "reuses": a[3], "live_blocks": a[4]}
