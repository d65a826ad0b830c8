"""Symmetric peer-HBM heap (python side).

The B200 replacement for three reference mechanisms at once: the EFS shared mount used to stage
data/initial weights (deeplearning.template:453-474), the pinned-host `comm_buf_` of KVStoreDist
(kvstore_dist.h:335-342) and the ZeroMQ data plane of ps-lite (zmq_van.h).  Every rank owns one arena that all
peers map, after which the fused kernels in csrc/cuda/collectives.cu address any peer's HBM with plain
loads/stores over NVLink 5 / NVSwitch.

Backends (DLCFN_SYMM_BACKEND = auto | vmm | torch | ipc; auto tries them in this order):
  vmm    the library's own plumbing (csrc/cuda/symm_vmm.cu): cuMemCreate arenas exported as file descriptors, passed
         between the ranks of the node over abstract AF_UNIX sockets (SCM_RIGHTS) and mapped with cuMemMap; NVLS multicast
         object created by rank 0 (cuMulticastCreate / AddDevice / BindMem) - no private torch API involved
  torch  torch.distributed._symmetric_memory (same mechanism, borrowed)
  ipc    one cudaMalloc arena per rank + CUDA-IPC handles through the torch.distributed store (no multicast)
`stage_put` / `stage_get` are the shared-staging API on top (initial weights, shared data: the EFS analogue).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_int, c_void_p

import torch
import torch.distributed as dist

from ..ops import native

_TYPESTR = {torch.uint8: "|u1", torch.float32: "<f4", torch.int32: "<i4", torch.int64: "<i8",
            torch.float16: "<f2"}


class _RawCuda:
    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}


def tensor_from_ptr(ptr: int, nbytes: int, device: torch.device) -> torch.Tensor:
    """uint8 view of raw device memory (no ownership)."""
    return torch.as_tensor(_RawCuda(ptr, nbytes), device=device)


def _agree(err, rank: int, world: int, group, what: str) -> None:
    """Collective: every rank reports whether its part of `what` worked; if any rank failed, ALL ranks raise the same error
    (a rank that raised alone would leave its peers waiting in the next collective)."""
    if world <= 1:
        if err is not None:
            raise err
        return
    errs: list = [None] * world
    dist.all_gather_object(errs, None if err is None else f"{type(err).__name__}: {err}", group=group)
    bad = [f"rank {r}: {e}" for r, e in enumerate(errs) if e]
    if bad:
        raise RuntimeError(f"{what} failed ({'; '.join(bad)})") from err


def exchange_fds(fd: int, rank: int, world: int, group, tag: str) -> list[int]:
    """Every rank contributes one file descriptor; returns the list of all ranks' descriptors as valid in THIS process.
    Descriptors cross process boundaries as SCM_RIGHTS ancillary data over abstract AF_UNIX sockets (names agreed through
    torch.distributed); -1 entries (a rank with nothing to share) stay -1.  A failure on any rank (peer gone, descriptor
    refused, DLCFN_FD_EXCHANGE_TIMEOUT seconds without a sender - default 120) raises on every rank."""
    import socket
    import uuid
    me = f"\0dlcfn-{tag}-{os.getpid()}-{uuid.uuid4().hex[:8]}"
    srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    srv.bind(me)
    srv.listen(world + 1)
    names: list = [None] * world
    dist.all_gather_object(names, (me, fd >= 0), group=group)
    out = [-1] * world
    out[rank] = fd
    err = None
    try:
        if fd >= 0:
            for p, (name, _) in enumerate(names):
                if p == rank:
                    continue
                with socket.socket(socket.AF_UNIX, socket.SOCK_STREAM) as c:     # closed on failure too: the peer sees EOF
                    c.connect(name)                 # queued in the listener's backlog until it accepts
                    socket.send_fds(c, [rank.to_bytes(4, "little")], [fd])
        expect = sum(1 for p, (_, has) in enumerate(names) if has and p != rank)
        tmo = float(os.environ.get("DLCFN_FD_EXCHANGE_TIMEOUT", "120"))
        srv.settimeout(tmo)
        for _ in range(expect):
            conn, _ = srv.accept()
            with conn:
                conn.settimeout(tmo)
                msg, fds, _, _ = socket.recv_fds(conn, 4, 1)
            if len(fds) != 1 or len(msg) != 4:
                raise RuntimeError("file-descriptor exchange: malformed message")
            out[int.from_bytes(msg, "little")] = fds[0]
    except Exception as e:                           # reported to everybody below
        err = e
    finally:
        srv.close()
    if err is not None:
        for p, f in enumerate(out):
            if p != rank and f >= 0:
                os.close(f)
    _agree(err, rank, world, group, f"file-descriptor exchange '{tag}'")
    return out


class SymmetricHeap:
    """One arena per rank, mapped by every peer.  Bump allocation; offsets are identical on all ranks
    as long as all ranks perform the same sequence of `alloc` calls (they do: the layout is derived
    from the model, which is replicated)."""

    CH_ALLREDUCE = 0
    CH_PS = 1
    CH_BCAST = 2
    CH_BARRIER = 3
    CH_ASYNC = 4  # uses 4 and 5
    ONESHOT_MAX = 16 << 10   # bytes: csrc/cuda/symm.cuh:kOneShotMax / kLLMax

    def __init__(self, heap_bytes: int, group: dist.ProcessGroup | None = None, device: torch.device | None = None):
        if not torch.cuda.is_available():
            raise native.NativeUnavailable("SymmetricHeap needs a CUDA device")
        self.group = group
        if dist.is_available() and dist.is_initialized():
            self.rank = dist.get_rank(group)
            self.world = dist.get_world_size(group)
        else:
            self.rank, self.world = 0, 1
        if self.world > 8:
            raise ValueError("one NVSwitch domain: at most 8 ranks")
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.heap_bytes = int(heap_bytes)
        L = native.lib()
        ctx = c_void_p()
        self.backend = "ipc"
        self._symm_tensor = None
        want = os.environ.get("DLCFN_SYMM_BACKEND", "auto")
        self._vmm = None
        with torch.cuda.device(self.device):
            if self.world > 1 and want in ("auto", "vmm"):
                try:
                    ctx = self._create_vmm_backend(L)
                    self.backend = "vmm"
                except Exception as e:          # no exportable VMM handles here (or fd passing refused): next backend
                    self._release_vmm()
                    if want == "vmm":
                        raise
                    self._why_not_vmm = repr(e)
                    ctx = c_void_p()
            if self.backend == "ipc" and self.world > 1 and want in ("auto", "torch"):
                try:
                    ctx = self._create_torch_backend(L)
                    self.backend = "torch-symm"
                except Exception as e:          # no symmetric-memory support here: CUDA-IPC arena instead
                    if want == "torch":
                        raise
                    self._why_not_torch = repr(e)
                    ctx = c_void_p()
            if self.backend == "ipc":
                native.check(L.dlcfn_symm_create(self.heap_bytes, self.rank, self.world, ctypes.byref(ctx)), "symm_create")
                if self.world > 1:
                    buf = ctypes.create_string_buffer(64)
                    native.check(L.dlcfn_symm_get_handle(ctx, buf), "symm_get_handle")
                    handles: list = [None] * self.world
                    dist.all_gather_object(handles, bytes(buf.raw), group=group)
                    for p, h in enumerate(handles):
                        if p != self.rank:
                            native.check(L.dlcfn_symm_open_peer(ctx, p, ctypes.create_string_buffer(h, 64)), "symm_open_peer")
            self._ctx = ctx
            native.check(L.dlcfn_symm_finalize(ctx), "symm_finalize")
        self.has_multicast = bool(L.dlcfn_symm_has_multicast(ctx))
        # auto: in-switch reduction (NVLS) pays off when several peers share the port (N > 2); at N = 2 the peer ld/st
        # kernel moves less data per link (measured 656 vs 400 GB/s at 256 MiB)
        algo = os.environ.get("DLCFN_ALLREDUCE_ALGO", "auto")
        self.allreduce_algo = 1 if (self.has_multicast and (algo == "nvls" or (algo == "auto" and self.world > 2))) else 0
        self._base = L.dlcfn_symm_heap_ptr(ctx, self.rank)
        self._bytes = tensor_from_ptr(self._base, self.heap_bytes, self.device)
        self._top = 0
        if self.world > 1:
            dist.barrier(group=group)

    # ------------------------------------------------------------------ own VMM / multicast plumbing
    def _exchange_fds(self, fd: int, tag: str) -> list[int]:
        return exchange_fds(fd, self.rank, self.world, self.group, tag)

    def _create_vmm_backend(self, L):
        """Arena from the driver's virtual-memory API (csrc/cuda/symm_vmm.cu).  Every phase ends in an agreement: a rank whose
        cuMem* / cuMulticast* call failed makes ALL ranks fall through to the next backend together."""
        from ctypes import c_ulonglong
        dev = self.device.index if self.device.index is not None else torch.cuda.current_device()
        sup = int(L.dlcfn_vmm_supported(dev))
        flags = [None] * self.world
        dist.all_gather_object(flags, sup, group=self.group)
        if not all(f & 1 for f in flags):
            raise RuntimeError(f"exportable VMM allocations are not supported on every rank (flags {flags})")
        want_mc = all(f & 2 for f in flags) and os.environ.get("DLCFN_SYMM_MULTICAST", "1") != "0"
        agree = lambda err, what: _agree(err, self.rank, self.world, self.group, what)    # noqa: E731

        def phase(what, fn):
            err = None
            try:
                fn()
            except Exception as e:
                err = e
            agree(err, what)

        st = {"total": 0, "gran": 0, "local": None, "peers": {}, "mc": None, "fds": []}
        self._vmm = st
        ptr, hdl, fd = c_void_p(), c_ulonglong(), c_int(-1)
        ctx = c_void_p()

        def alloc():
            gran = int(L.dlcfn_vmm_granularity(dev, self.world, int(want_mc)))
            if gran <= 0:
                raise RuntimeError("cuMemGetAllocationGranularity failed")
            total = ((1 << 20) + (self.heap_bytes + 255) // 256 * 256 + gran - 1) // gran * gran
            native.check(L.dlcfn_vmm_alloc(total, dev, gran, ctypes.byref(ptr), ctypes.byref(hdl), ctypes.byref(fd)), "vmm_alloc")
            st.update(total=total, gran=gran, local=(ptr.value, hdl.value), fds=[fd.value])
        phase("VMM arena allocation", alloc)
        total, gran = st["total"], st["gran"]
        fds = self._exchange_fds(fd.value, "arena")

        def import_peers():
            native.check(L.dlcfn_symm_create_external(ptr.value, total, self.rank, self.world, ctypes.byref(ctx)), "symm_create_external")
            for p, pfd in enumerate(fds):
                if p == self.rank:
                    continue
                pp, ph = c_void_p(), c_ulonglong()
                try:
                    native.check(L.dlcfn_vmm_import(pfd, total, dev, gran, ctypes.byref(pp), ctypes.byref(ph)), "vmm_import")
                finally:
                    os.close(pfd)
                st["peers"][p] = (pp.value, ph.value)
                native.check(L.dlcfn_symm_set_peer_ptr(ctx, p, pp.value), "symm_set_peer_ptr")
        phase("mapping the peers' arenas", import_peers)
        if want_mc:
            mc, mfd = c_ulonglong(), c_int(-1)

            def mc_create():
                if self.rank == 0:
                    native.check(L.dlcfn_vmm_mc_create(total, self.world, ctypes.byref(mc), ctypes.byref(mfd)), "vmm_mc_create")
            phase("multicast object creation", mc_create)
            mfds = self._exchange_fds(mfd.value if self.rank == 0 else -1, "mc")

            def mc_join():
                if self.rank != 0:
                    try:
                        native.check(L.dlcfn_vmm_mc_import(mfds[0], ctypes.byref(mc)), "vmm_mc_import")
                    finally:
                        os.close(mfds[0])
                else:
                    os.close(mfd.value)
                native.check(L.dlcfn_vmm_mc_add_device(mc.value, dev), "vmm_mc_add_device")
            phase("joining the multicast object", mc_join)      # also the rendezvous: every device has joined before anyone binds

            def mc_bind():
                mptr = c_void_p()
                native.check(L.dlcfn_vmm_mc_bind_map(mc.value, hdl.value, total, dev, gran, ctypes.byref(mptr)), "vmm_mc_bind_map")
                st["mc"] = (mptr.value, mc.value)
                native.check(L.dlcfn_symm_set_multicast(ctx, mptr.value), "symm_set_multicast")
            phase("binding the arena to the multicast object", mc_bind)
        os.close(fd.value)
        st["fds"] = []
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)
        return ctx

    # ------------------------------------------------------------------ shared staging (the EFS analogue)
    def stage_put(self, name: str, tensor: torch.Tensor) -> tuple[int, int]:
        """Publish `tensor` (this rank's copy) under `name` in the heap: allocates the same region on every rank (all ranks
        must call with the same name / size, like every allocation) and copies the local data in.  Peers read it with
        `stage_get(name, src_rank)` straight out of this rank's HBM - initial weights from rank 0, shared data shards ...
        (reference: files on the EFS mount every instance shares, deeplearning.template:453-474)."""
        reg = self.__dict__.setdefault("_staged", {})
        nb = tensor.numel() * tensor.element_size()
        if name not in reg:
            reg[name] = (self.alloc(nb), nb, tensor.dtype, tuple(tensor.shape))
        off, cap, _, _ = reg[name]
        if nb != cap:
            raise ValueError(f"staged object {name!r} was registered with {cap} bytes, got {nb}")
        self._bytes[off: off + nb].copy_(tensor.detach().contiguous().reshape(-1).view(torch.uint8))
        return off, nb

    def stage_get(self, name: str, src_rank: int, stream: torch.cuda.Stream | None = None) -> torch.Tensor:
        """A copy of what `src_rank` staged under `name`, read over NVLink from its arena (callers order it after the
        producer's stage_put with a barrier, e.g. `heap.barrier()`)."""
        off, nb, dtype, shape = self._staged[name]
        src = tensor_from_ptr(self.peer_heap_ptr(src_rank) + off, nb, self.device)
        out = torch.empty(nb, dtype=torch.uint8, device=self.device)
        with torch.cuda.stream(stream) if stream is not None else torch.cuda.stream(torch.cuda.current_stream()):
            out.copy_(src, non_blocking=True)
        return out.view(dtype).view(shape)

    def _create_torch_backend(self, L):
        """Arena from torch symmetric memory: same layout, every peer already mapped, plus (when the fabric
        supports it) an NVLS multicast mapping that the multimem all-reduce kernel uses."""
        import torch.distributed._symmetric_memory as symm_mem
        total = (1 << 20) + (self.heap_bytes + 255) // 256 * 256
        group = self.group if self.group is not None else dist.group.WORLD
        t = symm_mem.empty(total, dtype=torch.uint8, device=self.device)
        try:
            hdl = symm_mem.rendezvous(t, group=group)
        except TypeError:
            hdl = symm_mem.rendezvous(t, group.group_name)
        ptrs = list(hdl.buffer_ptrs)
        ctx = c_void_p()
        native.check(L.dlcfn_symm_create_external(ptrs[self.rank], total, self.rank, self.world, ctypes.byref(ctx)),
                     "symm_create_external")
        for p, ptr in enumerate(ptrs):
            if p != self.rank:
                native.check(L.dlcfn_symm_set_peer_ptr(ctx, p, ptr), "symm_set_peer_ptr")
        mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
        if mc:
            native.check(L.dlcfn_symm_set_multicast(ctx, mc), "symm_set_multicast")
        self._symm_tensor, self._symm_handle = t, hdl
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)
        return ctx

    # ------------------------------------------------------------------ allocation
    def alloc(self, nbytes: int, align: int = 256) -> int:
        off = (self._top + align - 1) // align * align
        if off + nbytes > self.heap_bytes:
            raise MemoryError(f"symmetric heap exhausted: need {off + nbytes}, have {self.heap_bytes}")
        self._top = off + (nbytes + 15) // 16 * 16
        return off

    def view(self, off: int, numel: int, dtype: torch.dtype) -> torch.Tensor:
        nb = numel * torch.empty((), dtype=dtype).element_size()
        if off < 0 or numel < 0 or off + nb > self.heap_bytes:
            raise IndexError(f"symmetric heap view [{off}, {off + nb}) outside the heap of {self.heap_bytes} bytes")
        return self._bytes[off: off + nb].view(dtype)

    def alloc_tensor(self, numel: int, dtype: torch.dtype) -> tuple[int, torch.Tensor]:
        esz = torch.empty((), dtype=dtype).element_size()
        off = self.alloc(numel * esz)
        t = self.view(off, numel, dtype)
        return off, t

    def peer_heap_ptr(self, peer: int) -> int:
        return native.lib().dlcfn_symm_heap_ptr(self._ctx, peer)

    # ------------------------------------------------------------------ collectives
    def barrier(self, stream: torch.cuda.Stream | None = None) -> None:
        native.check(native.lib().dlcfn_barrier(self._ctx, self.CH_BARRIER, native.stream_ptr(stream)), "barrier")
        native._count()

    def allreduce_(self, off: int, count: int, dtype: torch.dtype, scale: float = 1.0, nblocks: int = 0,
                   stream: torch.cuda.Stream | None = None, channel: int | None = None, dst_off: int | None = None,
                   algo: int | None = None) -> None:
        """Fused bucket scale + all-reduce (reference C11: Horovod fusion buffer + ncclAllReduce).  In place unless
        dst_off is given.  algo 0 = peer ld/st two-shot, 1 = NVLS multimem (default when a multicast mapping exists), 2 = one-shot
        (<= 16 KiB, one barrier; default at that size), 3 = flag-in-data "LL" (<= 16 KiB, no barrier; opt-in)."""
        ch = self.CH_ALLREDUCE if channel is None else channel
        a = self.allreduce_algo if algo is None else algo
        nbytes = count * (2 if dtype == torch.bfloat16 else 4)
        if algo is None and a == 1 and nbytes < (2 << 20):
            a = 0      # below ~2 MiB the switch round trip costs more than it saves (profiles/scale8/README.md)
        if algo is None and nbytes <= self.ONESHOT_MAX and self.world > 1 and os.environ.get("DLCFN_ALLREDUCE_ONESHOT", "1") != "0":
            # latency-bound: push + ONE flag barrier + local reduce (kallreduce_oneshot, validated on 2 and 8 B200s).
            # DLCFN_ALLREDUCE_ONESHOT=ll selects the flag-in-data form without any barrier (kallreduce_ll) - written after the
            # round's GPU budget was spent, covered by tests/mp/collectives_worker.py under DLCFN_EXPERIMENTAL=1, not yet run
            a = 3 if os.environ.get("DLCFN_ALLREDUCE_ONESHOT") == "ll" else 2
        native.check(native.lib().dlcfn_allreduce_ex(self._ctx, off, off if dst_off is None else dst_off, count,
                                                     native.dtype_code(dtype), float(scale), ch, nblocks, a,
                                                     native.stream_ptr(stream)), "allreduce")
        native._count()

    def broadcast_(self, off: int, nbytes: int, root: int, nblocks: int = 32,
                   stream: torch.cuda.Stream | None = None) -> None:
        native.check(native.lib().dlcfn_broadcast(self._ctx, off, nbytes, root, self.CH_BCAST, nblocks,
                                                  native.stream_ptr(stream)), "broadcast")
        native._count()

    def ps_sync_step(self, grad_off: int, weight_off: int, total: int, shard_lo: int, shard_hi: int,
                     worker_mask: int, master: torch.Tensor | None, s1: torch.Tensor | None,
                     s2: torch.Tensor | None, dtype: torch.dtype, kind: str, *, lr: float, wd: float = 0.0,
                     momentum: float = 0.0, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8,
                     rescale: float = 1.0, clip: float = -1.0, nblocks: int = 0,
                     stream: torch.cuda.Stream | None = None, algo: int | None = None,
                     pull_mask: int | None = None) -> None:
        """push (+) server reduce (+) optimizer (+) pull in one kernel (reference C6/C7/C2).  algo 1 routes the push
        through multimem.ld_reduce and the pull through multimem.st (ranks that are not workers must keep a zero
        gradient buffer); nblocks 0 = size-based grid."""
        a = self.allreduce_algo if algo is None else algo
        if a == 1 and not self.has_multicast:
            a = 0
        native.check(native.lib().dlcfn_ps_sync_step_ex(
            self._ctx, grad_off, weight_off, total, shard_lo, shard_hi, worker_mask, int(pull_mask or 0), native._ptr(master),
            native._ptr(s1), native._ptr(s2), native.dtype_code(dtype), native.OPT_KINDS[kind], lr, wd, momentum,
            beta1, beta2, eps, rescale, clip, self.CH_PS, int(nblocks), int(a), native.stream_ptr(stream)), "ps_sync_step")
        native._count()

    def ps_async_worker(self, server_mask: int, nblocks: int = 16, stream: torch.cuda.Stream | None = None) -> None:
        native.check(native.lib().dlcfn_ps_async_worker(self._ctx, server_mask, self.CH_ASYNC, nblocks,
                                                        native.stream_ptr(stream)), "ps_async_worker")
        native._count()

    def ps_async_server(self, grad_off: int, weight_off: int, shard_lo: int, shard_hi: int, worker_mask: int,
                        master: torch.Tensor, s1, s2, dtype: torch.dtype, kind: str, *, lr: float, wd: float = 0.0,
                        momentum: float = 0.0, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8,
                        rescale: float = 1.0, clip: float = -1.0, nblocks: int = 16,
                        stream: torch.cuda.Stream | None = None) -> None:
        native.check(native.lib().dlcfn_ps_async_server(
            self._ctx, grad_off, weight_off, shard_lo, shard_hi, worker_mask, native._ptr(master), native._ptr(s1),
            native._ptr(s2), native.dtype_code(dtype), native.OPT_KINDS[kind], lr, wd, momentum, beta1, beta2, eps,
            rescale, clip, self.CH_ASYNC, nblocks, native.stream_ptr(stream)), "ps_async_server")
        native._count()

    def check_error(self, stream: torch.cuda.Stream | None = None) -> None:
        e = c_int(0)
        native.check(native.lib().dlcfn_symm_check_error(self._ctx, native.stream_ptr(stream), ctypes.byref(e)),
                     "symm_check_error")
        if e.value:
            raise RuntimeError("a peer did not reach a device-side flag barrier within the timeout "
                               "(dead or stalled rank) - reference analogue: get_num_dead_node > 0")

    def set_debug(self, timeout_ms: int = 0, stress_ns: int = 0, drop_peer: int = -1) -> None:
        """Protocol test hooks: flag-wait timeout, per-CTA random delay before every signal (schedule fuzzing), and fault
        injection - stop signalling `drop_peer`, which must then surface a timeout through check_error / poll_error."""
        native.check(native.lib().dlcfn_symm_set_debug(self._ctx, int(timeout_ms), int(stress_ns), int(drop_peer)), "symm_set_debug")

    def poll_error(self, stream: torch.cuda.Stream | None = None) -> None:
        """Non-blocking dead-peer check: raises if the error word copied by the PREVIOUS call was set, then queues the next
        copy (pinned host word + event) behind the work already on `stream`.  The host never waits for the GPU here."""
        st = getattr(self, "_poll", None)
        if st is None:
            st = self._poll = {"host": torch.zeros(1, dtype=torch.int32).pin_memory(), "event": torch.cuda.Event(), "armed": False}
        if st["armed"]:
            st["event"].synchronize()            # recorded a polling interval ago: long complete
            if int(st["host"][0]) != 0:
                st["armed"] = False
                self.check_error(stream)         # clears the device word, raises
                raise RuntimeError("a peer did not reach a device-side flag barrier within the timeout (dead or stalled rank)")
        native.check(native.lib().dlcfn_symm_poll_error_async(self._ctx, st["host"].data_ptr(), native.stream_ptr(stream)),
                     "symm_poll_error_async")
        st["event"].record(stream if stream is not None else torch.cuda.current_stream())
        st["armed"] = True

    def _release_vmm(self) -> None:
        """Unmap the multicast view, the peers and the local arena and drop their handles (also after a half-built backend)."""
        st = getattr(self, "_vmm", None)
        self._vmm = None
        if not st:
            return
        L = native.lib()
        if st["mc"]:
            L.dlcfn_vmm_unmap_release(st["mc"][0], st["mc"][1], st["total"])
        for pp, ph in st["peers"].values():
            L.dlcfn_vmm_unmap_release(pp, ph, st["total"])
        if st["local"]:
            L.dlcfn_vmm_unmap_release(st["local"][0], st["local"][1], st["total"])
        for f in st.get("fds", []):
            try:
                os.close(f)
            except OSError:
                pass

    def close(self) -> None:
        if getattr(self, "_ctx", None):
            torch.cuda.synchronize(self.device)
            self._bytes = None
            native.lib().dlcfn_symm_destroy(self._ctx)
            self._release_vmm()
            self._ctx = None
            self._symm_tensor = None
