"""Control flow of SymmetricHeap._create_vmm_backend (parallel/symm.py) on the host: the driver entry points of
csrc/cuda/symm_vmm.cu are replaced by stand-ins that hand out temporary files as "exportable handles", the ranks are gloo
processes.  Checks the success path (every peer imported from the right descriptor, multicast object created by rank 0 and
joined by all) and that a failure injected on ONE rank in any phase raises on EVERY rank and releases what was mapped."""
import ctypes
import os
import sys
import tempfile

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from deeplearning_cfn_b200.parallel import symm  # noqa: E402

rank = int(os.environ.get("RANK", os.environ.get("OMPI_COMM_WORLD_RANK", "0")))
world = int(os.environ.get("WORLD_SIZE", os.environ.get("OMPI_COMM_WORLD_SIZE", "1")))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("MASTER_PORT", "29643")
os.environ["DLCFN_FD_EXCHANGE_TIMEOUT"] = "20"
dist.init_process_group("gloo", rank=rank, world_size=world)
torch.cuda.current_device = lambda: 0
torch.cuda.synchronize = lambda *a, **k: None


class FakeDriver:
    """Stand-in for the dlcfn_vmm_* / dlcfn_symm_* entry points used while the backend is built."""

    def __init__(self, fail_at=None):
        self.fail_at = fail_at                  # (function name, rank) that returns CUDA error 2
        self.files, self.released, self.imported, self.joined = [], [], {}, False

    def _rc(self, name):
        return 2 if self.fail_at == (name, rank) else 0

    def _file(self, text):
        f = tempfile.TemporaryFile()
        f.write(text.encode())
        f.flush()
        self.files.append(f)
        return os.dup(f.fileno())

    def dlcfn_vmm_supported(self, dev):
        return 3

    def dlcfn_vmm_granularity(self, dev, world_, mc):
        return 2 << 20

    def dlcfn_vmm_alloc(self, total, dev, gran, pptr, phdl, pfd):
        if self._rc("alloc"):
            return 2
        pptr._obj.value, phdl._obj.value, pfd._obj.value = 0x10000 * (rank + 1), 100 + rank, self._file(f"arena {rank} {total}")
        return 0

    def dlcfn_symm_create_external(self, ptr, total, rank_, world_, pctx):
        pctx._obj.value = 0xC0DE
        return 0

    def dlcfn_vmm_import(self, fd, total, dev, gran, pp, ph):
        if self._rc("import"):
            return 2
        who = int(os.pread(fd, 64, 0).decode().split()[1])
        pp._obj.value, ph._obj.value = 0x10000 * (who + 1) + 0x8, 200 + who
        self.imported[who] = pp._obj.value
        return 0

    def dlcfn_symm_set_peer_ptr(self, ctx, p, ptr):
        assert self.imported[p] == ptr
        return 0

    def dlcfn_vmm_mc_create(self, total, world_, pmc, pmfd):
        if self._rc("mc_create"):
            return 2
        pmc._obj.value, pmfd._obj.value = 555, self._file("multicast object")
        return 0

    def dlcfn_vmm_mc_import(self, fd, pmc):
        assert os.pread(fd, 64, 0) == b"multicast object"
        pmc._obj.value = 555
        return self._rc("mc_import")

    def dlcfn_vmm_mc_add_device(self, mc, dev):
        assert mc == 555
        self.joined = True
        return self._rc("mc_add_device")

    def dlcfn_vmm_mc_bind_map(self, mc, hdl, total, dev, gran, pmptr):
        if self._rc("mc_bind"):
            return 2
        pmptr._obj.value = 0xABC000
        return 0

    def dlcfn_symm_set_multicast(self, ctx, ptr):
        return 0

    def dlcfn_vmm_unmap_release(self, ptr, hdl, total):
        self.released.append(ptr)
        return 0


def make_heap():
    h = object.__new__(symm.SymmetricHeap)
    h.rank, h.world, h.group, h.device, h.heap_bytes, h._vmm = rank, world, None, torch.device("cpu"), 1 << 20, None
    return h


def open_fds():
    return len(os.listdir("/proc/self/fd"))


# ---- success
fd0 = open_fds()
drv = FakeDriver()
h = make_heap()
ctx = h._create_vmm_backend(drv)
assert ctx.value == 0xC0DE and sorted(drv.imported) == [p for p in range(world) if p != rank] and drv.joined
assert h._vmm["mc"] == (0xABC000, 555) and len(h._vmm["peers"]) == world - 1 and h._vmm["fds"] == []
for f in drv.files:
    f.close()
assert open_fds() == fd0, (open_fds(), fd0)          # every received / exported descriptor was closed again
symm.native.lib = lambda: drv                        # _release_vmm looks the library up itself
h._release_vmm()
assert len(drv.released) == world + 1 and h._vmm is None          # multicast view + peers + own arena

# ---- a failure on one rank in any phase raises on every rank, with the failing rank named, and releases the partial mappings
for name, who, phrase in (("alloc", 1, "VMM arena allocation"), ("import", world - 1, "mapping the peers' arenas"),
                          ("mc_create", 0, "multicast object creation"), ("mc_add_device", 1, "joining the multicast object"),
                          ("mc_bind", world - 1, "binding the arena to the multicast object")):
    drv = FakeDriver(fail_at=(name, who))
    symm.native.lib = lambda drv=drv: drv
    h = make_heap()
    try:
        h._create_vmm_backend(drv)
        raise SystemExit(f"rank {rank}: failure in {name} on rank {who} did not raise here")
    except RuntimeError as e:
        assert phrase in str(e) and f"rank {who}:" in str(e), (name, str(e))
    h._release_vmm()
    assert h._vmm is None
    for f in drv.files:
        f.close()
    dist.barrier()
print(f"rank {rank} VMM_FLOW_OK", flush=True)
dist.destroy_process_group()
