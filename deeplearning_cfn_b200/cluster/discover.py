"""Fabric probe: which GPUs exist, how they are linked, and the rank order derived from it.

In the reference, rank/hostfile discovery comes from SQS + AutoScaling + EC2 lookups
(dl_cfn_setup.py:209-280) and GPU-to-GPU structure is discovered separately inside MXNet by
`gpu_topology.h:133-196` (cudaDeviceGetP2PAttribute performance ranks -> link matrix -> Kernighan-Lin
trees).  On a B200 box both questions have one source of truth - the NVLink 5 / NVSwitch fabric - so
the probe feeds the provisioner (slot order, capacity) AND the communication planner
(`utils.rt.plan_trees`, the native K5 replacement; on NVSwitch it proves the matrix uniform and the
planner degenerates to "every peer is one hop").
"""
from __future__ import annotations

import os
import shutil
import subprocess
from dataclasses import dataclass, field


@dataclass
class Topology:
    num_gpus: int = 0
    names: list = field(default_factory=list)
    # link[i][j]: relative link weight between GPU i and j (0 = no P2P, 1 = PCIe-class, >=2 NVLink-class)
    link: list = field(default_factory=list)
    source: str = "none"
    multicast: bool | None = None          # NVLS multicast objects supported on every device (native probe)
    native_atomics: bool | None = None     # native atomics over every peer link

    def uniform(self) -> bool:
        """All off-diagonal links equal and non-zero (an NVSwitch domain)."""
        vals = {self.link[i][j] for i in range(self.num_gpus) for j in range(self.num_gpus) if i != j}
        return len(vals) <= 1 and (not vals or next(iter(vals)) > 0)

    def rank_order(self) -> list[int]:
        """GPU ids in rank order.  Uniform fabric: identity.  Otherwise greedy chain that always steps
        over the strongest remaining link, so neighbouring ranks share the best links (useful for rings)."""
        n = self.num_gpus
        if n == 0:
            return []
        if self.uniform():
            return list(range(n))
        order, left = [0], set(range(1, n))
        while left:
            cur = order[-1]
            nxt = max(sorted(left), key=lambda j: self.link[cur][j])
            order.append(nxt)
            left.discard(nxt)
        return order

    def summary(self) -> dict:
        return {"num_gpus": self.num_gpus, "names": self.names, "uniform": self.uniform() if self.num_gpus else None,
                "source": self.source, "multicast": self.multicast, "link": self.link}


def from_matrix(link: list, names: list | None = None, source: str = "synthetic") -> Topology:
    n = len(link)
    return Topology(n, names or [f"gpu{i}" for i in range(n)], [list(map(int, r)) for r in link], source)


def _probe_native() -> Topology | None:
    """The library's own probe (csrc/cuda/symm_vmm.cu:dlcfn_probe_fabric): cudaDeviceGetP2PAttribute access / performance
    rank / native atomics per pair (what gpu_topology.h:133-196 reads) plus CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED per
    device.  Link weight = 0 without peer access, otherwise 2 + (worst rank - this pair's rank): best links weigh most."""
    try:
        import ctypes

        import torch
        if not torch.cuda.is_available():
            return None
        from ..ops import native
        L = native.lib()
        n = torch.cuda.device_count()
        arr = ctypes.c_int * (n * n)
        access, perf, atom, mc = arr(), arr(), arr(), (ctypes.c_int * n)()
        got = L.dlcfn_probe_fabric(n, access, perf, atom, mc)
        if got != n:
            return None
        worst = max([perf[i * n + j] for i in range(n) for j in range(n) if i != j and access[i * n + j]] or [0])
        link = [[0 if i == j or not access[i * n + j] else 2 + (worst - perf[i * n + j]) for j in range(n)] for i in range(n)]
        topo = Topology(n, [torch.cuda.get_device_name(i) for i in range(n)], link, "dlcfn_probe_fabric")
        topo.multicast = bool(n) and all(mc[i] for i in range(n))
        topo.native_atomics = all(atom[i * n + j] for i in range(n) for j in range(n) if i != j) if n > 1 else None
        return topo
    except Exception:
        return None


def _probe_torch() -> Topology | None:
    try:
        import torch
        if not torch.cuda.is_available():
            return None
        n = torch.cuda.device_count()
        names = [torch.cuda.get_device_name(i) for i in range(n)]
        link = [[0] * n for _ in range(n)]
        for i in range(n):
            for j in range(n):
                if i != j and torch.cuda.can_device_access_peer(i, j):
                    link[i][j] = 2
        topo = Topology(n, names, link, "torch.cuda")
        _refine_with_nvidia_smi(topo)
        return topo
    except Exception:
        return None


def _refine_with_nvidia_smi(topo: Topology) -> None:
    """`nvidia-smi topo -m` distinguishes NV<k> (k NVLinks) from PIX/PHB/SYS (PCIe)."""
    exe = shutil.which("nvidia-smi")
    if exe is None:
        return
    try:
        out = subprocess.check_output([exe, "topo", "-m"], stderr=subprocess.DEVNULL, timeout=20).decode()
    except Exception:
        return
    rows = [l.split() for l in out.splitlines() if l.startswith("GPU")]
    for i, r in enumerate(rows[:topo.num_gpus]):
        for j, cell in enumerate(r[1:1 + topo.num_gpus]):
            if i == j:
                continue
            if cell.startswith("NV"):
                try:
                    topo.link[i][j] = 1 + int(cell[2:])
                except ValueError:
                    topo.link[i][j] = 2
            elif cell in ("PIX", "PXB", "PHB", "NODE", "SYS"):
                topo.link[i][j] = 1 if topo.link[i][j] else 0
    topo.source = "torch.cuda+nvidia-smi"


def probe_topology() -> Topology:
    """Probe order: explicit override (tests / CPU boxes) -> CUDA -> empty."""
    fake = os.environ.get("DLCFN_FAKE_GPUS")
    if fake is not None:
        n = int(fake)
        link = [[0 if i == j else 18 for j in range(n)] for i in range(n)]
        return Topology(n, [f"fake-b200-{i}" for i in range(n)], link, "DLCFN_FAKE_GPUS")
    t = _probe_native()
    if t is None:
        t = _probe_torch()
    elif t.uniform():
        _refine_with_nvidia_smi(t)      # NV<k> link counts for the report; the uniformity verdict is the native probe's
        if t.source != "dlcfn_probe_fabric":
            t.source = "dlcfn_probe_fabric+nvidia-smi"
    return t if t is not None else Topology()
