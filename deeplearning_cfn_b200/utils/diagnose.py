"""`diagnose` - one report about the box and the installation (reference MX/tools/diagnose.py: platform, python, package,
hardware, environment sections), extended with what matters for this runtime: the native libraries, the GPUs and their
peer / multicast capability, NCCL, and every DMLC_* / PS_* / MXNET_* / DLCFN_* / HOROVOD_* / NCCL_* variable that is set.

    python -m deeplearning_cfn_b200 diagnose [--json]
"""
from __future__ import annotations

import argparse
import json
import os
import platform
import shutil
import subprocess
import sys

_ENV_PREFIXES = ("DMLC_", "PS_", "MXNET_", "DLCFN_", "HOROVOD_", "NCCL_", "DEEPLEARNING_", "EFS_MOUNT", "CUDA_VISIBLE_DEVICES",
                 "MASTER_", "RANK", "WORLD_SIZE", "LOCAL_RANK", "LOCAL_WORLD_SIZE", "OMP_NUM_THREADS")


def _run(cmd: list[str], timeout: float = 10.0) -> str | None:
    try:
        return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout).stdout.strip()
    except Exception:
        return None


def collect() -> dict:
    info: dict = {"platform": {"system": platform.platform(), "machine": platform.machine(), "node": platform.node(),
                               "cpus": os.cpu_count()},
                  "python": {"version": sys.version.split()[0], "executable": sys.executable}}
    try:
        import torch
        t = {"version": torch.__version__, "cuda_build": torch.version.cuda, "cuda_available": torch.cuda.is_available(),
             "device_count": torch.cuda.device_count() if torch.cuda.is_available() else 0}
        try:
            t["nccl"] = ".".join(map(str, torch.cuda.nccl.version())) if torch.cuda.is_available() else None
        except Exception:
            t["nccl"] = None
        import torch.distributed as dist
        t["distributed_backends"] = [b for b in ("nccl", "gloo", "mpi") if getattr(dist, f"is_{b}_available")()]
        info["torch"] = t
        gpus = []
        if torch.cuda.is_available():
            for i in range(torch.cuda.device_count()):
                p = torch.cuda.get_device_properties(i)
                peers = [j for j in range(torch.cuda.device_count()) if j != i and torch.cuda.can_device_access_peer(i, j)]
                gpus.append({"index": i, "name": p.name, "sm": f"{p.major}.{p.minor}", "sms": p.multi_processor_count,
                             "memory_gib": round(p.total_memory / 2 ** 30, 1), "peer_access": peers})
        info["gpus"] = gpus
    except Exception as e:      # noqa: BLE001
        info["torch"] = {"error": repr(e)}
    pkg: dict = {}
    try:
        import deeplearning_cfn_b200 as pkg_mod
        pkg["path"] = os.path.dirname(pkg_mod.__file__)
        from ..ops import native
        pkg["cuda_library"] = {"path": str(native.lib_path()), "present": native.lib_path().exists()}
        try:
            native.lib()
            pkg["cuda_library"]["loads"] = True
        except Exception as e:      # noqa: BLE001
            pkg["cuda_library"]["loads"] = repr(e)
        from . import rt
        try:
            rt.lib()
            pkg["runtime_library"] = {"loads": True}
        except Exception as e:      # noqa: BLE001
            pkg["runtime_library"] = {"loads": repr(e)}
    except Exception as e:      # noqa: BLE001
        pkg["error"] = repr(e)
    info["package"] = pkg
    tools = {}
    for exe in ("nvcc", "nvidia-smi", "ncu", "mpirun", "ssh"):
        tools[exe] = shutil.which(exe)
    nv = _run(["nvcc", "--version"]) if tools["nvcc"] else None
    if nv:
        tools["nvcc_release"] = nv.splitlines()[-2] if len(nv.splitlines()) >= 2 else nv
    if tools["nvidia-smi"]:
        tools["driver"] = _run(["nvidia-smi", "--query-gpu=driver_version", "--format=csv,noheader"])
        topo = _run(["nvidia-smi", "topo", "-m"])
        if topo:
            info["nvlink_topology"] = topo.splitlines()[:12]
    info["tools"] = tools
    info["environment"] = {k: v for k, v in sorted(os.environ.items()) if k.startswith(_ENV_PREFIXES)}
    hostfile = os.environ.get("DEEPLEARNING_WORKERS_PATH")
    if hostfile and os.path.exists(hostfile):
        info["hostfile"] = open(hostfile).read().split()
    return info


def render(info: dict) -> str:
    out = []
    for section, body in info.items():
        out.append(f"----------{section}----------")
        if isinstance(body, dict):
            for k, v in body.items():
                out.append(f"{k:<22}: {v}")
        elif isinstance(body, list):
            out += [str(v) for v in body] or ["(none)"]
        else:
            out.append(str(body))
    return "\n".join(out)


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description="Diagnose the box and the installation")
    ap.add_argument("--json", action="store_true")
    a = ap.parse_args(argv)
    info = collect()
    print(json.dumps(info, indent=1) if a.json else render(info))
    return 0


if __name__ == "__main__":
    sys.exit(main())
