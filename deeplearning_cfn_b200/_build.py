"""In-tree native build: csrc/cuda -> _lib/libdlcfn_cuda.so (nvcc, sm_100a only),
csrc/rt -> _lib/libdlcfn_rt.so (g++, CPU runtime: transport / topology / engine).

The shared objects are git-ignored but travel with the tree to the GPU box, so nothing is
JIT-compiled there.  `python -m deeplearning_cfn_b200._build` rebuilds what is stale.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
LIB_DIR = PKG / "_lib"
CUDA_SRC = ROOT / "csrc" / "cuda"
RT_SRC = ROOT / "csrc" / "rt"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "--expt-relaxed-constexpr", "-cudart", "static",
]
CXX_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-fvisibility=hidden", "-pthread", "-Wall"]


def _nvcc() -> str | None:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    return cand if os.path.exists(cand) else None


def _stale(out: Path, deps: list[Path]) -> bool:
    if not out.exists():
        return True
    t = out.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps)


def _run(cmd: list[str]) -> None:
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError(f"native build failed: {cmd[0]} exited {r.returncode}")


def build_cuda(force: bool = False, verbose: bool = False) -> Path | None:
    """Compile every .cu under csrc/cuda for sm_100a into one shared object."""
    nvcc = _nvcc()
    out = LIB_DIR / "libdlcfn_cuda.so"
    if nvcc is None:
        return out if out.exists() else None
    srcs = sorted(CUDA_SRC.glob("*.cu"))
    hdrs = sorted(CUDA_SRC.glob("*.cuh"))
    if not srcs:     # a staged copy of the package without csrc/: nothing to build here
        return out if out.exists() else None
    LIB_DIR.mkdir(exist_ok=True)
    obj_dir = ROOT / "build" / "cuda"
    obj_dir.mkdir(parents=True, exist_ok=True)
    objs = []
    procs = []
    for s in srcs:
        o = obj_dir / (s.stem + ".o")
        objs.append(o)
        if force or _stale(o, [s] + hdrs):
            cmd = [nvcc, *NVCC_FLAGS, "-c", str(s), "-o", str(o)]
            if verbose:
                print(" ".join(cmd))
            procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        log, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(" ".join(cmd) + "\n" + log)
            raise RuntimeError("nvcc failed")
    if force or procs or _stale(out, objs):
        _run([nvcc, "-shared", *NVCC_FLAGS, "-o", str(out), *map(str, objs)])
    return out


def build_rt(force: bool = False, verbose: bool = False) -> Path | None:
    """Compile the CPU runtime (ps transport, topology planner, engine) into one shared object."""
    cxx = shutil.which("g++")
    out = LIB_DIR / "libdlcfn_rt.so"
    srcs = sorted(RT_SRC.glob("*.cpp"))
    if cxx is None or not srcs:
        return out if out.exists() else None
    hdrs = sorted(RT_SRC.glob("*.h"))
    LIB_DIR.mkdir(exist_ok=True)
    if force or _stale(out, srcs + hdrs):
        cmd = [cxx, *CXX_FLAGS, "-shared", "-o", str(out), *map(str, srcs)]
        if verbose:
            print(" ".join(cmd))
        _run(cmd)
    return out


def build_all(force: bool = False, verbose: bool = False) -> dict:
    return {"cuda": build_cuda(force, verbose), "rt": build_rt(force, verbose)}


if __name__ == "__main__":
    res = build_all(force="--force" in sys.argv, verbose=True)
    for k, v in res.items():
        print(f"{k}: {v}")
