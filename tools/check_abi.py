#!/usr/bin/env python
"""Compare every ctypes `argtypes` declaration with the C signature it binds (csrc/cuda/*.cu `DLCFN_API`, csrc/rt/*.cpp
`RT_API` / `DLCFN_RT_API`): number of parameters and their class (pointer, 32-bit int, 64-bit int, float, double).
A wrong count or width in a ctypes table only shows up as garbage arguments at run time - for the CUDA library that means on
the GPU box.  Runs on the host: both libraries dlopen without a GPU.

    python tools/check_abi.py
"""
from __future__ import annotations

import ctypes
import glob
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

_PTRISH = ("cudaStream_t", "CUstream")


def c_kind(decl: str) -> str:
    d = decl.strip()
    d = re.sub(r"/\*.*?\*/", "", d).strip()
    if "*" in d or "&" in d or any(t in d for t in _PTRISH):
        return "p"
    toks = [t for t in re.split(r"\s+", d) if t not in ("const", "unsigned", "signed", "volatile")]
    ty = " ".join(toks[:-1]) if len(toks) > 1 else toks[0]
    if "double" in ty:
        return "d"
    if "float" in ty:
        return "f"
    if "long long" in d or re.search(r"\blong\b", d) or ty in ("size_t", "uint64_t", "int64_t", "uintptr_t", "ssize_t"):
        return "q"
    if ty in ("int", "", "uint32_t", "int32_t", "uint16_t", "int16_t", "uint8_t", "int8_t", "char", "bool", "short") or d.startswith("unsigned"):
        return "i"
    return "p"          # typedef'd function pointers (callbacks), handles


RETURNS: dict[str, str] = {}


def c_signatures() -> dict[str, list[str]]:
    sigs = {}
    pat = re.compile(r"(?:DLCFN_API|DLCFN_RT_API|RT_API)\s+([\w\s\*:]+?)\b(\w+)\s*\(([^)]*)\)\s*\{", re.S)
    for path in glob.glob(os.path.join(ROOT, "csrc", "cuda", "*.cu")) + glob.glob(os.path.join(ROOT, "csrc", "rt", "*.cpp")):
        src = open(path).read()
        src = re.sub(r"//[^\n]*", "", src)
        for m in pat.finditer(src):
            ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
            params = [] if args in ("", "void") else [a for a in args.split(",")]
            sigs[name] = [c_kind(a) for a in params]
            RETURNS[name] = "v" if ret == "void" else c_kind(ret + " _")
    return sigs


def py_kind(t) -> str:
    if t is None:
        return "?"
    if t in (ctypes.c_void_p, ctypes.c_char_p, ctypes.c_wchar_p) or issubclass(t, (ctypes._Pointer, ctypes.Array)) or \
            hasattr(t, "_type_") and isinstance(getattr(t, "_type_"), type):
        return "p"
    if hasattr(t, "_argtypes_"):        # CFUNCTYPE callbacks
        return "p"
    if t in (ctypes.c_float,):
        return "f"
    if t in (ctypes.c_double,):
        return "d"
    return "q" if ctypes.sizeof(t) == 8 else "i"


def declared_libraries():
    """Load both libraries and run every module-level declaration helper."""
    libs = []
    from deeplearning_cfn_b200.ops import batchnorm, layernorm, native, nn, pool, roi, sparse
    L = native.lib()
    for mod in (batchnorm, layernorm, nn, pool, roi, sparse):
        fn = getattr(mod, "_declare", None)
        if fn is not None:
            try:
                fn()
            except TypeError:
                fn(L)
    libs.append(("cuda", L))
    import torch
    from deeplearning_cfn_b200.utils import recordio, rt, storage
    R = rt.lib()
    storage._L()                                   # host pool / shm tables
    recordio._rtlib()                              # RecordIO reader / writer / image iterator tables
    rt.reduce_sum_cpu([torch.ones(4), torch.ones(4)])          # declares dlcfn_reduce_sum_cpu on first use
    try:
        storage.DevicePool()                       # device pool tables (the constructor declares before touching CUDA)
    except Exception:
        pass
    libs.append(("rt", R))
    return libs


def main() -> int:
    sigs = c_signatures()
    problems, checked = [], 0
    for tag, L in declared_libraries():
        for name, kinds in sorted(sigs.items()):
            try:
                fn = getattr(L, name)
            except AttributeError:
                continue
            rk = RETURNS.get(name, "i")
            if rk in ("p", "q", "f", "d"):          # ctypes defaults to a 32-bit int return: pointers / 64-bit values truncate
                rt_ = fn.restype
                got = "v" if rt_ is None else py_kind(rt_)
                if got != rk:
                    problems.append(f"{name}: returns '{rk}' in C, restype is '{got}' in ctypes")
            at = fn.argtypes
            if at is None:
                continue
            checked += 1
            pk = [py_kind(t) for t in at]
            if len(pk) != len(kinds):
                problems.append(f"{name}: C takes {len(kinds)} parameters, ctypes declares {len(pk)}")
                continue
            for i, (a, b) in enumerate(zip(kinds, pk)):
                if a != b:
                    problems.append(f"{name}: parameter {i} is '{a}' in C, '{b}' in ctypes   (C {''.join(kinds)} / py {''.join(pk)})")
                    break
    for p in problems:
        print(p)
    print(f"{len(sigs)} exported functions parsed, {checked} ctypes tables compared, {len(problems)} mismatches")
    return 1 if problems else 0


if __name__ == "__main__":
    sys.exit(main())
