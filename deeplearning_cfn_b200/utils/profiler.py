"""Profiler: Chrome-trace JSON of named host ranges and device (CUDA-event) ranges, NVTX mirroring,
aggregate statistics, per-rank file names, env autostart.

Reference: the MXNet profiler (MX/src/profiler/profiler.cc:166-250 emits `"traceEvents"` JSON; the engine
wraps every op in a ProfileOperator; kvstore ops are named `KVStoreDistDefaultPush`, `KVStoreReduce`,
`KVStoreBCast`..., kvstore_dist.h:265,409,433; aggregate tables aggregate_stats.cc; python API
`profiler.set_config/set_state/dump/dumps`, Domains/Tasks/Counters/Markers profiler.py:197-450; env
`MXNET_PROFILER_AUTOSTART`; server-side profiling prefixes the file with `rank<r>_`, kvstore_dist_server.h:276-323).
Device ranges are measured with CUDA events on the launching stream (the timing rule of this project) and
resolved lazily at dump time so recording never synchronises.
"""
from __future__ import annotations

import contextlib
import json
import os
import threading
import time
from collections import defaultdict

import torch

_cfg = {"filename": "profile.json", "profile_all": True, "aggregate_stats": True, "running": False,
        "profile_symbolic": True, "profile_imperative": True, "profile_api": True, "profile_memory": True}
_lock = threading.Lock()
_events: list = []
_device_pending: list = []
_counters: dict = {}
_t0 = time.perf_counter()


def _rank() -> int:
    return int(os.environ.get("RANK", os.environ.get("DMLC_TASK_ID", "0")) or 0)


def set_config(filename: str = "profile.json", profile_all: bool = True, aggregate_stats: bool = True,
               profile_process: str = "worker", profile_symbolic: bool = True, profile_imperative: bool = True,
               profile_memory: bool = True, profile_api: bool = True, **_ignored) -> None:
    """The reference's switches (profiler.py:33-68): `profile_all` records every category; otherwise operators are
    recorded when `profile_symbolic` / `profile_imperative` is set, kvstore / API ranges with `profile_api`, storage
    counters with `profile_memory`.  User-defined Domains / Tasks / Markers are always recorded."""
    _cfg.update(filename=filename, profile_all=profile_all, aggregate_stats=aggregate_stats, profile_symbolic=profile_symbolic,
                profile_imperative=profile_imperative, profile_memory=profile_memory, profile_api=profile_api)
    if profile_process == "server":
        # the reference forwards the parameters to the servers (kSetProfilerParams); the file gets a rank prefix there
        _server_command("config:" + filename)


def _server_command(body: str) -> None:
    """kSetProfilerParams to every server (worker side of kvstore_dist.h:96-102); needs a joined dist kvstore."""
    from . import rt
    rt.lib().dlcfn_kv_send_command(rt.CMD_SET_PROFILER, body.encode(), len(body))


def set_state(state: str = "stop", profile_process: str = "worker") -> None:
    if profile_process == "server":
        _server_command("state:" + ("run" if state == "run" else "stop"))
        return
    _cfg["running"] = state == "run"


def is_running() -> bool:
    return _cfg["running"]


def category_enabled(category: str) -> bool:
    if _cfg["profile_all"]:
        return True
    c = category.lower()
    if c in ("operator", "train", "symbolic", "imperative"):
        return _cfg["profile_symbolic"] or _cfg["profile_imperative"]
    if c in ("kvstore", "api", "collective", "io"):
        return _cfg["profile_api"]
    if c in ("memory", "storage"):
        return _cfg["profile_memory"]
    return True             # user domains


def _now_us() -> float:
    return (time.perf_counter() - _t0) * 1e6


@contextlib.contextmanager
def scope(name: str, category: str = "operator", device: bool = False, stream=None):
    """Record one range.  device=True additionally brackets it with CUDA events on `stream`."""
    if not _cfg["running"] or not category_enabled(category):
        yield
        return
    use_nvtx = torch.cuda.is_available()
    if use_nvtx:
        torch.cuda.nvtx.range_push(name)
    ev = None
    if device and torch.cuda.is_available():
        s = stream or torch.cuda.current_stream()
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record(s)
    t = _now_us()
    try:
        yield
    finally:
        dur = _now_us() - t
        if ev is not None:
            ev[1].record(stream or torch.cuda.current_stream())
            with _lock:
                _device_pending.append((name, category, t, ev))
        with _lock:
            _events.append({"name": name, "cat": category, "ph": "X", "ts": t, "dur": dur, "pid": _rank(),
                            "tid": threading.get_ident() % 100000})
        if use_nvtx:
            torch.cuda.nvtx.range_pop()


def marker(name: str, scope_: str = "process") -> None:
    if _cfg["running"]:
        with _lock:
            _events.append({"name": name, "cat": "marker", "ph": "i", "ts": _now_us(), "pid": _rank(), "tid": 0,
                            "s": "p" if scope_ == "process" else "g"})


class Counter:
    def __init__(self, name: str, value: int = 0):
        self.name, self.value = name, value
        self._emit()

    def _emit(self):
        if _cfg["running"]:
            with _lock:
                _events.append({"name": self.name, "cat": "counter", "ph": "C", "ts": _now_us(), "pid": _rank(),
                                "args": {self.name: self.value}})

    def set_value(self, v):
        self.value = v
        self._emit()

    def increment(self, d=1):
        self.set_value(self.value + d)

    def decrement(self, d=1):
        self.set_value(self.value - d)


def _resolve_device() -> None:
    with _lock:
        pend, _device_pending[:] = list(_device_pending), []
    for name, cat, t, (a, b) in pend:
        b.synchronize()
        with _lock:
            _events.append({"name": name + " [device]", "cat": cat, "ph": "X", "ts": t, "dur": a.elapsed_time(b) * 1e3,
                            "pid": _rank(), "tid": 999999})


def dumps(reset: bool = False) -> str:
    """Aggregate table (count / total / min / max / avg per name), like profiler.dumps()."""
    _resolve_device()
    agg = defaultdict(lambda: [0, 0.0, float("inf"), 0.0])
    with _lock:
        for e in _events:
            if e.get("ph") == "X":
                a = agg[e["name"]]
                a[0] += 1
                a[1] += e["dur"]
                a[2] = min(a[2], e["dur"])
                a[3] = max(a[3], e["dur"])
        if reset:
            _events.clear()
    lines = [f"{'Name':48s} {'Count':>8s} {'Total(us)':>14s} {'Min(us)':>12s} {'Max(us)':>12s} {'Avg(us)':>12s}"]
    for name, (c, tot, mn, mx) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{name[:48]:48s} {c:8d} {tot:14.1f} {mn:12.1f} {mx:12.1f} {tot / c:12.1f}")
    return "\n".join(lines)


def dump(finished: bool = True, filename: str | None = None, profile_process: str = "worker") -> str:
    """Write the Chrome-trace file; non-zero ranks prefix the name with `rank<r>_`.  profile_process='server' asks
    every parameter server to write its own `rank<r>_<file>` (request timeline: push / pull per key)."""
    if profile_process == "server":
        _server_command("dump")
        return ""
    _resolve_device()
    fn = filename or _cfg["filename"]
    r = _rank()
    if r != 0 or os.environ.get("DMLC_ROLE") == "server":
        d, b = os.path.split(fn)
        fn = os.path.join(d, f"rank{r}_{b}")
    with _lock:
        data = {"traceEvents": list(_events), "displayTimeUnit": "ms"}
    with open(fn, "w") as f:
        json.dump(data, f)
    if finished:
        _cfg["running"] = False
    return fn


if os.environ.get("MXNET_PROFILER_AUTOSTART", os.environ.get("DLCFN_PROFILER_AUTOSTART", "0")) == "1":
    # MXNET_PROFILER_MODE (env_var.md:154-165): 0 = operators of symbolic execution only, 1 = everything
    _all = os.environ.get("MXNET_PROFILER_MODE", "1") != "0"
    set_config(filename=os.environ.get("DLCFN_PROFILER_FILE", "profile.json"), profile_all=_all, profile_symbolic=True,
               profile_imperative=_all, profile_api=_all, profile_memory=_all)
    set_state("run")


# ----------------------------------------------------------------------------- user-defined objects (profiler.py:197-450)
class Domain:
    """A named group of user ranges / counters / markers; shows up as the trace category."""

    def __init__(self, name: str):
        self.name = name

    def new_task(self, name: str) -> "Task":
        return Task(self, name)

    def new_frame(self, name: str) -> "Frame":
        return Frame(self, name)

    def new_counter(self, name: str, value: int | None = None) -> "Counter":
        c = Counter(f"{self.name}::{name}")
        if value is not None:
            c.set_value(value)
        return c

    def new_marker(self, name: str) -> "Marker":
        return Marker(self, name)

    def __str__(self):
        return self.name


class _Range:
    _kind = "range"

    def __init__(self, domain: Domain | None, name: str):
        self.domain, self.name = domain, name
        self._t = None

    def start(self) -> None:
        self._t = _now_us()
        if torch.cuda.is_available() and _cfg["running"]:
            torch.cuda.nvtx.range_push(self.name)

    def stop(self) -> None:
        if self._t is None:
            return
        if torch.cuda.is_available() and _cfg["running"]:
            torch.cuda.nvtx.range_pop()
        if _cfg["running"]:
            with _lock:
                _events.append({"name": self.name, "cat": f"{self.domain or 'user'}:{self._kind}", "ph": "X", "pid": _rank(),
                                "tid": 0, "ts": self._t, "dur": _now_us() - self._t})
        self._t = None

    def __enter__(self):
        self.start()
        return self

    def __exit__(self, *exc):
        self.stop()
        return False

    def __str__(self):
        return self.name


class Task(_Range):
    """A logical unit of work inside a Domain (may span threads)."""
    _kind = "task"


class Frame(_Range):
    """A repeated unit of work, e.g. one training iteration."""
    _kind = "frame"


class Event(_Range):
    """A range that belongs to no domain."""
    _kind = "event"

    def __init__(self, name: str):
        super().__init__(None, name)


class Marker:
    """An instant in time inside a Domain: `Marker(domain, 'epoch end').mark('process')`."""

    def __init__(self, domain: Domain, name: str):
        self.domain, self.name = domain, name

    def mark(self, scope: str = "process") -> None:
        if not _cfg["running"]:
            return
        with _lock:
            _events.append({"name": self.name, "cat": f"{self.domain}:marker", "ph": "i", "pid": _rank(), "tid": 0,
                            "ts": _now_us(), "s": {"global": "g", "process": "p", "thread": "t"}.get(scope, "p")})


def pause(profile_process: str = "worker") -> None:
    """Stop recording without closing the trace (profiler.py: pause / resume)."""
    if profile_process == "server":
        _server_command("state:stop")
    else:
        _cfg["running"] = False


def resume(profile_process: str = "worker") -> None:
    if profile_process == "server":
        _server_command("state:run")
    else:
        _cfg["running"] = True
