#!/usr/bin/env python
"""Run one test many times to see whether it is flaky (reference: MX/tools/flakiness_checker.py - `-n` trials, `-s` seed,
every trial with its own seed exported to the test process).

    python tools/flakiness_checker.py tests/test_kvstore.py::test_dist_sync_closed_form -n 20
    python tools/flakiness_checker.py test_kvstore.test_dist_sync_closed_form -n 20 -s 7      # module.test form of the reference
    python tools/flakiness_checker.py tests/test_gpu_gemm.py -n 5 -m gpu                      # a whole file, GPU tests

Each trial is its own pytest process (own process group, killed as a group on `--timeout`) with DLCFN_TEST_SEED /
MXNET_TEST_SEED set to `seed + trial`; tests/conftest.py seeds torch / numpy / random from it.  Prints one line per failing
trial with the seed to reproduce it and exits non-zero if any trial failed.
"""
from __future__ import annotations

import argparse
import os
import random
import signal
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def resolve(spec: str) -> str:
    """`file.py::test`, `tests/file.py`, or the reference's `module.test_name` / `dir/module.test_name` form."""
    if spec.endswith(".py") or "::" in spec:
        return spec
    mod, _, test = spec.rpartition(".")
    if not mod:
        mod, test = spec, ""
    path = mod.replace(".", "/") + ".py"
    for cand in (path, os.path.join("tests", path)):
        if os.path.exists(os.path.join(ROOT, cand)):
            return cand + (f"::{test}" if test else "")
    raise SystemExit(f"cannot find a test file for {spec!r}")


def run_trial(target: str, seed: int, marker: str | None, timeout: float, verbose: bool) -> tuple[bool, float, str]:
    env = dict(os.environ, DLCFN_TEST_SEED=str(seed), MXNET_TEST_SEED=str(seed), PYTHONHASHSEED=str(seed % 4294967295))
    cmd = [sys.executable, "-m", "pytest", target, "-x", "-q", "-p", "no:cacheprovider"]
    if marker:
        cmd += ["-m", marker]
    t0 = time.time()
    p = subprocess.Popen(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, start_new_session=True)
    try:
        out, _ = p.communicate(timeout=timeout)
        ok = p.returncode == 0
    except subprocess.TimeoutExpired:
        for sig in (signal.SIGTERM, signal.SIGKILL):
            try:
                os.killpg(p.pid, sig)
            except ProcessLookupError:
                break
            try:
                p.communicate(timeout=10)
                break
            except subprocess.TimeoutExpired:
                continue
        out, ok = f"timed out after {timeout} s", False
    if verbose:
        print(out)
    return ok, time.time() - t0, out


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description="Check a test for flakiness")
    ap.add_argument("test", help="file.py[::test] or module.test_name")
    ap.add_argument("-n", "--num-trials", type=int, default=10)
    ap.add_argument("-s", "--seed", type=int, default=None, help="base seed (default: random); trial i uses seed + i")
    ap.add_argument("-m", "--marker", default=None, help="pytest -m expression (e.g. gpu)")
    ap.add_argument("--timeout", type=float, default=900.0, help="seconds per trial")
    ap.add_argument("-v", "--verbose", action="store_true")
    a = ap.parse_args(argv)
    target = resolve(a.test)
    base = a.seed if a.seed is not None else random.SystemRandom().randrange(1 << 31)
    fails = []
    for i in range(a.num_trials):
        ok, dt, out = run_trial(target, base + i, a.marker, a.timeout, a.verbose)
        print(f"trial {i + 1}/{a.num_trials} seed {base + i}: {'ok' if ok else 'FAILED'} ({dt:.1f} s)", flush=True)
        if not ok:
            fails.append(base + i)
            tail = [l for l in out.splitlines() if l.strip()][-6:]
            print("    " + "\n    ".join(tail))
    print(f"{a.num_trials - len(fails)}/{a.num_trials} trials passed" + (f"; failing seeds: {fails} (re-run with -n 1 -s <seed>)" if fails else ""))
    return 1 if fails else 0


if __name__ == "__main__":
    sys.exit(main())
