import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run_job(cmd, env=None, timeout=300, cwd=None):
    """subprocess.run(capture_output=True, text=True) for commands that start whole jobs (launcher -> scheduler, servers,
    workers): the command gets its own session and on a timeout the entire process group is killed, so a hung job cannot
    leave processes behind that slow down or confuse later tests."""
    import signal
    import subprocess
    p = subprocess.Popen(cmd, env=env, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                         start_new_session=True)
    try:
        out, err = p.communicate(timeout=timeout)
    except subprocess.TimeoutExpired:
        # SIGTERM first: launchers that supervise their own children in other process groups (torchrun) take their
        # workers down on it; then SIGKILL whatever is left of the session
        for sig, grace in ((signal.SIGTERM, 15), (signal.SIGKILL, None)):
            try:
                os.killpg(p.pid, sig)
            except ProcessLookupError:
                break
            try:
                out, err = p.communicate(timeout=grace)
                break
            except subprocess.TimeoutExpired:
                continue
        else:
            out, err = p.communicate()
        return subprocess.CompletedProcess(cmd, -9, out, (err or "") + f"\n[tests] job killed after {timeout} s")
    return subprocess.CompletedProcess(cmd, p.returncode, out, err)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA GPU (run on the B200 box with `pytest -m gpu`)")
    config.addinivalue_line("markers", "multigpu: needs >= 2 GPUs")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
        ngpu = torch.cuda.device_count() if has_gpu else 0
    except Exception:
        has_gpu, ngpu = False, 0
    for item in items:
        # a kernel that never finishes must not sit on the GPU for the CPU tier's generous limit
        if "gpu" in item.keywords and item.get_closest_marker("timeout") is None:
            item.add_marker(pytest.mark.timeout(700 if "multigpu" in item.keywords else 300))
        if "gpu" in item.keywords and not has_gpu and os.environ.get("DLCFN_MOCK_GPU_DRYRUN") != "1":
            item.add_marker(pytest.mark.skip(reason="no CUDA device"))
        if "multigpu" in item.keywords and ngpu < 2:
            item.add_marker(pytest.mark.skip(reason="needs >= 2 GPUs"))


@pytest.fixture(autouse=True)
def _seed_from_environment(request):
    """The reference's `@with_seed()` (MX/tests/python/unittest/common.py): when DLCFN_TEST_SEED / MXNET_TEST_SEED is set
    (tools/flakiness_checker.py sets it per trial) every test starts from that seed and a failure names it; without it tests
    keep their own fixed seeds."""
    seed = os.environ.get("DLCFN_TEST_SEED") or os.environ.get("MXNET_TEST_SEED")
    if seed is None:
        yield
        return
    import random
    import numpy as np
    import torch
    s = int(seed)
    random.seed(s)
    np.random.seed(s % (1 << 32))
    torch.manual_seed(s)
    yield
    rep = getattr(request.node, "rep_call", None)
    if rep is not None and rep.failed:
        print(f"\n[tests] {request.node.nodeid} failed with DLCFN_TEST_SEED={s}")


@pytest.hookimpl(hookwrapper=True)
def pytest_runtest_makereport(item, call):
    outcome = yield
    rep = outcome.get_result()
    setattr(item, "rep_" + rep.when, rep)


# ---------------------------------------------------------------------------------------------------------------------
# DLCFN_MOCK_GPU_DRYRUN=1: execute the BODIES of the GPU-marked tests on the host with the stand-in GPU (tests/mockgpu.py)
# and `dev()` pointing at the CPU.  Numeric assertions are expected to fail (the stubbed kernels compute nothing) and are
# ignored; any other exception is a bug in a wrapper or in the test code itself that would otherwise first show up on the
# GPU box.      DLCFN_MOCK_GPU_DRYRUN=1 DLCFN_EXPERIMENTAL=1 python -m pytest tests/test_gpu_experimental.py -m gpu -q
_DRYRUN = os.environ.get("DLCFN_MOCK_GPU_DRYRUN") == "1"


@pytest.hookimpl(trylast=True)
def pytest_itemcollected(item):
    if _DRYRUN and "multigpu" in item.keywords:
        item.add_marker(pytest.mark.skip(reason="dry run: multi-process GPU tests are not walked"))


@pytest.hookimpl(hookwrapper=True)
def pytest_pyfunc_call(pyfuncitem):
    if not (_DRYRUN and "gpu" in pyfuncitem.keywords):
        yield
        return
    import torch
    from mockgpu import mock_gpu
    mod = pyfuncitem.module
    had_dev = hasattr(mod, "dev")
    old_dev = getattr(mod, "dev", None)
    if had_dev:
        mod.dev = lambda: torch.device("cpu")
    with mock_gpu():
        outcome = yield
    if had_dev:
        mod.dev = old_dev
    exc = outcome.excinfo
    if exc is not None and issubclass(exc[0], AssertionError):
        outcome.force_result(None)           # numerics are not what a dry run checks
