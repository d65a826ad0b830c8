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
        if "gpu" in item.keywords and not has_gpu:
            item.add_marker(pytest.mark.skip(reason="no CUDA device"))
        if "multigpu" in item.keywords and ngpu < 2:
            item.add_marker(pytest.mark.skip(reason="needs >= 2 GPUs"))
