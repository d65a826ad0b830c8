"""Fused peer-HBM collectives / parameter-server kernels at world size 2 (and more when available).  The file sorts last in the
GPU tier: these tests start multi-process jobs over several devices, and a box-level problem there (a busy peer GPU, a port) must
not stop the single-GPU kernel tests from running under `pytest -x`."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(nproc, script, port, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "mp", script)]
    from conftest import run_job
    return run_job(cmd, timeout=timeout, cwd=ROOT)


def test_collectives_world2():
    r = _run(2, "collectives_worker.py", 29611)
    assert r.returncode == 0 and "COLLECTIVES_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_collectives_all_gpus():
    n = torch.cuda.device_count()
    if n < 4:
        pytest.skip("needs >= 4 GPUs")
    r = _run(n, "collectives_worker.py", 29612)
    assert r.returncode == 0 and "COLLECTIVES_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_trainer_modes_agree_world2():
    r = _run(2, "trainer_worker.py", 29613)
    assert r.returncode == 0 and "TRAINER_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_kvstore_dist_device_world2():
    r = _run(2, "kv_device_worker.py", 29614)
    assert r.returncode == 0 and "KV_DEVICE_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_ps_resident_linear_pull_fused_gemm_world2():
    r = _run(2, "sharded_linear_worker.py", 29615)
    assert r.returncode == 0 and "SHARDED_LINEAR_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_bandwidth_tool_per_key_error_is_zero_world2():
    r = _run(2, "bandwidth_worker.py", 29616)
    assert r.returncode == 0 and "BANDWIDTH_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_flag_protocol_under_schedule_fuzzing_and_fault_injection_world2():
    r = _run(2, "protocol_worker.py", 29617)
    assert r.returncode == 0 and "PROTOCOL_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
