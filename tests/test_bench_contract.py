"""bench.py / __graft_entry__ contract, checked on the host: `--dry-cpu` runs the whole script (both timed loops, the
rank reduction, the JSON line) with tiny shapes so that an edit to bench.py cannot first fail on the GPU box."""
import json
import os
import sys

import pytest

from conftest import run_job

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REQUIRED = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "clocks", "gpu_launches"}


def _last_json(out: str) -> dict:
    lines = [l for l in out.splitlines() if l.startswith("{")]
    assert len(lines) == 1, f"bench must print exactly one JSON line, got {len(lines)}:\n{out[-2000:]}"
    return json.loads(lines[0])


def test_bench_dry_run_single_rank_prints_the_contract_line():
    r = run_job([sys.executable, "bench.py", "--dry-cpu", "--steps", "3", "--warmup", "1"], cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    rec = _last_json(r.stdout)
    assert REQUIRED <= set(rec), REQUIRED - set(rec)
    assert rec["metric"] == "resnet50_train_images_per_sec" and rec["n_gpus"] == 1 and rec["steps"] == 3 and rec["warmup"] == 1
    assert rec["higher_is_better"] is True and rec["scaling"] == "weak" and rec["value"] > 0
    assert abs(rec["value"] - rec["config"]["global_batch"] * 1e3 / rec["ms_per_step"]) / rec["value"] < 0.01
    assert {"model", "global_batch", "seq_len", "parallelism"} <= set(rec["config"])
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(rec["clocks"])
    e = rec["e2e"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(e)
    b, s = rec["config"]["per_gpu_batch"], rec["config"]["image"]
    assert e["h2d_bytes_per_step"] == b * s * s * 3 + b * 8 and e["d2h_bytes_per_step"] == 4      # uint8 pixels + int64 labels
    assert e["value"] != rec["value"]                       # measured by its own loop
    assert rec["last_loss"] == rec["last_loss"] and e["last_loss"] == e["last_loss"]            # finite, not NaN
    assert "dry_cpu" in rec                                  # a host run is labelled as such


def test_bench_dry_run_bert_and_reference_arm():
    r = run_job([sys.executable, "bench.py", "--dry-cpu", "--model", "bert-large", "--steps", "2", "--warmup", "1"], cwd=ROOT,
                timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    rec = _last_json(r.stdout)
    assert rec["metric"] == "bert-large_pretrain_sequences_per_sec" and rec["vs_baseline"] is None
    r = run_job([sys.executable, "bench.py", "--impl", "reference", "--gpus", "1", "--steps", "2", "--warmup", "1"], cwd=ROOT,
                timeout=120)
    assert r.returncode == 0
    rec = _last_json(r.stdout)
    assert rec["impl"] == "reference" and "unavailable" in rec and "\n" not in rec["unavailable"]


def test_bench_dry_run_two_ranks_under_torchrun():
    """The driver's multi-GPU launch line, on gloo: one JSON line from rank 0, whole-job value, max-over-ranks time."""
    env = dict(os.environ, OMP_NUM_THREADS="2")
    port = 29500 + (os.getpid() % 400)
    r = run_job([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                 "127.0.0.1", "--master-port", str(port), "bench.py", "--dry-cpu", "--gpus", "2", "--steps", "2", "--warmup", "1"],
                cwd=ROOT, env=env, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    rec = _last_json(r.stdout)
    assert rec["n_gpus"] == 2 and rec["config"]["global_batch"] == 2 * rec["config"]["per_gpu_batch"]
    assert abs(rec["value"] - rec["config"]["global_batch"] * 1e3 / rec["ms_per_step"]) / rec["value"] < 0.01
    assert rec["e2e"]["h2d_bytes_per_step"] == 2 * (rec["config"]["per_gpu_batch"] * (rec["config"]["image"] ** 2 * 3 + 8))


def test_graft_entry_has_build_and_smoke():
    sys.path.insert(0, ROOT)
    try:
        import __graft_entry__ as g
    finally:
        sys.path.pop(0)
    assert callable(g.build) and callable(g.smoke)


def test_no_undefined_names_in_code_that_only_runs_on_a_gpu():
    """GPU-only branches, gated experimental tests and tools never execute on the CPU test tier: a static pass over every
    module (tools/lint_names.py) keeps typos in them from first surfacing on the GPU box."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    try:
        import lint_names
    finally:
        sys.path.pop(0)
    files = lint_names.collect([os.path.join(ROOT, x) for x in ("deeplearning_cfn_b200", "tests", "tools", "bench.py", "__graft_entry__.py")])
    problems = [p for f in files for p in lint_names.check_file(f)]
    assert len(files) > 100 and not problems, "\n".join(problems)


def test_ctypes_tables_match_the_c_signatures():
    """tools/check_abi.py: parameter count / class and pointer-or-64-bit return types of every ctypes binding against the
    exported C functions of both native libraries (a wrong table only misbehaves at run time - for CUDA, on the GPU box)."""
    r = run_job([sys.executable, os.path.join(ROOT, "tools", "check_abi.py")], cwd=ROOT, timeout=300)
    assert r.returncode == 0 and " 0 mismatches" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
    compared = int(r.stdout.strip().splitlines()[-1].split(" exported functions parsed, ")[1].split(" ")[0])
    assert compared >= 100


def test_gpu_tier_test_bodies_run_on_the_stand_in_gpu():
    """DLCFN_MOCK_GPU_DRYRUN=1 executes the bodies of the GPU-marked experimental and model-zoo tests on the host (stubbed
    kernels, numeric assertions ignored): a typo, a wrong argument count or a bad constructor in a test that only ever runs
    on the GPU box shows up here."""
    env = dict(os.environ, DLCFN_MOCK_GPU_DRYRUN="1", DLCFN_EXPERIMENTAL="1", PYTHONPATH=ROOT)
    r = run_job([sys.executable, "-m", "pytest", "tests/test_gpu_experimental.py", "tests/test_models.py", "-m", "gpu", "-q", "--no-header",
                 "-p", "no:cacheprovider"], cwd=ROOT, env=env, timeout=900)
    tail = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else ""
    assert r.returncode == 0 and " passed" in tail and "failed" not in tail, r.stdout[-3000:] + r.stderr[-1000:]


def test_launch_time_divisors_of_the_gemm_producers_are_exact():
    """csrc/cuda/gemm_tcgen05.cu:FastDiv - the TMA producer / MMA issuer divide by launch-time constants (tiles, image rows,
    channel blocks) with a multiply-high + shift; the host export evaluates the same arithmetic.  Exact for every 0 <= n < 2^31."""
    import random
    from deeplearning_cfn_b200.ops import native
    L = native.lib()
    rnd = random.Random(7)
    ds = list(range(1, 600)) + [rnd.randrange(1, 1 << 24) for _ in range(300)] + [1 << k for k in range(31)] + \
        [(1 << k) + 1 for k in range(1, 30)] + [(1 << k) - 1 for k in range(2, 31)]
    top = (1 << 31) - 1
    for d in ds:
        ns = [0, 1, d - 1, d, d + 1, 2 * d - 1, 2 * d, top, top - 1, (top // d) * d, (top // d) * d - 1] + [rnd.randrange(0, 1 << 31) for _ in range(40)]
        for n in ns:
            if 0 <= n <= top:
                assert L.dlcfn_fastdiv_host(n, d) == n // d, (n, d)
