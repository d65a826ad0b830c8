"""Host model check of the device flag protocols (tools/protocol_model.py): every interleaving of the synchronisation skeletons
of the fused collectives / the staged pull is safe and live, every negative control (the protecting mechanism removed) and the
dropped-flag fault are caught.  The GPU-side counterparts are tests/mp/protocol_worker.py and tools/sanitize_collectives.sh."""
import importlib.util
import pathlib

import pytest

_spec = importlib.util.spec_from_file_location("protocol_model", pathlib.Path(__file__).resolve().parents[1] / "tools" / "protocol_model.py")
pm = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(pm)


@pytest.mark.parametrize("W", [2, 3])
@pytest.mark.parametrize("name", list(pm.PROTOCOLS))
def test_protocol_is_safe_and_live_and_its_control_is_rejected(name, W):
    K = pm.default_calls(name, W)
    v = pm.check(name, W, K)
    assert v.ok and v.states > 10, v.violation
    c = pm.check(name, W, K, control=True)
    assert not c.ok and c.trace, f"the checker accepted {name} without its protection ({pm.PROTOCOLS[name][3]})"
    # the counterexample is a real schedule: every step names a program of the model
    assert all(":" in step for step in c.trace)


@pytest.mark.parametrize("name", ["twoshot", "oneshot"])
def test_dropped_flag_store_blocks_the_victim_and_never_yields_a_wrong_value(name):
    """SymmDev::drop_peer in the model: the victim's wait never completes (on the device the spin timeout turns that into the
    error word, tests/mp/protocol_worker.py) - and no schedule lets anybody read a wrong value first."""
    f = pm.check(name, 3, 3, drop=(0, 1))
    assert not f.ok and f.violation.startswith("deadlock") and "rank 1 at ('wge'" in f.violation, f.violation


def test_parity_reuse_needs_three_calls_to_show():
    """Why the default call counts are >= 3 for the parity protocols: with two calls the one-slot control of the flag-in-data
    form still deadlocks (the second call overwrites the first), but a TWO-slot protocol can only be attacked by call k + 2."""
    assert pm.check("ll", 2, 2, control=True).ok is False
    assert pm.check("oneshot", 3, 3).ok and pm.check("ll", 3, 4).ok


def test_cli_table():
    assert pm.main(["--world", "2", "--only", "oneshot", "ps_async", "--markdown"]) == 0
