"""Batch-state checkpoint (circuit_batch_save_state / circuit_batch_load_state in include/phy_engine_b200.h).

A transient saved after one analyze() and resumed in a NEW batch (another handle, as after a restart) must continue bit-identically
with the batch that never stopped: solution rows, capacitor / inductor / junction companion state, per-instance parameters, the
clock, and the sub-batches of the pivot safety net all travel in the blob."""
import numpy as np
import pytest

import pe_b200 as pe
import workloads as wl
from test_parity import abi  # noqa: F401  (fixture)


def make(abi, nl, over, n_inst, dt, steps):  # noqa: F811
    c = pe.Circuit(nl, abi)
    c.set_analyze_type(pe.TR)
    c.set_tr(dt, dt * (steps - 0.5))
    b = c.batch(n_inst)
    for e, name, v in over:
        b.set_param(e, name, v)
    return c, b


CASES = {
    "rc-ladder": lambda rng, n: (wl.rc_ladder(60), lambda info: [(e, "r", wl.sweep_values(rng, 1e3, n)) for e in info["R"]] + [(e, "c", wl.sweep_values(rng, 1e-9, n)) for e in info["C"]], 1e-8),
    "rlc-ladder": lambda rng, n: (wl.rlc_ladder(12), lambda info: [], 1e-9),
    "diode-rc": lambda rng, n: (wl.diode_rc(), lambda info: [], 1e-7),
    "transformer-ratio": lambda rng, n: (wl.transformer_stage(), lambda info: [(info["TX"], "n", np.resize(np.array([4.0, 1e-14, 0.0, 2.0, 1e-20]), n))], 1e-7),
}


@pytest.mark.parametrize("case", sorted(CASES))
def test_resume_is_bit_identical(abi, case):  # noqa: F811
    n_inst, steps = 9, 5
    rng = np.random.default_rng(3)
    (nl, info), over_of, dt = CASES[case](rng, n_inst)
    over = over_of(info)
    # the batch that never stops
    c0, b0 = make(abi, nl, over, n_inst, dt, steps)
    assert b0.analyze(), c0.abi.last_error()
    assert b0.analyze(), c0.abi.last_error()
    want = b0.solution()
    # one call, save, forget the batch; a new handle loads and continues
    c1, b1 = make(abi, nl, over, n_inst, dt, steps)
    assert b1.analyze(), c1.abi.last_error()
    blob = b1.save_state()
    t1 = b1.tr_duration
    if case == "transformer-ratio":
        assert b1.rescue_info(1)["sub_batches"] >= 1  # the instances the static order does not suit live in a sub-batch
    b1.close()
    c2, b2 = make(abi, nl, [(e, name, np.full(n_inst, v[0])) for e, name, v in over], n_inst, dt, steps)  # same keys, other values: the blob's win
    b2.load_state(blob)
    assert b2.tr_duration == t1
    assert b2.analyze(), c2.abi.last_error()
    assert np.array_equal(b2.solution(), want)
    assert (b2.status() == 0).all()


def test_load_refuses_what_does_not_fit(abi):  # noqa: F811
    nl, info = wl.rc_ladder(20)
    c, b = make(abi, nl, [], 4, 1e-8, 3)
    assert b.analyze()
    blob = b.save_state()
    # another instance count, another netlist, a truncated blob, garbage
    _, other_n = make(abi, nl, [], 5, 1e-8, 3)
    nl2, _ = wl.rc_ladder(21)
    _, other_nl = make(abi, nl2, [], 4, 1e-8, 3)
    _, same = make(abi, nl, [], 4, 1e-8, 3)
    for target, data in ((other_n, blob), (other_nl, blob), (same, blob[: len(blob) // 2]), (same, b"\0" * 64)):
        with pytest.raises(RuntimeError):
            target.load_state(data)
    same.load_state(blob)  # the handle is still usable after the refusals
    assert same.analyze() and b.analyze()
    assert np.array_equal(same.solution(), b.solution())
