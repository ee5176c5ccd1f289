"""Pivot safety net (host/batch.cpp run_rescues, PE_F_GUARD in csrc/pe_b200_program.h).

The elimination order is static: chosen once, on the first instance's values.  Eigen's SparseLU, which the reference calls,
re-pivots in every solve (include/Eigen/src/SparseLU/SparseLU_pivotL.h:76-107), so an instance whose values make a pivot of that
order vanish is still solved there.  Here such an instance trips the device-side pivot guard (an entry of L more than 2^20
times an entry of its column's pivot) and is solved again in a sub-batch whose order is chosen on ITS values.  These tests sweep a transformer ratio (the pivot of the nominal order is the
ratio itself) through tiny values and zero and compare every instance with the compiled reference -- OP, a transient that is
continued by a second analyze(), an AC sweep -- and check what the guard costs where it is not needed (guard elision).
"""
import numpy as np
import pytest

import pe_b200 as pe
import refapi
import workloads as wl
from test_parity import abi, assert_close  # noqa: F401  (abi is a fixture)

RATIOS = np.array([4.0, 2.0, 1e-3, 1e-9, 1e-14, 0.0, 1e6, 1e12, -3.0, 1e-20, 5.0, 1e18])


def batch_of(abi, nl, at, over, n_inst, guard=None, tr=None, ac=None, probes=None):  # noqa: F811
    c = pe.Circuit(nl, abi)
    c.set_analyze_type(at)
    if tr:
        c.set_tr(*tr)
    b = c.batch(n_inst)
    if ac:
        b.set_ac_sweep(*ac)
    if guard is not None:
        b.set_pivot_guard(guard, -1)
    for e, name, v in over:
        b.set_param(e, name, v)
    if probes is not None:
        b.set_probes(probes)
    return c, b


def test_ratio_sweep_across_a_vanishing_pivot_op(ref, abi):  # noqa: F811
    nl, info = wl.transformer_stage()
    over = [(info["TX"], "n", RATIOS)]
    want = refapi.run_batch(nl, pe.OP, RATIOS.size, over)
    assert (want["ok"] == 1).all()  # the reference solves every instance, n = 0 included
    # without the safety net the static order is silently inaccurate on some instances and singular on one
    c0, b0 = batch_of(abi, nl, pe.OP, over, RATIOS.size, guard=0.0)
    assert not b0.analyze()
    st0 = b0.status()
    x0 = b0.solution()
    assert st0[5] == 2 and (np.delete(st0, 5) == 0).all()
    rel0 = np.abs(x0 - want["x"].real).max(axis=1) / np.abs(want["x"].real).max(axis=1)
    assert rel0[4] > 1e-6 and rel0[9] > 1.0  # status OK, values wrong: what the guard is for
    # with it every instance matches the reference
    c, b = batch_of(abi, nl, pe.OP, over, RATIOS.size)
    assert b.analyze(), c.abi.last_error()
    assert (b.status() == 0).all()
    assert_close(b.solution(), want["x"].real, "ratio sweep, OP")
    soa = np.zeros((b.n_unknowns(), RATIOS.size))
    b.solution_soa_into(soa.ctypes.data)
    assert_close(soa.T, want["x"].real, "ratio sweep, OP (device layout)")
    info_r = b.rescue_info(0)
    assert info_r["flagged"] >= 3 and info_r["rescued"] == info_r["flagged"] and info_r["unguarded"] == 0
    assert (b.newton_iters() == 1).all()
    # a parameter write drops the sub-batches; the next analyze() builds them again
    b.set_param(info["R"], "r", np.full(RATIOS.size, 2e3))
    over2 = over + [(info["R"], "r", np.full(RATIOS.size, 2e3))]
    want2 = refapi.run_batch(nl, pe.OP, RATIOS.size, over2)
    assert b.analyze(), c.abi.last_error()
    assert_close(b.solution(), want2["x"].real, "ratio sweep, OP, after a parameter write")


def test_ratio_sweep_transient_continues_in_the_sub_batch(ref, abi):  # noqa: F811
    nl, info = wl.transformer_stage()
    ratios = RATIOS[:8]
    over = [(info["TX"], "n", ratios)]
    dt, steps = 1e-7, 6
    c, b = batch_of(abi, nl, pe.TR, over, ratios.size, tr=(dt, dt * (steps - 0.5)), probes=[1, 2, 3])
    assert b.analyze(), c.abi.last_error()
    want = refapi.run_batch(nl, pe.TR, ratios.size, over, t_step=dt, t_stop=dt * (steps - 0.5))
    assert (want["ok"] == 1).all() and (want["solves"] == steps).all()
    assert_close(b.solution(), want["x"].real, "ratio sweep, TR")
    assert b.rescue_info(1)["rescued"] >= 2
    w = b.waveform(steps)  # [steps, probes, n_inst]: rows of the rescued instances come from the sub-batch
    x = b.solution()
    for k, u in enumerate([1, 2, 3]):
        assert np.array_equal(w[steps - 1, k], x[:, u])
    # second call: the rescued instances continue where THEIR transient stopped; every instance alone (an order of its own,
    # nothing to rescue) must give the same trajectory
    assert b.analyze(), c.abi.last_error()
    x2 = b.solution()
    for i in (0, 3, 4, 5):
        ci, bi = batch_of(abi, nl, pe.TR, [(info["TX"], "n", ratios[i:i + 1])], 1, tr=(dt, dt * (steps - 0.5)))
        assert bi.analyze() and bi.analyze()
        assert_close(x2[i], bi.solution()[0], f"continued transient of instance {i}")
    assert b.total_solves == steps * ratios.size  # per call: the rescued instances count the solves of their sub-batch


def test_ratio_sweep_ac(ref, abi):  # noqa: F811
    nl, info = wl.transformer_stage(vac=True)
    ratios = RATIOS[:8]
    over = [(info["TX"], "n", ratios)]
    sweep = (2, 1e3, 1e7, 5)  # log
    want = refapi.run_batch(nl, pe.AC, ratios.size, over, ac=sweep)
    assert (want["ok"] == 1).all()
    c, b = batch_of(abi, nl, pe.AC, over, ratios.size, ac=sweep)
    assert b.analyze(), c.abi.last_error()
    got = b.ac_solution()  # [n_inst, points, n] complex
    assert_close(got.real, want["x"].real, "ratio sweep, AC (re)")
    assert_close(got.imag, want["x"].imag, "ratio sweep, AC (im)")
    assert b.rescue_info(3)["rescued"] >= 2


def test_guard_elision_and_its_limits(abi):  # noqa: F811
    # rows that carry only positive conductances need no guard (symmetric, diagonally dominant: growth <= 2 in any order):
    # of the 102 pivots of a 100-section RC ladder only the source's keep it, and only one of those has a column of L to test --
    # the hot loop of config B pays nothing
    nl, info = wl.rc_ladder(100)
    c = pe.Circuit(nl, abi)
    c.set_analyze_type(pe.TR)
    c.set_tr(1e-8, 3e-8)
    b = c.batch(5)
    assert b.analyze(), c.abi.last_error()
    assert b.rescue_info(1)["guarded_pivots"] == 1
    # a negative resistance takes the argument away: every pivot is guarded again
    b.set_param(info["R"][3], "r", np.array([1e3, 1e3, -1e3, 1e3, 1e3]))
    assert b.analyze(), c.abi.last_error()
    assert b.rescue_info(1)["guarded_pivots"] >= 100  # every pivot that has a column of L
    # inductors put a branch row next to every node: nothing is provably safe
    nl2, _ = wl.rlc_ladder(20)
    c2 = pe.Circuit(nl2, abi)
    c2.set_analyze_type(pe.TR)
    c2.set_tr(1e-9, 3e-9)
    b2 = c2.batch(3)
    assert b2.analyze(), c2.abi.last_error()
    r = b2.rescue_info(1)
    assert r["guarded_pivots"] > 40 and r["flagged"] == 0


def test_single_instance_abi_takes_the_same_net(ref, abi):  # noqa: F811
    # circuit_analyze() (Part 1 of the ABI) runs a batch of one: its order is chosen on its own values, so there is nothing to
    # re-order, but the guard and the final unguarded round must not change what it returns
    for n in (4.0, 1e-14, 0.0):
        nl, info = wl.transformer_stage(n=n)
        rc = refapi.RefCircuit(nl)
        rc.set_analyze_type(pe.OP)
        assert rc.analyze()
        c = pe.Circuit(nl, abi)
        c.set_analyze_type(pe.OP)
        assert c.analyze(), c.abi.last_error()
        assert_close(c.solution().real, rc.solution().real, f"single instance, n = {n}")


def test_wide_ac_sweep_re_runs_only_the_flagged_points(ref, abi):  # noqa: F811
    # config D's circuit over seven decades: the order is chosen at the geometric mean of the sweep, a few per cent of the
    # points (the ends) exceed the multiplier bound and are solved again -- each as a one-point unit with its own omega and an
    # order chosen there -- not the whole sweep of their instance
    nl, info = wl.rlc_ladder(64)
    sweep = (pe.SWEEP_LOG, 1e3, 1e10, 256)
    c = pe.Circuit(nl, abi)
    c.set_analyze_type(pe.AC)
    b = c.batch(2)
    b.set_ac_sweep(*sweep)
    b.set_param(info["R"][0], "r", np.array([10.0, 12.0]))
    assert b.analyze(), c.abi.last_error()
    r = b.rescue_info(3)
    assert 0 < r["flagged"] < 2 * 256 // 4 and r["rescued"] == r["flagged"] and r["unguarded"] == 0 and r["sub_batches"] == 1
    assert (b.status() == 0).all() and (b.newton_iters() == 1).all() and b.total_solves == 2 * 256
    got = b.ac_solution()
    want = refapi.run_batch(nl, pe.AC, 2, [(info["R"][0], "r", np.array([10.0, 12.0]))], ac=sweep)
    assert_close(got.real, want["x"].real, "wide sweep (re)")
    assert_close(got.imag, want["x"].imag, "wide sweep (im)")
    # a sample of lanes (circuit_batch_ac_solution_lanes) reads the same values, rescued points included
    sel = np.array([0, 3, 200, 254, 255, 256, 300, 511])
    assert np.array_equal(b.ac_solution_lanes(sel), got.reshape(2 * 256, -1)[sel])
    # the same sweep again: the omega table stays on the device, the sub-batch is re-used, the results are the same bits
    assert b.analyze(), c.abi.last_error()
    assert np.array_equal(b.ac_solution(), got)
    r2 = b.rescue_info(3)
    assert r2["sub_batches"] == 1 and r2["flagged"] == r["flagged"]
    # another sweep: new table, new flags (the sub-batch of the old sweep numbers ITS points: it must go)
    b.set_ac_sweep(pe.SWEEP_LOG, 1e5, 1e7, 64)
    assert b.analyze(), c.abi.last_error()
    want2 = refapi.run_batch(nl, pe.AC, 2, [(info["R"][0], "r", np.array([10.0, 12.0]))], ac=(pe.SWEEP_LOG, 1e5, 1e7, 64))
    assert_close(b.ac_solution().real, want2["x"].real, "narrow sweep (re)")
    assert_close(b.ac_solution().imag, want2["x"].imag, "narrow sweep (im)")


def test_repeated_analysis_keeps_counters_and_status_consistent(ref, abi):  # noqa: F811
    # second and later calls on unchanged inputs take a short path (device-side counts only): status, per-lane counters, totals and
    # values must be what the first call reported
    nl, info = wl.transformer_stage()
    over = [(info["TX"], "n", RATIOS)]
    c, b = batch_of(abi, nl, pe.OP, over, RATIOS.size)
    assert b.analyze(), c.abi.last_error()
    first = (b.solution().copy(), b.status().copy(), b.newton_iters().copy(), b.total_solves, b.rescue_info(0))
    for _ in range(3):
        assert b.analyze(), c.abi.last_error()
        assert np.array_equal(b.solution(), first[0]) and np.array_equal(b.status(), first[1]) and np.array_equal(b.newton_iters(), first[2])
        assert b.total_solves == first[3] == RATIOS.size
    again = b.rescue_info(0)
    assert again["sub_batches"] == first[4]["sub_batches"] and again["flagged"] == first[4]["flagged"]  # nothing was re-built or re-flagged
    # a non-linear batch where one instance fails for good (the reference fails there too): the failure is reported every time
    nl2, info2 = wl.npn_resistor_biased()
    c2 = pe.Circuit(nl2, abi)
    c2.set_analyze_type(pe.OP)
    b2 = c2.batch(3)
    ok1 = b2.analyze()
    st1 = b2.status().copy()
    ok2 = b2.analyze()
    assert ok1 == ok2 and np.array_equal(b2.status(), st1)


def test_comparators_read_the_rescued_solution(abi):  # noqa: F811
    # analog -> digital boundary on top of the safety net: the comparator of a rescued instance compares THAT instance's
    # re-solved voltages (the main batch's rows for it are what the guard rejected)
    nl, info = wl.transformer_stage()
    cmp_ = nl.add(pe.COMPARATOR, 0.0, 5.0)
    vr = nl.add(pe.VDC, 1e-15)
    nl.wire(vr, 1, 0, 0)  # element 0 is the ground placeholder
    nl.wire(cmp_, 1, vr, 0)
    nl.wire(cmp_, 0, info["R"], 0)
    ratios = np.array([4.0, 1e-14, 0.0, 2.0, 1e-20, 20.0])
    c = pe.Circuit(nl, abi)
    c.set_analyze_type(pe.OP)
    b = c.batch(ratios.size)
    b.set_param(info["TX"], "n", ratios)
    assert b.analyze(), c.abi.last_error()
    assert b.rescue_info(0)["rescued"] >= 2
    v = b.solution()[:, c.pin_unknown(info["R"], 0)]
    assert v[1] > 1e-13 and 0.0 < v[4] < 1e-15  # the rescued instances carry their own tiny, accurate voltages
    assert (b.digital_clk().ravel() == (v >= 1e-15).astype(np.uint8)).all()
    assert b.digital_clk().ravel().tolist() == [1, 1, 0, 1, 0, 1]
