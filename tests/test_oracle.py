"""Pins the oracle: the compiled reference (oracle/_ref/libpe_ref.so, built by oracle/Makefile from /root/reference)
against every known answer the reference's own tests hold for this path, and the read-only probes of
oracle/ref_harness.cpp against the reference's own analyze().  CPU only."""
import math

import numpy as np
import pytest

import pe_b200 as pe
import refapi
import workloads as wl


def solo(nl, at, **kw):
    c = refapi.RefCircuit(nl)
    c.set_analyze_type(at)
    if "tr" in kw:
        c.set_tr(*kw["tr"])
    if "omega" in kw:
        c.set_ac_omega(kw["omega"])
    return c


def test_rc_step_matches_reference_test_bound(ref):
    # test/0005.models/rc_step_tr.cpp:59-61: vout(tau) = 1 - 1/e within 5e-3; SURVEY 8(c): 0.63027499952138644 after 100 steps
    nl, info = wl.rc_ladder(1, r=1e3, c=1e-9)
    c = solo(nl, pe.TR, tr=(1e-8, 1e-6))
    ok, n = c.analyze_counted()
    assert ok and n == 100
    vout = c.solution()[1].real
    assert abs(vout - (1.0 - math.exp(-1.0))) < 5e-3
    assert abs(vout - 0.63027499952138644) < 1e-15
    assert c.tr_duration == 1.0000000000000004e-06


def test_diode_op_trajectory(ref):
    # test/0011.nonlinear/op_pn_junction.cpp:28-29 (0.5 < Vd < 0.9) and the recorded solve_once trajectory
    nl, info = wl.diode_resistor()
    c = solo(nl, pe.OP)
    c.prepare()
    want = [0.99999999961337649, 0.99999998608866025, 0.99999712300284538, 0.99951739457696998, 0.94370916686840356, 0.62949799338322621,
            0.62944165509863270, 0.62944159766137553, 0.62944159766131580, 0.62944159766131591]
    for w in want:
        assert c.solve_once()
        assert abs(c.solution()[1].real - w) < 1e-15
    c2 = solo(nl, pe.OP)
    ok, n = c2.analyze_counted()
    assert ok and n == 7
    assert 0.5 < c2.solution()[1].real < 0.9
    assert c2.solution()[1].real == 0.62944165509863270


def test_counted_analyze_is_analyze(ref):
    # the harness' counting Newton driver returns bit-identical state to the reference's own circuit_analyze()
    nl, info = wl.diode_ladder(8)
    a = solo(nl, pe.OP)
    assert a.analyze()
    b = solo(nl, pe.OP)
    ok, n = b.analyze_counted()
    assert ok and n >= 2
    assert (a.solution() == b.solution()).all()


def test_ac_omega_known_answer(ref):
    # test/0012.ac/ac_omega.cpp:30-32: RC low-pass at omega = 1/(RC): 0.6 < |Vout| < 0.8; exact (0.5, -0.5)
    nl = pe.Netlist()
    g = nl.ground()
    src = nl.add(pe.VAC, 1.0, 50.0, 0.0)
    r = nl.add(pe.R, 1e3)
    cap = nl.add(pe.C, 1e-6)
    nl.wire(src, 1, g, 0)
    nl.wire(src, 0, r, 0)
    nl.wire(r, 1, cap, 0)
    nl.wire(cap, 1, g, 0)
    c = solo(nl, pe.AC, omega=1000.0)
    assert c.analyze()
    vout = c.solution()[1]
    assert 0.6 < abs(vout) < 0.8
    assert abs(vout - (0.5 - 0.5j)) < 1e-15


def test_reference_failure_contract(ref):
    # resistor-biased NPN does not converge in the reference (64 iterations, analyze() false)
    nl, info = wl.npn_resistor_biased()
    c = solo(nl, pe.OP)
    ok, n = c.analyze_counted()
    assert not ok and n == 64


def test_run_batch_equals_single_runs(ref):
    nl, info = wl.rc_ladder(5)
    rng = np.random.default_rng(0)
    over = [(info["R"][2], "r", rng.uniform(500, 2000, 3))]
    r = refapi.run_batch(nl, pe.TR, 3, over, t_step=1e-8, t_stop=1e-7, threads=2)
    for i in range(3):
        c = solo(nl, pe.TR, tr=(1e-8, 1e-7))
        assert c.set_param(info["R"][2], "r", float(over[0][2][i])) == 0
        assert c.analyze()
        assert (c.solution() == r["x"][i]).all()


def test_golden_file_is_current(ref):
    """the committed golden vectors are what the compiled reference produces today (spot check of two cases)"""
    import os

    import golden_cases

    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cases.npz"))
    for name in ("diode_op_mc", "rc_ladder_tr_24x37"):
        nl, over, kw = golden_cases.build(name)
        r = refapi.run_batch(nl, golden_cases.CASES[name]["at"], golden_cases.CASES[name]["n_inst"], over, **kw)
        assert (r["x"] == gold[name + "/x"]).all()
        assert (r["solves"] == gold[name + "/solves"]).all()
