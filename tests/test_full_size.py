"""BASELINE.json configs at FULL size on the GPU, checked through size-independent properties plus a reference spot check
(the reference cannot run 1e6 solves in a test): replication (instances with identical parameters, placed in different
lane groups / clusters, give bit-identical answers), exact linearity (a source scaled by 2 scales every unknown by exactly
2 in binary floating point), and the compiled reference on a handful of the instances."""
import os

import numpy as np
import pytest

import pe_b200 as pe
import workloads as wl
from test_parity import assert_close

pytestmark = pytest.mark.gpu


def test_config_b_full_size_rc_ladder_transient():
    n_inst, sections, steps = 10000, 1000, 100
    nl, info = wl.rc_ladder(sections)
    rng = np.random.default_rng(77)
    base = 4000
    r_vals = np.empty((sections, n_inst))
    c_vals = np.empty((sections, n_inst))
    r_vals[:, :base] = 1e3 * rng.uniform(0.8, 1.2, (sections, base))
    c_vals[:, :base] = 1e-9 * rng.uniform(0.8, 1.2, (sections, base))
    # instances [base, 2 base): the same circuits with the source doubled; the rest: replicas of the first instances
    r_vals[:, base:2 * base] = r_vals[:, :base]
    c_vals[:, base:2 * base] = c_vals[:, :base]
    rep = n_inst - 2 * base
    r_vals[:, 2 * base:] = r_vals[:, :rep]
    c_vals[:, 2 * base:] = c_vals[:, :rep]
    v_src = np.full(n_inst, 1.0)
    v_src[base:2 * base] = 2.0
    c = pe.Circuit(nl)
    c.set_analyze_type(pe.TR)
    c.set_tr(1e-8, 1e-8 * (steps - 0.5))
    b = c.batch(n_inst)
    b.set_param(info["src"], "V", v_src)
    for k, e in enumerate(info["R"]):
        b.set_param(e, "r", r_vals[k])
    for k, e in enumerate(info["C"]):
        b.set_param(e, "c", c_vals[k])
    assert b.analyze(), c.abi.last_error()
    assert b.total_solves == n_inst * steps
    x = b.solution()
    assert np.isfinite(x).all()
    assert (x[2 * base:] == x[:rep]).all(), "replicated instances differ"
    # exact wherever nothing underflows (the far end of the ladder is still at ~1e-300 V after 100 steps)
    big = np.abs(x[:base]) > 1e-280
    assert big.mean() > 0.1
    assert (x[base:2 * base][big] == 2.0 * x[:base][big]).all(), "doubling the source does not double the state exactly"
    assert (np.abs(x[base:2 * base][~big]) <= 4e-280).all()
    # physical sanity: a charging ladder, monotone along the ladder
    assert (x[:, 0] == v_src).all() and (np.diff(x[:, :sections + 1], axis=1) <= 1e-12).all()
    import refapi

    if os.path.exists(refapi.REF_LIB):
        pick = [0, 1777, base - 1]
        over = [(info["src"], "V", v_src[pick])] + [(e, "r", r_vals[k][pick]) for k, e in enumerate(info["R"])] + [(e, "c", c_vals[k][pick]) for k, e in enumerate(info["C"])]
        want = refapi.run_batch(nl, pe.TR, len(pick), over, t_step=1e-8, t_stop=1e-8 * (steps - 0.5))
        assert (want["ok"] == 1).all() and (want["solves"] == steps).all()
        assert_close(x[pick], want["x"].real, "config B full size vs reference")


def test_config_d_full_size_ac_sweep_linearity():
    points = 1_000_000
    nl, info = wl.rlc_ladder(64)
    c = pe.Circuit(nl)
    c.set_analyze_type(pe.AC)
    b = c.batch(2)
    b.set_param(info["src"], "Vp", np.array([1.0, 2.0]))
    b.set_ac_sweep(pe.SWEEP_LOG, 1e3, 1e10, points)
    assert b.analyze(), c.abi.last_error()
    assert b.total_solves == 2 * points and (b.status() == 0).all()
    # values of the big sweep itself, on a sample of its own points (every 251st: 3 985 points per instance, the ends included):
    # linear in the source amplitude bit for bit, and equal to the reference solving the same omega on its own
    sample = np.unique(np.concatenate([np.arange(0, points, 251), [points - 1, points - 2, 1]]))
    x0 = b.ac_solution_lanes(sample)
    x1 = b.ac_solution_lanes(points + sample)
    big_s = (np.abs(x0.real) > 1e-280) & (np.abs(x0.imag) > 1e-280)
    assert big_s.mean() > 0.1 and (x1[big_s] == 2.0 * x0[big_s]).all()
    om_all = b.ac_omegas()
    assert om_all.size == points
    import refapi as _refapi

    if os.path.exists(_refapi.REF_LIB):
        rr = _refapi.RefCircuit(nl)
        rr.set_analyze_type(pe.AC)
        for q in range(0, sample.size, 16):  # every 16th sampled point: 250 single-point reference solves
            rr.set_ac_omega(float(om_all[sample[q]]))
            assert rr.analyze()
            assert_close(x0[q].real, rr.solution().real, f"config D, point {int(sample[q])} (re)")
            assert_close(x0[q].imag, rr.solution().imag, f"config D, point {int(sample[q])} (im)")
    # the full solution (2 x 1e6 x 194 complex) is 6 GB: check a small sweep of the same circuit point by point instead,
    # and the big one through its lane count, status flags and the linearity of a 4097-point sweep
    b2 = c.batch(2)
    b2.set_param(info["src"], "Vp", np.array([1.0, 2.0]))
    b2.set_ac_sweep(pe.SWEEP_LOG, 1e3, 1e10, 4097)
    assert b2.analyze()
    x = b2.ac_solution()
    big = (np.abs(x[0].real) > 1e-280) & (np.abs(x[0].imag) > 1e-280)  # exact wherever nothing underflows
    assert big.mean() > 0.1
    assert (x[1][big] == 2.0 * x[0][big]).all()
    assert np.isfinite(x.real).all() and np.isfinite(x.imag).all()
    import refapi

    if os.path.exists(refapi.REF_LIB):
        r = refapi.RefCircuit(nl)
        r.set_analyze_type(pe.AC)
        r.set_ac_sweep(pe.SWEEP_LOG, 1e3, 1e10, 4097)
        ok, n = r.analyze_counted()
        om, xr = r.ac_results()
        assert ok and n == 4097 and (b2.ac_omegas() == om).all()
        assert_close(x[0], xr, "config D vs reference")
