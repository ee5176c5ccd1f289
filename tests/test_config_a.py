"""Config A of BASELINE.json (benchmark/series_parallel.cpp: a ring of random resistors with random node merges, one DC
solve): the reduce-and-core path (host/frontal.cpp, csrc/pe_b200_frontal.cu) against the compiled reference.

  * seeded clones at n ~ 3e3 and 1e4 unknowns, single instance and Monte-Carlo batches (emulator and GPU)
  * -m gpu: the full size (100 003 resistors, 9 000 merges, 91 003 unknowns) against the reference's own solution and
    against Kirchhoff's current law evaluated independently from the netlist (a size-independent property)
"""
import ctypes as ct

import numpy as np
import pytest

import pe_b200 as pe
import refapi
import workloads as wl
from test_parity import abi, assert_close  # noqa: F401  (abi is a fixture)


@pytest.fixture
def frontal_from(abi):  # noqa: F811
    abi.lib.phy_engine_b200_set_frontal_min.argtypes = [ct.c_size_t]

    def set_min(n):
        assert abi.lib.phy_engine_b200_set_frontal_min(n) == 0

    yield set_min
    abi.lib.phy_engine_b200_set_frontal_min(20000)


@pytest.mark.parametrize("n_ring,n_merge", [(3000, 270), (10000, 900)])
def test_series_parallel_matches_reference(ref, abi, frontal_from, n_ring, n_merge):  # noqa: F811
    frontal_from(500)
    nl, info = wl.series_parallel(n_ring, n_merge, seed=5)
    rc = refapi.RefCircuit(nl)
    rc.set_analyze_type(pe.DC)
    ok, n = rc.analyze_counted()
    assert ok and n == 1
    c = pe.Circuit(nl, abi)
    c.set_analyze_type(pe.DC)
    assert c.analyze(), c.abi.last_error()
    assert_close(c.solution(), rc.solution().real, "series_parallel, single instance")
    b = c.batch(1)
    assert b.analyze()
    fi = b.frontal_info()
    assert b.last_kernel() == 3 and fi is not None
    assert fi["unknowns"] == len(rc.solution()) and fi["eliminated"] + fi["core_rows"] == fi["unknowns"]
    assert fi["levels"] <= 12  # the chains between two junctions halve per level
    assert fi["core_rows"] < 0.12 * fi["unknowns"]


def test_series_parallel_monte_carlo_batch(ref, abi, frontal_from):  # noqa: F811
    frontal_from(500)
    n_ring, n_merge, n_inst = 2500, 220, 6
    nl, info = wl.series_parallel(n_ring, n_merge, seed=8)
    rng = np.random.default_rng(3)
    picks = [info["res"][k] for k in rng.choice(len(info["res"]), 400, replace=False)]
    over = [(e, "r", rng.uniform(1e-5, 1e5, n_inst)) for e in picks] + [(info["V"], "V", rng.uniform(1.0, 5.0, n_inst))]
    want = refapi.run_batch(nl, pe.DC, n_inst, over)
    assert (want["ok"] == 1).all()
    c = pe.Circuit(nl, abi)
    c.set_analyze_type(pe.DC)
    b = c.batch(n_inst)
    for e, name, v in over:
        b.set_param(e, name, v)
    assert b.analyze(), c.abi.last_error()
    assert b.last_kernel() == 3
    assert (b.newton_iters() == 1).all() and (b.status() == 0).all()
    assert_close(b.solution(), want["x"].real, "series_parallel Monte Carlo")


def test_small_or_mixed_circuits_keep_the_program_kernels(abi, frontal_from):  # noqa: F811
    frontal_from(500)
    nl, info = wl.rc_ladder(700)  # capacitors: not a resistor network
    c = pe.Circuit(nl, abi)
    c.set_analyze_type(pe.DC)
    b = c.batch(2)
    assert b.analyze(), c.abi.last_error()
    assert b.last_kernel() != 3 and b.frontal_info() is None
    frontal_from(20000)
    nl, info = wl.series_parallel(1500, 100, seed=2)  # below the size threshold
    c = pe.Circuit(nl, abi)
    c.set_analyze_type(pe.DC)
    b = c.batch(2)
    assert b.analyze(), c.abi.last_error()
    assert b.last_kernel() != 3


@pytest.mark.gpu
def test_config_a_full_size(ref):
    abi = pe.product()  # noqa: F811
    nl, info = wl.series_parallel(100_000, 9_000, seed=1)
    c = pe.Circuit(nl, abi)
    c.set_analyze_type(pe.DC)
    b = c.batch(1)
    assert b.analyze(), c.abi.last_error()
    fi = b.frontal_info()
    assert b.last_kernel() == 3 and fi["unknowns"] > 90_000
    x = b.solution()[0]
    # (1) Kirchhoff's current law at every node, computed from the netlist alone: sum over the resistors at a node of
    # (v_node - v_other) / r = current injected by the source branch
    n_nodes = fi["unknowns"] - 1
    e, w, p = nl.arrays()
    res, rv = np.array(info["res"]), info["res_values"]
    pin_node = {}
    for ele in list(info["res"]) + [info["V"]]:
        for pin in (0, 1):
            pin_node[(ele, pin)] = c.pin_unknown(ele, pin)
    na = np.array([pin_node[(int(r), 0)] for r in res])
    nb = np.array([pin_node[(int(r), 1)] for r in res])
    va = np.where(na >= 0, x[np.maximum(na, 0)], 0.0)
    vb = np.where(nb >= 0, x[np.maximum(nb, 0)], 0.0)
    cur = (va - vb) / rv
    net = np.zeros(n_nodes)
    mag = np.zeros(n_nodes)
    np.add.at(net, na[na >= 0], cur[na >= 0])
    np.add.at(net, nb[nb >= 0], -cur[nb >= 0])
    np.add.at(mag, na[na >= 0], np.abs(cur[na >= 0]))
    np.add.at(mag, nb[nb >= 0], np.abs(cur[nb >= 0]))
    i_src = x[c.branch_unknown(info["V"])]
    vp, vq = pin_node[(info["V"], 0)], pin_node[(info["V"], 1)]
    net[vp] += i_src
    net[vq] -= i_src
    mag[vp] += abs(i_src)
    mag[vq] += abs(i_src)
    assert np.abs(x[vp] - x[vq] - 3.0) < 1e-12  # the source row
    # Resistances from U(1e-5, 1e5) Ohm: a node between micro-ohm resistors sees currents that are differences of nearly equal
    # voltages, so its OWN balance cannot be better than eps * |v| / r.  The reference's solution of this netlist gives: worst
    # node 1.4e-5 of the node's currents, normwise (worst imbalance over the largest current) 4.2e-10, median node 5.5e-12,
    # 99 % of the nodes below 1.5e-9 (measured with oracle/_ref, 126 s).  The bars below are what a backward-stable solve meets.
    rel = np.abs(net) / np.maximum(mag, 1e-300)
    assert float(np.abs(net).max() / mag.max()) <= 1e-9, float(np.abs(net).max() / mag.max())
    assert float(np.median(rel)) <= 1e-10 and float(np.quantile(rel, 0.99)) <= 1e-8, (float(np.median(rel)), float(np.quantile(rel, 0.99)), float(rel.max()))
    # (2) the reference's own solution of the same netlist (Eigen SparseLU on one host core: ~1.5 minutes).  Resistances
    # drawn from U(1e-5, 1e5) Ohm make the system ill-conditioned (kappa ~ 1e10): two backward-stable
    # solvers agree to ~kappa * eps relative to the LARGEST voltage, not component by component -- nodes a few micro-ohms from
    # ground sit at nanovolts.  The bar is therefore 1e-9 of the largest |x| for every unknown (measured 5.9e-10: 1.8e-9 V on
    # voltages of up to 3 V), and 8 digits component-wise for every unknown above 1e-3 of it.  Which of the two solutions is the
    # better one the Kirchhoff check above tells: normwise imbalance 8e-11 here, 4e-10 for the reference.
    rc = refapi.RefCircuit(nl, fast=True)
    rc.set_analyze_type(pe.DC)
    ok, n = rc.analyze_counted()
    assert ok
    xr = rc.solution().real
    scale = np.abs(xr).max()
    err = np.abs(x - xr)
    assert err.max() <= 1e-9 * scale, (float(err.max()), float(scale))
    big = np.abs(xr) >= 1e-3 * scale
    assert big.sum() > 1000
    rel_big = np.abs(x[big] - xr[big]) / np.abs(xr[big])
    assert float(rel_big.max()) <= 1e-8, float(rel_big.max())
