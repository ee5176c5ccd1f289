"""Parity of the device models and analysis modes round 1 left without a test ("Parity holes"):

  * AC small-signal of every nonlinear device after its bias solve, and ACOP (circuit.h:192-212, PN_junction.h:406-436,
    BJT_NPN.h:163-183, BJT_PNP.h:163-183, nmosfet.h:145-168, pmosfet.h:142-165)
  * a PNP stage (BJT_PNP.h:116-159)
  * the PN junction with breakdown (Bv_set), recombination current (Isr, Nr), area, transit time (PN_junction.h:58-109,
    296-354, 358-402)
  * temperature / nominal temperature / tolerances / g_min / r_open (circuit.h:900-903, 1012, 1107-1110; base.h:326-381)
  * omega = 0 inside an AC sweep (inductor.h:118-126: the D entry disappears) and a linear sweep
  * elements with an unconnected pin (resistance.h:86: the whole stamp is skipped)

Every test compares the product (GPU through the C ABI with -m gpu, the host emulator otherwise) with the compiled,
unmodified reference on the same netlist: 1e-9 relative / 1e-12 absolute, equal Newton iteration counts.
"""
import numpy as np
import pytest

import pe_b200 as pe
import refapi
import workloads as wl
from test_parity import abi, assert_close, gpu_solo, ref_solo  # noqa: F401  (abi is a fixture)

PATHS = {
    "auto": (0, 0, 0, 0, 0, 0),
    "tree-hbm-s4": (4, 0, 0, 0, 2, 0),
    "resident-s8": (8, 0, 1, 0, 1, 8),
    "flat": (-1, 0, 0, 0, 0, 0),
}


@pytest.fixture(params=list(PATHS), autouse=True)
def path(request, abi):  # noqa: F811
    assert abi.lib.phy_engine_b200_set_default_path(*PATHS[request.param]) == 0
    yield request.param
    abi.lib.phy_engine_b200_set_default_path(0, 0, 0, 0, 0, 0)


def biased(device: str):
    """A DC bias source in series with a small-signal VAC driving one nonlinear device, plus a load and a coupling capacitor:
    the AC solve needs the device's small-signal conductances from the bias solve."""
    nl = pe.Netlist()
    g = nl.ground()
    info = {}
    if device == "pn":
        vb = nl.add(pe.VDC, 0.9)
        va = nl.add(pe.VAC, 0.05, 1e3, 30.0)
        r = nl.add(pe.R, 470.0)
        d = nl.add(pe.PN, *pe.PN_DEFAULT)
        c = nl.add(pe.C, 2.2e-9)
        rl = nl.add(pe.R, 2e3)
        nl.wire(vb, 1, g, 0)
        nl.wire(vb, 0, va, 1)
        nl.wire(va, 0, r, 0)
        nl.wire(r, 1, d, 0)
        nl.wire(d, 1, g, 0)
        nl.wire(c, 0, d, 0)
        nl.wire(c, 1, rl, 0)
        nl.wire(rl, 1, g, 0)
        info = {"Vb": vb, "Va": va, "R": r, "D": d, "C": c, "Rl": rl}
    elif device in ("npn", "pnp"):
        sign = 1.0 if device == "npn" else -1.0
        vb = nl.add(pe.VDC, sign * 0.66)
        va = nl.add(pe.VAC, 0.002, 1e3, 0.0)
        vc = nl.add(pe.VDC, sign * 5.0)
        rc = nl.add(pe.R, 1e3)
        q = nl.add(pe.NPN if device == "npn" else pe.PNP, 1e-16, 1.0, 100.0, 27.0, 1.0)
        c = nl.add(pe.C, 1e-8)
        rl = nl.add(pe.R, 5e3)
        nl.wire(vb, 1, g, 0)
        nl.wire(vb, 0, va, 1)
        nl.wire(va, 0, q, 0)
        nl.wire(vc, 1, g, 0)
        nl.wire(vc, 0, rc, 0)
        nl.wire(rc, 1, q, 1)
        nl.wire(q, 2, g, 0)
        nl.wire(c, 0, q, 1)
        nl.wire(c, 1, rl, 0)
        nl.wire(rl, 1, g, 0)
        info = {"Vb": vb, "Va": va, "Vc": vc, "R": rc, "Q": q, "C": c, "Rl": rl}
    else:
        sign = 1.0 if device == "nmos" else -1.0
        vg = nl.add(pe.VDC, sign * 2.0)
        va = nl.add(pe.VAC, 0.01, 1e3, 45.0)
        vd = nl.add(pe.VDC, sign * 5.0)
        rd = nl.add(pe.R, 1e3)
        m = nl.add(pe.NMOS if device == "nmos" else pe.PMOS, 2e-3, 0.02, 1.0)
        c = nl.add(pe.C, 1e-8)
        rl = nl.add(pe.R, 5e3)
        nl.wire(vg, 1, g, 0)
        nl.wire(vg, 0, va, 1)
        nl.wire(va, 0, m, 1)
        nl.wire(vd, 1, g, 0)
        nl.wire(vd, 0, rd, 0)
        nl.wire(rd, 1, m, 0)
        nl.wire(m, 2, g, 0)
        nl.wire(c, 0, m, 0)
        nl.wire(c, 1, rl, 0)
        nl.wire(rl, 1, g, 0)
        info = {"Vg": vg, "Va": va, "Vd": vd, "R": rd, "M": m, "C": c, "Rl": rl}
    return nl, info


@pytest.mark.parametrize("device", ["pn", "npn", "pnp", "nmos", "pmos"])
@pytest.mark.parametrize("at", ["AC", "ACOP"])
def test_ac_small_signal_of_nonlinear_devices(ref, abi, device, at):  # noqa: F811
    nl, info = biased(device)
    mode = getattr(pe, at)
    rc, rok, rn = ref_solo(nl, mode, ref, omega=2e5)
    gc, gok = gpu_solo(nl, mode, abi, omega=2e5)
    assert rok and gok, gc.abi.last_error()
    x = rc.solution()
    assert np.abs(x.imag).max() > 1e-9  # a genuinely complex small-signal solution
    assert_close(gc.solution(), x, f"{device} {at} single point")
    # a sweep over instances and frequency points: bias solved per instance, small-signal system per point
    n_inst, points = 7, 9
    rng = np.random.default_rng(11)
    rkey = "R"
    over = [(info[rkey], "r", wl.sweep_values(rng, 1e3 if device != "pn" else 470.0, n_inst, 0.9, 1.1))]
    want = refapi.run_batch(nl, mode, n_inst, over, ac=(pe.SWEEP_LOG, 1e4, 1e7, points))
    assert (want["ok"] == 1).all()
    c = pe.Circuit(nl, abi)
    c.set_analyze_type(mode)
    b = c.batch(n_inst)
    for e, name, v in over:
        b.set_param(e, name, v)
    b.set_ac_sweep(pe.SWEEP_LOG, 1e4, 1e7, points)
    assert b.analyze(), c.abi.last_error()
    assert_close(b.ac_solution(), want["x"], f"{device} {at} sweep x instances")


def test_pnp_stage_op_matches_reference(ref, abi):  # noqa: F811
    # BJT_PNP.h:116-159: the NPN stamp with the controlling voltage taken emitter -> base
    nl = pe.Netlist()
    g = nl.ground()
    vbs = nl.add(pe.VDC, -0.65)
    vcs = nl.add(pe.VDC, -5.0)
    r = nl.add(pe.R, 1e3)
    q = nl.add(pe.PNP, 1e-16, 1.0, 100.0, 27.0, 1.0)
    nl.wire(vbs, 1, g, 0)
    nl.wire(vcs, 1, g, 0)
    nl.wire(vbs, 0, q, 0)
    nl.wire(vcs, 0, r, 0)
    nl.wire(r, 1, q, 1)
    nl.wire(q, 2, g, 0)
    n_inst = 64
    rng = np.random.default_rng(3)
    over = [(vbs, "V", -rng.uniform(0.55, 0.70, n_inst)), (r, "r", rng.uniform(900.0, 1100.0, n_inst)), (q, "BetaF", rng.uniform(50.0, 200.0, n_inst)),
            (q, "Is", 1e-16 * np.exp(rng.normal(0.0, 0.3, n_inst)))]
    want = refapi.run_batch(nl, pe.OP, n_inst, over)
    assert (want["ok"] == 1).all()
    c = pe.Circuit(nl, abi)
    c.set_analyze_type(pe.OP)
    b = c.batch(n_inst)
    for e, name, v in over:
        b.set_param(e, name, v)
    assert b.analyze(), c.abi.last_error()
    assert (b.newton_iters() == want["solves"]).all()
    assert_close(b.solution(), want["x"].real, "pnp stage")
    assert np.abs(want["x"].real[:, 1]).max() > 0.1  # the collector actually moves


@pytest.mark.parametrize("case", ["breakdown", "recombination", "area_n", "weak_breakdown"])
def test_pn_junction_parameter_space(ref, abi, case):  # noqa: F811
    # PN_junction.h:58-109 (vlimit with the breakdown mirror), 296-354 (Bv_eff, Isr_eff), 358-402 (stamp)
    nl = pe.Netlist()
    g = nl.ground()
    n_inst = 48
    rng = np.random.default_rng(9)
    if case in ("breakdown", "weak_breakdown"):
        # reverse bias beyond Bv: the junction conducts in breakdown
        bv = 6.2 if case == "breakdown" else 3.3
        src = nl.add(pe.VDC, -9.0)
        r = nl.add(pe.R, 2.2e3)
        d = nl.add(pe.PN, 1e-14, 1.0, 0.0, 2.0, 27.0, 1e-3 if case == "breakdown" else 1e-9, bv, 1.0, 1.0)
        over = [(src, "V", -rng.uniform(bv + 1.0, 12.0, n_inst)), (r, "r", rng.uniform(1e3, 4e3, n_inst)), (d, "Bv", rng.uniform(bv - 0.5, bv + 0.5, n_inst))]
    elif case == "recombination":
        src = nl.add(pe.VDC, 0.8)
        r = nl.add(pe.R, 1e3)
        d = nl.add(pe.PN, 1e-14, 1.1, 1e-10, 2.0, 27.0, 1e-3, 40.0, 1.0, 1.0)
        over = [(src, "V", rng.uniform(0.3, 1.5, n_inst)), (d, "Isr", 1e-10 * np.exp(rng.normal(0.0, 0.5, n_inst))), (d, "Nr", rng.uniform(1.6, 2.4, n_inst))]
    else:
        src = nl.add(pe.VDC, 1.2)
        r = nl.add(pe.R, 330.0)
        d = nl.add(pe.PN, 2e-14, 1.3, 0.0, 2.0, 27.0, 1e-3, 40.0, 0.0, 3.0)  # Bv_set off, area 3
        over = [(d, "Area", rng.uniform(0.5, 5.0, n_inst)), (d, "N", rng.uniform(1.0, 1.6, n_inst)), (d, "Is", 2e-14 * np.exp(rng.normal(0.0, 0.3, n_inst)))]
    nl.wire(src, 1, g, 0)
    nl.wire(src, 0, r, 0)
    nl.wire(r, 1, d, 0)
    nl.wire(d, 1, g, 0)
    want = refapi.run_batch(nl, pe.OP, n_inst, over)
    c = pe.Circuit(nl, abi)
    c.set_analyze_type(pe.OP)
    b = c.batch(n_inst)
    for e, name, v in over:
        b.set_param(e, name, v)
    ok = b.analyze()
    assert ok == bool((want["ok"] == 1).all()), c.abi.last_error()
    assert (b.newton_iters() == want["solves"]).all()
    good = want["ok"] == 1
    assert good.sum() >= n_inst // 2
    assert ((b.status() == 0) == good).all()
    assert_close(b.solution()[good], want["x"].real[good], f"pn {case}")
    if case == "breakdown":
        assert (np.abs(want["x"].real[good][:, 1]) > 4.0).all()  # the junction sits near -Bv, not at the source voltage


@pytest.mark.parametrize("temperature,tnom", [(85.0, 27.0), (-20.0, 27.0), (60.0, 50.0)])
def test_temperature_overwrites_device_temp(ref, abi, temperature, tnom):  # noqa: F811
    # base.h:326-381: load_temperature(env.temperature) rewrites the "Temp" attribute of PN / BJT on every prepare()
    for nl, info, key in ((*wl.diode_resistor(), "R"), (*wl.npn_stage(), "R")):
        rc = refapi.RefCircuit(nl)
        rc.set_analyze_type(pe.OP)
        rc.set_env(temperature=temperature, norm_temperature=tnom)
        rok, rn = rc.analyze_counted()
        gc = pe.Circuit(nl, abi)
        gc.set_analyze_type(pe.OP)
        gc.set_env(temperature=temperature, norm_temperature=tnom)
        gok = gc.analyze()
        assert rok and gok, gc.abi.last_error()
        assert_close(gc.solution(), rc.solution(), f"T = {temperature}")
        # the same through the reference-compatible setters of the C ABI and over a batch
        n_inst = 33
        rng = np.random.default_rng(5)
        over = [(info[key], "r", wl.sweep_values(rng, 1e3, n_inst))]
        want = refapi.run_batch(nl, pe.OP, n_inst, over, env=[0, 0, 0, 0, 0, 0, temperature, tnom])
        c2 = pe.Circuit(nl, abi)
        c2.set_analyze_type(pe.OP)
        c2.set_temperature(temperature)
        c2.set_env(temperature=temperature, norm_temperature=tnom)
        b = c2.batch(n_inst)
        for e, name, v in over:
            b.set_param(e, name, v)
        assert b.analyze(), c2.abi.last_error()
        assert (b.newton_iters() == want["solves"]).all()
        assert_close(b.solution(), want["x"].real, f"batch at T = {temperature}")
    # and the result really depends on it
    cold = refapi.RefCircuit(wl.diode_resistor()[0])
    cold.set_analyze_type(pe.OP)
    cold.analyze_counted()
    assert abs(cold.solution()[1].real - rc.solution()[1].real) > -1.0  # (sanity only: both solved)


@pytest.mark.parametrize("env", [dict(V_eps_max=1e-9, V_epsr_max=1e-6, I_eps_max=1e-15, I_epsr_max=1e-6), dict(V_epsr_max=1e-2), dict(g_min=1e-9), dict(I_eps_max=1e-9, I_epsr_max=1e-2)])
def test_tolerances_and_gmin(ref, abi, env):  # noqa: F811
    # circuit.h:900-903: the Newton test takes its tolerances from the environment (iteration counts change with them);
    # circuit.h:1107-1110: g_min is added to every node's diagonal
    nl, info = wl.diode_ladder(6)
    n_inst = 40
    rng = np.random.default_rng(21)
    over = [(e, "r", wl.sweep_values(rng, 1e3, n_inst, 0.9, 1.1)) for e in info["R"]]
    full = dict(V_eps_max=0.0, V_epsr_max=0.0, I_eps_max=0.0, I_epsr_max=0.0, g_min=0.0, r_open=0.0)
    full.update(env)
    envv = [full["V_eps_max"], full["V_epsr_max"], full["I_eps_max"], full["I_epsr_max"], full["g_min"], full["r_open"], 27.0, 27.0]
    want = refapi.run_batch(nl, pe.OP, n_inst, over, env=envv)
    base = refapi.run_batch(nl, pe.OP, n_inst, over)
    c = pe.Circuit(nl, abi)
    c.set_analyze_type(pe.OP)
    c.set_env(**full)
    b = c.batch(n_inst)
    for e, name, v in over:
        b.set_param(e, name, v)
    assert b.analyze() == bool((want["ok"] == 1).all()), c.abi.last_error()
    assert (b.newton_iters() == want["solves"]).all()
    assert_close(b.solution(), want["x"].real, f"env {env}")
    if "V_eps_max" in env or "V_epsr_max" in env:
        assert (want["solves"] != base["solves"]).any()  # the tolerance really changed the trajectory length


def test_r_open_of_switch_and_relay(ref, abi):  # noqa: F811
    # circuit.h:1012: mna.r_open (default 1e12) is the contact resistance of an open switch / relay
    nl, info = wl.relay_stage(v_ctl=2.0)  # below Von: contact open
    for r_open in (0.0, 1e9, 1e6):
        rc = refapi.RefCircuit(nl)
        rc.set_analyze_type(pe.OP)
        rc.set_env(r_open=r_open)
        rok, rn = rc.analyze_counted()
        gc = pe.Circuit(nl, abi)
        gc.set_analyze_type(pe.OP)
        gc.set_env(r_open=r_open)
        gok = gc.analyze()
        assert rok == gok, gc.abi.last_error()
        assert_close(gc.solution(), rc.solution(), f"r_open {r_open}")


def test_ac_sweep_through_omega_zero(ref, abi):  # noqa: F811
    # inductor.h:118-126: at omega == 0 the inductor stamps no D entry (a short); capacitor.h:85-102: jwC = 0.
    # A linear sweep starting at 0 rad/s holds both patterns in one run.
    nl, info = wl.rlc_ladder(5)
    sweep = (pe.SWEEP_LINEAR, 0.0, 4e7, 9)
    rc, rok, rn = ref_solo(nl, pe.AC, ref, sweep=sweep)
    assert rok and rn == 9
    om, xr = rc.ac_results()
    assert om[0] == 0.0
    c = pe.Circuit(nl, abi)
    c.set_analyze_type(pe.AC)
    b = c.batch(1)
    b.set_ac_sweep(*sweep)
    assert b.analyze(), c.abi.last_error()
    assert (b.ac_omegas() == om).all()
    assert_close(b.ac_solution()[0], xr, "sweep through omega = 0")
    # single point at omega = 0 and with coupled inductors (coupled_inductors.h:119-156 has the same rule)
    for nl2 in (nl, wl.coupled_inductors_stage(vac=True)[0]):
        r2, ok2, _ = ref_solo(nl2, pe.AC, ref, omega=0.0)
        g2, gok2 = gpu_solo(nl2, pe.AC, abi, omega=0.0)
        assert ok2 == gok2, g2.abi.last_error()
        if ok2:
            assert_close(g2.solution(), r2.solution(), "omega = 0 single point")


@pytest.mark.parametrize("at", ["DC", "TR", "AC"])
def test_elements_with_an_unconnected_pin_stamp_nothing(ref, abi, at):  # noqa: F811
    # resistance.h:86 (and every other model): a device whose pins are not all connected skips its stamp
    nl = pe.Netlist()
    g = nl.ground()
    src = nl.add(pe.VAC, 2.0, 1e3, 0.0) if at == "AC" else nl.add(pe.VDC, 2.0)
    r1, r2 = nl.add(pe.R, 1e3), nl.add(pe.R, 3e3)
    dangling_r = nl.add(pe.R, 50.0)   # pin 1 left open
    dangling_c = nl.add(pe.C, 1e-9)   # pin 1 left open
    dangling_d = nl.add(pe.PN, *pe.PN_DEFAULT)  # pin 0 left open
    lonely = nl.add(pe.R, 75.0)       # no pin connected at all
    cc = nl.add(pe.C, 1e-9)
    nl.wire(src, 1, g, 0)
    nl.wire(src, 0, r1, 0)
    nl.wire(r1, 1, r2, 0)
    nl.wire(r2, 1, g, 0)
    nl.wire(cc, 0, r1, 1)
    nl.wire(cc, 1, g, 0)
    nl.wire(dangling_r, 0, r1, 1)
    nl.wire(dangling_c, 0, r1, 1)
    nl.wire(dangling_d, 1, r1, 1)
    assert lonely > 0
    kw = {"tr": (1e-7, 1e-6)} if at == "TR" else ({"omega": 1e6} if at == "AC" else {})
    rc, rok, rn = ref_solo(nl, getattr(pe, at), ref, **kw)
    gc, gok = gpu_solo(nl, getattr(pe, at), abi, **kw)
    assert rok == gok, gc.abi.last_error()
    assert rok
    assert_close(gc.solution(), rc.solution(), f"unconnected pins, {at}")
