"""Seeded golden cases shared by tests/golden/make_golden.py (which asks the compiled reference for the answers) and
tests/test_golden.py (which replays them on the product).  Inputs are regenerated from the seeds, only the reference's
outputs are stored in tests/golden/cases.npz."""
import numpy as np

import pe_b200 as pe
import workloads as wl

CASES = {
    # config B shape: RC ladder transient, element-wise R_i / C_i sweep
    "rc_ladder_tr_24x37": {"at": pe.TR, "n_inst": 37},
    "rc_ladder_tr_300x8": {"at": pe.TR, "n_inst": 8},
    # config C: Newton operating points
    "diode_op_mc": {"at": pe.OP, "n_inst": 130},
    "diode_ladder16_op": {"at": pe.OP, "n_inst": 33},
    "npn_stage_op": {"at": pe.OP, "n_inst": 40},
    "cmos_stage_op": {"at": pe.OP, "n_inst": 40},
    "npn_resistor_biased_fails": {"at": pe.OP, "n_inst": 3},
    # config D: RLC ladder AC sweep (instances x frequency points)
    "rlc_ladder_ac_16": {"at": pe.AC, "n_inst": 3},
    # nonlinear transient with the diffusion capacitance companion
    "diode_tr_tt": {"at": pe.TR, "n_inst": 9},
    # full bridge rectifier (element 54)
    "bridge_rectifier_op": {"at": pe.OP, "n_inst": 24},
    # the rest of SURVEY 8(a) row a4 and 8(f) row 4: transformers, coupled inductors, relay, generators
    "transformer_tr": {"at": pe.TR, "n_inst": 12},
    "center_tap_dc": {"at": pe.DC, "n_inst": 12},
    "coupled_inductors_tr": {"at": pe.TR, "n_inst": 12},
    "coupled_inductors_ac": {"at": pe.AC, "n_inst": 4},
    "relay_tr": {"at": pe.TR, "n_inst": 16},
    "pulse_generator_tr": {"at": pe.TR, "n_inst": 10},
    "triangle_generator_tr": {"at": pe.TR, "n_inst": 10},
}


def build(name):
    """-> (netlist, overrides [(element, attribute, values[n_inst])], run kwargs)"""
    n = CASES[name]["n_inst"]
    if name.startswith("rc_ladder_tr_"):
        sections = int(name.split("_")[-1].split("x")[0])
        nl, info = wl.rc_ladder(sections)
        rng = np.random.default_rng(100 + sections)
        over = [(e, "r", wl.sweep_values(rng, 1e3, n)) for e in info["R"]] + [(e, "c", wl.sweep_values(rng, 1e-9, n)) for e in info["C"]]
        return nl, over, {"t_step": 1e-8, "t_stop": 3e-7}
    if name == "diode_op_mc":
        nl, info = wl.diode_resistor(n_diodes=2, v=3.0)
        rng = np.random.default_rng(11)
        over = [(info["R"], "r", wl.sweep_values(rng, 1e3, n, 0.95, 1.05))]
        for d in info["D"]:
            over.append((d, "Is", 1e-14 * np.exp(0.3 * rng.standard_normal(n))))
            over.append((d, "N", rng.uniform(1.0, 1.2, n)))
        return nl, over, {}
    if name == "diode_ladder16_op":
        nl, info = wl.diode_ladder(16)
        rng = np.random.default_rng(5)
        return nl, [(e, "r", wl.sweep_values(rng, 1e3, n, 0.95, 1.05)) for e in info["R"]], {}
    if name == "npn_stage_op":
        nl, info = wl.npn_stage()
        rng = np.random.default_rng(21)
        return nl, [(info["Vb"], "V", rng.uniform(0.55, 0.70, n)), (info["R"], "r", rng.uniform(900.0, 1100.0, n)),
                    (info["Q"], "BetaF", rng.uniform(50.0, 200.0, n))], {}
    if name == "cmos_stage_op":
        nl, info = wl.cmos_stage()
        rng = np.random.default_rng(22)
        return nl, [(info["Vg"], "V", rng.uniform(1.2, 3.0, n)), (info["R"], "r", rng.uniform(900.0, 1100.0, n))], {}
    if name == "npn_resistor_biased_fails":
        nl, info = wl.npn_resistor_biased()
        return nl, [(info["Rc"], "r", np.array([900.0, 1000.0, 1100.0]))], {}
    if name == "rlc_ladder_ac_16":
        nl, info = wl.rlc_ladder(16)
        rng = np.random.default_rng(2)
        over = [(e, "r", wl.sweep_values(rng, 10.0, n)) for e in info["R"]] + [(e, "L", wl.sweep_values(rng, 1e-6, n)) for e in info["L"]]
        return nl, over, {"ac": (pe.SWEEP_LOG, 1e4, 1e9, 25)}
    if name == "diode_tr_tt":
        nl, info = wl.diode_resistor(v=2.0)
        rng = np.random.default_rng(9)
        return nl, [(info["D"][0], "tt", rng.uniform(1e-9, 1e-8, n)), (info["R"], "r", wl.sweep_values(rng, 1e3, n))], {"t_step": 1e-9, "t_stop": 2e-8}
    if name == "bridge_rectifier_op":
        nl, info = wl.bridge_rectifier()
        rng = np.random.default_rng(17)
        return nl, [(info["V"], "V", rng.uniform(-8.0, 8.0, n)), (info["R"], "r", rng.uniform(200.0, 5000.0, n))], {}
    if name == "transformer_tr":
        nl, info = wl.transformer_stage(vac=True)
        rng = np.random.default_rng(61)
        return nl, [(info["TX"], "n", rng.uniform(0.5, 8.0, n)), (info["R"], "r", rng.uniform(500.0, 2000.0, n))], {"t_step": 1e-7, "t_stop": 3e-6}
    if name == "center_tap_dc":
        nl, info = wl.center_tap_stage(vac=False)
        rng = np.random.default_rng(62)
        return nl, [(info["TX"], "n_total", rng.uniform(0.5, 6.0, n)), (info["R2"], "r", rng.uniform(500.0, 4000.0, n))], {}
    if name in ("coupled_inductors_tr", "coupled_inductors_ac"):
        nl, info = wl.coupled_inductors_stage(vac=True)
        rng = np.random.default_rng(63)
        over = [(info["K"], "L1", rng.uniform(5e-4, 2e-3, n)), (info["K"], "L2", rng.uniform(2e-4, 8e-4, n)), (info["K"], "k", rng.uniform(0.5, 0.99, n))]
        return nl, over, ({"t_step": 1e-7, "t_stop": 4e-6} if name.endswith("_tr") else {"ac": (pe.SWEEP_LOG, 1e4, 1e8, 9)})
    if name == "relay_tr":
        nl, info = wl.relay_stage(vac=True)
        rng = np.random.default_rng(64)
        return nl, [(info["Vctl"], "Vp", rng.uniform(4.0, 9.0, n)), (info["R"], "r", rng.uniform(200.0, 900.0, n))], {"t_step": 1e-7, "t_stop": 8e-6}
    if name in ("pulse_generator_tr", "triangle_generator_tr"):
        nl, info = wl.generator_rc("pulse" if name.startswith("pulse") else "triangle")
        rng = np.random.default_rng(65)
        return nl, [(info["G"], "freq", rng.uniform(1.5e5, 4e5, n)), (info["G"], "phase", rng.uniform(0.0, 6.0, n)), (info["R"], "r", rng.uniform(500.0, 2000.0, n))], {"t_step": 5e-8, "t_stop": 1.2e-5}
    raise KeyError(name)
