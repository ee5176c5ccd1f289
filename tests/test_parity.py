"""Parity: the solve path, called through the C ABI, against the compiled reference on the same netlists.

backend "gpu" (-m gpu, on the B200): libphyengine_b200.so, i.e. the CUDA kernels.  These are the parity tests proper.
backend "emu" (CPU suite): tests/emu/libpe_emu.so = the same host side with the kernel's interpreter replayed on the host
(test infrastructure; checks the symbolic phase and the warp-stream schedules, including a race detector).

Tolerance (BASELINE.json north_star): node voltages and branch currents within 1e-9 relative / 1e-12 absolute in
FP64, identical Newton iteration counts (solve_once-equivalents per analyze()).
"""
import numpy as np
import pytest

import pe_b200 as pe
import refapi
import workloads as wl

RTOL, ATOL = 1e-9, 1e-12


@pytest.fixture(params=[pytest.param("gpu", marks=pytest.mark.gpu), "emu"])
def abi(request):
    if request.param == "gpu":
        return pe.product()
    import emuapi

    return emuapi.emulator()


# Every test runs once per solve path (set as the process-wide default every new batch starts from):
#   auto            the product default: shared-memory resident kernel for small circuits, tree-streaming (HBM
#                   workspace, one warp per sub-tree) otherwise
#   tree-hbm-s4     tree-streaming kernel forced, 4 sub-tree warps per 32 lanes
#   tree-hbm-s16    ... 16 sub-tree warps, L2 operand prefetch on
#   tree-hbm-s8j2   ... 8 sub-tree warps, 64 lanes per CTA (two per thread), fused elimination steps
#   tree-hbm-s4j4   ... 4 sub-tree warps, 128 lanes per CTA (four per thread: 1 KB workspace rows)
#   tree-hbm-s32j4  ... 32 sub-tree warps split over a thread-block cluster of two CTAs (two SMs per 128-lane group)
#   resident-s8     shared-memory kernel forced, 8 word streams per instance (4 instances per warp), fused elimination steps
#   resident-s32j2  ... 32 streams, 2 instances per CTA handled by the same thread (vector loads)
#   flat            flat HBM-streaming kernel (the first-generation path), one warp per 32 instances
#   flat-g4         ... 4 sub-tree warps per 32 instances
#   stream          the stream kernel required wherever it applies (linear circuits of more than 64 unknowns: one warp per lane
#                   group, one word stream, generated TMA-fed tiles); everything else takes the default path
PATHS = {
    "auto": (0, 0, 0, 0, 0, 0),
    "tree-hbm-s4": (4, 0, 0, 0, 2, 0),
    "tree-hbm-s16": (16, 0, 0, 0, 2, 1),
    "tree-hbm-s8j2": (8, 0, 2, 0, 2, 8),
    "tree-hbm-s4j4": (4, 0, 4, 0, 2, 0),
    "tree-hbm-s32j4": (32, 0, 4, 0, 2, 0),
    "resident-s8": (8, 0, 1, 0, 1, 8),
    "resident-s32j2": (32, 2, 2, 0, 1, 0),
    "flat": (-1, 0, 0, 0, 0, 0),
    "flat-g4": (-1, 0, 0, 4, 0, 0),
    "stream": (0, 0, 0, 0, 0, 64),
}


# tests whose circuits the stream kernel applies to (elsewhere the "stream" path is the "auto" path again)
STREAM_TESTS = {"test_rc_ladder_batch_sweep", "test_tr_resume_continues_like_reference", "test_random_links_dc", "test_linear_zoo_every_linear_element"}


@pytest.fixture(params=list(PATHS), autouse=True)
def path(request, abi):
    if request.param == "stream" and request.node.originalname not in STREAM_TESTS:
        pytest.skip("the stream kernel does not apply to this circuit")
    rc = abi.lib.phy_engine_b200_set_default_path(*PATHS[request.param])
    assert rc == 0
    yield request.param
    abi.lib.phy_engine_b200_set_default_path(0, 0, 0, 0, 0, 0)


def assert_close(got, want, what=""):
    got = np.asarray(got)
    want = np.asarray(want)
    err = np.abs(got - want)
    tol = ATOL + RTOL * np.maximum(np.abs(got), np.abs(want))
    bad = err > tol
    if bad.any():
        i = np.unravel_index(np.argmax(err - tol), err.shape)
        raise AssertionError(f"{what}: {bad.sum()} of {bad.size} values differ; worst at {i}: got {got[i]!r} want {want[i]!r} (err {err[i]:.3e}, tol {tol[i]:.3e})")


def ref_solo(nl, at, ref, **kw):
    c = refapi.RefCircuit(nl)
    c.set_analyze_type(at)
    if "tr" in kw:
        c.set_tr(*kw["tr"])
    if "omega" in kw:
        c.set_ac_omega(kw["omega"])
    if "sweep" in kw:
        c.set_ac_sweep(*kw["sweep"])
    ok, n = c.analyze_counted()
    return c, ok, n


def gpu_solo(nl, at, abi, **kw):
    c = pe.Circuit(nl, abi)
    c.set_analyze_type(at)
    if "tr" in kw:
        c.set_tr(*kw["tr"])
    if "omega" in kw:
        c.set_ac_omega(kw["omega"])
    if "sweep" in kw:
        c.set_ac_sweep(*kw["sweep"])
    ok = c.analyze()
    return c, ok


def test_rc_step_tr_matches_reference(ref, abi):
    # test/0005.models/rc_step_tr.cpp: 1 V, 1 kOhm, 1 nF, dt 1e-8, 100 steps; oracle vout = 0.63027499952138644
    nl, info = wl.rc_ladder(1)
    rc, rok, rn = ref_solo(nl, pe.TR, ref, tr=(1e-8, 1e-6))
    gc, gok = gpu_solo(nl, pe.TR, abi, tr=(1e-8, 1e-6))
    assert rok and gok, gc.abi.last_error()
    assert rn == 100
    assert abs(rc.solution()[1].real - 0.63027499952138644) < 1e-15
    assert_close(gc.solution(), rc.solution(), "rc step")
    # sampled through the reference-compatible circuit_sample_u8
    gv, gvo, gi, gco = gc.sample()
    rv, rvo, ri, rco = rc.sample()
    assert (gvo == rvo).all() and (gco == rco).all()
    assert_close(gv, rv, "sample voltages")
    assert_close(gi, ri, "sample currents")


def test_tr_resume_continues_like_reference(ref, abi):
    nl, _ = wl.rc_ladder(3)
    rc, _, _ = ref_solo(nl, pe.TR, ref, tr=(1e-8, 3e-7))
    gc, gok = gpu_solo(nl, pe.TR, abi, tr=(1e-8, 3e-7))
    assert gok
    assert_close(gc.solution(), rc.solution(), "first analyze")
    for _ in range(2):  # analyze() again continues the transient (circuit.h:242)
        assert rc.analyze() and gc.analyze()
        assert_close(gc.solution(), rc.solution(), "resumed analyze")


@pytest.mark.parametrize("n_sections,n_inst", [(8, 33), (50, 64), (200, 40)])
def test_rc_ladder_batch_sweep(ref, abi, path, n_sections, n_inst):
    nl, info = wl.rc_ladder(n_sections)
    rng = np.random.default_rng(7)
    over = []
    for e in info["R"]:
        over.append((e, "r", wl.sweep_values(rng, 1e3, n_inst)))
    for e in info["C"]:
        over.append((e, "c", wl.sweep_values(rng, 1e-9, n_inst)))
    want = refapi.run_batch(nl, pe.TR, n_inst, over, t_step=1e-8, t_stop=2e-7)
    c = pe.Circuit(nl, abi)
    c.set_analyze_type(pe.TR)
    c.set_tr(1e-8, 2e-7)
    b = c.batch(n_inst)
    for e, name, v in over:
        b.set_param(e, name, v)
    assert b.analyze(), c.abi.last_error()
    if path == "flat-g4" and n_sections >= 200:
        assert b.program_info(pe.MODE_TR)["warps"] == 4
    if path.startswith("tree-hbm"):
        ri = b.resident_info(pe.MODE_TR)
        assert ri["resident"] == 1 and ri["hbm"] == 1 and ri["last_I"] == 32 * ri["last_J"]
    if path.startswith("resident"):
        ri = b.resident_info(pe.MODE_TR)
        assert ri["resident"] == 1 and ri["last_S"] == ri["streams"] and ri["last_I"] > 0
        if path == "resident-s32j2":
            assert (ri["last_S"], ri["last_I"], ri["last_J"]) == (32, 2, 2)
    assert (want["ok"] == 1).all()
    assert_close(b.solution(), want["x"].real, "ladder state")
    assert b.total_solves == int(want["solves"].sum())
    assert (b.status() == 0).all()


def test_diode_op_iteration_trajectory(ref, abi):
    # test/0011.nonlinear/op_pn_junction.cpp; SURVEY.md 8(c): 7 iterations, Vd = 0.62944165509863270
    nl, info = wl.diode_resistor()
    rc, rok, rn = ref_solo(nl, pe.OP, ref)
    gc, gok = gpu_solo(nl, pe.OP, abi)
    assert rok and gok
    assert rn == 7
    assert abs(rc.solution()[1].real - 0.62944165509863270) < 1e-15
    b = gc.batch(1)
    assert b.analyze()
    assert int(b.newton_iters()[0]) == 7
    assert_close(gc.solution(), rc.solution(), "diode op")
    assert_close(b.solution()[0], rc.solution().real, "diode op (batch of 1)")


def test_diode_monte_carlo(ref, abi):
    n_inst = 257
    nl, info = wl.diode_resistor(n_diodes=2, v=3.0)
    rng = np.random.default_rng(11)
    over = [(info["R"], "r", wl.sweep_values(rng, 1e3, n_inst, 0.95, 1.05))]
    for d in info["D"]:
        over.append((d, "Is", 1e-14 * np.exp(0.3 * rng.standard_normal(n_inst))))
        over.append((d, "N", rng.uniform(1.0, 1.2, n_inst)))
    want = refapi.run_batch(nl, pe.OP, n_inst, over)
    c = pe.Circuit(nl, abi)
    c.set_analyze_type(pe.OP)
    b = c.batch(n_inst)
    for e, name, v in over:
        b.set_param(e, name, v)
    assert b.analyze(), c.abi.last_error()
    assert (b.newton_iters() == want["solves"]).all()
    assert_close(b.solution(), want["x"].real, "diode MC")


def test_diode_ladder_newton(ref, abi):
    n_inst = 64
    nl, info = wl.diode_ladder(16)
    rng = np.random.default_rng(5)
    over = [(e, "r", wl.sweep_values(rng, 1e3, n_inst, 0.95, 1.05)) for e in info["R"]]
    want = refapi.run_batch(nl, pe.OP, n_inst, over)
    c = pe.Circuit(nl, abi)
    c.set_analyze_type(pe.OP)
    b = c.batch(n_inst)
    for e, name, v in over:
        b.set_param(e, name, v)
    assert b.analyze(), c.abi.last_error()
    assert (b.newton_iters() == want["solves"]).all()
    assert_close(b.solution(), want["x"].real, "diode ladder")


def test_bjt_and_mos_op(ref, abi):
    for nl, info, sweeps in (
        (*wl.npn_stage(), [("Vb", "V", 0.55, 0.70), ("R", "r", 900.0, 1100.0)]),
        (*wl.cmos_stage(), [("Vg", "V", 1.2, 3.0), ("R", "r", 900.0, 1100.0)]),
        (*wl.cmos_stage(with_pmos=False), [("Vg", "V", 0.5, 3.0)]),
    ):
        n_inst = 96
        rng = np.random.default_rng(3)
        over = [(info[k], name, rng.uniform(lo, hi, n_inst)) for k, name, lo, hi in sweeps]
        want = refapi.run_batch(nl, pe.OP, n_inst, over)
        c = pe.Circuit(nl, abi)
        c.set_analyze_type(pe.OP)
        b = c.batch(n_inst)
        for e, name, v in over:
            b.set_param(e, name, v)
        ok = b.analyze()
        assert ok == bool((want["ok"] == 1).all())
        assert (b.newton_iters() == want["solves"]).all()
        good = want["ok"] == 1
        assert_close(b.solution()[good], want["x"].real[good], "transistor stage")


def test_failure_parity_resistor_biased_npn(ref, abi):
    # the reference itself does not converge here (64 iterations, analyze() false)
    nl, info = wl.npn_resistor_biased()
    rc, rok, rn = ref_solo(nl, pe.OP, ref)
    assert not rok and rn == 64
    c = pe.Circuit(nl, abi)
    c.set_analyze_type(pe.OP)
    b = c.batch(3)
    assert not b.analyze()
    assert (b.status() != 0).all()
    assert (b.newton_iters() == 64).all()


def test_ac_single_point(ref, abi):
    # test/0012.ac/ac_omega.cpp shape: R-C low pass at the corner frequency -> (0.5, -0.5)
    nl = pe.Netlist()
    g = nl.ground()
    s = nl.add(pe.VAC, 1.0, 50.0, 0.0)
    r = nl.add(pe.R, 1e3)
    cc = nl.add(pe.C, 1e-6)
    nl.wire(s, 1, g, 0)
    nl.wire(s, 0, r, 0)
    nl.wire(r, 1, cc, 0)
    nl.wire(cc, 1, g, 0)
    rc, rok, _ = ref_solo(nl, pe.AC, ref, omega=1e3)
    gc, gok = gpu_solo(nl, pe.AC, abi, omega=1e3)
    assert rok and gok
    assert_close(gc.solution(), rc.solution(), "ac point")
    assert abs(rc.solution()[1] - (0.5 - 0.5j)) < 1e-12


@pytest.mark.parametrize("n_sections,points", [(2, 17), (8, 200), (64, 64)])
def test_ac_log_sweep_rlc(ref, abi, path, n_sections, points):
    nl, info = wl.rlc_ladder(n_sections)
    sweep = (pe.SWEEP_LOG, 1e3, 1e10, points)
    rc, rok, rn = ref_solo(nl, pe.AC, ref, sweep=sweep)
    assert rok and rn == points
    om, xr = rc.ac_results()
    c = pe.Circuit(nl, abi)
    c.set_analyze_type(pe.AC)
    b = c.batch(1)
    b.set_ac_sweep(*sweep)
    assert b.analyze(), c.abi.last_error()
    assert (b.ac_omegas() == om).all()  # cumulative-product omegas are bit-identical
    assert_close(b.ac_solution()[0], xr, "ac sweep")
    assert b.total_solves == points


def test_ac_sweep_with_instances(ref, abi):
    n_inst, points = 5, 40
    nl, info = wl.rlc_ladder(4)
    rng = np.random.default_rng(2)
    over = [(e, "r", wl.sweep_values(rng, 10.0, n_inst)) for e in info["R"]] + [(e, "L", wl.sweep_values(rng, 1e-6, n_inst)) for e in info["L"]]
    want = refapi.run_batch(nl, pe.AC, n_inst, over, ac=(pe.SWEEP_LINEAR, 1e5, 1e8, points))
    c = pe.Circuit(nl, abi)
    c.set_analyze_type(pe.AC)
    b = c.batch(n_inst)
    for e, name, v in over:
        b.set_param(e, name, v)
    b.set_ac_sweep(pe.SWEEP_LINEAR, 1e5, 1e8, points)
    assert b.analyze(), c.abi.last_error()
    assert_close(b.ac_solution(), want["x"], "ac sweep x instances")


def test_random_links_dc(ref, abi):
    # test/0013.cuda/cuda_random_links_correctness.cu shape (chain + random 1 kOhm chords), scaled down
    nl, info = wl.random_links(256, 64, seed=3)
    rc, rok, _ = ref_solo(nl, pe.DC, ref)
    gc, gok = gpu_solo(nl, pe.DC, abi)
    assert rok and gok
    assert_close(gc.solution(), rc.solution(), "random links")


def test_diode_tr_with_transit_time(ref, abi):
    nl, info = wl.diode_resistor(v=2.0)
    n_inst = 16
    rng = np.random.default_rng(9)
    over = [(info["D"][0], "tt", rng.uniform(1e-9, 1e-8, n_inst)), (info["R"], "r", wl.sweep_values(rng, 1e3, n_inst))]
    want = refapi.run_batch(nl, pe.TR, n_inst, over, t_step=1e-9, t_stop=2e-8)
    c = pe.Circuit(nl, abi)
    c.set_analyze_type(pe.TR)
    c.set_tr(1e-9, 2e-8)
    b = c.batch(n_inst)
    for e, name, v in over:
        b.set_param(e, name, v)
    assert b.analyze(), c.abi.last_error()
    assert (b.newton_iters() == want["solves"]).all()
    assert_close(b.solution(), want["x"].real, "diode TR")


@pytest.mark.gpu
def test_no_cpu_fallback_and_native_launches():
    n0 = pe.launch_count()
    nl, _ = wl.rc_ladder(2)
    c, ok = gpu_solo(nl, pe.DC, pe.product())
    assert ok
    assert pe.launch_count() > n0  # the answer came from our kernels


def test_emulator_saw_no_races_or_unbalanced_barriers():
    """runs last in this file: the warp streams of every program replayed above were race-free between barriers"""
    import emuapi

    assert emuapi.races() == 0
    assert emuapi.unbalanced() == 0


def test_full_bridge_rectifier_op_and_tr(ref, abi):
    # element 54 = 4 default PN junctions bound to A, B, +, - (full_bridge_rectifier.h:28-90)
    nl, info = wl.bridge_rectifier()
    n_inst = 40
    rng = np.random.default_rng(17)
    over = [(info["V"], "V", rng.uniform(-8.0, 8.0, n_inst)), (info["R"], "r", rng.uniform(200.0, 5000.0, n_inst))]
    for at, kw in ((pe.OP, {}), (pe.TR, {"t_step": 1e-6, "t_stop": 5e-6})):
        want = refapi.run_batch(nl, at, n_inst, over, **kw)
        c = pe.Circuit(nl, abi)
        c.set_analyze_type(at)
        if kw:
            c.set_tr(kw["t_step"], kw["t_stop"])
        b = c.batch(n_inst)
        for e, name, v in over:
            b.set_param(e, name, v)
        ok = b.analyze()
        assert ok == bool((want["ok"] == 1).all())
        good = want["ok"] == 1
        assert good.sum() >= n_inst // 2
        assert (b.newton_iters() == want["solves"]).all()
        assert_close(b.solution()[good], want["x"].real[good], "bridge rectifier")


@pytest.mark.parametrize("at", ["DC", "TR", "TROP", "AC"])
def test_linear_zoo_every_linear_element(ref, abi, at):
    # R, C, L, VDC/VAC, IDC, IAC, VCCS, VCVS, CCCS, CCVS, op-amp, switch in one netlist
    n_inst = 19
    nl, info = wl.linear_zoo(vac=at in ("TR", "TROP", "AC"))
    rng = np.random.default_rng(23)
    over = [(info["R"][0], "r", rng.uniform(500.0, 2000.0, n_inst)), (info["G"], "G", rng.uniform(5e-4, 2e-3, n_inst)), (info["E"], "Mu", rng.uniform(1.0, 3.0, n_inst)),
            (info["F"], "alpha", rng.uniform(1.0, 5.0, n_inst)), (info["H"], "r", rng.uniform(100.0, 400.0, n_inst)), (info["L"], "L", rng.uniform(5e-4, 2e-3, n_inst)),
            (info["C"], "C", rng.uniform(5e-8, 2e-7, n_inst)), (info["IDC"], "I", rng.uniform(5e-4, 2e-3, n_inst))]
    code = {"DC": pe.DC, "TR": pe.TR, "TROP": pe.TROP, "AC": pe.AC}[at]
    kw = {"t_step": 2e-7, "t_stop": 4e-6} if at in ("TR", "TROP") else ({"ac": (pe.SWEEP_LOG, 1e4, 1e7, 13)} if at == "AC" else {})
    want = refapi.run_batch(nl, code, n_inst, over, **kw)
    assert (want["ok"] == 1).all()
    c = pe.Circuit(nl, abi)
    c.set_analyze_type(code)
    if at in ("TR", "TROP"):
        c.set_tr(kw["t_step"], kw["t_stop"])
    b = c.batch(n_inst)
    for e, name, v in over:
        b.set_param(e, name, v)
    if at == "AC":
        b.set_ac_sweep(*kw["ac"])
    assert b.analyze(), c.abi.last_error()
    if at == "AC":
        assert_close(b.ac_solution(), want["x"], "linear zoo AC")
    else:
        assert_close(b.solution(), want["x"].real, "linear zoo " + at)
        assert b.total_solves == int(want["solves"].sum())


@pytest.mark.parametrize("at", ["DC", "TR", "AC"])
def test_ideal_transformer(ref, abi, at):
    # transformer.h:66-99 (two branch rows: Vp - n Vs = 0, Is + n Ip = 0), turns ratio swept per instance
    n_inst = 21
    nl, info = wl.transformer_stage(vac=at != "DC")
    rng = np.random.default_rng(31)
    over = [(info["TX"], "n", rng.uniform(0.5, 8.0, n_inst)), (info["R"], "r", rng.uniform(500.0, 2000.0, n_inst)), (info["Rs"], "r", rng.uniform(20.0, 100.0, n_inst))]
    code = {"DC": pe.DC, "TR": pe.TR, "AC": pe.AC}[at]
    kw = {"t_step": 1e-7, "t_stop": 3e-6} if at == "TR" else ({"ac": (pe.SWEEP_LOG, 1e4, 1e8, 11)} if at == "AC" else {})
    want = refapi.run_batch(nl, code, n_inst, over, **kw)
    assert (want["ok"] == 1).all()
    c = pe.Circuit(nl, abi)
    c.set_analyze_type(code)
    if at == "TR":
        c.set_tr(kw["t_step"], kw["t_stop"])
    b = c.batch(n_inst)
    for e, name, v in over:
        b.set_param(e, name, v)
    if at == "AC":
        b.set_ac_sweep(*kw["ac"])
    assert b.analyze(), c.abi.last_error()
    if at == "AC":
        assert_close(b.ac_solution(), want["x"], "transformer AC")
    else:
        x = b.solution()
        assert_close(x, want["x"].real, "transformer " + at)
        assert b.total_solves == int(want["solves"].sum())
        if at == "DC":
            # the ideal transformer relation itself: V(P) - V(Q) = n (V(S) - V(T))
            up, uq = c.pin_unknown(info["TX"], 0), c.pin_unknown(info["TX"], 1)
            us, ut = c.pin_unknown(info["TX"], 2), c.pin_unknown(info["TX"], 3)
            vq = 0.0 if uq < 0 else x[:, uq]
            assert_close(x[:, up] - vq, over[0][2] * (x[:, us] - x[:, ut]), "Vp = n Vs")


@pytest.mark.parametrize("at", ["DC", "TR", "AC"])
def test_center_tap_transformer(ref, abi, at):
    # transformer_center_tap.h:72-125 (n_half = 2 n_total, three branch rows), ratio swept per instance
    n_inst = 17
    nl, info = wl.center_tap_stage(vac=at != "DC")
    rng = np.random.default_rng(37)
    over = [(info["TX"], "n_total", rng.uniform(0.5, 6.0, n_inst)), (info["R1"], "r", rng.uniform(500.0, 2000.0, n_inst)), (info["R2"], "r", rng.uniform(500.0, 4000.0, n_inst))]
    code = {"DC": pe.DC, "TR": pe.TR, "AC": pe.AC}[at]
    kw = {"t_step": 1e-7, "t_stop": 3e-6} if at == "TR" else ({"ac": (pe.SWEEP_LOG, 1e4, 1e8, 11)} if at == "AC" else {})
    want = refapi.run_batch(nl, code, n_inst, over, **kw)
    assert (want["ok"] == 1).all()
    c = pe.Circuit(nl, abi)
    c.set_analyze_type(code)
    if at == "TR":
        c.set_tr(kw["t_step"], kw["t_stop"])
    b = c.batch(n_inst)
    for e, name, v in over:
        b.set_param(e, name, v)
    if at == "AC":
        b.set_ac_sweep(*kw["ac"])
    assert b.analyze(), c.abi.last_error()
    if at == "AC":
        assert_close(b.ac_solution(), want["x"], "centre-tap transformer AC")
    else:
        assert_close(b.solution(), want["x"].real, "centre-tap transformer " + at)
        assert b.total_solves == int(want["solves"].sum())
    # the solo (reference-compatible) path with the constant-folded ratio
    rc, rok, _ = ref_solo(nl, code, ref, **({"tr": (1e-7, 3e-6)} if at == "TR" else ({"omega": 1e6} if at == "AC" else {})))
    gc, gok = gpu_solo(nl, code, abi, **({"tr": (1e-7, 3e-6)} if at == "TR" else ({"omega": 1e6} if at == "AC" else {})))
    assert rok and gok, gc.abi.last_error()
    assert_close(gc.solution(), rc.solution(), "centre-tap transformer solo " + at)


@pytest.mark.parametrize("at", ["OP", "TR"])
def test_relay_hysteresis(ref, abi, at):
    # relay.h:74-105 -- the contact follows the coil voltage of the previous solve with hysteresis
    # (Von 5 V, Voff 3 V); instances sit on both sides of the thresholds, the transient drives the coil through both of them
    n_inst = 24
    nl, info = wl.relay_stage(vac=at == "TR")
    rng = np.random.default_rng(41)
    over = [(info["R"], "r", rng.uniform(200.0, 900.0, n_inst)), (info["Rb"], "r", rng.uniform(4e3, 2e4, n_inst))]
    if at == "OP":
        over.append((info["Vctl"], "V", np.linspace(2.0, 9.0, n_inst)))
    else:
        over.append((info["Vctl"], "Vp", rng.uniform(4.0, 9.0, n_inst)))
    code = pe.OP if at == "OP" else pe.TR
    kw = {"t_step": 1e-7, "t_stop": 8e-6} if at == "TR" else {}
    want = refapi.run_batch(nl, code, n_inst, over, **kw)
    assert (want["ok"] == 1).all()
    c = pe.Circuit(nl, abi)
    c.set_analyze_type(code)
    if at == "TR":
        c.set_tr(kw["t_step"], kw["t_stop"])
    b = c.batch(n_inst)
    for e, name, v in over:
        b.set_param(e, name, v)
    assert b.analyze(), c.abi.last_error()
    x = b.solution()
    assert_close(x, want["x"].real, "relay " + at)
    assert (b.newton_iters() == want["solves"]).all(), "Newton iteration counts differ from the reference"
    if at == "OP":
        ub = c.pin_unknown(info["RY"], 3)
        closed = x[:, ub] > 2.5
        assert closed.any() and (~closed).any()  # both contact states occur in the batch


def test_nonlinear_tr_with_plain_capacitor(ref, abi):
    # regression: a companion update folded into the iter section is skipped after the first Newton iteration; the skip
    # must still consume the op's per-column rows (shared-memory kernel, several streams per warp)
    n_inst = 20
    nl, info = wl.diode_rc(vac=True)
    rng = np.random.default_rng(43)
    over = [(info["R"], "r", rng.uniform(200.0, 900.0, n_inst)), (info["C"], "C", rng.uniform(1e-9, 4e-9, n_inst)), (info["V"], "Vp", rng.uniform(4.0, 9.0, n_inst))]
    want = refapi.run_batch(nl, pe.TR, n_inst, over, t_step=1e-7, t_stop=4e-6)
    assert (want["ok"] == 1).all()
    c = pe.Circuit(nl, abi)
    c.set_analyze_type(pe.TR)
    c.set_tr(1e-7, 4e-6)
    b = c.batch(n_inst)
    for e, name, v in over:
        b.set_param(e, name, v)
    assert b.analyze(), c.abi.last_error()
    assert_close(b.solution(), want["x"].real, "diode + RC transient")
    assert (b.newton_iters() == want["solves"]).all(), "Newton iteration counts differ from the reference"


@pytest.mark.parametrize("at", ["DC", "TR", "TROP", "AC"])
def test_coupled_inductors(ref, abi, at):
    # coupled_inductors.h:92-246 (DC shorts, trapezoidal 2x2 Thevenin companion, AC -j omega [[L1 M],[M L2]])
    n_inst = 18
    nl, info = wl.coupled_inductors_stage(vac=at != "DC")
    rng = np.random.default_rng(47)
    over = [(info["K"], "L1", rng.uniform(5e-4, 2e-3, n_inst)), (info["K"], "L2", rng.uniform(2e-4, 8e-4, n_inst)), (info["K"], "k", rng.uniform(0.5, 0.99, n_inst)),
            (info["R"], "r", rng.uniform(500.0, 2000.0, n_inst))]
    code = {"DC": pe.DC, "TR": pe.TR, "TROP": pe.TROP, "AC": pe.AC}[at]
    kw = {"t_step": 1e-7, "t_stop": 4e-6} if at in ("TR", "TROP") else ({"ac": (pe.SWEEP_LOG, 1e4, 1e8, 11)} if at == "AC" else {})
    want = refapi.run_batch(nl, code, n_inst, over, **kw)
    assert (want["ok"] == 1).all()
    c = pe.Circuit(nl, abi)
    c.set_analyze_type(code)
    if at in ("TR", "TROP"):
        c.set_tr(kw["t_step"], kw["t_stop"])
    b = c.batch(n_inst)
    for e, name, v in over:
        b.set_param(e, name, v)
    if at == "AC":
        b.set_ac_sweep(*kw["ac"])
    assert b.analyze(), c.abi.last_error()
    if at == "AC":
        assert_close(b.ac_solution(), want["x"], "coupled inductors AC")
    else:
        assert_close(b.solution(), want["x"].real, "coupled inductors " + at)
        assert b.total_solves == int(want["solves"].sum())
    # constant-folded M on the solo path
    skw = {"tr": (1e-7, 4e-6)} if at in ("TR", "TROP") else ({"omega": 1e6} if at == "AC" else {})
    rc, rok, _ = ref_solo(nl, code, ref, **skw)
    gc, gok = gpu_solo(nl, code, abi, **skw)
    assert rok and gok, gc.abi.last_error()
    assert_close(gc.solution(), rc.solution(), "coupled inductors solo " + at)


@pytest.mark.parametrize("kind", ["sawtooth", "square", "pulse", "triangle"])
@pytest.mark.parametrize("at", ["DC", "TR", "AC"])
def test_time_domain_generators(ref, abi, kind, at):
    # generator/{sawtooth,square,pulse,triangle}.h -- E(k) = f(tr_duration); DC = f(0); AC drives 0 V
    n_inst = 16
    nl, info = wl.generator_rc(kind)
    rng = np.random.default_rng(53)
    over = [(info["G"], "Vh", rng.uniform(2.0, 5.0, n_inst)), (info["G"], "freq", rng.uniform(1.5e5, 4e5, n_inst)), (info["G"], "phase", rng.uniform(0.0, 6.0, n_inst)),
            (info["R"], "r", rng.uniform(500.0, 2000.0, n_inst))]
    if kind in ("square", "pulse"):
        over.append((info["G"], "duty", rng.uniform(0.2, 0.8, n_inst)))
    code = {"DC": pe.DC, "TR": pe.TR, "AC": pe.AC}[at]
    kw = {"t_step": 5e-8, "t_stop": 1.2e-5} if at == "TR" else ({"ac": (pe.SWEEP_LOG, 1e4, 1e8, 7)} if at == "AC" else {})
    want = refapi.run_batch(nl, code, n_inst, over, **kw)
    assert (want["ok"] == 1).all()
    c = pe.Circuit(nl, abi)
    c.set_analyze_type(code)
    if at == "TR":
        c.set_tr(kw["t_step"], kw["t_stop"])
    b = c.batch(n_inst)
    for e, name, v in over:
        b.set_param(e, name, v)
    if at == "AC":
        b.set_ac_sweep(*kw["ac"])
    assert b.analyze(), c.abi.last_error()
    if at == "AC":
        assert_close(b.ac_solution(), want["x"], kind + " generator AC")
    else:
        assert_close(b.solution(), want["x"].real, kind + " generator " + at)
        assert b.total_solves == int(want["solves"].sum())
