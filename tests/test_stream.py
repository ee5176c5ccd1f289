"""The stream kernel (host/stream.cpp, csrc/pe_b200_stream.cu): one warp per lane group, one-stream programs generated as
TMA-fed tiles.  Through the C ABI on the B200 (-m gpu) and, with the same generated source compiled for the host, on the
emulator (its bulk copies run at issue time and every ordering rule of the ring is checked).

* bit-identical with the word interpreter running the same one-stream program (the generated code restates every op)
* 1e-9 / 1e-12 against the compiled reference, across a resumed transient, with waveform probes, ragged lane counts
* the DC program of a large linear circuit runs through it too; circuits it does not cover take the other kernels
"""
import ctypes as ct

import numpy as np
import pytest

import pe_b200 as pe
import refapi
import workloads as wl
from test_parity import abi, assert_close  # noqa: F401  (abi is a fixture)

STREAM, NO_STREAM_NO_JIT = 64, 128 + 32


def emu_counters(abi):  # noqa: F811
    lib = abi.lib
    if not hasattr(lib, "pe_emu_stream_errors"):
        return None
    lib.pe_emu_stream_errors.restype = ct.c_uint64
    lib.pe_emu_stream_launches.restype = ct.c_uint64
    return int(lib.pe_emu_stream_launches()), int(lib.pe_emu_stream_errors())


def ladder_batch(abi, n_sections, n_inst, tuning, steps, resident=None, seed=1, probes=None):  # noqa: F811
    nl, info = wl.rc_ladder(n_sections)
    c = pe.Circuit(nl, abi)
    c.set_analyze_type(pe.TR)
    c.set_tr(1e-8, 1e-8 * (steps - 0.5))
    b = c.batch(n_inst)
    if resident:
        b.set_resident(*resident)
        b.set_workspace(2)
    b.set_tuning(tuning)
    rng = np.random.default_rng(seed)
    over = [(e, "r", wl.sweep_values(rng, 1e3, n_inst)) for e in info["R"]] + [(e, "c", wl.sweep_values(rng, 1e-9, n_inst)) for e in info["C"]]
    for e, name, v in over:
        b.set_param(e, name, v)
    if probes is not None:
        b.set_probes(probes)
    return nl, c, b, over


@pytest.mark.parametrize("n_sections,n_inst,steps", [(160, 33, 4), (300, 40, 6), (1000, 70, 3)])
def test_stream_kernel_is_bit_identical_to_the_interpreter(ref, abi, n_sections, n_inst, steps):  # noqa: F811
    before = emu_counters(abi)
    nl, c1, b1, over = ladder_batch(abi, n_sections, n_inst, NO_STREAM_NO_JIT, steps, resident=(1, 0, 1))
    assert b1.analyze(), c1.abi.last_error()
    assert b1.last_kernel() == 0
    x1 = b1.solution()
    nl, c2, b2, _ = ladder_batch(abi, n_sections, n_inst, STREAM, steps)
    assert b2.analyze(), c2.abi.last_error()
    assert b2.last_kernel() == 2  # the stream kernel ran
    x2 = b2.solution()
    assert np.array_equal(x1, x2)
    assert b1.total_solves == b2.total_solves == n_inst * steps
    # a second analyze() continues the transient (circuit.h:242): state written by the stream kernel is complete
    assert b1.analyze() and b2.analyze()
    assert np.array_equal(b1.solution(), b2.solution())
    # and both agree with the compiled reference
    want = refapi.run_batch(nl, pe.TR, n_inst, over, t_step=1e-8, t_stop=1e-8 * (steps - 0.5))
    assert (want["solves"] == steps).all()
    assert_close(x2, want["x"].real, "stream kernel vs reference")
    after = emu_counters(abi)
    if after is not None:
        assert after[0] >= before[0] + 2  # two stream launches on the emulator
        assert after[1] == before[1]  # no ordering violation: fences, ring reuse, tile accounting


def test_stream_kernel_waveform_probes_and_ragged_batch(ref, abi):  # noqa: F811
    n_sections, n_inst, steps = 120, 45, 8
    probes = [1, n_sections // 2, n_sections]
    nl, c, b, over = ladder_batch(abi, n_sections, n_inst, STREAM, steps, probes=probes)
    assert b.analyze(), c.abi.last_error()
    assert b.last_kernel() == 2
    w = b.waveform(steps)  # [steps, probes, n_inst]
    nl, c0, b0, _ = ladder_batch(abi, n_sections, n_inst, NO_STREAM_NO_JIT, steps, resident=(1, 0, 1), probes=probes)  # the same one-stream program, interpreted
    assert b0.analyze()
    assert np.array_equal(w, b0.waveform(steps))
    # the last row of the waveform is the final state
    x = b.solution()
    for k, u in enumerate(probes):
        assert np.array_equal(w[steps - 1, k], x[:, u])
    want = refapi.run_batch(nl, pe.TR, n_inst, over, t_step=1e-8, t_stop=1e-8 * (steps - 0.5))
    assert_close(x, want["x"].real, "ragged batch")


def test_stream_kernel_runs_the_dc_program_too(ref, abi):  # noqa: F811
    # OP of a large linear circuit takes the stream kernel as well (the DC program of the one-stream compile: no companion
    # models, one solve); a load resistor per section makes the DC solution non-trivial
    n_sections, n_inst = 150, 37
    nl, info = wl.rc_ladder(n_sections)
    g = 0  # element 0 is the ground placeholder
    loads = []
    for e in info["C"]:
        r = nl.add(pe.R, 47e3)
        nl.wire(r, 0, e, 0)
        nl.wire(r, 1, g, 0)
        loads.append(r)
    rng = np.random.default_rng(4)
    # every like element is swept: a partial sweep would make neighbouring elimination steps differ (broadcast constants next
    # to per-instance rows) and the program would not be periodic enough for the stream kernel
    over = [(e, "r", wl.sweep_values(rng, 1e3, n_inst)) for e in info["R"]] + [(e, "r", wl.sweep_values(rng, 47e3, n_inst)) for e in loads]
    want = refapi.run_batch(nl, pe.OP, n_inst, over)
    assert (want["ok"] == 1).all()
    c = pe.Circuit(nl, abi)
    c.set_analyze_type(pe.OP)
    b = c.batch(n_inst)
    b.set_tuning(STREAM)
    for e, name, v in over:
        b.set_param(e, name, v)
    assert b.analyze(), c.abi.last_error()
    assert b.last_kernel() == 2
    assert (b.status() == 0).all() and (b.newton_iters() == 1).all()
    x = b.solution()
    assert np.ptp(x[:, 1:n_sections]) > 0.05  # the loads pull the far end down
    assert_close(x, want["x"].real, "OP through the stream kernel")


def test_stream_is_not_taken_where_it_does_not_apply(abi):  # noqa: F811
    # nonlinear circuits, small circuits and AC run the other kernels even when the stream kernel is requested
    for nl, at in ((wl.diode_ladder(40)[0], pe.OP), (wl.rc_ladder(8)[0], pe.TR), (wl.rlc_ladder(40)[0], pe.AC)):
        c = pe.Circuit(nl, abi)
        c.set_analyze_type(at)
        if at == pe.TR:
            c.set_tr(1e-8, 3e-8)
        if at == pe.AC:
            c.set_ac_omega(1e6)
        b = c.batch(33)
        b.set_tuning(STREAM)
        assert b.analyze(), c.abi.last_error()
        assert b.last_kernel() != 2


def test_stream_kernel_covers_inductors(ref, abi):  # noqa: F811
    # IND_STEP (inductor.h:134-160) in the generated tiles: step response of an RLC ladder, every R / L / C swept
    n_sections, n_inst, steps, dt = 80, 37, 5, 1e-9
    nl, info = wl.rlc_ladder_dc(n_sections)
    rng = np.random.default_rng(2)
    over = ([(e, "r", wl.sweep_values(rng, 10.0, n_inst)) for e in info["R"]] + [(e, "L", wl.sweep_values(rng, 1e-6, n_inst)) for e in info["L"]] +
            [(e, "c", wl.sweep_values(rng, 1e-9, n_inst)) for e in info["C"]])
    got = {}
    for name, tuning, resident in (("interpreter", NO_STREAM_NO_JIT, (1, 0, 1)), ("stream", STREAM, None)):
        c = pe.Circuit(nl, abi)
        c.set_analyze_type(pe.TR)
        c.set_tr(dt, dt * (steps - 0.5))
        b = c.batch(n_inst)
        if resident:
            b.set_resident(*resident)
            b.set_workspace(2)
        b.set_tuning(tuning)
        for e, name_, v in over:
            b.set_param(e, name_, v)
        assert b.analyze(), c.abi.last_error()
        assert b.last_kernel() == (2 if name == "stream" else 0)
        got[name] = b.solution()
        assert b.analyze(), c.abi.last_error()  # the inductor's companion state written by the stream kernel is complete
        got[name + "2"] = b.solution()
    assert np.array_equal(got["interpreter"], got["stream"]) and np.array_equal(got["interpreter2"], got["stream2"])
    want = refapi.run_batch(nl, pe.TR, n_inst, over, t_step=dt, t_stop=dt * (steps - 0.5))
    assert (want["ok"] == 1).all()
    assert_close(got["stream"], want["x"].real, "RLC step response through the stream kernel")


def test_stream_kernel_covers_sinusoidal_sources(ref, abi):  # noqa: F811
    # a VAC / IAC source is re-evaluated at the time of every solve (VSIN, VAC.h:176): the generated tiles read the step's time
    # from the kernel's context.  RLC ladder driven by a 200 kHz sine, every R / L / C swept, across a resumed transient
    n_sections, n_inst, steps, dt = 80, 37, 6, 1e-9
    nl, info = wl.rlc_ladder(n_sections)
    rng = np.random.default_rng(2)
    over = ([(e, "r", wl.sweep_values(rng, 10.0, n_inst)) for e in info["R"]] + [(e, "L", wl.sweep_values(rng, 1e-6, n_inst)) for e in info["L"]] +
            [(e, "c", wl.sweep_values(rng, 1e-9, n_inst)) for e in info["C"]])
    got = {}
    for name, tuning, resident in (("interpreter", NO_STREAM_NO_JIT, (1, 0, 1)), ("stream", STREAM, None)):
        c = pe.Circuit(nl, abi)
        c.set_analyze_type(pe.TR)
        c.set_tr(dt, dt * (steps - 0.5))
        b = c.batch(n_inst)
        if resident:
            b.set_resident(*resident)
            b.set_workspace(2)
        b.set_tuning(tuning)
        for e, name_, v in over:
            b.set_param(e, name_, v)
        assert b.analyze(), c.abi.last_error()
        assert b.last_kernel() == (2 if name == "stream" else 0)
        got[name] = b.solution()
        assert b.analyze(), c.abi.last_error()  # the second call starts at t = 6 ns: the source's phase continues
        got[name + "2"] = b.solution()
    assert np.array_equal(got["interpreter"], got["stream"]) and np.array_equal(got["interpreter2"], got["stream2"])
    assert np.abs(got["stream2"] - got["stream"]).max() > 0.0
    want = refapi.run_batch(nl, pe.TR, n_inst, over, t_step=dt, t_stop=dt * (steps - 0.5))
    assert (want["ok"] == 1).all()
    assert_close(got["stream"], want["x"].real, "sine-driven RLC ladder through the stream kernel")


def test_stream_kernel_covers_the_time_domain_generators(ref, abi):  # noqa: F811
    # a pulse generator (generator/pulse.h) into an RC ladder: GEN_EVAL reads the time of the solve like VSIN does; 40 steps of
    # 10 ns cross the rising edge, the plateau and the falling edge of the 400 ns period
    n_sections, n_inst, steps, dt = 160, 33, 40, 1e-8
    nl, info = wl.pulse_rc_ladder(n_sections)
    rng = np.random.default_rng(6)
    over = [(e, "r", wl.sweep_values(rng, 1e3, n_inst)) for e in info["R"]] + [(e, "c", wl.sweep_values(rng, 1e-9, n_inst)) for e in info["C"]]
    got = {}
    for name, tuning, resident in (("interpreter", NO_STREAM_NO_JIT, (1, 0, 1)), ("stream", STREAM, None)):
        c = pe.Circuit(nl, abi)
        c.set_analyze_type(pe.TR)
        c.set_tr(dt, dt * (steps - 0.5))
        b = c.batch(n_inst)
        if resident:
            b.set_resident(*resident)
            b.set_workspace(2)
        b.set_tuning(tuning)
        for e, name_, v in over:
            b.set_param(e, name_, v)
        b.set_probes([0])
        assert b.analyze(), c.abi.last_error()
        assert b.last_kernel() == (2 if name == "stream" else 0)
        got[name] = b.solution()
        got[name + "w"] = b.waveform(steps)
    assert np.array_equal(got["interpreter"], got["stream"]) and np.array_equal(got["interpreterw"], got["streamw"])
    w = got["streamw"][:, 0, 0]
    assert w.max() > 3.9 and w.min() < -0.9  # the source node saw both levels of the pulse
    want = refapi.run_batch(nl, pe.TR, n_inst, over, t_step=dt, t_stop=dt * (steps - 0.5))
    assert (want["ok"] == 1).all()
    assert_close(got["stream"], want["x"].real, "pulse-driven RC ladder through the stream kernel")


def test_stream_request_on_an_uncovered_program_falls_back(ref, abi):  # noqa: F811
    # coupled inductors (KIND_STEP) are not covered by the generator: the program is rejected at compile time, the default
    # geometry runs and the caller gets the same answer
    nl, info = wl.coupled_inductors_stage()
    rc = refapi.RefCircuit(nl)
    rc.set_analyze_type(pe.TR)
    rc.set_tr(1e-8, 2e-7)
    assert rc.analyze()
    c = pe.Circuit(nl, abi)
    c.set_analyze_type(pe.TR)
    c.set_tr(1e-8, 2e-7)
    b = c.batch(3)
    b.set_tuning(STREAM)
    assert b.analyze(), c.abi.last_error()
    assert b.last_kernel() != 2
    assert_close(b.solution()[0], rc.solution().real, "coupled inductors next to the stream request")


def test_build_flags_reach_the_compiler(tmp_path):
    # the stream module is compiled at run time with the lane-group geometry as -D flags; a flag that silently does not arrive
    # gives a kernel built for another row width (this happened once: an "illegal memory access" in a kernel that was fine)
    import os
    import shutil

    if shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"):
        pytest.skip("nvcc not available")
    lib = pe.product().lib
    lib.pe_b200_stream_build.argtypes = [ct.c_char_p, ct.c_char_p, ct.c_char_p, ct.c_int, ct.c_int, ct.c_char_p, ct.c_size_t]
    src = tmp_path / "probe.inc"
    src.write_text("#define PE_STREAM_TILES 0u\n#define PE_STREAM_STAGE_ROWS 1u\n"
                   "static_assert(PE_SGL == 16 && PE_SJ == 1, \"build flags did not reach the compiler\");\n"
                   "__device__ __forceinline__ void pe_stream_iter(sk_ctx& k, uint32_t& fm) { (void)k; (void)fm; }\n")
    csrc = os.path.join(os.path.dirname(os.path.abspath(pe.__file__)), "csrc").encode()
    log = ct.create_string_buffer(4096)
    assert lib.pe_b200_stream_build(str(src).encode(), str(tmp_path / "ok.cubin").encode(), csrc, 1, 16, log, 4096) == 0, log.value.decode()
    assert os.path.getsize(tmp_path / "ok.cubin") > 0
    assert lib.pe_b200_stream_build(str(src).encode(), str(tmp_path / "bad.cubin").encode(), csrc, 1, 32, log, 4096) != 0
    assert b"build flags did not reach the compiler" in log.value
