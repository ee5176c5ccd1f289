"""The specialised tree-streaming kernel (host/jit.cpp: the iter section of a program compiled to straight-line sm_100a code
at run time) against the word interpreter it replaces -- bit for bit -- and against the compiled reference.

CPU part: the generator (which programs it covers, what it emits).  GPU part: parity through the C ABI."""
import ctypes as ct
import os

import numpy as np
import pytest

import pe_b200 as pe
import workloads as wl
from test_parity import assert_close

JIT, NO_JIT = 16, 32  # circuit_batch_set_tuning: bit 4 requires the specialised kernel, bit 5 forbids it


def _rc_batch(sections, n_inst, streams, tuning, seed=5, steps=20, abi=None):
    nl, info = wl.rc_ladder(sections)
    rng = np.random.default_rng(seed)
    c = pe.Circuit(nl, abi) if abi is not None else pe.Circuit(nl)
    c.set_analyze_type(pe.TR)
    c.set_tr(1e-8, 1e-8 * (steps - 0.5))
    b = c.batch(n_inst)
    b.set_resident(streams, 0, 4)
    b.set_workspace(2)
    b.set_tuning(tuning)
    items = [(e, "r") for e in info["R"]] + [(e, "c") for e in info["C"]]
    vals = np.ascontiguousarray(np.array([1e3] * sections + [1e-9] * sections)[:, None] * rng.uniform(0.8, 1.2, (2 * sections, n_inst)))
    b.set_params(b.param_table(items), vals.ctypes.data)
    return nl, info, c, b, vals


def _source(b, mode):
    lib = b.lib
    lib.circuit_batch_jit_source.restype = ct.c_size_t
    lib.circuit_batch_jit_source.argtypes = [ct.c_void_p, ct.c_int, ct.c_char_p, ct.c_size_t]
    n = lib.circuit_batch_jit_source(b.h, mode, None, 0)
    buf = ct.create_string_buffer(n + 1)
    lib.circuit_batch_jit_source(b.h, mode, buf, n)
    return buf.raw[:n].decode()


def test_generator_covers_linear_transient_and_shares_isomorphic_subtrees():
    _, _, c, b, _ = _rc_batch(256, 128, 8, 0)
    b.compile_host()
    src = _source(b, pe.MODE_TR)
    assert "pe_jit_iter" in src and "jcap(" in src and "jrcp(" in src
    n_funcs = src.count("__device__ __forceinline__ uint32_t jf")
    n_calls = src.count("fm |= jf")
    assert 0 < n_funcs < 8 * 8  # (stream, phase) functions are shared between isomorphic sub-trees
    assert n_calls >= n_funcs
    assert src.count("group_sync<CL>();") == b.resident_info(pe.MODE_TR)["iter_phases"] - 1


def test_generator_declines_nonlinear_programs():
    nl, info = wl.diode_ladder(4)
    c = pe.Circuit(nl)
    c.set_analyze_type(pe.OP)
    b = c.batch(64)
    b.compile_host()
    assert _source(b, pe.MODE_DC) == ""


@pytest.mark.gpu
@pytest.mark.parametrize("sections,streams,n_inst", [(48, 4, 300), (300, 32, 260)])
def test_specialised_kernel_is_bit_identical_to_the_interpreter(sections, streams, n_inst):
    _, _, c0, b0, vals = _rc_batch(sections, n_inst, streams, NO_JIT)
    assert b0.analyze(), c0.abi.last_error()
    assert b0.last_kernel() == 0
    _, _, c1, b1, _ = _rc_batch(sections, n_inst, streams, JIT)
    assert b1.analyze(), c1.abi.last_error()
    assert b1.last_kernel() == 1
    x0, x1 = b0.solution(), b1.solution()
    assert np.isfinite(x1).all()
    assert (x0 == x1).all(), "specialised kernel differs from the interpreter"
    assert b1.total_solves == b0.total_solves == n_inst * 20
    assert (b1.status() == 0).all()


@pytest.mark.gpu
def test_specialised_kernel_against_reference(ref):
    import refapi

    sections, n_inst, steps = 64, 130, 20
    nl, info, c, b, vals = _rc_batch(sections, n_inst, 8, JIT, seed=9, steps=steps)
    assert b.analyze(), c.abi.last_error()
    assert b.last_kernel() == 1
    pick = [0, 57, 129]
    over = [(e, "r", vals[k][pick]) for k, e in enumerate(info["R"])] + [(e, "c", vals[sections + k][pick]) for k, e in enumerate(info["C"])]
    want = refapi.run_batch(nl, pe.TR, len(pick), over, t_step=1e-8, t_stop=1e-8 * (steps - 0.5))
    assert (want["ok"] == 1).all()
    assert_close(b.solution()[pick], want["x"].real, "specialised kernel vs reference")


@pytest.mark.gpu
def test_second_analyze_continues_the_transient_with_the_specialised_kernel():
    # circuit.h:233-289: a second analyze() continues from tr_duration with the stored companion state
    _, _, c0, b0, _ = _rc_batch(48, 200, 4, NO_JIT)
    _, _, c1, b1, _ = _rc_batch(48, 200, 4, JIT)
    for _ in range(2):
        assert b0.analyze() and b1.analyze()
    assert b1.last_kernel() == 1
    assert (b0.solution() == b1.solution()).all()
    assert b0.tr_duration == b1.tr_duration
