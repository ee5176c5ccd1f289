"""Mixed-signal boundary, analog -> digital (SURVEY.md 8a row a16, config E shape): the comparators of a flash ADC,
evaluated for every instance on the device (circuit_batch_digital_clk), against (1) the reference's own
update_digital_clk through its C ABI on a few instances and (2) the comparator rule vA >= vB applied to the reference's
analog solution for the whole batch.  Bit-exact."""
import ctypes as ct

import numpy as np
import pytest

import pe_b200 as pe
import refapi
import workloads as wl
from test_parity import PATHS, assert_close


@pytest.fixture(params=[pytest.param("gpu", marks=pytest.mark.gpu), "emu"])
def abi(request):
    if request.param == "gpu":
        return pe.product()
    import emuapi

    return emuapi.emulator()


def reference_comparator_bits(nl, info, vin, r_vals):
    """the reference's digital states of the comparator outputs for ONE instance: circuit_analyze (DC) + circuit_digital_clk
    + circuit_sample_digital_state_u8 (dll_api.h:220-235)"""
    c = refapi.RefCircuit(nl)
    c.set_analyze_type(pe.DC)
    assert c.set_param(info["Vin"], "V", float(vin)) == 0
    for e, v in zip(info["R"], r_vals):
        assert c.set_param(e, "r", float(v)) == 0
    assert c.analyze()
    lib = c.abi.lib
    lib.circuit_digital_clk.argtypes = [ct.c_void_p]
    assert lib.circuit_digital_clk(c.h) == 0
    SZ, PSZ, PD = ct.c_size_t, ct.POINTER(ct.c_size_t), ct.POINTER(ct.c_double)
    lib.circuit_sample_digital_state_u8.argtypes = [ct.c_void_p, PSZ, PSZ, SZ, PD, PSZ, PD, PSZ, ct.POINTER(ct.c_uint8), PSZ]
    n = c.comp_size
    vo, co, dg = (SZ * (n + 1))(), (SZ * (n + 1))(), (SZ * (n + 1))()
    volt, cur, dig = (ct.c_double * 256)(), (ct.c_double * 64)(), (ct.c_uint8 * 256)()
    assert lib.circuit_sample_digital_state_u8(c.h, c._vp, c._cp, n, volt, vo, cur, co, dig, dg) == 0
    bits = []
    for cm in info["CMP"]:
        k = nl.component_index(cm)
        bits.append(int(dig[dg[k] + 2]))  # pin o of the comparator: 0 false, 1 true (2 X, 3 Z)
    return np.array(bits, dtype=np.uint8)


def test_flash_adc_comparators_match_the_reference(ref, abi):
    n_inst = 257
    nl, info = wl.flash_adc(16)
    rng = np.random.default_rng(31)
    vin = rng.uniform(0.0, 5.0, n_inst)
    r_vals = [1e3 * rng.uniform(0.97, 1.03, n_inst) for _ in info["R"]]
    over = [(info["Vin"], "V", vin)] + [(e, "r", v) for e, v in zip(info["R"], r_vals)]
    want = refapi.run_batch(nl, pe.DC, n_inst, over)
    assert (want["ok"] == 1).all()
    c = pe.Circuit(nl, abi)
    c.set_analyze_type(pe.DC)
    b = c.batch(n_inst)
    for e, name, v in over:
        b.set_param(e, name, v)
    assert b.analyze(), c.abi.last_error()
    x = b.solution()
    assert x.shape[1] == 19  # 17 analog nodes + 2 source branches: the comparator pins add none
    assert_close(x, want["x"].real, "adc ladder")
    bits = b.digital_clk()
    assert bits.shape == (n_inst, 15)
    # (2) the comparator rule on the reference's analog solution, whole batch
    vin_u = c.pin_unknown(info["Vin"], 0)
    taps = [c.pin_unknown(info["R"][k], 0) for k in range(1, 16)]
    expect = (want["x"].real[:, [vin_u]] >= want["x"].real[:, taps]).astype(np.uint8)
    assert (bits == expect).all()
    # thermometer code: monotone, and its height is the quantised input
    assert (np.diff(bits.astype(int), axis=1) <= 0).all()
    # (1) the reference's own digital engine on a few instances
    for i in (0, 1, 100, n_inst - 1):
        rb = reference_comparator_bits(nl, info, vin[i], [v[i] for v in r_vals])
        assert (rb == bits[i]).all(), (i, rb, bits[i])


@pytest.mark.gpu
def test_flash_adc_full_size_4096_instances():
    nl, info = wl.flash_adc(16)
    n_inst = 4096
    rng = np.random.default_rng(5)
    vin = rng.uniform(0.0, 5.0, n_inst)
    c = pe.Circuit(nl)
    c.set_analyze_type(pe.DC)
    b = c.batch(n_inst)
    b.set_param(info["Vin"], "V", vin)
    assert b.analyze()
    bits = b.digital_clk()
    # nominal ladder: tap k = 5 k / 16 V (random inputs: no tie within 1e-9)
    expect = (vin[:, None] >= 5.0 * np.arange(1, 16)[None, :] / 16.0).astype(np.uint8)
    safe = np.abs(vin[:, None] - 5.0 * np.arange(1, 16)[None, :] / 16.0).min(axis=1) > 1e-9
    assert (bits[safe] == expect[safe]).all()


def product_comparator_bits(abi, nl, info, vin, r_vals, which="state"):
    """the same reference sequence -- circuit_analyze, circuit_digital_clk, circuit_sample_* -- on the product's Part-1 ABI"""
    c = pe.Circuit(nl, abi)
    c.set_analyze_type(pe.DC)
    assert c.set_param(info["Vin"], "V", float(vin)) == 0
    for e, v in zip(info["R"], r_vals):
        assert c.set_param(e, "r", float(v)) == 0
    SZ, PSZ, PD = ct.c_size_t, ct.POINTER(ct.c_size_t), ct.POINTER(ct.c_double)
    lib = c.abi.lib
    lib.circuit_digital_clk.argtypes = [ct.c_void_p]
    n = c.comp_size
    vo, co, dg = (SZ * (n + 1))(), (SZ * (n + 1))(), (SZ * (n + 1))()
    volt, cur, dig = (ct.c_double * 256)(), (ct.c_double * 64)(), (ct.c_uint8 * 256)()
    fn = lib.circuit_sample_digital_state_u8 if which == "state" else lib.circuit_sample_u8
    fn.argtypes = [ct.c_void_p, PSZ, PSZ, SZ, PD, PSZ, PD, PSZ, ct.POINTER(ct.c_uint8), PSZ]
    before = None
    if which == "state":
        assert fn(c.h, c._vp, c._cp, n, volt, vo, cur, co, dig, dg) == 0
        before = [int(dig[dg[nl.component_index(cm)] + 2]) for cm in info["CMP"]]
    assert c.analyze(), c.abi.last_error()
    assert lib.circuit_digital_clk(c.h) == 0
    assert fn(c.h, c._vp, c._cp, n, volt, vo, cur, co, dig, dg) == 0
    bits = np.array([int(dig[dg[nl.component_index(cm)] + 2]) for cm in info["CMP"]], dtype=np.uint8)
    pins_ab = [int(dig[dg[nl.component_index(cm)] + j]) for cm in info["CMP"] for j in (0, 1)]
    return bits, before, pins_ab


def test_single_instance_digital_clk_follows_the_reference_sequence(ref, abi):
    # the Part-1 ABI accepted comparators but circuit_digital_clk was a no-op and the samples always read 0 / X
    nl, info = wl.flash_adc(16)
    r_nom = [1e3] * len(info["R"])
    for vin in (0.0, 0.31, 1.7, 2.5000001, 4.99, 5.0):
        want = reference_comparator_bits(nl, info, vin, r_nom)
        got, before, pins_ab = product_comparator_bits(abi, nl, info, vin, r_nom, "state")
        assert (got == want).all(), (vin, got, want)
        assert all(v == 2 for v in before)  # indeterminate until the first tick
        assert all(v == 2 for v in pins_ab)  # analog pins report X (dll_api.h:224-226)
        got8, _, pins8 = product_comparator_bits(abi, nl, info, vin, r_nom, "u8")
        assert (got8 == want).all() and not any(pins8)


D2A_PROBE = r"""
import ctypes as ct, sys
import pe_b200 as pe
from pe_b200 import Netlist
which = sys.argv[1]
nl = Netlist(); g = nl.ground()
va = nl.add(pe.VDC, 1.0); vb = nl.add(pe.VDC, 0.5)
cmp_ = nl.add(pe.COMPARATOR, 0.0, 5.0)
rl, ra, rb = nl.add(pe.R, 1e3), nl.add(pe.R, 1e3), nl.add(pe.R, 1e3)
for a, pa, b, pb in ((va, 1, g, 0), (vb, 1, g, 0), (va, 0, cmp_, 0), (vb, 0, cmp_, 1), (ra, 0, va, 0), (ra, 1, g, 0), (rb, 0, vb, 0), (rb, 1, g, 0),
                     (cmp_, 2, rl, 0), (rl, 1, g, 0)):  # the comparator's output pin sits on an ANALOG node (a resistor hangs on it)
    nl.wire(a, pa, b, pb)
if which == "reference":
    import refapi
    c = refapi.RefCircuit(nl)
else:
    import emuapi
    c = pe.Circuit(nl, emuapi.emulator())
c.set_analyze_type(pe.DC)
assert c.analyze()
c.abi.lib.circuit_digital_clk.argtypes = [ct.c_void_p]
rc = c.abi.lib.circuit_digital_clk(c.h)
print("clk", rc, flush=True)
if which == "reference":
    ok = c.analyze()  # circuit.h:509 + 1015-1022: the drivers collected by the tick become voltage-source branches here
    print("second analyze", ok, flush=True)
else:
    print("error:", c.abi.last_error(), flush=True)
"""


def test_comparator_on_an_analog_node_is_refused_where_the_reference_dies(ref, tmp_path):
    """VERDICT r01 item 7 (digital -> analog drivers, comparator.h:88-95 / circuit.h:1015-1022).  Through its own C ABI the
    reference does not survive that path: update_table_digital_clk visits the comparator once per hybrid node it touches
    (circuit.h:316-333), each visit appends the same (level, node) driver, and the next analyze() dies.  There is no reference
    behaviour to be drop-in for, so circuit_digital_clk refuses the netlist with an error instead of guessing one."""
    import os
    import subprocess
    import sys

    script = tmp_path / "d2a_probe.py"
    script.write_text(D2A_PROBE)
    env = dict(os.environ)
    here = os.path.dirname(os.path.abspath(__file__))
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(os.path.dirname(here), "phy-engine_b200"), here, env.get("PYTHONPATH", "")])
    r = subprocess.run([sys.executable, str(script), "reference"], capture_output=True, text=True, env=env, timeout=120)
    assert "clk 0" in r.stdout  # the tick itself succeeds
    assert r.returncode != 0 and "second analyze" not in r.stdout, (r.returncode, r.stdout, r.stderr[-300:])  # ... the solve after it does not return
    p = subprocess.run([sys.executable, str(script), "product"], capture_output=True, text=True, env=env, timeout=120)
    assert p.returncode == 0, p.stderr[-400:]
    assert "clk 1" in p.stdout and "analog" in p.stdout  # refused, with a message that names the analog output node
