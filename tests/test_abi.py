"""The C-ABI shared library loads and exports every symbol include/phy_engine_b200.h declares (no compute calls)."""
import ctypes as ct
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "phy_engine_b200.h")
LIB = os.path.join(ROOT, "phy-engine_b200", "libphyengine_b200.so")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    names = re.findall(r"\b([a-z_][a-z0-9_]*)\s*\(", text)
    keep = [n for n in names if n.startswith(("circuit_", "create_circuit", "destroy_circuit", "analyze_circuit", "phy_engine_", "verilog_", "pl_", "pe_to_pl_"))]
    return sorted(set(keep))


def test_library_exports_every_declared_symbol():
    lib = ct.CDLL(LIB)
    syms = declared_symbols()
    assert len(syms) >= 55 + 72  # Part 1 + Part 2 + the 72 exports of dll_api.h:252-440 (Part 3)
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, f"declared in include/phy_engine_b200.h but not exported: {missing}"


def test_reference_c_abi_names_are_kept():
    # the analog-path entry points of the reference's dll_api.h keep their names
    lib = ct.CDLL(LIB)
    for s in ("create_circuit", "create_circuit_ex", "destroy_circuit", "circuit_set_analyze_type", "circuit_set_tr", "circuit_set_ac_omega",
              "circuit_set_temperature", "circuit_set_tnom", "circuit_set_model_double_by_name", "circuit_set_model_digital", "circuit_analyze",
              "circuit_digital_clk", "circuit_sample_layout", "circuit_sample", "circuit_sample_u8", "circuit_sample_digital_state_u8", "analyze_circuit",
              "phy_engine_last_error", "phy_engine_clear_error", "phy_engine_string_free"):
        assert hasattr(lib, s), s


def test_error_convention_without_compute():
    lib = ct.CDLL(LIB)
    lib.phy_engine_last_error.restype = ct.c_char_p
    lib.circuit_analyze.argtypes = [ct.c_void_p]
    assert lib.circuit_analyze(None) == 1  # null handle -> 1 (dll_main.cpp:2141-2259 convention)
    lib.circuit_set_tr.argtypes = [ct.c_void_p, ct.c_double, ct.c_double]
    assert lib.circuit_set_tr(None, 1e-6, 1e-3) == 1
    lib.circuit_batch_create.restype = ct.c_void_p
    lib.circuit_batch_create.argtypes = [ct.c_void_p, ct.c_size_t]
    assert lib.circuit_batch_create(None, 4) is None
    assert b"null" in lib.phy_engine_last_error()


REF_PY = "/root/reference/python"
EMU = os.path.join(ROOT, "tests", "emu", "libpe_emu.so")


def _reference_package(lib_path):
    """The reference's own Python binding (python/phy_engine/_ffi.py), loaded fresh against `lib_path`."""
    import importlib
    import sys

    for m in [k for k in sys.modules if k == "phy_engine" or k.startswith("phy_engine.")]:
        del sys.modules[m]
    sys.path.insert(0, REF_PY)
    os.environ["PHY_ENGINE_LIB"] = lib_path
    try:
        return importlib.import_module("phy_engine")
    finally:
        sys.path.remove(REF_PY)


import pytest  # noqa: E402


@pytest.mark.skipif(not os.path.isdir(REF_PY), reason="the reference's Python package is only present where /root/reference is")
def test_reference_python_binding_loads_the_product_library():
    # python/phy_engine/_ffi.py:41-420 binds every export of dll_api.h when it loads the library: all of them must be there
    pkg = _reference_package(LIB)
    from phy_engine import _ffi

    lib = _ffi.load_library()
    assert lib is not None
    assert os.path.samefile(lib._name, LIB)
    # a call outside the hot path fails with the reference's convention and a message saying why
    lib.verilog_runtime_create.restype = ct.c_void_p
    assert lib.verilog_runtime_create(b"module top; endmodule", 21, b"", 0, None, None, 0) is None
    assert "outside the B200 hot path" in _ffi.last_error(lib)
    # the synthesis options are plain process-wide values with the reference's defaults (dll_main.cpp:58-61)
    assert lib.verilog_synth_get_opt_level() == 0 and lib.verilog_synth_get_loop_unroll_limit() == 64
    assert lib.verilog_synth_get_allow_inout() and not lib.verilog_synth_get_assume_binary_inputs()
    lib.verilog_synth_set_opt_level(2)
    assert lib.verilog_synth_get_opt_level() == 2
    lib.verilog_synth_set_opt_level(0)
    assert hasattr(pkg, "Circuit")


@pytest.mark.skipif(not os.path.isdir(REF_PY) or not os.path.exists(EMU), reason="needs the reference's Python package and the host emulator")
def test_reference_python_circuit_runs_the_smoke_netlist():
    # test/0008.dll/dll_main_smoke.cpp through the reference's own phy_engine.Circuit, bound to this host side (the emulator
    # library = the product's C ABI with the kernels replayed on the host; the same flow on the B200: tests/test_dll_smoke.py)
    pkg = _reference_package(EMU)
    from phy_engine.circuit import AnalyzeType, Element, ElementCode, Wire

    c = pkg.Circuit([Element(ElementCode.GROUND), Element(ElementCode.VDC, (5.0,)), Element(ElementCode.RESISTOR, (1000.0,))],
                    [Wire(1, 0, 2, 0), Wire(2, 1, 0, 0), Wire(1, 1, 0, 0)])
    assert c.component_count == 2
    c.set_analyze_type(AnalyzeType.DC)
    c.analyze()
    vo, co, do = c.sample_layout()
    assert vo == [0, 2, 4]
    s = c.sample()
    v = [x for comp in s.components for x in comp.pin_voltages]
    assert all(abs(a - b) <= 1e-6 for a, b in zip(v, (5.0, 0.0, 5.0, 0.0)))
    i = [x for comp in s.components for x in comp.branch_currents]
    assert len(i) == 1 and abs(abs(i[0]) - 5e-3) <= 1e-9
    c.set_model_double_by_name(1, "r", 500.0)
    c.analyze()
    s2 = c.sample()
    i2 = [x for comp in s2.components for x in comp.branch_currents]
    assert abs(abs(i2[0]) - 1e-2) <= 1e-9
    c.close()
