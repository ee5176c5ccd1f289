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
    keep = [n for n in names if n.startswith(("circuit_", "create_circuit", "destroy_circuit", "analyze_circuit", "phy_engine_"))]
    return sorted(set(keep))


def test_library_exports_every_declared_symbol():
    lib = ct.CDLL(LIB)
    syms = declared_symbols()
    assert len(syms) >= 55
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
