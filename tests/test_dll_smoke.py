"""test/0008.dll/dll_main_smoke.cpp of the reference, re-stated call for call through the raw C ABI — once against the
compiled reference (oracle/_ref) and once against the product (CUDA on the GPU box, the host emulator in the CPU suite):
same create_circuit arguments, same analyze_circuit call, same layout {0, 2, 4} and 1e-6 voltage checks
(dll_main_smoke.cpp:36-113), plus circuit_sample_layout / circuit_sample_u8 (dll_api.h:186-218)."""
import ctypes as ct

import numpy as np
import pytest

import pe_b200 as pe

SZ = ct.c_size_t
PSZ = ct.POINTER(SZ)
PD = ct.POINTER(ct.c_double)
PI = ct.POINTER(ct.c_int)


def run_smoke(lib):
    lib.create_circuit.restype = ct.c_void_p
    lib.create_circuit.argtypes = [PI, SZ, PI, SZ, PD, ct.POINTER(PSZ), ct.POINTER(PSZ), PSZ]
    lib.destroy_circuit.argtypes = [ct.c_void_p, PSZ, PSZ]
    lib.circuit_set_analyze_type.argtypes = [ct.c_void_p, ct.c_uint32]
    lib.analyze_circuit.argtypes = [ct.c_void_p, PSZ, PSZ, SZ, PI, PSZ, PD, SZ, PD, PSZ, PD, PSZ, ct.POINTER(ct.c_bool), PSZ]
    lib.circuit_sample_layout.argtypes = [ct.c_void_p, PSZ, PSZ, SZ, PSZ, PSZ, PSZ]
    elements = (ct.c_int * 3)(0, 4, 1)  # ground, VDC, R
    wires = (ct.c_int * 12)(1, 0, 2, 0, 2, 1, 0, 0, 1, 1, 0, 0)
    props = (ct.c_double * 2)(5.0, 1000.0)
    vp, cp, cs = PSZ(), PSZ(), SZ(0)
    c = lib.create_circuit(elements, 3, wires, 12, props, ct.byref(vp), ct.byref(cp), ct.byref(cs))
    assert c and vp and cp
    assert cs.value == 2
    # dll_main_smoke.cpp:67-68 casts the handle and calls set_analyze_type(DC); the C ABI equivalent:
    assert lib.circuit_set_analyze_type(c, pe.DC) == 0
    voltage = (ct.c_double * 16)()
    current = (ct.c_double * 16)()
    digital = (ct.c_bool * 16)()
    vo, co, do = (SZ * 3)(), (SZ * 3)(), (SZ * 3)()
    rc = lib.analyze_circuit(c, vp, cp, cs.value, None, None, None, 0, voltage, vo, current, co, digital, do)
    assert rc == 0
    assert list(vo) == [0, 2, 4]
    assert abs(voltage[0] - 5.0) <= 1e-6 and abs(voltage[1]) <= 1e-6 and abs(voltage[2] - 5.0) <= 1e-6 and abs(voltage[3]) <= 1e-6
    vo2, co2, do2 = (SZ * 3)(), (SZ * 3)(), (SZ * 3)()
    assert lib.circuit_sample_layout(c, vp, cp, cs.value, vo2, co2, do2) == 0
    out = {"voltage": [voltage[i] for i in range(4)], "vo": list(vo), "co": list(co), "do": list(do), "co2": list(co2), "do2": list(do2),
           "current": [current[i] for i in range(co[2])]}
    lib.destroy_circuit(c, vp, cp)
    return out


@pytest.fixture(params=[pytest.param("gpu", marks=pytest.mark.gpu), "emu"])
def product_lib(request):
    if request.param == "gpu":
        return pe.product().lib
    import emuapi

    return emuapi.emulator().lib


def test_dll_main_smoke_on_the_reference(ref):
    run_smoke(ref.lib)


def test_dll_main_smoke_on_the_product(product_lib):
    run_smoke(product_lib)


def test_dll_main_smoke_same_answers(ref, product_lib):
    a, b = run_smoke(ref.lib), run_smoke(product_lib)
    for k in ("vo", "co", "do", "co2", "do2"):
        assert a[k] == b[k], k
    assert np.allclose(a["voltage"], b["voltage"], rtol=1e-9, atol=1e-12)
    assert np.allclose(a["current"], b["current"], rtol=1e-9, atol=1e-12)  # the VDC branch current: -5 mA
