"""Parity against the committed golden vectors (tests/golden/cases.npz, produced by the compiled reference with
tests/golden/make_golden.py).  Needs neither /root/reference nor oracle/_ref at run time.

backend "gpu": the CUDA kernels through the C ABI (-m gpu).  backend "emu": the host replay of the kernel programs
(CPU suite).  Tolerance: 1e-9 relative / 1e-12 absolute, identical solve_once counts, identical ok / failed flags."""
import os

import numpy as np
import pytest

import golden_cases
import pe_b200 as pe
from test_parity import PATHS, assert_close

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cases.npz"))


@pytest.fixture(params=[pytest.param("gpu", marks=pytest.mark.gpu), "emu"])
def abi(request):
    if request.param == "gpu":
        return pe.product()
    import emuapi

    return emuapi.emulator()


@pytest.fixture(params=list(PATHS), autouse=True)
def path(request, abi):
    assert abi.lib.phy_engine_b200_set_default_path(*PATHS[request.param]) == 0
    yield request.param
    abi.lib.phy_engine_b200_set_default_path(0, 0, 0, 0, 0, 0)


@pytest.mark.parametrize("name", list(golden_cases.CASES))
def test_golden_case(abi, name):
    case = golden_cases.CASES[name]
    nl, over, kw = golden_cases.build(name)
    want_x, want_solves, want_ok = GOLD[name + "/x"], GOLD[name + "/solves"], GOLD[name + "/ok"]
    c = pe.Circuit(nl, abi)
    c.set_analyze_type(case["at"])
    if "t_step" in kw:
        c.set_tr(kw["t_step"], kw["t_stop"])
    b = c.batch(case["n_inst"])
    for e, attr, v in over:
        b.set_param(e, attr, v)
    if "ac" in kw:
        b.set_ac_sweep(*kw["ac"])
    ok = b.analyze()
    assert ok == bool((want_ok == 1).all())
    good = want_ok == 1
    if case["at"] == pe.AC:
        assert_close(b.ac_solution(), want_x, name)
        assert b.total_solves == int(want_solves.sum())
        return
    st = b.status().reshape(case["n_inst"])
    assert ((st == 0) == good).all(), "failed-instance flags differ from the reference"
    assert (b.newton_iters() == want_solves).all(), "solve_once counts differ from the reference"
    assert_close(b.solution()[good], want_x.real[good], name)
