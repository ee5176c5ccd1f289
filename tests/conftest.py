"""pytest configuration: markers, library builds, shared fixtures.

`-m "not gpu"` runs here (no GPU): oracle-vs-golden, host logic, ABI surface, kernel-program emulation.
`-m gpu` runs on a B200: parity of the CUDA path (through the C ABI) against the compiled reference / golden vectors.
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "phy-engine_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _make(directory, *targets):
    subprocess.run(["make", "-C", os.path.join(ROOT, directory), "-j8", *targets], check=True, stdout=subprocess.DEVNULL)


@pytest.fixture(scope="session", autouse=True)
def _built():
    # product library (nvcc cross-compiles without a GPU); prebuilt .so travels to the GPU box
    if not os.path.exists(os.path.join(ROOT, "phy-engine_b200", "libphyengine_b200.so")) or os.path.exists("/usr/local/cuda/bin/nvcc"):
        _make("phy-engine_b200")
    # compiled reference: only where the reference sources exist; the GPU box uses the prebuilt file
    if os.path.isdir("/root/reference") :
        _make("oracle")
    if os.path.isdir(os.path.join(ROOT, "tests", "emu")):
        _make("tests/emu")
    yield


@pytest.fixture(scope="session")
def ref():
    import refapi

    if not os.path.exists(refapi.REF_LIB):
        pytest.skip("compiled reference (oracle/_ref/libpe_ref.so) not available")
    return refapi.reference()


def has_gpu():
    import pe_b200

    try:
        return pe_b200.device_count() > 0
    except Exception:
        return False
