#!/usr/bin/env python
"""Build (or fetch from jit_cache/) the specialised kernel of the bench workload without a GPU: nvcc cross-compiles, the
cubin travels to the GPU box next to the library.  usage: jit_prebuild.py [sections] [streams] [--source out.inc]"""
import ctypes as ct
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "phy-engine_b200"))
import numpy as np  # noqa: E402
import pe_b200 as pe  # noqa: E402
import workloads as wl  # noqa: E402


def prebuild(sections=1000, streams=0, source=None, instances=128):
    nl, info = wl.rc_ladder(sections)
    c = pe.Circuit(nl)
    c.set_analyze_type(pe.TR)
    c.set_tr(1e-8, 1e-8 * 99.5)
    b = c.batch(instances)
    if streams:
        b.set_resident(streams, 0, 4)
    b.set_workspace(2)
    items = [(e, "r") for e in info["R"]] + [(e, "c") for e in info["C"]]
    # per-instance parameters are part of the compiled program's layout (the values do not matter here)
    vals = np.ones((len(items), instances))
    b.set_params(b.param_table(items), vals.ctypes.data)
    b.compile_host()
    lib = b.lib
    lib.circuit_batch_jit_source.restype = ct.c_size_t
    lib.circuit_batch_jit_source.argtypes = [ct.c_void_p, ct.c_int, ct.c_char_p, ct.c_size_t]
    lib.circuit_batch_jit_build.argtypes = [ct.c_void_p, ct.c_int, ct.c_int]
    n = lib.circuit_batch_jit_source(b.h, pe.MODE_TR, None, 0)
    if source:
        buf = ct.create_string_buffer(n + 1)
        lib.circuit_batch_jit_source(b.h, pe.MODE_TR, buf, n)
        open(source, "wb").write(buf.raw[:n])
    info_r = b.resident_info(pe.MODE_TR)
    cl = 2 if info_r["streams"] * 32 > 512 else 1
    t0 = time.time()
    rc = lib.circuit_batch_jit_build(b.h, pe.MODE_TR, cl)
    return rc, n, time.time() - t0, (b.abi.last_error() if rc else "")


if __name__ == "__main__":
    a = [x for x in sys.argv[1:] if not x.startswith("--")]
    src = None
    if "--source" in sys.argv:
        src = sys.argv[sys.argv.index("--source") + 1]
        a = [x for x in a if x != src]
    rc, n, dt, err = prebuild(int(a[0]) if a else 1000, int(a[1]) if len(a) > 1 else 0, src)
    print(f"rc={rc} source={n} bytes build={dt:.1f}s {err}")
