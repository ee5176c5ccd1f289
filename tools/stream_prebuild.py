#!/usr/bin/env python
"""Build (or fetch from jit_cache/) the stream kernel's module of the bench workload without a GPU: nvcc cross-compiles, the
cubin travels to the GPU box next to the library.  usage: stream_prebuild.py [sections] [instances] [--source out.inc]"""
import ctypes as ct
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "phy-engine_b200"))
import numpy as np  # noqa: E402
import pe_b200 as pe  # noqa: E402
import workloads as wl  # noqa: E402


def prebuild(sections=1000, instances=10000, source=None):
    nl, info = wl.rc_ladder(sections)
    c = pe.Circuit(nl)
    c.set_analyze_type(pe.TR)
    c.set_tr(1e-8, 1e-8 * 99.5)
    b = c.batch(instances)
    items = [(e, "r") for e in info["R"]] + [(e, "c") for e in info["C"]]
    # per-instance parameters are part of the compiled program's layout (the values do not matter here)
    vals = np.ones((len(items), instances))
    b.set_params(b.param_table(items), vals.ctypes.data)
    b.compile_host()
    lib = b.lib
    lib.circuit_batch_stream_source.restype = ct.c_size_t
    lib.circuit_batch_stream_source.argtypes = [ct.c_void_p, ct.c_int, ct.c_char_p, ct.c_size_t, ct.POINTER(ct.c_uint64)]
    lib.circuit_batch_stream_build.argtypes = [ct.c_void_p, ct.c_int]
    stats = (ct.c_uint64 * 8)()
    n = lib.circuit_batch_stream_source(b.h, pe.MODE_TR, None, 0, stats)
    if source and n:
        buf = ct.create_string_buffer(n + 1)
        lib.circuit_batch_stream_source(b.h, pe.MODE_TR, buf, n, None)
        open(source, "wb").write(buf.raw[:n])
    t0 = time.time()
    rc = lib.circuit_batch_stream_build(b.h, pe.MODE_TR)
    keys = ("tiles", "stage_rows", "loops", "loop_ops", "ops", "rows_fetched", "copies", "rows_stored")
    return rc, n, time.time() - t0, (b.abi.last_error() if rc else ""), dict(zip(keys, [int(v) for v in stats]))


if __name__ == "__main__":
    a = [x for x in sys.argv[1:] if not x.startswith("--")]
    src = None
    if "--source" in sys.argv:
        src = sys.argv[sys.argv.index("--source") + 1]
        a = [x for x in a if x != src]
    rc, n, dt, err, stats = prebuild(int(a[0]) if a else 1000, int(a[1]) if len(a) > 1 else 10000, src)
    print(f"rc={rc} source={n} bytes build={dt:.1f}s {stats} {err}")
