#!/usr/bin/env python
"""GPU check of the stream kernel: bit-identity with the word interpreter on the same one-stream program, parity with the
default (tree-scheduled) path, across a resumed transient.  usage: gpu_stream_check.py [sections] [steps] [instances] [J]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "phy-engine_b200"))
import numpy as np  # noqa: E402
import pe_b200 as pe  # noqa: E402
import workloads as wl  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 100
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
n_inst = int(sys.argv[3]) if len(sys.argv) > 3 else 300
J = int(sys.argv[4]) if len(sys.argv) > 4 else 0


def run(tuning, resident=None):
    nl, info = wl.rc_ladder(N)
    c = pe.Circuit(nl)
    c.set_analyze_type(pe.TR)
    c.set_tr(1e-8, 1e-8 * (steps - 0.5))
    b = c.batch(n_inst)
    if resident:
        b.set_resident(*resident)
        b.set_workspace(2)
    elif J:
        b.set_resident(0, 0, J)
    b.set_tuning(tuning)
    rng = np.random.default_rng(1)
    for e in info["R"]:
        b.set_param(e, "r", wl.sweep_values(rng, 1e3, n_inst))
    for e in info["C"]:
        b.set_param(e, "c", wl.sweep_values(rng, 1e-9, n_inst))
    t0 = time.time()
    ok = b.analyze()
    dt = time.time() - t0
    print(f"tuning {tuning}: ok {ok} kernel {b.last_kernel()} {dt:.2f} s", "" if ok else c.abi.last_error(), flush=True)
    x1 = b.solution()
    ok2 = b.analyze()
    return x1, b.solution(), b.total_solves, ok and ok2


xa1, xa2, sa, oka = run(128 + 32, (1, 0, max(J, 1)))  # word interpreter on the one-stream program
xs1, xs2, ss, oks = run(64)  # stream kernel required
xr1, xr2, sr, okr = run(128 + 32)  # default geometry, interpreter
bad = 0
for a, b_, n, exact in ((xa1, xs1, "one-stream interpreter vs stream", True), (xa2, xs2, "... resumed", True), (xr1, xs1, "default path vs stream", False), (xr2, xs2, "... resumed", False)):
    err = np.abs(a - b_)
    tol = 1e-12 + 1e-9 * np.maximum(np.abs(a), np.abs(b_))
    same = np.array_equal(a, b_)
    okk = same if exact else bool((err <= tol).all())
    bad += 0 if okk else 1
    print(f"{n}: max abs diff {err.max():.3e} bit-identical {same} -> {'ok' if okk else 'FAIL'}")
print("solves", sa, ss, sr, "all ok", oka and oks and okr)
sys.exit(1 if bad or not (oka and oks and okr) or not (sa == ss == sr) else 0)
