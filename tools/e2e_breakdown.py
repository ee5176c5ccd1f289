#!/usr/bin/env python
"""Where the end-to-end step of bench.py spends its time: wall clock around each C-ABI call (device synchronised after each)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "phy-engine_b200"))
import numpy as np, torch
import pe_b200 as pe, workloads as wl

n_inst, sections, steps = 10000, 1000, 100
nl, info = wl.rc_ladder(sections)
rng = np.random.default_rng(1)
items = [(e, "r") for e in info["R"]] + [(e, "c") for e in info["C"]]
P = len(items)
hv = torch.empty((P, n_inst), dtype=torch.float64).pin_memory()
hv.numpy()[:] = np.array([1e3] * sections + [1e-9] * sections)[:, None] * rng.uniform(0.8, 1.2, (P, n_inst))
hx = torch.empty((sections + 2, n_inst), dtype=torch.float64).pin_memory()
c = pe.Circuit(nl); c.set_analyze_type(pe.TR); c.set_tr(1e-8, 1e-8 * (steps - 0.5))
b = c.batch(n_inst); b.set_stream(torch.cuda.current_stream().cuda_stream)
t = b.param_table(items); b.set_params(t, hv.data_ptr()); b.prepare()
acc = {}
def tick(name, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); acc.setdefault(name, []).append((time.perf_counter() - t0) * 1e3); return r
for it in range(5):
    tick("set_params (H2D 160 MB)", lambda: b.set_params(t, hv.data_ptr()))
    tick("reset_state", lambda: b.reset_state())
    tick("analyze (kernel + status)", lambda: b.analyze())
    tick("solution_soa (D2H 80 MB)", lambda: b.solution_soa_into(hx.data_ptr()))
    tick("total_solves", lambda: b.total_solves)
for k, v in acc.items():
    print(f"{k:32s} {np.mean(v[1:]):8.3f} ms")
