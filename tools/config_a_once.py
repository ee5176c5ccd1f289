#!/usr/bin/env python
"""Config A at full size, one warm-up analyze() and one measured one (for launch lists under ncu; no reference run)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "phy-engine_b200"))
import pe_b200 as pe  # noqa: E402
import workloads as wl  # noqa: E402

nl, info = wl.series_parallel(100_000, 9_000, seed=1)
c = pe.Circuit(nl)
c.set_analyze_type(pe.DC)
b = c.batch(1)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    assert b.analyze(), c.abi.last_error()
print(b.frontal_info())
