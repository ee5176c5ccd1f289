#!/usr/bin/env python
"""Generates tests/golden/cases.npz from the compiled, unmodified reference (oracle/_ref/libpe_ref.so).

Run in the build container (where /root/reference exists and `make -C oracle` has produced the library):
    python tests/golden/make_golden.py
Each case = a seeded workload of phy-engine_b200/workloads.py, the per-instance parameter overrides, and the reference's
answers (final state of every unknown, solve_once counts, ok flags).  tests/test_golden.py replays the cases on the
product (CUDA on the GPU box, the host emulator in the CPU suite) without needing the reference at run time.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "phy-engine_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import golden_cases  # noqa: E402
import refapi  # noqa: E402


def main():
    out = {}
    for name, case in golden_cases.CASES.items():
        nl, over, kw = golden_cases.build(name)
        r = refapi.run_batch(nl, case["at"], case["n_inst"], over, **kw)
        out[name + "/x"] = r["x"]
        out[name + "/solves"] = r["solves"]
        out[name + "/ok"] = r["ok"]
        print(f"{name}: x{r['x'].shape} solves {int(r['solves'].sum())} ok {int((r['ok'] == 1).sum())}/{case['n_inst']}")
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "cases.npz"), **out)


if __name__ == "__main__":
    main()
