#!/usr/bin/env python
"""Per-phase op counts of the tree-scheduled program of config B (no GPU needed): how long the critical stream of every
phase is, and how many streams work in it.  usage: phase_stats.py [streams] [sections]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools")); sys.path.insert(0, os.path.join(ROOT, "phy-engine_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, pe_b200 as pe, workloads as wl, rdis, emuapi

S = int(sys.argv[1]) if len(sys.argv) > 1 else 32
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
nl, info = wl.rc_ladder(N)
c = pe.Circuit(nl, emuapi.emulator())  # host emulator: packs the program exactly as a launch does
c.set_analyze_type(pe.TR)
c.set_tr(1e-8, 1e-8 * 1.5)
b = c.batch(128)
b.set_resident(S, 0, 4)
b.set_workspace(2)
if len(sys.argv) > 3:
    b.set_tuning(int(sys.argv[3]))
items = [(e, "r") for e in info["R"]] + [(e, "c") for e in info["C"]]
vals = np.ascontiguousarray(np.array([1e3] * len(info["R"]) + [1e-9] * len(info["C"]))[:, None] * np.ones((1, 128)))
b.set_params(b.param_table(items), vals.ctypes.data)
b.analyze()
w, so = rdis.program(b, pe.MODE_TR)


def walk1(w, off):
    """one stream per warp: 32-word lines start with two prefetch bitmaps, PE_OP_SKIP pads to the next line"""
    pc = int(off)
    while True:
        if pc % 32 == 0:
            pc += 2
        h = int(w[pc]); op = h & 0xff
        if op == 0:
            return
        if op == 1:
            yield pc, "BAR", 0, 0, 1
            pc += 1
            continue
        if op == 4:
            pc = (pc // 32 + 1) * 32
            continue
        if op in (2, 3):
            rows = 2 + ((h >> 8) & 0x1f) + ((h >> 13) & 0x1f) + ((h >> 18) & 0x3f)
        else:
            rows = (h >> 8) & 0x1f
        yield pc, rdis.NAMES.get(op, f"op{op}"), rows, 0, 2 + rows
        pc += 2 + rows

print("resident info", b.resident_info(pe.MODE_TR))
for sec in (1, 2):
    phases = {}
    for wv in range(so.shape[1]):
        if so[sec, wv] == 0xffffffff:
            continue
        ph = 0
        for pc, name, rows, pcol, ln in walk1(w, so[sec, wv]):
            if name == "BAR":
                ph += 1
                continue
            if name.startswith("op4"):
                continue
            d = phases.setdefault(ph, {})
            d.setdefault(wv, []).append((name, rows))
    print("section", sec)
    if not phases:
        continue
    tot_crit = 0; tot_ops = 0
    for ph in sorted(phases):
        d = phases[ph]
        cnt = {wv: len(v) for wv, v in d.items()}
        words = {wv: sum(r for _, r in v) for wv, v in d.items()}
        crit = max(cnt.values()); tot_crit += crit; tot_ops += sum(cnt.values())
        kinds = {}
        for v in d.values():
            for n, _ in v:
                kinds[n] = kinds.get(n, 0) + 1
        print(f"  phase {ph:2d}: streams {len(cnt):3d}  ops max {crit:4d}  sum {sum(cnt.values()):6d}  operand rows max {max(words.values()):5d}  {kinds}")
    print(f"  critical path {tot_crit} ops, total {tot_ops} ops, parallel efficiency {tot_ops / (tot_crit * so.shape[1]):.3f}")
