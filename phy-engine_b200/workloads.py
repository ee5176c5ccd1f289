"""Synthetic netlists of the BASELINE.json configs, in create_circuit() wire format.

Every generator returns (Netlist, info) where info names the elements whose parameters are swept per instance.
The reference builds the same shapes through its C++ API in benchmark/ and test/ (cited per function); here they are
expressed once in the C-ABI wire format so that the product and the compiled reference consume identical inputs.
"""
from __future__ import annotations

import numpy as np

import pe_b200 as pe
from pe_b200 import Netlist


def rc_ladder(n_sections: int = 1000, r: float = 1e3, c: float = 1e-9, v: float = 1.0):
    """Config B: VDC -> n x (series R_i, shunt C_i to ground).  n nodes... n_sections+1 nodes + 1 branch unknowns.
    Same topology as test/0005.models/rc_step_tr.cpp generalised to n sections."""
    nl = Netlist()
    g = nl.ground()
    src = nl.add(pe.VDC, v)
    nl.wire(src, 1, g, 0)
    rs, cs = [], []
    prev = (src, 0)
    for _ in range(n_sections):
        ri = nl.add(pe.R, r)
        ci = nl.add(pe.C, c)
        nl.wire(prev[0], prev[1], ri, 0)
        nl.wire(ri, 1, ci, 0)
        nl.wire(ci, 1, g, 0)
        prev = (ri, 1)
        rs.append(ri)
        cs.append(ci)
    return nl, {"src": src, "R": rs, "C": cs}


def rlc_ladder(n_sections: int = 64, r: float = 10.0, l: float = 1e-6, c: float = 1e-9, vp: float = 1.0):
    """Config D: VAC -> n x (series R, series L, shunt C).  2n+1 nodes + n+1 branches (test/0012.ac/ac_omega.cpp shape)."""
    nl = Netlist()
    g = nl.ground()
    src = nl.add(pe.VAC, vp, 50.0, 0.0)
    nl.wire(src, 1, g, 0)
    rs, ls, cs = [], [], []
    prev = (src, 0)
    for _ in range(n_sections):
        ri = nl.add(pe.R, r)
        li = nl.add(pe.L, l)
        ci = nl.add(pe.C, c)
        nl.wire(prev[0], prev[1], ri, 0)
        nl.wire(ri, 1, li, 0)
        nl.wire(li, 1, ci, 0)
        nl.wire(ci, 1, g, 0)
        prev = (li, 1)
        rs.append(ri)
        ls.append(li)
        cs.append(ci)
    return nl, {"src": src, "R": rs, "L": ls, "C": cs}


def rlc_ladder_dc(n_sections: int = 64, r: float = 10.0, l: float = 1e-6, c: float = 1e-9, v: float = 1.0):
    """Step response of an RLC ladder: VDC -> n x (series R, series L, shunt C).  2n + 1 nodes, n + 1 branches; the transient
    program is DOT / CAP_STEP / IND_STEP ops only (no time-dependent source), which the stream kernel covers."""
    nl = Netlist()
    g = nl.ground()
    src = nl.add(pe.VDC, v)
    nl.wire(src, 1, g, 0)
    prev = (src, 0)
    R, L, C = [], [], []
    for _ in range(n_sections):
        er, el, ec = nl.add(pe.R, r), nl.add(pe.L, l), nl.add(pe.C, c)
        nl.wire(prev[0], prev[1], er, 0)
        nl.wire(er, 1, el, 0)
        nl.wire(el, 1, ec, 0)
        nl.wire(ec, 1, g, 0)
        prev = (el, 1)
        R.append(er)
        L.append(el)
        C.append(ec)
    return nl, {"V": src, "R": R, "L": L, "C": C}


def diode_resistor(v: float = 1.0, r: float = 1e3, n_diodes: int = 1):
    """Config C (c1): V - R - n series diodes to ground (test/0011.nonlinear/op_pn_junction.cpp for n = 1)."""
    nl = Netlist()
    g = nl.ground()
    src = nl.add(pe.VDC, v)
    rr = nl.add(pe.R, r)
    nl.wire(src, 1, g, 0)
    nl.wire(src, 0, rr, 0)
    prev = (rr, 1)
    ds = []
    for _ in range(n_diodes):
        d = nl.add(pe.PN, *pe.PN_DEFAULT)
        nl.wire(prev[0], prev[1], d, 0)
        prev = (d, 1)
        ds.append(d)
    nl.wire(prev[0], prev[1], g, 0)
    return nl, {"src": src, "R": rr, "D": ds}


def diode_ladder(n_stages: int = 16, v: float = 5.0, r: float = 1e3):
    """Config C (c4): V -> n x (series R_i, shunt diode_i to ground): n+1 nodes + 1 branch, n non-linear devices."""
    nl = Netlist()
    g = nl.ground()
    src = nl.add(pe.VDC, v)
    nl.wire(src, 1, g, 0)
    prev = (src, 0)
    rs, ds = [], []
    for _ in range(n_stages):
        ri = nl.add(pe.R, r)
        di = nl.add(pe.PN, *pe.PN_DEFAULT)
        nl.wire(prev[0], prev[1], ri, 0)
        nl.wire(ri, 1, di, 0)
        nl.wire(di, 1, g, 0)
        prev = (ri, 1)
        rs.append(ri)
        ds.append(di)
    return nl, {"src": src, "R": rs, "D": ds}


def npn_stage(vb: float = 0.65, vcc: float = 5.0, rc: float = 1e3):
    """Config C (c2): NPN with a voltage-driven base, RC to VCC, emitter grounded (the
    resistor-biased stage does not converge in the reference)."""
    nl = Netlist()
    g = nl.ground()
    vbs = nl.add(pe.VDC, vb)
    vcs = nl.add(pe.VDC, vcc)
    r = nl.add(pe.R, rc)
    q = nl.add(pe.NPN, 1e-16, 1.0, 100.0, 27.0, 1.0)
    nl.wire(vbs, 1, g, 0)
    nl.wire(vcs, 1, g, 0)
    nl.wire(vbs, 0, q, 0)
    nl.wire(vcs, 0, r, 0)
    nl.wire(r, 1, q, 1)
    nl.wire(q, 2, g, 0)
    return nl, {"Vb": vbs, "Vcc": vcs, "R": r, "Q": q}


def npn_resistor_biased(vcc: float = 5.0, rb: float = 1e5, rc: float = 1e3):
    """Failure-parity case: the reference's un-limited BJT diverges on this."""
    nl = Netlist()
    g = nl.ground()
    vcs = nl.add(pe.VDC, vcc)
    r_b = nl.add(pe.R, rb)
    r_c = nl.add(pe.R, rc)
    q = nl.add(pe.NPN, 1e-16, 1.0, 100.0, 27.0, 1.0)
    nl.wire(vcs, 1, g, 0)
    nl.wire(vcs, 0, r_b, 0)
    nl.wire(vcs, 0, r_c, 0)
    nl.wire(r_b, 1, q, 0)
    nl.wire(r_c, 1, q, 1)
    nl.wire(q, 2, g, 0)
    return nl, {"Vcc": vcs, "Rb": r_b, "Rc": r_c, "Q": q}


def cmos_stage(vdd: float = 5.0, vg: float = 2.0, rd: float = 1e3, with_pmos: bool = True):
    """Config C (c3): NMOS common source; load = PMOS (gate grounded-ish via Vgp) in parallel with RD."""
    nl = Netlist()
    g = nl.ground()
    vd = nl.add(pe.VDC, vdd)
    vgs = nl.add(pe.VDC, vg)
    r = nl.add(pe.R, rd)
    mn = nl.add(pe.NMOS, 1e-3, 0.01, 1.0)
    nl.wire(vd, 1, g, 0)
    nl.wire(vgs, 1, g, 0)
    nl.wire(vd, 0, r, 0)
    nl.wire(r, 1, mn, 0)
    nl.wire(vgs, 0, mn, 1)
    nl.wire(mn, 2, g, 0)
    info = {"Vdd": vd, "Vg": vgs, "R": r, "MN": mn}
    if with_pmos:
        vgp = nl.add(pe.VDC, vdd - 2.0)
        mp = nl.add(pe.PMOS, 5e-4, 0.02, 1.0)
        nl.wire(vgp, 1, g, 0)
        nl.wire(mp, 2, vd, 0)  # source at VDD
        nl.wire(mp, 1, vgp, 0)
        nl.wire(mp, 0, mn, 0)  # drains tied
        info.update({"Vgp": vgp, "MP": mp})
    return nl, info


def random_links(n_nodes: int = 1000, n_links: int = 10, seed: int = 1, r: float = 1e3):
    """benchmark/0001.models/100000_random_links_cpu.cpp shape: chain of n 1 kOhm resistors from a 1 V source plus
    random chord resistors (seeded here; the reference benchmark draws them from a non-deterministic engine)."""
    rng = np.random.default_rng(seed)
    nl = Netlist()
    g = nl.ground()
    src = nl.add(pe.VDC, 1.0)
    nl.wire(src, 1, g, 0)
    chain = []
    prev = (src, 0)
    for _ in range(n_nodes):
        ri = nl.add(pe.R, r)
        nl.wire(prev[0], prev[1], ri, 0)
        prev = (ri, 1)
        chain.append(ri)
    load = nl.add(pe.R, r)
    nl.wire(prev[0], prev[1], load, 0)
    nl.wire(load, 1, g, 0)
    links = []
    for _ in range(n_links):
        a, b = rng.integers(0, n_nodes, size=2)
        if a == b:
            continue
        lk = nl.add(pe.R, r)
        nl.wire(lk, 0, chain[a], 1)
        nl.wire(lk, 1, chain[b], 1)
        links.append(lk)
    return nl, {"src": src, "chain": chain, "links": links, "load": load}


def sweep_values(rng: np.random.Generator, nominal: float, n_inst: int, lo: float = 0.8, hi: float = 1.2) -> np.ndarray:
    return nominal * rng.uniform(lo, hi, size=n_inst)


def bridge_rectifier(v: float = 5.0, r_load: float = 1e3, r_src: float = 10.0):
    """Full bridge rectifier (element 54 = four default PN junctions) between a source with series resistance and a
    resistive load (model/models/non-linear/full_bridge_rectifier.h)."""
    nl = Netlist()
    g = nl.ground()
    src = nl.add(pe.VDC, v)
    rs = nl.add(pe.R, r_src)
    br = nl.add(pe.BRIDGE)
    rl = nl.add(pe.R, r_load)
    nl.wire(src, 1, g, 0)
    nl.wire(src, 0, rs, 0)
    nl.wire(rs, 1, br, 0)   # A
    nl.wire(br, 1, g, 0)    # B
    nl.wire(br, 2, rl, 0)   # +
    nl.wire(br, 3, rl, 1)   # -
    return nl, {"V": src, "Rs": rs, "B": br, "R": rl}


def transformer_stage(vac: bool = False, n: float = 4.0):
    """Source with series resistance -> ideal transformer (element 14, model/models/linear/transformer.h) -> RC load on
    the secondary; the secondary's low side is grounded through a resistor so that every node has a DC path."""
    nl = Netlist()
    g = nl.ground()
    src = nl.add(pe.VAC, 3.0, 2e5, 30.0) if vac else nl.add(pe.VDC, 3.0)
    rs = nl.add(pe.R, 50.0)
    tx = nl.add(pe.TRANSFORMER, n)
    rl = nl.add(pe.R, 1e3)
    cl = nl.add(pe.C, 1e-8)
    rg = nl.add(pe.R, 10.0)
    nl.wire(src, 1, g, 0)
    nl.wire(src, 0, rs, 0)
    nl.wire(rs, 1, tx, 0)   # P
    nl.wire(tx, 1, g, 0)    # Q
    nl.wire(tx, 2, rl, 0)   # S
    nl.wire(tx, 3, rg, 0)   # T
    nl.wire(rg, 1, g, 0)
    nl.wire(rl, 1, tx, 3)
    nl.wire(cl, 0, tx, 2)
    nl.wire(cl, 1, tx, 3)
    return nl, {"V": src, "Rs": rs, "TX": tx, "R": rl, "C": cl}


def center_tap_stage(vac: bool = False, n_total: float = 2.0):
    """Source with series resistance -> centre-tapped transformer (element 16, transformer_center_tap.h) -> two unequal
    resistive half loads (one with a shunt capacitor) against the grounded centre tap."""
    nl = Netlist()
    g = nl.ground()
    src = nl.add(pe.VAC, 3.0, 2e5, 30.0) if vac else nl.add(pe.VDC, 3.0)
    rs = nl.add(pe.R, 50.0)
    tx = nl.add(pe.TRANSFORMER_CT, n_total)
    r1, r2 = nl.add(pe.R, 1e3), nl.add(pe.R, 2.2e3)
    c1 = nl.add(pe.C, 1e-8)
    rg = nl.add(pe.R, 5.0)
    nl.wire(src, 1, g, 0)
    nl.wire(src, 0, rs, 0)
    nl.wire(rs, 1, tx, 0)   # P
    nl.wire(tx, 1, g, 0)    # Q
    nl.wire(tx, 3, rg, 0)   # CT through a small resistor to ground
    nl.wire(rg, 1, g, 0)
    nl.wire(tx, 2, r1, 0)   # S1
    nl.wire(r1, 1, tx, 3)
    nl.wire(c1, 0, tx, 2)
    nl.wire(c1, 1, tx, 3)
    nl.wire(tx, 4, r2, 0)   # S2
    nl.wire(r2, 1, tx, 3)
    return nl, {"V": src, "Rs": rs, "TX": tx, "R1": r1, "R2": r2, "C": c1}


def diode_rc(vac: bool = True):
    """A diode charging an RC load next to an independent resistive divider: a non-linear transient whose plain capacitor
    takes the folded (first Newton iteration only) companion update."""
    nl = Netlist()
    g = nl.ground()
    src = nl.add(pe.VAC, 8.0, 2.5e5, 0.0) if vac else nl.add(pe.VDC, 5.0)
    ra, rb = nl.add(pe.R, 1e3), nl.add(pe.R, 9e3)
    d = nl.add(pe.PN, *pe.PN_DEFAULT)
    rl = nl.add(pe.R, 470.0)
    cl = nl.add(pe.C, 2e-9)
    nl.wire(src, 1, g, 0)
    nl.wire(src, 0, ra, 0)
    nl.wire(ra, 1, rb, 0)
    nl.wire(rb, 1, g, 0)
    nl.wire(d, 0, ra, 1)
    nl.wire(d, 1, rl, 0)
    nl.wire(rl, 1, g, 0)
    nl.wire(cl, 0, d, 1)
    nl.wire(cl, 1, g, 0)
    return nl, {"V": src, "Ra": ra, "Rb": rb, "D": d, "R": rl, "C": cl}


def coupled_inductors_stage(vac: bool = False):
    """Source with series resistance -> winding 1 of a pair of coupled inductors (element 15, coupled_inductors.h); winding 2
    drives an RC load, its low side is grounded through a small resistor."""
    nl = Netlist()
    g = nl.ground()
    src = nl.add(pe.VAC, 3.0, 2e5, 30.0) if vac else nl.add(pe.VDC, 3.0)
    rs = nl.add(pe.R, 50.0)
    kl = nl.add(pe.COUPLED_INDUCTORS, 1e-3, 4e-4, 0.95)
    rl = nl.add(pe.R, 1e3)
    cl = nl.add(pe.C, 1e-8)
    rg = nl.add(pe.R, 10.0)
    nl.wire(src, 1, g, 0)
    nl.wire(src, 0, rs, 0)
    nl.wire(rs, 1, kl, 0)   # P1
    nl.wire(kl, 1, g, 0)    # P2
    nl.wire(kl, 2, rl, 0)   # S1
    nl.wire(kl, 3, rg, 0)   # S2
    nl.wire(rg, 1, g, 0)
    nl.wire(rl, 1, kl, 3)
    nl.wire(cl, 0, kl, 2)
    nl.wire(cl, 1, kl, 3)
    return nl, {"V": src, "Rs": rs, "K": kl, "R": rl, "C": cl}


def generator_rc(kind: str = "square"):
    """A time-domain generator (elements 20-23, model/models/generator/*.h) driving an RC low-pass."""
    nl = Netlist()
    g = nl.ground()
    if kind == "sawtooth":
        src = nl.add(pe.GEN_SAWTOOTH, 4.0, -1.0, 2.5e5, 0.7)
    elif kind == "square":
        src = nl.add(pe.GEN_SQUARE, 4.0, -1.0, 2.5e5, 0.3, 0.7)
    elif kind == "pulse":
        src = nl.add(pe.GEN_PULSE, 4.0, -1.0, 2.5e5, 0.6, 0.7, 4e-7, 6e-7)
    else:
        src = nl.add(pe.GEN_TRIANGLE, 4.0, -1.0, 2.5e5, 0.7)
    r = nl.add(pe.R, 1e3)
    c = nl.add(pe.C, 1e-9)
    nl.wire(src, 1, g, 0)
    nl.wire(src, 0, r, 0)
    nl.wire(r, 1, c, 0)
    nl.wire(c, 1, g, 0)
    return nl, {"G": src, "R": r, "C": c}


def pulse_rc_ladder(n_sections: int = 100, r: float = 1e3, c: float = 1e-9):
    """A pulse generator (element 22, generator/pulse.h) into an RC ladder: the interconnect-delay shape of config B with a
    time-dependent source (one GEN_EVAL per solve next to the DOT / CAP_STEP ops)."""
    nl = Netlist()
    g = nl.ground()
    src = nl.add(pe.GEN_PULSE, 4.0, -1.0, 2.5e6, 0.6, 0.7, 4e-8, 6e-8)
    nl.wire(src, 1, g, 0)
    prev = (src, 0)
    R, C = [], []
    for _ in range(n_sections):
        er, ec = nl.add(pe.R, r), nl.add(pe.C, c)
        nl.wire(prev[0], prev[1], er, 0)
        nl.wire(er, 1, ec, 0)
        nl.wire(ec, 1, g, 0)
        prev = (er, 1)
        R.append(er)
        C.append(ec)
    return nl, {"G": src, "R": R, "C": C}


def relay_stage(vac: bool = False, v_ctl: float = 6.0):
    """A relay (element 18, controller/relay.h): the coil hangs on a resistive divider driven by a control source (DC, or a
    sine that crosses both hysteresis thresholds), the contact switches a 5 V supply onto an RC load."""
    nl = Netlist()
    g = nl.ground()
    ctl = nl.add(pe.VAC, 8.0, 2.5e5, 0.0) if vac else nl.add(pe.VDC, v_ctl)
    ra, rb = nl.add(pe.R, 1e3), nl.add(pe.R, 9e3)
    ry = nl.add(pe.RELAY, 5.0, 3.0)
    vs = nl.add(pe.VDC, 5.0)
    rl = nl.add(pe.R, 470.0)
    cl = nl.add(pe.C, 2e-9)
    nl.wire(ctl, 1, g, 0)
    nl.wire(ctl, 0, ra, 0)
    nl.wire(ra, 1, rb, 0)
    nl.wire(rb, 1, g, 0)
    nl.wire(ry, 0, ra, 1)   # C+ on the divider tap
    nl.wire(ry, 1, g, 0)    # C-
    nl.wire(vs, 1, g, 0)
    nl.wire(vs, 0, ry, 2)   # A
    nl.wire(ry, 3, rl, 0)   # B
    nl.wire(rl, 1, g, 0)
    nl.wire(cl, 0, ry, 3)
    nl.wire(cl, 1, g, 0)
    return nl, {"Vctl": ctl, "Ra": ra, "Rb": rb, "RY": ry, "Vs": vs, "R": rl, "C": cl}


def linear_zoo(vac: bool = False):
    """One netlist with every in-scope linear element: R, C, L, VDC / VAC, IDC, IAC, VCCS, VCVS, CCCS, CCVS, op-amp and a
    closed + an open single-pole switch (stamps of SURVEY.md Appendix A).  Every node has a DC path to ground."""
    nl = Netlist()
    g = nl.ground()
    src = nl.add(pe.VAC, 5.0, 1e5, 30.0) if vac else nl.add(pe.VDC, 5.0)
    r1, r2 = nl.add(pe.R, 1e3), nl.add(pe.R, 2e3)
    nl.wire(src, 1, g, 0)
    nl.wire(src, 0, r1, 0)
    nl.wire(r1, 1, r2, 0)  # n2
    nl.wire(r2, 1, g, 0)
    # VCCS: pins S, T, P, Q; controlled by V(n2)
    gm = nl.add(pe.VCCS, 1e-3)
    r3 = nl.add(pe.R, 1e3)
    nl.wire(gm, 2, r1, 1)
    nl.wire(gm, 3, g, 0)
    nl.wire(gm, 0, r3, 0)  # n3
    nl.wire(gm, 1, g, 0)
    nl.wire(r3, 1, g, 0)
    # VCVS (x2 of n2) driving R4 - sense branch of the CCCS - R5
    ev = nl.add(pe.VCVS, 2.0)
    r4, r5 = nl.add(pe.R, 1e3), nl.add(pe.R, 1e3)
    nl.wire(ev, 2, r1, 1)
    nl.wire(ev, 3, g, 0)
    nl.wire(ev, 0, r4, 0)  # n4
    nl.wire(ev, 1, g, 0)
    fc = nl.add(pe.CCCS, 3.0)  # pins S, T (output), P, Q (sense)
    nl.wire(r4, 1, fc, 2)  # n5
    nl.wire(fc, 3, r5, 0)  # n6
    nl.wire(r5, 1, g, 0)
    r6 = nl.add(pe.R, 500.0)
    nl.wire(fc, 0, r6, 0)  # n7
    nl.wire(fc, 1, g, 0)
    nl.wire(r6, 1, g, 0)
    # CCVS sensing the current through R7 (fed from n3)
    hv = nl.add(pe.CCVS, 250.0)  # pins S, T (output), P, Q (sense)
    r7, r8 = nl.add(pe.R, 2e3), nl.add(pe.R, 1e3)
    nl.wire(r7, 0, r3, 0)
    nl.wire(r7, 1, hv, 2)  # n8
    nl.wire(hv, 3, g, 0)
    nl.wire(hv, 0, r8, 0)  # n9
    nl.wire(hv, 1, g, 0)
    nl.wire(r8, 1, g, 0)
    # inverting op-amp stage on n2: pins +, -, OUT+, OUT-
    oa = nl.add(pe.OPAMP, 1e5)
    ri, rf = nl.add(pe.R, 1e3), nl.add(pe.R, 4.7e3)
    nl.wire(ri, 0, r1, 1)
    nl.wire(ri, 1, oa, 1)  # inverting input
    nl.wire(oa, 0, g, 0)
    nl.wire(oa, 3, g, 0)
    nl.wire(rf, 0, oa, 1)
    nl.wire(rf, 1, oa, 2)  # output
    rl = nl.add(pe.R, 1e4)
    nl.wire(rl, 0, oa, 2)
    nl.wire(rl, 1, g, 0)
    # switches: closed one in series with R9 from n1, open one bridging R9
    s1, s2 = nl.add(pe.SWITCH, 1.0), nl.add(pe.SWITCH, 0.0)
    r9, r10 = nl.add(pe.R, 1e3), nl.add(pe.R, 1e3)
    nl.wire(s1, 0, src, 0)
    nl.wire(s1, 1, r9, 0)
    nl.wire(r9, 1, r10, 0)
    nl.wire(r10, 1, g, 0)
    nl.wire(s2, 0, r9, 0)
    nl.wire(s2, 1, r9, 1)
    # RLC tail with current sources
    l1, c1, r11 = nl.add(pe.L, 1e-3), nl.add(pe.C, 1e-7), nl.add(pe.R, 100.0)
    idc = nl.add(pe.IDC, 1e-3)
    iac = nl.add(pe.IAC, 2e-3, 1e5, -45.0)
    nl.wire(l1, 0, r10, 0)
    nl.wire(l1, 1, r11, 0)  # n_l
    nl.wire(r11, 1, g, 0)
    nl.wire(c1, 0, l1, 1)
    nl.wire(c1, 1, g, 0)
    nl.wire(idc, 0, g, 0)
    nl.wire(idc, 1, l1, 1)
    nl.wire(iac, 0, g, 0)
    nl.wire(iac, 1, l1, 1)
    return nl, {"src": src, "R": [r1, r2, r3, r4, r5, r6, r7, r8, ri, rf, rl, r9, r10, r11], "G": gm, "E": ev, "F": fc, "H": hv, "OA": oa, "L": l1, "C": c1, "IDC": idc,
                "IAC": iac}


def flash_adc(levels: int = 16, vref: float = 5.0, r: float = 1e3):
    """Config E shape (test/0028.16b_adc): Vref across `levels` equal resistors (levels taps incl. the top), an input
    source Vin, and levels - 1 comparators (A = vin, B = tap k, output on a pure digital node).
    17 analog nodes + 2 source branches = 19 unknowns for levels = 16."""
    nl = Netlist()
    g = nl.ground()
    vr = nl.add(pe.VDC, vref)
    vin = nl.add(pe.VDC, 2.5)
    nl.wire(vr, 1, g, 0)
    nl.wire(vin, 1, g, 0)
    rs = [nl.add(pe.R, r) for _ in range(levels)]
    nl.wire(rs[0], 0, g, 0)
    for k in range(1, levels):
        nl.wire(rs[k - 1], 1, rs[k], 0)  # tap k
    nl.wire(rs[-1], 1, vr, 0)
    cmps = []
    for k in range(1, levels):
        cm = nl.add(pe.COMPARATOR, 0.0, 5.0)
        nl.wire(cm, 0, vin, 0)
        nl.wire(cm, 1, rs[k], 0)  # tap k = junction of rs[k-1] and rs[k]
        nl.wire(cm, 2, cm, 2)  # a net of its own for the output pin: a pure digital node
        cmps.append(cm)
    return nl, {"Vref": vr, "Vin": vin, "R": rs, "CMP": cmps}


def series_parallel(n_ring: int = 100_000, n_merge: int = 9_000, seed: int = 1, v: float = 3.0):
    """Config A: a seeded clone of benchmark/series_parallel.cpp (the reference draws from a non-deterministic engine,
    series_parallel.cpp:8): ground - R2 - R3_0 - ... - R3_{n-1} - node4, VDC between node2 (+) and node4 (-), R1 from node2 to
    ground, R4 left unconnected as in the source, resistances ~ U(1e-5, 1e5) Ohm; then n_merge random pairs of chain nodes are
    merged (merge_node = a wire between them in the C ABI's netlist format).  Returns (netlist, info)."""
    rng = np.random.default_rng(seed)
    nl = Netlist()
    g = nl.ground()
    v12 = rng.uniform(1e-5, 1e5, 2)
    r1 = nl.add(pe.R, float(v12[0]))
    r2 = nl.add(pe.R, float(v12[1]))
    r4 = nl.add(pe.R, float(rng.uniform(1e-5, 1e5)))  # never connected
    src = nl.add(pe.VDC, v)
    nl.wire(r1, 1, g, 0)
    nl.wire(r2, 0, g, 0)
    vals = rng.uniform(1e-5, 1e5, n_ring)
    chain = []
    prev = r2
    for i in range(n_ring):
        r3 = nl.add(pe.R, float(vals[i]))
        nl.wire(r3, 0, prev, 1)
        chain.append(r3)
        prev = r3
    nl.wire(src, 0, r1, 0)
    nl.wire(src, 1, prev, 1)
    pairs = rng.integers(0, n_ring, size=(n_merge, 2))
    for a, b in pairs:
        if a != b:
            nl.wire(chain[int(a)], 0, chain[int(b)], 0)
    return nl, {"V": src, "R1": r1, "R2": r2, "R4": r4, "chain": chain, "res": [r1, r2] + chain, "res_values": np.concatenate((v12, vals))}
