"""TEST INFRASTRUCTURE: ctypes access to the compiled, unmodified reference (oracle/_ref/libpe_ref.so).

Built by oracle/Makefile from /root/reference (src/dll_main.cpp + the read-only probes of oracle/ref_harness.cpp).
Only tests/, bench.py's cpu_baseline / --impl reference legs and __graft_entry__.smoke() use it, as the checker.
"""
from __future__ import annotations

import ctypes as ct
import os

import numpy as np

import pe_b200
from pe_b200 import CAbi, CircuitBase, Netlist, _p, _PD, _PI, _PSZ, _SZ

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libpe_ref.so")
REF_LIB_FAST = os.path.join(ROOT, "oracle", "_ref", "libpe_ref_fast.so")

_ref = {}


def reference(fast: bool = False) -> CAbi:
    path = REF_LIB_FAST if fast else REF_LIB
    if path in _ref:
        return _ref[path]
    abi = CAbi(path)
    lib = abi.lib
    V = ct.c_void_p
    lib.ref_set_env.argtypes = [V, _PD]
    lib.ref_set_ac_sweep.argtypes = [V, ct.c_int, ct.c_double, ct.c_double, _SZ]
    lib.ref_counts.argtypes = [V, _PSZ, _PSZ]
    lib.ref_get_solution.argtypes = [V, _PD]
    lib.ref_ac_result_count.restype = _SZ
    lib.ref_ac_result_count.argtypes = [V]
    lib.ref_ac_results.argtypes = [V, _PD, _PD]
    lib.ref_tr_duration.restype = ct.c_double
    lib.ref_tr_duration.argtypes = [V]
    lib.ref_reset.argtypes = [V]
    lib.ref_prepare.argtypes = [V]
    lib.ref_solve_once.argtypes = [V]
    lib.ref_analyze_counted.argtypes = [V, ct.POINTER(ct.c_uint64)]
    lib.ref_pin_node_index.restype = ct.c_longlong
    lib.ref_pin_node_index.argtypes = [V, _SZ, _SZ, _SZ]
    lib.ref_branch_index.restype = ct.c_longlong
    lib.ref_branch_index.argtypes = [V, _SZ, _SZ, _SZ]
    lib.ref_mna_nnz.restype = _SZ
    lib.ref_mna_nnz.argtypes = [V]
    lib.ref_mna_dump.argtypes = [V, _PI, _PI, _PD, _PD]
    lib.ref_sizeof_model_base.restype = _SZ
    lib.ref_run_batch.argtypes = [_PI, _SZ, _PI, _SZ, _PD, ct.c_uint32, ct.c_double, ct.c_double, ct.c_int, ct.c_double, ct.c_double, _SZ, _PD,
                                  _SZ, _SZ, _PSZ, ct.POINTER(ct.c_char_p), _PD, ct.c_int, _PD, ct.POINTER(ct.c_uint64), _PI, _PD]
    _ref[path] = abi
    return abi


class RefCircuit(CircuitBase):
    """A phy_engine::circult of the compiled reference, driven through its own C ABI + read-only probes."""

    def __init__(self, nl: Netlist, fast: bool = False):
        super().__init__(reference(fast), nl)

    def set_env(self, V_eps_max=0.0, V_epsr_max=0.0, I_eps_max=0.0, I_epsr_max=0.0, g_min=0.0, r_open=0.0, temperature=27.0, norm_temperature=27.0):
        a = np.array([V_eps_max, V_epsr_max, I_eps_max, I_epsr_max, g_min, r_open, temperature, norm_temperature], dtype=np.float64)
        self._rc(self.abi.lib.ref_set_env(self.h, _p(a, _PD)), "ref_set_env")

    def set_ac_sweep(self, sweep, w0, w1, points):
        self._rc(self.abi.lib.ref_set_ac_sweep(self.h, sweep, w0, w1, points), "ref_set_ac_sweep")

    def counts(self):
        a, b = ct.c_size_t(0), ct.c_size_t(0)
        self.abi.lib.ref_counts(self.h, ct.byref(a), ct.byref(b))
        return a.value, b.value

    def solution(self) -> np.ndarray:
        n = sum(self.counts())
        x = np.zeros(2 * max(n, 1))
        self.abi.lib.ref_get_solution(self.h, _p(x, _PD))
        return x[0:2 * n:2] + 1j * x[1:2 * n:2]

    def ac_results(self):
        k = int(self.abi.lib.ref_ac_result_count(self.h))
        n = sum(self.counts())
        om = np.zeros(max(k, 1))
        x = np.zeros((max(k, 1), n, 2))
        self.abi.lib.ref_ac_results(self.h, _p(om, _PD), _p(x, _PD))
        return om[:k], (x[..., 0] + 1j * x[..., 1])[:k]

    @property
    def tr_duration(self) -> float:
        return float(self.abi.lib.ref_tr_duration(self.h))

    def reset(self):
        self.abi.lib.ref_reset(self.h)

    def prepare(self):
        self.abi.lib.ref_prepare(self.h)

    def solve_once(self) -> bool:
        return self.abi.lib.ref_solve_once(self.h) == 0

    def analyze_counted(self):
        n = ct.c_uint64(0)
        rc = self.abi.lib.ref_analyze_counted(self.h, ct.byref(n))
        return rc == 0, n.value

    def pin_unknown(self, ele: int, pin: int) -> int:
        v, c = self.pos(ele)
        return int(self.abi.lib.ref_pin_node_index(self.h, v, c, pin))

    def branch_unknown(self, ele: int, br: int = 0) -> int:
        v, c = self.pos(ele)
        b = int(self.abi.lib.ref_branch_index(self.h, v, c, br))
        return b if b < 0 else self.counts()[0] + b

    def mna(self):
        """(rows, cols, vals complex, z complex) of the system stamped by the last solve_once()"""
        nnz = int(self.abi.lib.ref_mna_nnz(self.h))
        n = sum(self.counts())
        r = np.zeros(max(nnz, 1), dtype=np.int32)
        c = np.zeros(max(nnz, 1), dtype=np.int32)
        v = np.zeros(2 * max(nnz, 1))
        z = np.zeros(2 * max(n, 1))
        self.abi.lib.ref_mna_dump(self.h, _p(r, _PI), _p(c, _PI), _p(v, _PD), _p(z, _PD))
        return r[:nnz], c[:nnz], (v[0::2] + 1j * v[1::2])[:nnz], (z[0::2] + 1j * z[1::2])[:n]


def run_batch(nl: Netlist, at: int, n_inst: int, overrides=None, t_step=1e-6, t_stop=1e-6, ac=(0, 0.0, 0.0, 0), env=None, threads=None, n_unknowns=None,
              fast=False):
    """The reference's own analyze() over n_inst independent instances (fresh circuit each), `threads` workers.

    overrides: list of (element index, attribute name, values[n_inst]).
    Returns dict(x[n_inst, (points,) n] complex, solves[n_inst], ok[n_inst], seconds)."""
    abi = reference(fast)
    overrides = overrides or []
    e, w, p = nl.arrays()
    n_over = len(overrides)
    comp = np.array([nl.component_index(o[0]) for o in overrides] or [0], dtype=np.uintp)
    names = (ct.c_char_p * max(n_over, 1))(*[o[1].encode() for o in overrides] or [b""])
    vals = np.ascontiguousarray(np.concatenate([np.asarray(o[2], dtype=np.float64).reshape(n_inst) for o in overrides]) if overrides else np.zeros(1))
    envv = np.array(env if env is not None else [0, 0, 0, 0, 0, 0, 27.0, 27.0], dtype=np.float64)
    if n_unknowns is None:
        c = RefCircuit(nl, fast)
        c.set_analyze_type(pe_b200.DC)
        c.prepare()
        n_unknowns = sum(c.counts())
        c.close()
    sweep, w0, w1, points = ac
    npts = points if (at in (pe_b200.AC, pe_b200.ACOP) and sweep != 0 and points > 1) else 1
    x = np.zeros((n_inst, npts, n_unknowns, 2))
    solves = np.zeros(n_inst, dtype=np.uint64)
    ok = np.zeros(n_inst, dtype=np.int32)
    sec = ct.c_double(0.0)
    threads = threads or os.cpu_count() or 1
    abi.lib.ref_run_batch(_p(e, _PI), e.size, _p(w, _PI), len(nl.wires), _p(p, _PD), at, t_step, t_stop, sweep, w0, w1, points, _p(envv, _PD), n_inst, n_over,
                          _p(comp, _PSZ), names, _p(vals, _PD), threads, _p(x, _PD), _p(solves, ct.POINTER(ct.c_uint64)), _p(ok, _PI), ct.byref(sec))
    xc = x[..., 0] + 1j * x[..., 1]
    if npts == 1:
        xc = xc[:, 0, :]
    return {"x": xc, "solves": solves, "ok": ok, "seconds": sec.value, "threads": threads}
