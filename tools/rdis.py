#!/usr/bin/env python
"""Disassembler / statistics for packed resident programs (tooling for DESIGN.md and tuning)."""
import ctypes as ct
import sys
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "phy-engine_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pe_b200 as pe  # noqa: E402

NAMES = {0: "END", 1: "BAR", 2: "DOT", 3: "CDOT", 10: "RECIP", 11: "MUL", 12: "SUB", 13: "COPY", 14: "VSIN", 15: "SINCOS", 16: "MUL2DIV", 17: "KMUT", 20: "CAP_STEP",
         21: "IND_STEP", 22: "RELAY_EVAL", 23: "KIND_STEP", 24: "GEN_EVAL", 30: "PN_PREP", 31: "PN_EVAL", 32: "PN_STEP", 33: "PN_ACCAP", 40: "BJT_PREP", 41: "BJT_EVAL", 50: "NMOS_EVAL", 51: "PMOS_EVAL"}


def program(b, mode):
    lib = b.lib
    lib.circuit_batch_program_words.restype = ct.c_size_t
    lib.circuit_batch_program_words.argtypes = [ct.c_void_p, ct.c_int]
    lib.circuit_batch_program_copy.argtypes = [ct.c_void_p, ct.c_int, ct.c_void_p]
    lib.circuit_batch_resident_secoff.restype = ct.c_size_t
    lib.circuit_batch_resident_secoff.argtypes = [ct.c_void_p, ct.c_int, ct.c_void_p]
    n = lib.circuit_batch_program_words(b.h, mode)
    w = np.zeros(n, dtype=np.uint32)
    lib.circuit_batch_program_copy(b.h, mode, w.ctypes.data_as(ct.c_void_p))
    k = lib.circuit_batch_resident_secoff(b.h, mode, None)
    so = np.zeros(k, dtype=np.uint32)
    lib.circuit_batch_resident_secoff(b.h, mode, so.ctypes.data_as(ct.c_void_p))
    return w, so.reshape(6, -1)  # rows 0-2: main stream of prep / step / iter, rows 3-5: side stream


def walk(w, off, C):
    """yields (offset, opname, n_rows, n_percol_rows, length) for one warp stream"""
    pc = int(off)
    while True:
        h = int(w[pc])
        op = h & 0xff
        if op == 0:
            return
        if op == 1:
            yield pc, "BAR", 0, 0, 1
            pc += 1
            continue
        mask = int(w[pc + 1])
        if op == 2 or op == 3:
            rows = 2 + ((h >> 8) & 0x1f) + ((h >> 13) & 0x1f) + ((h >> 18) & 0x3f)
        else:
            rows = (h >> 8) & 0x1f
        pcol = bin(mask).count("1")
        ln = 2 + rows + pcol * (C - 1)
        yield pc, NAMES.get(op, f"op{op}"), rows, pcol, ln
        pc += ln


def stats(b, mode, ig):
    w, so = program(b, mode)
    C = max(1, 32 // ig)
    out = []
    for sec in range(3):
        for wv in range(so.shape[1]):
            if so[sec, wv] == 0xffffffff:
                continue
            ops = list(walk(w, so[sec, wv], C))
            nb = sum(1 for o in ops if o[1] == "BAR")
            nv = len(ops) - nb
            rows = sum(o[2] for o in ops)
            pcol = sum(o[3] for o in ops)
            words = sum(o[4] for o in ops)
            out.append((sec, wv, int(so[sec, wv]), nv, nb, rows, pcol, words))
    return out
