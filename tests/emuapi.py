"""TEST INFRASTRUCTURE: loads tests/emu/libpe_emu.so (host side of the product + host replay of the kernel interpreter)."""
import ctypes as ct
import os

import pe_b200

EMU_LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu", "libpe_emu.so")
_emu = None


def emulator():
    global _emu
    if _emu is None:
        _emu = pe_b200.bind_full_abi(pe_b200.CAbi(EMU_LIB))
        _emu.lib.pe_emu_races.restype = ct.c_uint64
        _emu.lib.pe_emu_unbalanced_barriers.restype = ct.c_uint64
    return _emu


def races():
    return int(emulator().lib.pe_emu_races())


def unbalanced():
    return int(emulator().lib.pe_emu_unbalanced_barriers())
