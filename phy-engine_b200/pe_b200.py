"""ctypes binding over the C ABI of libphyengine_b200.so (include/phy_engine_b200.h).

This is the thin Python caller used by tests/, bench.py and __graft_entry__.py; the product is the shared
library.  `Netlist` builds the wire-format arrays of create_circuit() (reference: include/phy_engine/dll_api.h:51-150,
src/dll_main.cpp:1530-1956), `CAbi` wraps the reference-compatible Part 1 of the ABI for ANY library that exports
it (the product library here, the compiled reference in tests/), `Circuit` / `Batch` wrap handles of the product.

There is no CPU path: if the shared library is missing, loading raises; if no CUDA device is visible, analyze()
returns an error (see phy_engine_last_error()).
"""
from __future__ import annotations

import ctypes as ct
import os
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libphyengine_b200.so")

# phy_engine_element_code (dll_api.h:51-135), in-scope subset
GROUND, R, C, L, VDC, VAC, IDC, IAC, VCCS, VCVS, CCCS, CCVS, SWITCH, PN = range(14)
COMPARATOR = 19  # pins A, B (analog), o (digital output); properties Ll, Hl
OPAMP = 17
GEN_SAWTOOTH, GEN_SQUARE, GEN_PULSE, GEN_TRIANGLE = 20, 21, 22, 23  # generators: Vh, Vl, freq [, duty], phase(rad) [, tr, tf]
COUPLED_INDUCTORS = 15  # pins P1, P2 (winding 1), S1, S2 (winding 2); properties L1, L2, k
RELAY = 18  # pins C+, C- (coil), A, B (contact); properties Von, Voff
TRANSFORMER_CT = 16  # centre-tapped transformer: pins P, Q, S1, CT, S2; property n_total = Vp / V(S1 - S2)
TRANSFORMER = 14  # ideal transformer: pins P, Q (primary), S, T (secondary); property n = Vp / Vs
NPN, PNP, NMOS, PMOS = 50, 51, 52, 53
BRIDGE = 54  # full bridge rectifier (pins A, B, +, -)

# analyze_type (circuits/analyze.h)
OP, DC, AC, ACOP, TR, TROP = range(6)
SWEEP_SINGLE, SWEEP_LINEAR, SWEEP_LOG = range(3)
MODE_DC, MODE_TR, MODE_TROP, MODE_AC = range(4)

PROPS = {GROUND: 0, R: 1, C: 1, L: 1, VDC: 1, VAC: 3, IDC: 1, IAC: 3, VCCS: 1, VCVS: 1, CCCS: 1, CCVS: 1, SWITCH: 1,
         PN: 9, OPAMP: 1, NPN: 5, PNP: 5, NMOS: 3, PMOS: 3, BRIDGE: 0, COMPARATOR: 2, TRANSFORMER: 1, TRANSFORMER_CT: 1, RELAY: 2, COUPLED_INDUCTORS: 3, GEN_SAWTOOTH: 4, GEN_SQUARE: 5, GEN_PULSE: 7, GEN_TRIANGLE: 4}
PINS = {GROUND: 1, R: 2, C: 2, L: 2, VDC: 2, VAC: 2, IDC: 2, IAC: 2, VCCS: 4, VCVS: 4, CCCS: 4, CCVS: 4, SWITCH: 2,
        PN: 2, OPAMP: 4, NPN: 3, PNP: 3, NMOS: 3, PMOS: 3, BRIDGE: 4, COMPARATOR: 3, TRANSFORMER: 4, TRANSFORMER_CT: 5, RELAY: 4, COUPLED_INDUCTORS: 4, GEN_SAWTOOTH: 2, GEN_SQUARE: 2, GEN_PULSE: 2, GEN_TRIANGLE: 2}
BRANCHES = {GROUND: 0, R: 0, C: 0, L: 1, VDC: 1, VAC: 1, IDC: 0, IAC: 0, VCCS: 0, VCVS: 1, CCCS: 1, CCVS: 2, SWITCH: 1,
            PN: 0, OPAMP: 1, NPN: 0, PNP: 0, NMOS: 0, PMOS: 0, BRIDGE: 0, COMPARATOR: 0, TRANSFORMER: 2, TRANSFORMER_CT: 3, RELAY: 1, COUPLED_INDUCTORS: 2, GEN_SAWTOOTH: 1, GEN_SQUARE: 1, GEN_PULSE: 1, GEN_TRIANGLE: 1}
# defaults of the PN junction's 9 positional properties (PN_junction.h:22-33): Is N Isr Nr Temp Ibv Bv Bv_set Area
PN_DEFAULT = (1e-14, 1.0, 0.0, 2.0, 27.0, 1e-3, 40.0, 1.0, 1.0)


class PhyEngineError(RuntimeError):
    pass


@dataclass
class Netlist:
    """Wire-format netlist: element codes, positional properties, (ele1, pin1, ele2, pin2) wires."""

    elements: list = field(default_factory=list)
    props: list = field(default_factory=list)
    wires: list = field(default_factory=list)

    def add(self, code: int, *props: float) -> int:
        if len(props) != PROPS[code]:
            raise ValueError(f"element code {code} takes {PROPS[code]} properties, got {len(props)}")
        self.elements.append(int(code))
        self.props.extend(float(p) for p in props)
        return len(self.elements) - 1

    def ground(self) -> int:
        return self.add(GROUND)

    def wire(self, e1: int, p1: int, e2: int, p2: int) -> None:
        self.wires.extend((int(e1), int(p1), int(e2), int(p2)))

    def component_index(self, ele: int) -> int:
        """index into vec_pos/chunk_pos (components = non-ground elements in order)"""
        cache = self.__dict__.get("_comp_index")
        if cache is None or len(cache) != len(self.elements):
            # prefix count of the non-ground elements (netlists of 1e5 elements ask for this per element)
            cache = np.concatenate(([0], np.cumsum(np.asarray(self.elements) != GROUND)))[:-1]
            self.__dict__["_comp_index"] = cache
        return int(cache[ele])

    def arrays(self):
        e = np.asarray(self.elements, dtype=np.int32)
        w = np.asarray(self.wires if self.wires else [0], dtype=np.int32)
        p = np.asarray(self.props if self.props else [0.0], dtype=np.float64)
        return e, w, p


_SZ = ct.c_size_t
_PSZ = ct.POINTER(ct.c_size_t)
_PD = ct.POINTER(ct.c_double)
_PI = ct.POINTER(ct.c_int)


def _p(a, t):
    return a.ctypes.data_as(t)


class CAbi:
    """Part 1 of the ABI (identical in the reference library and in the product)."""

    def __init__(self, path: str):
        if not os.path.exists(path):
            raise PhyEngineError(f"shared library not found: {path} (build it: python -c 'import __graft_entry__ as g; g.build()')")
        self.path = path
        self.lib = lib = ct.CDLL(path, mode=ct.RTLD_LOCAL)
        lib.phy_engine_last_error.restype = ct.c_char_p
        lib.create_circuit.restype = ct.c_void_p
        lib.create_circuit.argtypes = [_PI, _SZ, _PI, _SZ, _PD, ct.POINTER(_PSZ), ct.POINTER(_PSZ), _PSZ]
        lib.destroy_circuit.restype = None
        lib.destroy_circuit.argtypes = [ct.c_void_p, _PSZ, _PSZ]
        lib.circuit_set_analyze_type.argtypes = [ct.c_void_p, ct.c_uint32]
        lib.circuit_set_tr.argtypes = [ct.c_void_p, ct.c_double, ct.c_double]
        lib.circuit_set_ac_omega.argtypes = [ct.c_void_p, ct.c_double]
        lib.circuit_set_temperature.argtypes = [ct.c_void_p, ct.c_double]
        lib.circuit_set_tnom.argtypes = [ct.c_void_p, ct.c_double]
        lib.circuit_set_model_double_by_name.argtypes = [ct.c_void_p, _SZ, _SZ, ct.c_char_p, _SZ, ct.c_double]
        lib.circuit_analyze.argtypes = [ct.c_void_p]
        lib.circuit_digital_clk.argtypes = [ct.c_void_p]
        lib.circuit_sample_layout.argtypes = [ct.c_void_p, _PSZ, _PSZ, _SZ, _PSZ, _PSZ, _PSZ]
        lib.circuit_sample_u8.argtypes = [ct.c_void_p, _PSZ, _PSZ, _SZ, _PD, _PSZ, _PD, _PSZ, ct.POINTER(ct.c_uint8), _PSZ]

    def last_error(self) -> str:
        s = self.lib.phy_engine_last_error()
        return s.decode("utf-8", "replace") if s else ""


class CircuitBase:
    """A circuit handle of a Part-1 library.  Subclassed by `Circuit` (product) and tests' reference wrapper."""

    def __init__(self, abi: CAbi, nl: Netlist):
        self.abi = abi
        self.nl = nl
        e, w, p = nl.arrays()
        self._keep = (e, w, p)
        vp, cp, cs = _PSZ(), _PSZ(), ct.c_size_t(0)
        n_w = len(nl.wires)
        self.h = abi.lib.create_circuit(_p(e, _PI), e.size, _p(w, _PI), n_w, _p(p, _PD), ct.byref(vp), ct.byref(cp), ct.byref(cs))
        if not self.h:
            raise PhyEngineError("create_circuit failed: " + abi.last_error())
        self._vp, self._cp, self.comp_size = vp, cp, cs.value
        self.vec_pos = [vp[i] for i in range(self.comp_size)]
        self.chunk_pos = [cp[i] for i in range(self.comp_size)]

    def close(self):
        if getattr(self, "h", None):
            self.abi.lib.destroy_circuit(self.h, self._vp, self._cp)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _rc(self, rc, what):
        if rc != 0:
            raise PhyEngineError(f"{what} failed (rc={rc}): {self.abi.last_error()}")

    def pos(self, ele: int):
        k = self.nl.component_index(ele)
        return self.vec_pos[k], self.chunk_pos[k]

    def set_analyze_type(self, at: int):
        self._rc(self.abi.lib.circuit_set_analyze_type(self.h, at), "circuit_set_analyze_type")

    def set_tr(self, t_step: float, t_stop: float):
        self._rc(self.abi.lib.circuit_set_tr(self.h, t_step, t_stop), "circuit_set_tr")

    def set_ac_omega(self, omega: float):
        self._rc(self.abi.lib.circuit_set_ac_omega(self.h, omega), "circuit_set_ac_omega")

    def set_temperature(self, t: float):
        self._rc(self.abi.lib.circuit_set_temperature(self.h, t), "circuit_set_temperature")

    def set_param(self, ele: int, name: str, value: float) -> int:
        v, c = self.pos(ele)
        b = name.encode()
        return self.abi.lib.circuit_set_model_double_by_name(self.h, v, c, b, len(b), float(value))

    def analyze(self) -> bool:
        return self.abi.lib.circuit_analyze(self.h) == 0

    def sample(self):
        """(voltage[pins], voltage_ord, current[branches], current_ord) through circuit_sample_u8"""
        n = self.comp_size
        vo = np.zeros(n + 1, dtype=np.uintp)
        co = np.zeros(n + 1, dtype=np.uintp)
        dg = np.zeros(n + 1, dtype=np.uintp)
        lib = self.abi.lib
        self._rc(lib.circuit_sample_layout(self.h, self._vp, self._cp, n, _p(vo, _PSZ), _p(co, _PSZ), _p(dg, _PSZ)), "circuit_sample_layout")
        v = np.zeros(max(int(vo[n]), 1))
        i = np.zeros(max(int(co[n]), 1))
        d = np.zeros(max(int(dg[n]), 1), dtype=np.uint8)
        self._rc(lib.circuit_sample_u8(self.h, self._vp, self._cp, n, _p(v, _PD), _p(vo, _PSZ), _p(i, _PD), _p(co, _PSZ), _p(d, ct.POINTER(ct.c_uint8)), _p(dg, _PSZ)),
                 "circuit_sample_u8")
        return v[: int(vo[n])], vo, i[: int(co[n])], co


_product = None


def product() -> CAbi:
    """Load libphyengine_b200.so once and declare the additive (Part 2/3) entry points."""
    global _product
    if _product is None:
        _product = bind_full_abi(CAbi(LIB_PATH))
    return _product


def bind_full_abi(abi: CAbi) -> CAbi:
    """declare Parts 2 and 3 of include/phy_engine_b200.h on a loaded library"""
    lib = abi.lib
    V = ct.c_void_p
    PU32 = ct.POINTER(ct.c_uint32)
    PI32 = ct.POINTER(ct.c_int32)
    PI64 = ct.POINTER(ct.c_int64)
    lib.circuit_set_env.argtypes = [V, _PD]
    lib.circuit_set_ac_sweep.argtypes = [V, ct.c_int, ct.c_double, ct.c_double, _SZ]
    lib.circuit_unknown_count.argtypes = [V, _PSZ, _PSZ]
    lib.circuit_pin_unknown.restype = ct.c_longlong
    lib.circuit_pin_unknown.argtypes = [V, _SZ, _SZ, _SZ]
    lib.circuit_branch_unknown.restype = ct.c_longlong
    lib.circuit_branch_unknown.argtypes = [V, _SZ, _SZ, _SZ]
    lib.circuit_get_solution.argtypes = [V, _PD, _PD]
    lib.circuit_batch_create.restype = V
    lib.circuit_batch_create.argtypes = [V, _SZ]
    lib.circuit_batch_destroy.restype = None
    lib.circuit_batch_destroy.argtypes = [V]
    lib.circuit_batch_set_device.argtypes = [V, ct.c_int]
    lib.circuit_batch_set_stream.argtypes = [V, V]
    lib.circuit_batch_set_param.argtypes = [V, _SZ, _SZ, ct.c_char_p, _SZ, _PD]
    lib.circuit_batch_set_ac_sweep.argtypes = [V, ct.c_int, ct.c_double, ct.c_double, _SZ]
    lib.circuit_batch_set_params.argtypes = [V, _SZ, _PSZ, _PSZ, ct.POINTER(ct.c_char_p), ct.c_void_p]
    lib.circuit_batch_solution_soa.argtypes = [V, ct.c_void_p]
    lib.circuit_batch_set_probes.argtypes = [V, _PSZ, _SZ]
    lib.circuit_batch_set_subtree_warps.argtypes = [V, ct.c_int]
    lib.circuit_batch_set_resident.argtypes = [V, ct.c_int, ct.c_int, ct.c_int]
    lib.phy_engine_b200_set_default_path.argtypes = [ct.c_int, ct.c_int, ct.c_int, ct.c_int, ct.c_int, ct.c_uint]
    lib.circuit_batch_set_tuning.argtypes = [V, ct.c_uint]
    lib.circuit_batch_last_kernel.argtypes = [V]
    lib.circuit_batch_last_kernel.restype = ct.c_int
    lib.circuit_batch_stream_info.argtypes = [V, ct.c_int, ct.POINTER(ct.c_int64)]
    lib.circuit_batch_stream_info.restype = ct.c_int
    lib.circuit_batch_set_workspace.argtypes = [V, ct.c_int]
    lib.circuit_batch_digital_clk.argtypes = [V]
    lib.circuit_batch_comparator_count.restype = _SZ
    lib.circuit_batch_comparator_count.argtypes = [V]
    lib.circuit_batch_comparator_states.argtypes = [V, ct.POINTER(ct.c_uint8)]
    lib.circuit_batch_set_chunks.argtypes = [V, ct.c_int]
    lib.circuit_batch_resident_info.argtypes = [V, ct.c_int, ct.POINTER(ct.c_int64)]
    for f in ("circuit_batch_prepare", "circuit_batch_reset_state", "circuit_batch_analyze", "circuit_batch_compile_host"):
        getattr(lib, f).argtypes = [V]
    lib.circuit_batch_lanes.restype = _SZ
    lib.circuit_batch_lanes.argtypes = [V]
    lib.circuit_batch_points.restype = _SZ
    lib.circuit_batch_points.argtypes = [V]
    lib.circuit_batch_total_solves.restype = ct.c_uint64
    lib.circuit_batch_total_solves.argtypes = [V]
    lib.circuit_batch_tr_duration.restype = ct.c_double
    lib.circuit_batch_tr_duration.argtypes = [V]
    lib.circuit_batch_solution.argtypes = [V, _PD]
    lib.circuit_batch_ac_solution.argtypes = [V, _PD]
    lib.circuit_batch_ac_omegas.argtypes = [V, _PD]
    lib.circuit_batch_status.argtypes = [V, PI32]
    lib.circuit_batch_newton_iters.argtypes = [V, PU32]
    lib.circuit_batch_waveform.argtypes = [V, _PD]
    lib.circuit_batch_stats.argtypes = [V, ct.c_int, _PSZ, _PSZ, _PSZ, _PSZ, _PSZ, _PSZ]
    lib.circuit_batch_param_device_ptr.argtypes = [V, _SZ, _SZ, ct.c_char_p, _SZ, ct.POINTER(_PD)]
    lib.circuit_batch_solution_device_ptr.argtypes = [V, ct.POINTER(_PD), _PSZ]
    lib.circuit_batch_program_words.restype = _SZ
    lib.circuit_batch_program_words.argtypes = [V, ct.c_int]
    lib.circuit_batch_program_copy.argtypes = [V, ct.c_int, PU32]
    lib.circuit_batch_const_count.restype = _SZ
    lib.circuit_batch_const_count.argtypes = [V]
    lib.circuit_batch_const_copy.argtypes = [V, _PD]
    lib.circuit_batch_program_info.argtypes = [V, ct.c_int, PI64]
    lib.circuit_batch_swept_slot.restype = ct.c_longlong
    lib.circuit_batch_swept_slot.argtypes = [V, _SZ, _SZ, ct.c_char_p, _SZ]
    lib.circuit_batch_swept_values.argtypes = [V, ct.c_longlong, _PD]
    lib.phy_engine_b200_device_count.restype = ct.c_int
    lib.phy_engine_b200_launch_count.restype = ct.c_uint64
    lib.phy_engine_b200_aux_launch_count.restype = ct.c_uint64
    lib.phy_engine_b200_timing.restype = None
    lib.phy_engine_b200_timing.argtypes = [ct.c_int]
    lib.phy_engine_b200_kernel_ms.restype = ct.c_double
    return abi


def device_count() -> int:
    return product().lib.phy_engine_b200_device_count()


def launch_count() -> int:
    return int(product().lib.phy_engine_b200_launch_count())


def aux_launch_count() -> int:
    return int(product().lib.phy_engine_b200_aux_launch_count())


def kernel_timing(on: bool) -> None:
    product().lib.phy_engine_b200_timing(1 if on else 0)


def kernel_ms() -> float:
    """device time of the solve kernels launched since the last call (CUDA events on the launching stream)"""
    return float(product().lib.phy_engine_b200_kernel_ms())


class Circuit(CircuitBase):
    """Handle of the B200 library (circuit_* of include/phy_engine_b200.h)."""

    def __init__(self, nl: Netlist, abi: CAbi | None = None):
        super().__init__(abi or product(), nl)

    def set_env(self, V_eps_max=0.0, V_epsr_max=0.0, I_eps_max=0.0, I_epsr_max=0.0, g_min=0.0, r_open=0.0, temperature=27.0, norm_temperature=27.0):
        a = np.array([V_eps_max, V_epsr_max, I_eps_max, I_epsr_max, g_min, r_open, temperature, norm_temperature], dtype=np.float64)
        self._rc(self.abi.lib.circuit_set_env(self.h, _p(a, _PD)), "circuit_set_env")

    def set_ac_sweep(self, sweep: int, w0: float, w1: float, points: int):
        self._rc(self.abi.lib.circuit_set_ac_sweep(self.h, sweep, w0, w1, points), "circuit_set_ac_sweep")

    def unknown_count(self):
        a, b = ct.c_size_t(0), ct.c_size_t(0)
        self._rc(self.abi.lib.circuit_unknown_count(self.h, ct.byref(a), ct.byref(b)), "circuit_unknown_count")
        return a.value, b.value

    def pin_unknown(self, ele: int, pin: int) -> int:
        v, c = self.pos(ele)
        return int(self.abi.lib.circuit_pin_unknown(self.h, v, c, pin))

    def branch_unknown(self, ele: int, br: int = 0) -> int:
        v, c = self.pos(ele)
        return int(self.abi.lib.circuit_branch_unknown(self.h, v, c, br))

    def solution(self):
        n = sum(self.unknown_count())
        re, im = np.zeros(max(n, 1)), np.zeros(max(n, 1))
        self._rc(self.abi.lib.circuit_get_solution(self.h, _p(re, _PD), _p(im, _PD)), "circuit_get_solution")
        return re[:n] + 1j * im[:n]

    def batch(self, n_instances: int) -> "Batch":
        return Batch(self, n_instances)


class Batch:
    """circuit_batch_* : B independent instances of one netlist, solved together on the GPU."""

    def __init__(self, circuit: Circuit, n_instances: int):
        self.c = circuit
        self.lib = circuit.abi.lib
        self.n_inst = int(n_instances)
        self.h = self.lib.circuit_batch_create(circuit.h, self.n_inst)
        if not self.h:
            raise PhyEngineError("circuit_batch_create failed: " + circuit.abi.last_error())

    def close(self):
        if getattr(self, "h", None):
            self.lib.circuit_batch_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _rc(self, rc, what):
        if rc != 0:
            raise PhyEngineError(f"{what} failed (rc={rc}): {self.c.abi.last_error()}")

    def set_device(self, dev: int):
        self._rc(self.lib.circuit_batch_set_device(self.h, dev), "circuit_batch_set_device")

    def set_stream(self, stream_ptr: int):
        self._rc(self.lib.circuit_batch_set_stream(self.h, ct.c_void_p(stream_ptr)), "circuit_batch_set_stream")

    def set_subtree_warps(self, g: int):
        self._rc(self.lib.circuit_batch_set_subtree_warps(self.h, g), "circuit_batch_set_subtree_warps")

    def set_resident(self, streams: int = 0, instances_per_cta: int = 0, instances_per_thread: int = 0):
        """Resident (shared-memory) solve path: streams -1 = never, 0 = automatic, else a power of two."""
        self._rc(self.lib.circuit_batch_set_resident(self.h, streams, instances_per_cta, instances_per_thread), "circuit_batch_set_resident")

    def set_workspace(self, where: int):
        """0 = automatic, 1 = shared memory (resident kernel), 2 = HBM (tree-streaming kernel)"""
        self._rc(self.lib.circuit_batch_set_workspace(self.h, where), "circuit_batch_set_workspace")

    def set_chunks(self, chunks: int):
        """tree-streaming kernel: chunks of the time loop for dynamic scheduling (0 = automatic, 1 = static)"""
        self._rc(self.lib.circuit_batch_set_chunks(self.h, chunks), "circuit_batch_set_chunks")

    def set_tuning(self, flags: int):
        """bit 0 L2 operand prefetch, bit 1 two lines ahead, bit 2 no L1 re-fetch of results, bit 3 fused elimination steps, bit 4 / 5 require / forbid the specialised kernel, bit 6 / 7 require / forbid the stream kernel"""
        self._rc(self.lib.circuit_batch_set_tuning(self.h, flags), "circuit_batch_set_tuning")

    def last_kernel(self) -> int:
        """0 = word interpreter, 1 = specialised (run-time compiled) tree-streaming kernel, 2 = stream kernel"""
        return int(self.lib.circuit_batch_last_kernel(self.h))

    def frontal_info(self):
        """reduce-and-core path of the last analyze(): None when it was not taken"""
        v = (ct.c_int64 * 11)()
        self.lib.circuit_batch_frontal_info.argtypes = [ct.c_void_p, ct.POINTER(ct.c_int64)]
        if self.lib.circuit_batch_frontal_info(self.h, v) != 0:
            return None
        keys = ("unknowns", "eliminated", "levels", "core_rows", "edges", "launches", "ld_core", "core_edges", "reduce_us", "lu_us", "subst_us")
        return {k: int(x) for k, x in zip(keys, v)}

    def save_state(self) -> bytes:
        self.lib.circuit_batch_save_state.argtypes = [ct.c_void_p, ct.c_void_p, ct.c_size_t, ct.POINTER(ct.c_size_t)]
        n = ct.c_size_t(0)
        self._rc(self.lib.circuit_batch_save_state(self.h, None, 0, ct.byref(n)), "circuit_batch_save_state")
        buf = ct.create_string_buffer(n.value)
        self._rc(self.lib.circuit_batch_save_state(self.h, buf, n.value, ct.byref(n)), "circuit_batch_save_state")
        return buf.raw[: n.value]

    def load_state(self, blob: bytes):
        self.lib.circuit_batch_load_state.argtypes = [ct.c_void_p, ct.c_char_p, ct.c_size_t]
        self._rc(self.lib.circuit_batch_load_state(self.h, blob, len(blob)), "circuit_batch_load_state")

    def set_pivot_guard(self, guard: float = -1.0, rounds: int = -1):
        self.lib.circuit_batch_set_pivot_guard.argtypes = [ct.c_void_p, ct.c_double, ct.c_int]
        self._rc(self.lib.circuit_batch_set_pivot_guard(self.h, guard, rounds), "circuit_batch_set_pivot_guard")

    def rescue_info(self, mode: int) -> dict:
        v = (ct.c_int64 * 6)()
        self.lib.circuit_batch_rescue_info.argtypes = [ct.c_void_p, ct.c_int, ct.POINTER(ct.c_int64)]
        self._rc(self.lib.circuit_batch_rescue_info(self.h, mode, v), "circuit_batch_rescue_info")
        keys = ("guarded_pivots", "flagged", "rescued", "unguarded", "sub_runs", "sub_batches")
        return {k: int(x) for k, x in zip(keys, v)}

    def stream_info(self, mode: int) -> dict:
        v = (ct.c_int64 * 6)()
        self._rc(self.lib.circuit_batch_stream_info(self.h, mode, v), "circuit_batch_stream_info")
        keys = ("last_kernel", "warps_per_cta", "ring_stages", "smem_per_cta", "tiles_per_solve", "stage_rows")
        return {k: int(x) for k, x in zip(keys, v)}

    def digital_clk(self) -> np.ndarray:
        """comparator states [n_instances, n_comparators] (vA >= vB) after the last analyze()"""
        self._rc(self.lib.circuit_batch_digital_clk(self.h), "circuit_batch_digital_clk")
        n = int(self.lib.circuit_batch_comparator_count(self.h))
        out = np.zeros((self.n_inst, max(n, 1)), dtype=np.uint8)
        self._rc(self.lib.circuit_batch_comparator_states(self.h, out.ctypes.data_as(ct.POINTER(ct.c_uint8))), "circuit_batch_comparator_states")
        return out[:, :n]

    def resident_info(self, mode: int) -> dict:
        v = (ct.c_int64 * 13)()
        self._rc(self.lib.circuit_batch_resident_info(self.h, mode, v), "circuit_batch_resident_info")
        keys = ("resident", "streams", "smem_slots", "I", "J", "io_entries", "last_S", "last_I", "last_J", "iter_phases", "words", "max_warp_words", "hbm")
        return {k: int(x) for k, x in zip(keys, v)}

    def set_param(self, ele: int, name: str, values):
        a = np.ascontiguousarray(values, dtype=np.float64)
        if a.shape != (self.n_inst,):
            raise ValueError("values must have shape (n_instances,)")
        v, c = self.c.pos(ele)
        b = name.encode()
        self._rc(self.lib.circuit_batch_set_param(self.h, v, c, b, len(b), _p(a, _PD)), f"circuit_batch_set_param({name})")

    def set_ac_sweep(self, sweep: int, w0: float, w1: float, points: int):
        self._rc(self.lib.circuit_batch_set_ac_sweep(self.h, sweep, w0, w1, points), "circuit_batch_set_ac_sweep")

    def set_ac_slice(self, first: int, count: int):
        """solve only the points [first, first + count) of the sweep (one rank's shard); count = 0: all"""
        self.lib.circuit_batch_set_ac_slice.argtypes = [ct.c_void_p, ct.c_size_t, ct.c_size_t]
        self._rc(self.lib.circuit_batch_set_ac_slice(self.h, first, count), "circuit_batch_set_ac_slice")

    def param_table(self, items):
        """pre-resolve [(element, attribute name)] for set_params()"""
        pos = [self.c.pos(e) for e, _ in items]
        vp = np.array([p[0] for p in pos], dtype=np.uintp)
        cp = np.array([p[1] for p in pos], dtype=np.uintp)
        names = (ct.c_char_p * len(items))(*[n.encode() for _, n in items])
        return (len(items), vp, cp, names)

    def set_params(self, table, values_ptr: int):
        """values_ptr: host address of float64 [n_params][n_instances] (pinned memory makes the copy a plain DMA)"""
        n, vp, cp, names = table
        self._rc(self.lib.circuit_batch_set_params(self.h, n, _p(vp, _PSZ), _p(cp, _PSZ), names, ct.c_void_p(values_ptr)), "circuit_batch_set_params")

    def solution_soa_into(self, host_ptr: int):
        """final state in the device-native [n_unknowns][n_instances] layout, written to host address host_ptr"""
        self._rc(self.lib.circuit_batch_solution_soa(self.h, ct.c_void_p(host_ptr)), "circuit_batch_solution_soa")

    def set_probes(self, unknowns):
        a = np.ascontiguousarray(unknowns, dtype=np.uintp)
        self._probes = a.size
        self._rc(self.lib.circuit_batch_set_probes(self.h, _p(a, _PSZ), a.size), "circuit_batch_set_probes")

    def prepare(self):
        self._rc(self.lib.circuit_batch_prepare(self.h), "circuit_batch_prepare")

    def reset_state(self):
        self._rc(self.lib.circuit_batch_reset_state(self.h), "circuit_batch_reset_state")

    def compile_host(self):
        self._rc(self.lib.circuit_batch_compile_host(self.h), "circuit_batch_compile_host")

    def analyze(self) -> bool:
        return self.lib.circuit_batch_analyze(self.h) == 0

    @property
    def lanes(self) -> int:
        return int(self.lib.circuit_batch_lanes(self.h))

    @property
    def points(self) -> int:
        return int(self.lib.circuit_batch_points(self.h))

    @property
    def total_solves(self) -> int:
        return int(self.lib.circuit_batch_total_solves(self.h))

    @property
    def tr_duration(self) -> float:
        return float(self.lib.circuit_batch_tr_duration(self.h))

    def n_unknowns(self) -> int:
        return sum(self.c.unknown_count())

    def solution(self) -> np.ndarray:
        x = np.zeros((self.n_inst, max(self.n_unknowns(), 1)))
        self._rc(self.lib.circuit_batch_solution(self.h, _p(x, _PD)), "circuit_batch_solution")
        return x[:, : self.n_unknowns()]

    def ac_solution(self) -> np.ndarray:
        n = self.n_unknowns()
        x = np.zeros((self.lanes, n, 2))
        self._rc(self.lib.circuit_batch_ac_solution(self.h, _p(x, _PD)), "circuit_batch_ac_solution")
        return (x[..., 0] + 1j * x[..., 1]).reshape(self.n_inst, self.points, n)

    def ac_solution_lanes(self, lanes) -> np.ndarray:
        """complex solution [len(lanes), n] of selected lanes (lane = instance * points + point) of the last AC analyze()"""
        lanes = np.ascontiguousarray(lanes, dtype=np.uintp)
        n = self.n_unknowns()
        x = np.zeros((lanes.size, n, 2))
        self.lib.circuit_batch_ac_solution_lanes.argtypes = [ct.c_void_p, ct.POINTER(ct.c_size_t), ct.c_size_t, _PD]
        self._rc(self.lib.circuit_batch_ac_solution_lanes(self.h, _p(lanes, ct.POINTER(ct.c_size_t)), lanes.size, _p(x, _PD)), "circuit_batch_ac_solution_lanes")
        return x[..., 0] + 1j * x[..., 1]

    def ac_omegas(self) -> np.ndarray:
        om = np.zeros(max(self.points, 1))
        self._rc(self.lib.circuit_batch_ac_omegas(self.h, _p(om, _PD)), "circuit_batch_ac_omegas")
        return om[: self.points]

    def status(self) -> np.ndarray:
        st = np.zeros(max(self.lanes, 1), dtype=np.int32)
        self._rc(self.lib.circuit_batch_status(self.h, _p(st, ct.POINTER(ct.c_int32))), "circuit_batch_status")
        return st[: self.lanes]

    def newton_iters(self) -> np.ndarray:
        sv = np.zeros(max(self.lanes, 1), dtype=np.uint32)
        self._rc(self.lib.circuit_batch_newton_iters(self.h, _p(sv, ct.POINTER(ct.c_uint32))), "circuit_batch_newton_iters")
        return sv[: self.lanes]

    def waveform(self, steps: int) -> np.ndarray:
        w = np.zeros((steps, self._probes, self.n_inst))
        self._rc(self.lib.circuit_batch_waveform(self.h, _p(w, _PD)), "circuit_batch_waveform")
        return w

    def stats(self, mode: int) -> dict:
        v = [ct.c_size_t(0) for _ in range(6)]
        self._rc(self.lib.circuit_batch_stats(self.h, mode, *[ct.byref(x) for x in v]), "circuit_batch_stats")
        keys = ("n_unknowns", "nnz_a", "nnz_lu", "n_fma", "n_lane_slots", "n_inst_slots")
        return {k: x.value for k, x in zip(keys, v)}

    def param_device_ptr(self, ele: int, name: str) -> int:
        v, c = self.c.pos(ele)
        b = name.encode()
        p = _PD()
        self._rc(self.lib.circuit_batch_param_device_ptr(self.h, v, c, b, len(b), ct.byref(p)), "circuit_batch_param_device_ptr")
        return ct.cast(p, ct.c_void_p).value or 0

    def solution_device_ptr(self):
        p, s = _PD(), ct.c_size_t(0)
        self._rc(self.lib.circuit_batch_solution_device_ptr(self.h, ct.byref(p), ct.byref(s)), "circuit_batch_solution_device_ptr")
        return (ct.cast(p, ct.c_void_p).value or 0), s.value

    # ---- introspection of the symbolic phase (host only) ----
    def program(self, mode: int) -> np.ndarray:
        n = int(self.lib.circuit_batch_program_words(self.h, mode))
        out = np.zeros(max(n, 1), dtype=np.uint32)
        self._rc(self.lib.circuit_batch_program_copy(self.h, mode, _p(out, ct.POINTER(ct.c_uint32))), "circuit_batch_program_copy")
        return out[:n]

    def constants(self) -> np.ndarray:
        n = int(self.lib.circuit_batch_const_count(self.h))
        out = np.zeros(max(n, 1))
        self._rc(self.lib.circuit_batch_const_copy(self.h, _p(out, _PD)), "circuit_batch_const_copy")
        return out[:n]

    def program_info(self, mode: int) -> dict:
        info = np.zeros(64, dtype=np.int64)
        self._rc(self.lib.circuit_batch_program_info(self.h, mode, _p(info, ct.POINTER(ct.c_int64))), "circuit_batch_program_info")
        keys = ("cplx", "structurally_singular", "n_lane_slots", "omega_slot", "n_inst_slots", "dt_slot", "x_slot0", "n_unknowns", "warps",
                "n_real_lane_slots", "n_leaves", "n_leaf_rows", "n_top_rows", "max_warp_words", "nnz_a", "nnz_lu")
        d = dict(zip(keys, (int(v) for v in info[:16])))
        for s, name in enumerate(("prep", "step", "iter")):
            d[name] = [int(v) for v in info[16 + 16 * s: 32 + 16 * s]]
        return d
