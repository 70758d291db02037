"""CPU oracle for the gsim hot path — TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package; the product package
``digilogic_b200`` never does.  PARITY UNPINNED: see the header of
``oracle/gsim_oracle.c`` (gsim 1.1.4 is an un-vendored crates.io dependency of the
reference, ``crates/digilogic_gsim/Cargo.toml:15``; the reference has no test that
evaluates a circuit).

ctypes binding over ``oracle/_build/libgsim_oracle.so`` (built by ``oracle/Makefile``).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libgsim_oracle.so")

AND, OR, XOR, NAND, NOR, XNOR, NOT, MUX, BUF, SLICE, MERGE, REG, HGATE, ADD, SUB = range(15)
SHL, LSHR, ASHR = "shl", "lshr", "ashr"
CMP_EQ, CMP_NE, CMP_LTU, CMP_GTU, CMP_LEU, CMP_GEU, CMP_LTS, CMP_GTS, CMP_LES, CMP_GES = range(10)
RUN_OK, RUN_MAX_STEPS, RUN_CONFLICT = 0, 1, 2
ERR_NAMES = {
    0: "Ok", 1: "TooManyComponents", 2: "InvalidWireId", 3: "WireWidthMismatch",
    4: "WireWidthIncompatible", 5: "OffsetOutOfRange", 6: "TooFewInputs", 7: "InvalidInputCount",
}


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc). Building the checker is not using it."""
    srcs = [os.path.join(_HERE, f) for f in ("gsim_oracle.c", "bitsliced.c", "gsim_oracle.h")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        vp, u8, u32, u64, i64 = C.c_void_p, C.c_uint8, C.c_uint32, C.c_uint64, C.c_int64
        pu32, pu64, pu8 = C.POINTER(u32), C.POINTER(u64), C.POINTER(u8)
        sig = {
            "go_set_config": (None, [C.c_int] * 4),
            "go_builder_new": (vp, []),
            "go_builder_free": (None, [vp]),
            "go_add_wire": (i64, [vp, u8]),
            "go_add_wires": (C.c_int, [vp, u8, u32, pu32]),
            "go_builder_wire_width": (C.c_int, [vp, u32, pu8]),
            "go_add_gate": (C.c_int, [vp, C.c_int, pu32, u32, u32, pu32]),
            "go_add_not": (C.c_int, [vp, u32, u32, pu32]),
            "go_add_mux": (C.c_int, [vp, pu32, u32, u32, u32, pu32]),
            "go_add_buffer": (C.c_int, [vp, u32, u32, u32, pu32]),
            "go_set_buffer_z_passthrough": (None, [C.c_int]),
            "go_set_copy_z_passthrough": (None, [C.c_int]),
            "go_add_register": (C.c_int, [vp, u32, u32, u32, u32, pu32]),
            "go_add_horizontal_gate": (C.c_int, [vp, C.c_int, u32, u32, pu32]),
            "go_add_arith": (C.c_int, [vp, C.c_int, u32, u32, u32, pu32]),
            "go_add_mul": (C.c_int, [vp, u32, u32, u32, pu32]),
            "go_add_compare": (C.c_int, [vp, C.c_int, u32, u32, u32, pu32]),
            "go_add_shift": (C.c_int, [vp, C.c_int, u32, u32, u32, pu32]),
            "go_add_slice": (C.c_int, [vp, u32, u32, u32, pu32]),
            "go_add_merge": (C.c_int, [vp, pu32, u32, u32, pu32]),
            "go_add_gates2": (C.c_int, [vp, vp, vp, vp, vp, u32]),
            "go_build": (vp, [vp]),
            "go_sim_free": (None, [vp]),
            "go_wire_width": (C.c_int, [vp, u32, pu8]),
            "go_run_sim": (C.c_int, [vp, u64]),
            "go_last_steps": (u64, [vp]),
            "go_last_comp_evals": (u64, [vp]),
            "go_set_wire_drive": (C.c_int, [vp, u32, pu32, pu32, u32]),
            "go_get_wire_state": (C.c_int, [vp, u32, pu32, pu32]),
            "go_logic_op": (None, [C.c_int, u32, u32, u32, u32, pu32, pu32]),
            "go_lanes_new": (vp, [vp, u32]),
            "go_lanes_free": (None, [vp]),
            "go_lanes_set_drive_bit": (C.c_int, [vp, u32, u32, vp, vp]),
            "go_lanes_get_bit": (C.c_int, [vp, u32, u32, vp, vp]),
            "go_lanes_get_all": (None, [vp, vp, vp]),
            "go_lanes_run": (u64, [vp, u64, u32, vp, vp]),
            "bs_new": (vp, [vp, u32]),
            "bs_free": (None, [vp]),
            "bs_is_topological": (C.c_int, [vp]),
            "bs_last_applications": (u64, [vp]),
            "bs_set_drive": (None, [vp, u32, vp, vp]),
            "bs_get": (None, [vp, u32, vp, vp]),
            "bs_get_all": (None, [vp, vp, vp]),
            "bs_run": (None, [vp, u64, vp, vp]),
            "bs_eval_dag": (C.c_int, [vp, vp]),
            "bs_logic_op": (None, [C.c_int, u64, u64, u64, u64, pu64, pu64]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class OracleError(Exception):
    def __init__(self, code):
        super().__init__(ERR_NAMES.get(code, str(code)))
        self.code = code


class Builder:
    """gsim ``SimulatorBuilder`` restated (names follow the call sites in
    reference crates/digilogic_gsim/src/lib.rs:88-190)."""

    def __init__(self):
        self._h = lib().go_builder_new()

    def __del__(self):
        if getattr(self, "_h", None):
            lib().go_builder_free(self._h)
            self._h = None

    def add_wire(self, width=1):
        r = lib().go_add_wire(self._h, width)
        if r < 0:
            raise OracleError(1)
        return int(r)

    def add_wires(self, n, width=1):
        first = C.c_uint32()
        rc = lib().go_add_wires(self._h, width, n, C.byref(first))
        if rc:
            raise OracleError(rc)
        return first.value

    def get_wire_width(self, wire):
        out = C.c_uint8()
        rc = lib().go_builder_wire_width(self._h, wire, C.byref(out))
        if rc:
            raise OracleError(rc)
        return out.value

    def _gate(self, kind, inputs, output):
        arr = (C.c_uint32 * len(inputs))(*inputs)
        cell = C.c_uint32()
        rc = lib().go_add_gate(self._h, kind, arr, len(inputs), output, C.byref(cell))
        if rc:
            raise OracleError(rc)
        return cell.value

    def add_and_gate(self, inputs, output): return self._gate(AND, inputs, output)
    def add_or_gate(self, inputs, output): return self._gate(OR, inputs, output)
    def add_xor_gate(self, inputs, output): return self._gate(XOR, inputs, output)
    def add_nand_gate(self, inputs, output): return self._gate(NAND, inputs, output)
    def add_nor_gate(self, inputs, output): return self._gate(NOR, inputs, output)
    def add_xnor_gate(self, inputs, output): return self._gate(XNOR, inputs, output)

    def add_gate(self, kind, inputs, output):
        if kind == NOT:
            return self.add_not_gate(inputs[0], output)
        return self._gate(kind, inputs, output)

    def add_not_gate(self, input, output):
        cell = C.c_uint32()
        rc = lib().go_add_not(self._h, input, output, C.byref(cell))
        if rc:
            raise OracleError(rc)
        return cell.value

    def add_multiplexer(self, inputs, select, output):
        arr = (C.c_uint32 * len(inputs))(*inputs)
        cell = C.c_uint32()
        rc = lib().go_add_mux(self._h, arr, len(inputs), select, output, C.byref(cell))
        if rc:
            raise OracleError(rc)
        return cell.value

    def add_buffer(self, input, enable, output):
        cell = C.c_uint32()
        rc = lib().go_add_buffer(self._h, input, enable, output, C.byref(cell))
        if rc:
            raise OracleError(rc)
        return cell.value

    def add_horizontal_gate(self, kind, input, output):
        cell = C.c_uint32()
        rc = lib().go_add_horizontal_gate(self._h, kind, input, output, C.byref(cell))
        if rc:
            raise OracleError(rc)
        return cell.value

    def add_horizontal_and_gate(self, input, output): return self.add_horizontal_gate(AND, input, output)
    def add_horizontal_or_gate(self, input, output): return self.add_horizontal_gate(OR, input, output)
    def add_horizontal_xor_gate(self, input, output): return self.add_horizontal_gate(XOR, input, output)
    def add_horizontal_nand_gate(self, input, output): return self.add_horizontal_gate(NAND, input, output)
    def add_horizontal_nor_gate(self, input, output): return self.add_horizontal_gate(NOR, input, output)
    def add_horizontal_xnor_gate(self, input, output): return self.add_horizontal_gate(XNOR, input, output)

    def _arith(self, kind, a, b, output):
        cell = C.c_uint32()
        rc = lib().go_add_arith(self._h, kind, a, b, output, C.byref(cell))
        if rc:
            raise OracleError(rc)
        return cell.value

    def add_add(self, input_a, input_b, output): return self._arith(ADD, input_a, input_b, output)
    def add_sub(self, input_a, input_b, output): return self._arith(SUB, input_a, input_b, output)

    def add_mul(self, input_a, input_b, output):
        cell = C.c_uint32()
        rc = lib().go_add_mul(self._h, input_a, input_b, output, C.byref(cell))
        if rc:
            raise OracleError(rc)
        return cell.value

    def add_compare(self, op, input_a, input_b, output):
        """op: CMP_EQ .. CMP_GES (gsim add_compare_{equal,not_equal,less_than,...}{,_signed})."""
        cell = C.c_uint32()
        rc = lib().go_add_compare(self._h, op, input_a, input_b, output, C.byref(cell))
        if rc:
            raise OracleError(rc)
        return cell.value

    def add_shift(self, kind, input, amount, output):
        """kind: SHL / LSHR / ASHR (gsim add_left_shift / add_logical_right_shift / add_arithmetic_right_shift)."""
        cell = C.c_uint32()
        rc = lib().go_add_shift(self._h, {SHL: 17, LSHR: 18, ASHR: 19}[kind], input, amount, output, C.byref(cell))
        if rc:
            raise OracleError(rc)
        return cell.value

    def add_left_shift(self, input, amount, output): return self.add_shift(SHL, input, amount, output)
    def add_logical_right_shift(self, input, amount, output): return self.add_shift(LSHR, input, amount, output)
    def add_arithmetic_right_shift(self, input, amount, output): return self.add_shift(ASHR, input, amount, output)

    def add_register(self, data_in, data_out, enable, clock):
        cell = C.c_uint32()
        rc = lib().go_add_register(self._h, data_in, data_out, enable, clock, C.byref(cell))
        if rc:
            raise OracleError(rc)
        return cell.value

    def add_slice(self, input, offset, output):
        cell = C.c_uint32()
        rc = lib().go_add_slice(self._h, input, offset, output, C.byref(cell))
        if rc:
            raise OracleError(rc)
        return cell.value

    def add_merge(self, inputs, output):
        arr = (C.c_uint32 * len(inputs))(*inputs)
        cell = C.c_uint32()
        rc = lib().go_add_merge(self._h, arr, len(inputs), output, C.byref(cell))
        if rc:
            raise OracleError(rc)
        return cell.value

    def add_gates2(self, kinds, in0, in1, out):
        kinds = np.ascontiguousarray(kinds, np.uint8)
        in0 = np.ascontiguousarray(in0, np.uint32)
        in1 = np.ascontiguousarray(in1, np.uint32)
        out = np.ascontiguousarray(out, np.uint32)
        rc = lib().go_add_gates2(self._h, _p(kinds), _p(in0), _p(in1), _p(out), len(kinds))
        if rc:
            raise OracleError(rc)

    def build(self):
        return Simulator(lib().go_build(self._h))


class Simulator:
    """gsim ``Simulator`` restated: one stimulus vector per instance."""

    def __init__(self, h):
        self._h = h

    def __del__(self):
        if getattr(self, "_h", None):
            lib().go_sim_free(self._h)
            self._h = None

    def get_wire_width(self, wire):
        out = C.c_uint8()
        rc = lib().go_wire_width(self._h, wire, C.byref(out))
        if rc:
            raise OracleError(rc)
        return out.value

    def set_wire_drive(self, wire, plane0, plane1):
        """plane0/plane1: ints (bit i = bit i of the wire) — LogicState::from_bit_planes."""
        w = self.get_wire_width(wire)
        n = (w + 31) // 32
        p0 = (C.c_uint32 * n)(*[(plane0 >> (32 * i)) & 0xFFFFFFFF for i in range(n)])
        p1 = (C.c_uint32 * n)(*[(plane1 >> (32 * i)) & 0xFFFFFFFF for i in range(n)])
        rc = lib().go_set_wire_drive(self._h, wire, p0, p1, n)
        if rc:
            raise OracleError(rc)

    def run_sim(self, max_steps):
        return lib().go_run_sim(self._h, max_steps)

    @property
    def last_steps(self):
        return lib().go_last_steps(self._h)

    def get_wire_state(self, wire):
        w = self.get_wire_width(wire)
        n = (w + 31) // 32
        p0 = (C.c_uint32 * n)()
        p1 = (C.c_uint32 * n)()
        rc = lib().go_get_wire_state(self._h, wire, p0, p1)
        if rc:
            raise OracleError(rc)
        mask = (1 << w) - 1
        return (sum(int(p0[i]) << (32 * i) for i in range(n)) & mask,
                sum(int(p1[i]) << (32 * i) for i in range(n)) & mask)


class Lanes:
    """L independent scalar simulators of one netlist, driven through 64-lane words."""

    def __init__(self, builder: Builder, n_lanes: int):
        self.n_lanes = n_lanes
        self.words = (n_lanes + 63) // 64
        self._h = lib().go_lanes_new(builder._h, n_lanes)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().go_lanes_free(self._h)
            self._h = None

    def set_drive(self, wire, value, valid, bit=0):
        value = np.ascontiguousarray(value, np.uint64)
        valid = np.ascontiguousarray(valid, np.uint64)
        assert value.size == self.words and valid.size == self.words
        rc = lib().go_lanes_set_drive_bit(self._h, wire, bit, _p(value), _p(valid))
        if rc:
            raise OracleError(rc)

    def get(self, wire, bit=0):
        value = np.zeros(self.words, np.uint64)
        valid = np.zeros(self.words, np.uint64)
        rc = lib().go_lanes_get_bit(self._h, wire, bit, _p(value), _p(valid))
        if rc:
            raise OracleError(rc)
        return value, valid

    def get_all(self, n_wires):
        value = np.zeros((n_wires, self.words), np.uint64)
        valid = np.zeros((n_wires, self.words), np.uint64)
        lib().go_lanes_get_all(self._h, _p(value), _p(valid))
        return value, valid

    def run(self, max_steps, threads=1):
        """Returns (status[lane] u8, steps[lane] u64, total component evaluations)."""
        status = np.zeros(self.n_lanes, np.uint8)
        steps = np.zeros(self.n_lanes, np.uint64)
        evals = lib().go_lanes_run(self._h, max_steps, threads, _p(status), _p(steps))
        return status, steps, int(evals)


class BitSliced:
    """Dense 64-lane bit-sliced model (oracle/bitsliced.c); 1-bit wires only."""

    def __init__(self, builder: Builder, n_lanes: int, n_wires: int):
        self.n_lanes = n_lanes
        self.n_wires = n_wires
        self.words = (n_lanes + 63) // 64
        self._h = lib().bs_new(builder._h, n_lanes)
        if not self._h:
            raise ValueError("bit-sliced model needs 1-bit wires")

    def __del__(self):
        if getattr(self, "_h", None):
            lib().bs_free(self._h)
            self._h = None

    @property
    def topological(self):
        return bool(lib().bs_is_topological(self._h))

    @property
    def last_applications(self):
        return lib().bs_last_applications(self._h)

    def set_drive(self, wire, value, valid):
        value = np.ascontiguousarray(value, np.uint64)
        valid = np.ascontiguousarray(valid, np.uint64)
        assert value.size == self.words and valid.size == self.words
        lib().bs_set_drive(self._h, wire, _p(value), _p(valid))

    def get(self, wire):
        value = np.zeros(self.words, np.uint64)
        valid = np.zeros(self.words, np.uint64)
        lib().bs_get(self._h, wire, _p(value), _p(valid))
        return value, valid

    def get_all(self):
        value = np.zeros((self.n_wires, self.words), np.uint64)
        valid = np.zeros((self.n_wires, self.words), np.uint64)
        lib().bs_get_all(self._h, _p(value), _p(valid))
        return value, valid

    def run(self, max_steps):
        """Returns (conflict[words], unsettled[words]) lane masks."""
        conflict = np.zeros(self.words, np.uint64)
        unsettled = np.zeros(self.words, np.uint64)
        lib().bs_run(self._h, max_steps, _p(conflict), _p(unsettled))
        return conflict, unsettled

    def eval_dag(self):
        conflict = np.zeros(self.words, np.uint64)
        rc = lib().bs_eval_dag(self._h, _p(conflict))
        if rc:
            raise ValueError("netlist is not a single-driver DAG in creation order")
        return conflict


def status_from_masks(conflict, unsettled, n_lanes):
    """Per-lane SimulationRunResult ordinals from the two lane masks."""
    lanes = np.arange(n_lanes)
    c = (conflict[lanes // 64] >> (lanes % 64).astype(np.uint64)) & np.uint64(1)
    u = (unsettled[lanes // 64] >> (lanes % 64).astype(np.uint64)) & np.uint64(1)
    return np.where(c == 1, RUN_CONFLICT, np.where(u == 1, RUN_MAX_STEPS, RUN_OK)).astype(np.uint8)
