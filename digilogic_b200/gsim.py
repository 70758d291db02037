"""Host-side mirror of the gsim API surface the reference uses (names, argument meaning and
error behaviour as at the call sites in crates/digilogic_gsim/src/lib.rs:12-230), over the C ABI."""
from __future__ import annotations

import ctypes as C
import enum

import numpy as np

from . import _ffi
from ._ffi import SimInfo

AND, OR, XOR, NAND, NOR, XNOR, NOT, MUX = range(8)
CMP_EQ, CMP_NE, CMP_LTU, CMP_GTU, CMP_LEU, CMP_GEU, CMP_LTS, CMP_GTS, CMP_LES, CMP_GES = range(10)


class AddComponentError(enum.IntEnum):
    """gsim::AddComponentError (variants named at crates/digilogic_gsim/src/lib.rs:55-66)."""
    TooManyComponents = 1
    InvalidWireId = 2
    WireWidthMismatch = 3
    WireWidthIncompatible = 4
    OffsetOutOfRange = 5
    TooFewInputs = 6
    InvalidInputCount = 7


class SimulationRunResult(enum.IntEnum):
    """gsim::SimulationRunResult (crates/digilogic_gsim/src/lib.rs:68-74)."""
    Ok = 0
    MaxStepsReached = 1
    Err = 2


class B200Error(RuntimeError):
    """A failure of the engine itself (no device, CUDA error, bad argument)."""

    def __init__(self, code, where=""):
        self.code = int(code)
        names = {16: "NullPointer", 17: "InvalidArgument", 18: "NoDevice", 19: "CudaError", 20: "OutOfMemory",
                 21: "OutOfRange", 22: "OutOfWires"}
        detail = _ffi.last_error()
        super().__init__(f"{where}: {names.get(self.code, self.code)}" + (f" ({detail})" if detail else ""))


class ComponentError(Exception):
    def __init__(self, code):
        self.error = AddComponentError(code)
        super().__init__(self.error.name)


def _check(rc, where):
    if rc == 0:
        return
    if 1 <= rc <= 7:
        raise ComponentError(rc)
    raise B200Error(rc, where)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class LogicState:
    """Two-plane value: bit i of plane0 = value, of plane1 = valid (Z=(0,0) X=(1,0) 0=(0,1) 1=(1,1))."""
    __slots__ = ("plane0", "plane1")

    def __init__(self, plane0: int, plane1: int):
        self.plane0, self.plane1 = int(plane0), int(plane1)

    @classmethod
    def from_bit_planes(cls, plane0_words, plane1_words):
        """LogicState::from_bit_planes(&[u32], &[u32]) (lib.rs:209)."""
        p0 = sum(int(w) << (32 * i) for i, w in enumerate(plane0_words))
        p1 = sum(int(w) << (32 * i) for i, w in enumerate(plane1_words))
        return cls(p0, p1)

    @classmethod
    def from_bool(cls, v: bool):
        return cls(1 if v else 0, 1)

    @classmethod
    def from_int(cls, v: int, width: int):
        return cls(v & ((1 << width) - 1), (1 << width) - 1)

    HIGH_Z = None  # filled below
    UNDEFINED = None

    def to_bit_planes(self, width: int):
        """LogicState::to_bit_planes(width) (lib.rs:230): ceil(width/32) u32 words per plane."""
        n = (width + 31) // 32
        return ([(self.plane0 >> (32 * i)) & 0xFFFFFFFF for i in range(n)],
                [(self.plane1 >> (32 * i)) & 0xFFFFFFFF for i in range(n)])

    def to_str(self, width: int) -> str:
        out = []
        for i in reversed(range(width)):
            v, m = (self.plane0 >> i) & 1, (self.plane1 >> i) & 1
            out.append("ZX01"[2 * m + v])
        return "".join(out)

    def __eq__(self, other):
        return isinstance(other, LogicState) and (self.plane0, self.plane1) == (other.plane0, other.plane1)

    def __repr__(self):
        return f"LogicState(plane0={self.plane0:#x}, plane1={self.plane1:#x})"


LogicState.HIGH_Z = LogicState(0, 0)
LogicState.UNDEFINED = LogicState(1, 0)


class SimulatorBuilder:
    """gsim::SimulatorBuilder (lib.rs:12)."""

    def __init__(self):
        self._lib = _ffi.lib()
        self._h = self._lib.b2_builder_new()
        if not self._h:
            raise MemoryError

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.b2_builder_free(self._h)
            self._h = None

    def add_wire(self, width: int = 1):
        """-> WireId, or None when out of ids (lib.rs:135-139)."""
        out = C.c_uint32()
        rc = self._lib.b2_add_wire(self._h, width, C.byref(out))
        if rc == 22:
            return None
        _check(rc, "add_wire")
        return out.value

    def add_wires(self, n: int, width: int = 1) -> int:
        out = C.c_uint32()
        _check(self._lib.b2_add_wires(self._h, width, n, C.byref(out)), "add_wires")
        return out.value

    def get_wire_width(self, wire: int) -> int:
        out = C.c_uint8()
        _check(self._lib.b2_builder_wire_width(self._h, wire, C.byref(out)), "get_wire_width")
        return out.value

    def _gate(self, kind, inputs, output):
        arr = (C.c_uint32 * len(inputs))(*inputs)
        cell = C.c_uint32()
        _check(self._lib.b2_add_gate(self._h, kind, arr, len(inputs), output, C.byref(cell)), "add_gate")
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
        _check(self._lib.b2_add_not(self._h, input, output, C.byref(cell)), "add_not_gate")
        return cell.value

    def add_multiplexer(self, inputs, select, output):
        arr = (C.c_uint32 * len(inputs))(*inputs)
        cell = C.c_uint32()
        _check(self._lib.b2_add_mux(self._h, arr, len(inputs), select, output, C.byref(cell)), "add_multiplexer")
        return cell.value

    def add_buffer(self, input, enable, output):
        """gsim ``SimulatorBuilder::add_buffer``: tri-state buffer (no call site in the reference yet)."""
        cell = C.c_uint32()
        _check(self._lib.b2_add_buffer(self._h, input, enable, output, C.byref(cell)), "add_buffer")
        return cell.value

    def add_add(self, input_a, input_b, output):
        """gsim ``SimulatorBuilder::add_add``: wrapping sum; Undefined if any input bit is not valid."""
        cell = C.c_uint32()
        _check(self._lib.b2_add_add(self._h, input_a, input_b, output, C.byref(cell)), "add_add")
        return cell.value

    def add_sub(self, input_a, input_b, output):
        """gsim ``SimulatorBuilder::add_sub``: wrapping difference; Undefined if any input bit is not valid."""
        cell = C.c_uint32()
        _check(self._lib.b2_add_sub(self._h, input_a, input_b, output, C.byref(cell)), "add_sub")
        return cell.value

    def add_mul(self, input_a, input_b, output):
        """gsim ``SimulatorBuilder::add_mul``: wrapping product; Undefined if any input bit is not valid."""
        cell = C.c_uint32()
        _check(self._lib.b2_add_mul(self._h, input_a, input_b, output, C.byref(cell)), "add_mul")
        return cell.value

    def add_compare(self, op, input_a, input_b, output):
        """gsim ``add_compare_*``: op = CMP_EQ .. CMP_GES; one-bit output, Undefined if any input bit is not valid."""
        cell = C.c_uint32()
        _check(self._lib.b2_add_compare(self._h, op, input_a, input_b, output, C.byref(cell)), "add_compare")
        return cell.value

    def add_compare_equal(self, a, b, o): return self.add_compare(CMP_EQ, a, b, o)
    def add_compare_not_equal(self, a, b, o): return self.add_compare(CMP_NE, a, b, o)
    def add_compare_less_than(self, a, b, o): return self.add_compare(CMP_LTU, a, b, o)
    def add_compare_greater_than(self, a, b, o): return self.add_compare(CMP_GTU, a, b, o)
    def add_compare_less_than_or_equal(self, a, b, o): return self.add_compare(CMP_LEU, a, b, o)
    def add_compare_greater_than_or_equal(self, a, b, o): return self.add_compare(CMP_GEU, a, b, o)
    def add_compare_less_than_signed(self, a, b, o): return self.add_compare(CMP_LTS, a, b, o)
    def add_compare_greater_than_signed(self, a, b, o): return self.add_compare(CMP_GTS, a, b, o)
    def add_compare_less_than_or_equal_signed(self, a, b, o): return self.add_compare(CMP_LES, a, b, o)
    def add_compare_greater_than_or_equal_signed(self, a, b, o): return self.add_compare(CMP_GES, a, b, o)

    def add_shift(self, kind, input, amount, output):
        """kind: "shl" / "lshr" / "ashr" (gsim add_left_shift / add_logical_right_shift / add_arithmetic_right_shift)."""
        cell = C.c_uint32()
        _check(self._lib.b2_add_shift(self._h, {"shl": 0, "lshr": 1, "ashr": 2}[kind], input, amount, output, C.byref(cell)), "add_shift")
        return cell.value

    def add_left_shift(self, input, amount, output): return self.add_shift("shl", input, amount, output)
    def add_logical_right_shift(self, input, amount, output): return self.add_shift("lshr", input, amount, output)
    def add_arithmetic_right_shift(self, input, amount, output): return self.add_shift("ashr", input, amount, output)

    def add_horizontal_gate(self, kind, input, output):
        """gsim ``add_horizontal_{and,or,xor,nand,nor,xnor}_gate``: the gate folded over all bits of one wire."""
        cell = C.c_uint32()
        _check(self._lib.b2_add_horizontal_gate(self._h, kind, input, output, C.byref(cell)), "add_horizontal_gate")
        return cell.value

    def add_horizontal_and_gate(self, input, output): return self.add_horizontal_gate(AND, input, output)
    def add_horizontal_or_gate(self, input, output): return self.add_horizontal_gate(OR, input, output)
    def add_horizontal_xor_gate(self, input, output): return self.add_horizontal_gate(XOR, input, output)
    def add_horizontal_nand_gate(self, input, output): return self.add_horizontal_gate(NAND, input, output)
    def add_horizontal_nor_gate(self, input, output): return self.add_horizontal_gate(NOR, input, output)
    def add_horizontal_xnor_gate(self, input, output): return self.add_horizontal_gate(XNOR, input, output)

    def add_register(self, data_in, data_out, enable, clock):
        """gsim ``SimulatorBuilder::add_register``: loads data_in on a rising clock edge while enable is 1."""
        cell = C.c_uint32()
        _check(self._lib.b2_add_register(self._h, data_in, data_out, enable, clock, C.byref(cell)), "add_register")
        return cell.value

    def add_slice(self, input, offset, output):
        """gsim ``SimulatorBuilder::add_slice``: output = input[offset .. offset + width(output))."""
        cell = C.c_uint32()
        _check(self._lib.b2_add_slice(self._h, input, offset, output, C.byref(cell)), "add_slice")
        return cell.value

    def add_merge(self, inputs, output):
        """gsim ``SimulatorBuilder::add_merge``: output = inputs concatenated, inputs[0] in the low bits."""
        arr = (C.c_uint32 * len(inputs))(*inputs)
        cell = C.c_uint32()
        _check(self._lib.b2_add_merge(self._h, arr, len(inputs), output, C.byref(cell)), "add_merge")
        return cell.value

    def add_gates2(self, kinds, in0, in1, out):
        """Bulk add of one/two-input gates (batch extension)."""
        kinds = np.ascontiguousarray(kinds, np.uint8)
        in0 = np.ascontiguousarray(in0, np.uint32)
        in1 = np.ascontiguousarray(in1, np.uint32)
        out = np.ascontiguousarray(out, np.uint32)
        _check(self._lib.b2_add_gates2(self._h, _p(kinds), _p(in0), _p(in1), _p(out), len(kinds)), "add_gates2")

    OPT_MUX_Z_PASSTHROUGH, OPT_INITIAL_UNDEFINED, OPT_BUFFER_Z_PASSTHROUGH, OPT_COPY_Z_PASSTHROUGH = 1, 2, 3, 4

    def set_option(self, option: int, value: int):
        """b2_builder_set_option: the reconstructed points of gsim's behaviour, one switch each."""
        _check(self._lib.b2_builder_set_option(self._h, option, value), "builder.set_option")

    def build(self) -> "Simulator":
        """SimulatorBuilder::build() (lib.rs:127); leaves this builder empty."""
        h = self._lib.b2_build(self._h)
        if not h:
            raise MemoryError("b2_build failed")
        return Simulator(h)


class Simulator:
    """gsim::Simulator, plus the batch-lane extension."""

    def __init__(self, h):
        self._lib = _ffi.lib()
        self._h = h

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.b2_sim_free(self._h)
            self._h = None

    def clone(self) -> "Simulator":
        """b2_sim_clone: another simulator of the same compiled netlist (shared host arrays and device copy)."""
        h = self._lib.b2_sim_clone(self._h)
        if not h:
            raise MemoryError("b2_sim_clone failed")
        return Simulator(h)

    # ---- gsim surface
    def get_wire_width(self, wire: int) -> int:
        out = C.c_uint8()
        _check(self._lib.b2_wire_width(self._h, wire, C.byref(out)), "get_wire_width")
        return out.value

    def set_wire_drive(self, wire: int, state: LogicState):
        width = self.get_wire_width(wire)
        p0, p1 = state.to_bit_planes(width)
        a0, a1 = (C.c_uint32 * len(p0))(*p0), (C.c_uint32 * len(p1))(*p1)
        _check(self._lib.b2_set_wire_drive(self._h, wire, a0, a1, len(p0)), "set_wire_drive")

    def run_sim(self, max_steps: int) -> SimulationRunResult:
        rc = self._lib.b2_run_sim(self._h, max_steps)
        if rc < 0:
            raise B200Error(-rc, "run_sim")
        return SimulationRunResult(rc)

    def get_wire_state(self, wire: int) -> LogicState:
        width = self.get_wire_width(wire)
        n = (width + 31) // 32
        a0, a1 = (C.c_uint32 * n)(), (C.c_uint32 * n)()
        _check(self._lib.b2_get_wire_state(self._h, wire, a0, a1, n), "get_wire_state")
        mask = (1 << width) - 1
        s = LogicState.from_bit_planes(a0, a1)
        return LogicState(s.plane0 & mask, s.plane1 & mask)

    # ---- batch-lane extension
    @property
    def info(self) -> SimInfo:
        out = SimInfo()
        _check(self._lib.b2_sim_get_info(self._h, C.byref(out)), "get_info")
        return out

    @property
    def pair_info(self) -> tuple:
        """(launches, children, unstored rows) of the level-pair plan the last levelised run used; zeros = per-level kernels."""
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        _check(self._lib.b2_sim_get_pair_info(self._h, C.byref(a), C.byref(b), C.byref(c)), "get_pair_info")
        return a.value, b.value, c.value

    @property
    def hot_rows(self) -> list:
        """Store rows the unit-delay work-list kernel stages in shared memory (b2_sim_get_hot_rows)."""
        rows, n = (C.c_uint32 * 4)(), C.c_uint32()
        _check(self._lib.b2_sim_get_hot_rows(self._h, rows, C.byref(n)), "get_hot_rows")
        return [int(rows[i]) for i in range(n.value)]

    def plan_pairs(self, keep_wires) -> dict:
        """The level-pair plan for `keep_wires` as designated outputs, computed on the host (b2_sim_plan_pairs)."""
        keep = np.ascontiguousarray(keep_wires, np.uint32)
        out = (C.c_uint64 * 6)()
        _check(self._lib.b2_sim_plan_pairs(self._h, _p(keep), keep.size, out), "plan_pairs")
        return dict(zip(("launches", "items", "children", "unstored", "dropped_later", "dropped_now"), (int(x) for x in out)))

    def set_device(self, device: int):
        _check(self._lib.b2_sim_set_device(self._h, device), "set_device")

    def set_stream(self, stream_handle: int | None):
        _check(self._lib.b2_sim_set_stream(self._h, C.c_void_p(stream_handle or 0)), "set_stream")

    OPT_RESIDENT_KERNEL = 1
    OPT_VALID_ELISION = 2
    OPT_LEVEL_GRAPH = 3
    OPT_EVAL_UNROLL = 4
    OPT_FRONTIER = 5
    OPT_STEP_BUDGET_INCLUSIVE = 6
    OPT_DEVICE_LOOP = 7
    OPT_CONFLICT_NEEDS_DISAGREE = 8
    OPT_L2_HINTS = 9
    OPT_PERIOD2_SKIP = 10
    OPT_PDL = 11
    OPT_PAIR_FUSION = 12
    OPT_HOT_STAGING = 13

    def set_option(self, option: int, value: int):
        _check(self._lib.b2_sim_set_option(self._h, option, value), "set_option")

    def synchronize(self):
        _check(self._lib.b2_sim_synchronize(self._h), "synchronize")

    def set_lanes(self, n_lanes: int):
        _check(self._lib.b2_sim_set_lanes(self._h, n_lanes), "set_lanes")

    def max_lanes(self, bytes_: int = 0) -> int:
        return self._lib.b2_sim_max_lanes(self._h, bytes_)

    @property
    def lane_words(self) -> int:
        return self.info.lane_words

    def set_wire_drive_lanes(self, wire, value, valid, bit=0, lane_word0=0):
        value = np.ascontiguousarray(value, np.uint64)
        valid = np.ascontiguousarray(valid, np.uint64)
        assert value.size == valid.size
        _check(self._lib.b2_set_wire_drive_lanes(self._h, wire, bit, lane_word0, value.size, _p(value), _p(valid)),
               "set_wire_drive_lanes")

    def set_drives_lanes(self, wires, value, valid):
        """value/valid: [n_wires, words] uint64 host arrays (numpy, or pinned torch tensors' numpy views)."""
        wires = np.ascontiguousarray(wires, np.uint32)
        assert value.shape == valid.shape and value.shape[0] == wires.size
        _check(self._lib.b2_set_drives_lanes(self._h, _p(wires), wires.size, _p(value), _p(valid), value.strides[0] // 8),
               "set_drives_lanes")

    def set_drives_lanes_ptr(self, wires, value_ptr, valid_ptr, stride_words):
        wires = np.ascontiguousarray(wires, np.uint32)
        _check(self._lib.b2_set_drives_lanes(self._h, _p(wires), wires.size, C.c_void_p(value_ptr), C.c_void_p(valid_ptr),
                                             stride_words), "set_drives_lanes")

    def fill_stimulus(self, wires, seed, lane_word_base=0, mode=0):
        wires = np.ascontiguousarray(wires, np.uint32)
        _check(self._lib.b2_fill_stimulus(self._h, _p(wires), wires.size, seed, lane_word_base, mode), "fill_stimulus")

    def run_sim_lanes(self, max_steps: int, masks: bool = True):
        """-> (worst SimulationRunResult, conflict[words], unsettled[words])."""
        words = max(self.info.lane_words, 1)
        if masks:
            conflict = np.zeros(words, np.uint64)
            unsettled = np.zeros(words, np.uint64)
            rc = self._lib.b2_run_sim_lanes(self._h, max_steps, _p(conflict), _p(unsettled))
        else:
            conflict = unsettled = None
            rc = self._lib.b2_run_sim_lanes(self._h, max_steps, None, None)
        if rc < 0:
            raise B200Error(-rc, "run_sim_lanes")
        return SimulationRunResult(rc), conflict, unsettled

    def run_sim_lanes_async(self, max_steps: int) -> SimulationRunResult:
        rc = self._lib.b2_run_sim_lanes_async(self._h, max_steps)
        if rc < 0:
            raise B200Error(-rc, "run_sim_lanes_async")
        return SimulationRunResult(rc)

    def get_wire_lanes(self, wire, bit=0, lane_word0=0, n_words=None):
        n_words = self.info.lane_words - lane_word0 if n_words is None else n_words
        value = np.zeros(n_words, np.uint64)
        valid = np.zeros(n_words, np.uint64)
        _check(self._lib.b2_get_wire_lanes(self._h, wire, bit, lane_word0, n_words, _p(value), _p(valid)), "get_wire_lanes")
        return value, valid

    def gather_wires_host(self, wires, out_value=None, out_valid=None):
        wires = np.ascontiguousarray(wires, np.uint32)
        words = self.info.lane_words
        if out_value is None:
            out_value = np.zeros((wires.size, words), np.uint64)
            out_valid = np.zeros((wires.size, words), np.uint64)
        _check(self._lib.b2_gather_wires_host(self._h, _p(wires), wires.size, _p(out_value), _p(out_valid),
                                              out_value.strides[0] // 8), "gather_wires_host")
        return out_value, out_valid

    def gather_wires_ptr(self, wires, value_ptr, valid_ptr, stride_words, device: bool, async_: bool = False):
        wires = np.ascontiguousarray(wires, np.uint32)
        fn = self._lib.b2_gather_wires_device if device else (
            self._lib.b2_gather_wires_host_async if async_ else self._lib.b2_gather_wires_host)
        _check(fn(self._h, _p(wires), wires.size, C.c_void_p(value_ptr), C.c_void_p(valid_ptr), stride_words), "gather_wires")

    def pack_sim_state(self, wires):
        """-> (bit_len, plane0 bytes, plane1 bytes) in the netcode SimState layout."""
        wires = np.ascontiguousarray(wires, np.uint32)
        total = sum(self.get_wire_width(int(w)) for w in wires) if wires.size < 4096 else None
        cap = (total // 8 + 1) if total is not None else (wires.size * 255) // 8 + 1
        p0 = np.zeros(cap, np.uint8)
        p1 = np.zeros(cap, np.uint8)
        bits = C.c_uint64()
        _check(self._lib.b2_pack_sim_state(self._h, _p(wires), wires.size, _p(p0), _p(p1), cap, C.byref(bits)), "pack_sim_state")
        n = bits.value // 8 + 1
        return bits.value, p0[:n].tobytes(), p1[:n].tobytes()


def kernel_launches() -> int:
    return _ffi.lib().b2_kernel_launches()
