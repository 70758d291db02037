"""`B200Server`: the `SimServer` implementation (trait at reference
crates/digilogic_netcode/src/server.rs:43-102), sibling of `GsimServer`
(crates/digilogic_gsim/src/lib.rs:101-239).  The state machine itself is compiled C++ behind
`b2_server_*`; this class is the harness-side binding with the trait's method names.
`ServerResult<T>` maps to: return T, or raise `ServerErrorException(ServerError.X)`."""
from __future__ import annotations

import ctypes as C
import enum

import numpy as np

from . import _ffi
from .sim_state import SimState


class ServerError(enum.IntEnum):
    """crates/digilogic_netcode/src/lib.rs:67-85 (ordinal order)."""
    Unsupported = 0
    Other = 1
    InvalidState = 2
    OutOfResources = 3
    InvalidNetId = 4
    WidthMismatch = 5
    WidthIncompatible = 6
    OutOfRange = 7
    InvalidInputCount = 8
    MaxStepsReached = 9
    DriverConflict = 10


class ServerErrorException(Exception):
    def __init__(self, error: ServerError):
        self.error = error
        super().__init__(error.name)


def _res(rc):
    if rc != 0:
        raise ServerErrorException(ServerError(rc - 1))


_KINDS = {"and": 0, "or": 1, "xor": 2, "nand": 3, "nor": 4, "xnor": 5}


class B200Server:
    def __init__(self):
        self._lib = _ffi.lib()
        self._h = self._lib.b2_server_new()

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.b2_server_free(self._h)
            self._h = None

    def max_clients(self) -> int:
        return self._lib.b2_server_max_clients(self._h)

    def client_connected(self, client_id: int):
        self._lib.b2_server_client_connected(self._h, client_id)

    def client_disconnected(self, client_id: int):
        self._lib.b2_server_client_disconnected(self._h, client_id)

    def begin_build(self, client_id: int):
        _res(self._lib.b2_server_begin_build(self._h, client_id))

    def end_build(self, client_id: int):
        _res(self._lib.b2_server_end_build(self._h, client_id))

    def add_net(self, client_id: int, width: int) -> int:
        out = C.c_uint32()
        _res(self._lib.b2_server_add_net(self._h, client_id, width, C.byref(out)))
        return out.value

    def _gate(self, kind, client_id, width, inputs, output):
        arr = (C.c_uint32 * len(inputs))(*inputs)
        cell = C.c_uint32()
        _res(self._lib.b2_server_add_gate(self._h, client_id, _KINDS[kind], width, arr, len(inputs), output, C.byref(cell)))
        return cell.value

    def add_and_gate(self, client_id, width, inputs, output): return self._gate("and", client_id, width, inputs, output)
    def add_or_gate(self, client_id, width, inputs, output): return self._gate("or", client_id, width, inputs, output)
    def add_xor_gate(self, client_id, width, inputs, output): return self._gate("xor", client_id, width, inputs, output)
    def add_nand_gate(self, client_id, width, inputs, output): return self._gate("nand", client_id, width, inputs, output)
    def add_nor_gate(self, client_id, width, inputs, output): return self._gate("nor", client_id, width, inputs, output)
    def add_xnor_gate(self, client_id, width, inputs, output): return self._gate("xnor", client_id, width, inputs, output)

    def add_not_gate(self, client_id, width, input, output):
        cell = C.c_uint32()
        _res(self._lib.b2_server_add_not_gate(self._h, client_id, width, input, output, C.byref(cell)))
        return cell.value

    def add_mux(self, client_id, width, inputs, output):
        """inputs[0] is the select net (crates/digilogic_gsim/src/lib.rs:185-187)."""
        arr = (C.c_uint32 * len(inputs))(*inputs)
        cell = C.c_uint32()
        _res(self._lib.b2_server_add_mux(self._h, client_id, width, arr, len(inputs), output, C.byref(cell)))
        return cell.value

    def set_net_drive(self, client_id, net, bit_plane_0: bytes, bit_plane_1: bytes):
        b0 = (C.c_uint8 * max(len(bit_plane_0), 1)).from_buffer_copy(bytes(bit_plane_0) or b"\0")
        b1 = (C.c_uint8 * max(len(bit_plane_1), 1)).from_buffer_copy(bytes(bit_plane_1) or b"\0")
        _res(self._lib.b2_server_set_net_drive(self._h, client_id, net, b0, len(bit_plane_0), b1, len(bit_plane_1)))

    def eval(self, client_id, max_steps: int):
        _res(self._lib.b2_server_eval(self._h, client_id, max_steps))

    def get_net_state(self, client_id, net):
        """-> (width, plane0, plane1): the two 32-byte scratch planes, bytes past ceil(width/8) stale."""
        width = C.c_uint8()
        p0, p1 = C.POINTER(C.c_uint8)(), C.POINTER(C.c_uint8)()
        _res(self._lib.b2_server_get_net_state(self._h, client_id, net, C.byref(width), C.byref(p0), C.byref(p1)))
        return width.value, bytes(p0[:32]), bytes(p1[:32])

    # -- Adapter::sim_state (server.rs:375-383), the per-net way and the bulk way
    def sim_state_per_net(self, client_id, nets) -> SimState:
        st = SimState()
        for net in nets:
            width, p0, p1 = self.get_net_state(client_id, net)
            st.push_net(width, p0, p1)
        return st

    def sim_state(self, client_id, nets, total_bits_hint: int | None = None) -> SimState:
        nets = np.ascontiguousarray(nets, np.uint32)
        cap = (total_bits_hint if total_bits_hint is not None else 255 * max(nets.size, 1)) // 8 + 1
        p0, p1 = np.zeros(cap, np.uint8), np.zeros(cap, np.uint8)
        bits = C.c_uint64()
        _res(self._lib.b2_server_sim_state(self._h, client_id, nets.ctypes.data_as(C.c_void_p), nets.size,
                                           p0.ctypes.data_as(C.c_void_p), p1.ctypes.data_as(C.c_void_p), cap, C.byref(bits)))
        n = bits.value // 8 + 1
        return SimState.from_planes(bits.value, p0[:n].tobytes(), p1[:n].tobytes())
