"""TEST HARNESS (not part of the product package).  The caller side of the drop-in boundary, restated: the netcode `Adapter` and
`process_message` (reference crates/digilogic_netcode/src/server.rs:233-470) and the msgpack wire
form of `ClientMessage` / `ServerMessage` / `SimState` (crates/digilogic_netcode/src/lib.rs:87-179),
so that a client session can be replayed byte for byte against any `SimServer` implementation —
here `B200Server`.  The UDP/renet transport itself stays in the reference and is out of scope.

Wire form: `rmp_serde::to_vec` (server.rs:16,21; client.rs:28) = compact msgpack — structs as
positional arrays, unit enum variants as their name, other variants as a one-entry map
{name: payload} with struct-variant fields as an array, newtype structs (`NetId`) as their inner
value, `#[serde(with = "serde_bytes")]` fields as `bin`, plain `Vec<u8>` as an array of ints.
(to be confirmed against one recorded packet wherever the GUI can run.)
"""
from __future__ import annotations

import msgpack

from digilogic_b200.server import ServerError, ServerErrorException
from digilogic_b200.sim_state import SimState

GATE_KINDS = {"AddAndGate": "add_and_gate", "AddOrGate": "add_or_gate", "AddXorGate": "add_xor_gate",
              "AddNandGate": "add_nand_gate", "AddNorGate": "add_nor_gate", "AddXnorGate": "add_xnor_gate"}


# ------------------------------------------------------------------------------ wire form

def encode_client_message(msg_id: int, kind: str, **fields) -> bytes:
    """ClientMessage{id, kind} (lib.rs:159-163); `kind` is a ClientMessageKind variant name."""
    order = {
        "BeginBuild": None, "EndBuild": None, "QueryReport": None, "QueryUpdate": None,
        "AddNet": ["width"], "AddNotGate": ["width", "input", "output"], "AddMux": ["width", "inputs", "output"],
        "SetNetDrive": ["net", "bit_plane_0", "bit_plane_1"], "Eval": ["max_steps"],
        **{k: ["width", "inputs", "output"] for k in GATE_KINDS},
    }[kind]
    if order is None:
        payload = kind
    else:
        vals = []
        for f in order:
            v = fields[f]
            vals.append(list(v) if isinstance(v, (bytes, bytearray, tuple)) else v)   # Vec<u8> / Vec<NetId>: arrays
        payload = {kind: vals}
    return msgpack.packb([msg_id, payload], use_bin_type=True)


def decode_client_message(data: bytes):
    """-> (id, kind name, {field: value})."""
    msg_id, payload = msgpack.unpackb(data, raw=False, strict_map_key=False)
    if isinstance(payload, str):
        return msg_id, payload, {}
    (kind, vals), = payload.items()
    names = {"AddNet": ["width"], "AddNotGate": ["width", "input", "output"], "AddMux": ["width", "inputs", "output"],
             "SetNetDrive": ["net", "bit_plane_0", "bit_plane_1"], "Eval": ["max_steps"],
             **{k: ["width", "inputs", "output"] for k in GATE_KINDS}}[kind]
    return msg_id, kind, dict(zip(names, vals))


def encode_sim_state(st: SimState):
    return [st.order, st.bit_len, st.bit_plane_0, st.bit_plane_1]        # serde_bytes -> bin


def encode_server_message(kind: str, **fields) -> bytes:
    """ServerMessage (lib.rs:87-93): Error{id, error} | Ready | BuildingFinished | Report(SimState)."""
    if kind in ("Ready", "BuildingFinished"):
        return msgpack.packb(kind, use_bin_type=True)
    if kind == "Error":
        return msgpack.packb({"Error": [fields["id"], ServerError(fields["error"]).name]}, use_bin_type=True)
    if kind == "Report":
        return msgpack.packb({"Report": encode_sim_state(fields["sim_state"])}, use_bin_type=True)
    raise ValueError(kind)


def decode_server_message(data: bytes):
    payload = msgpack.unpackb(data, raw=False, strict_map_key=False)
    if isinstance(payload, str):
        return payload, {}
    (kind, vals), = payload.items()
    if kind == "Error":
        return kind, {"id": vals[0], "error": ServerError[vals[1]]}
    order, bit_len, p0, p1 = vals
    return kind, {"sim_state": SimState(order, bit_len, bytes(p0), bytes(p1))}


# ------------------------------------------------------------------------------ Adapter

class Adapter:
    """`Adapter<S: SimServer>` (server.rs:233-384): maps the client's 0-based net ids to the engine's
    ids, records cell ids, and assembles the `SimState` of all nets after an eval."""

    def __init__(self, inner, bulk_sim_state: bool = True):
        self.inner = inner
        self.client_state = {}      # client id -> {"net_map": [...], "cell_map": [...], "order": 0}
        self.bulk_sim_state = bulk_sim_state and hasattr(inner, "sim_state")

    def _state(self, client_id):
        try:
            return self.client_state[client_id]
        except KeyError:            # the reference expect()s here (server.rs:262-275)
            raise ServerErrorException(ServerError.Other)

    @staticmethod
    def _net(state, net):           # the reference indexes the Vec and would panic (server.rs:141-148)
        if not 0 <= net < len(state["net_map"]):
            raise ServerErrorException(ServerError.InvalidNetId)
        return state["net_map"][net]

    def client_connected(self, client_id):
        self.client_state[client_id] = {"net_map": [], "cell_map": [], "order": 0}
        self.inner.client_connected(client_id)

    def client_disconnected(self, client_id):
        self.client_state.pop(client_id, None)
        self.inner.client_connected(client_id)      # sic — server.rs:296-299 calls client_connected

    def begin_build(self, client_id):
        st = self._state(client_id)
        st["net_map"].clear()
        st["cell_map"].clear()
        self.inner.begin_build(client_id)

    def end_build(self, client_id):
        self.inner.end_build(client_id)

    def add_net(self, client_id, width):
        net_id = self.inner.add_net(client_id, width)
        self._state(client_id)["net_map"].append(net_id)

    def add_gate(self, method, client_id, width, inputs, output):
        st = self._state(client_id)
        ins = [self._net(st, i) for i in inputs]
        out = self._net(st, output)
        st["cell_map"].append(getattr(self.inner, method)(client_id, width, ins, out))

    def add_not_gate(self, client_id, width, input, output):
        st = self._state(client_id)
        st["cell_map"].append(self.inner.add_not_gate(client_id, width, self._net(st, input), self._net(st, output)))

    def set_net_drive(self, client_id, net, bit_plane_0, bit_plane_1):
        st = self._state(client_id)
        self.inner.set_net_drive(client_id, self._net(st, net), bytes(bit_plane_0), bytes(bit_plane_1))

    def eval(self, client_id, max_steps):
        self.inner.eval(client_id, max_steps)

    def sim_state(self, client_id) -> SimState:
        """server.rs:375-383 — per net through get_net_state + push_net, or (bulk_sim_state) in one
        device pass through the engine's b2_server_sim_state; both give the same bytes."""
        st = self._state(client_id)
        nets = st["net_map"]
        if self.bulk_sim_state:
            out = self.inner.sim_state(client_id, nets) if nets else SimState()
            out.order = st["order"]
            return out
        out = SimState(order=st["order"])
        for net in nets:
            width, p0, p1 = self.inner.get_net_state(client_id, net)
            out.push_net(width, p0, p1)
        return out


def process_message(adapter: Adapter, client_id: int, data: bytes) -> list[bytes]:
    """`process_message` + the error reply of the poll loop (server.rs:386-470,513-529): one encoded
    ClientMessage in, the encoded ServerMessages it causes out."""
    msg_id, kind, f = decode_client_message(data)
    out = []
    try:
        if kind == "BeginBuild":
            adapter.begin_build(client_id)
        elif kind == "EndBuild":
            adapter.end_build(client_id)
            out.append(encode_server_message("BuildingFinished"))
        elif kind == "AddNet":
            adapter.add_net(client_id, f["width"])
        elif kind in GATE_KINDS:
            adapter.add_gate(GATE_KINDS[kind], client_id, f["width"], f["inputs"], f["output"])
        elif kind == "AddNotGate":
            adapter.add_not_gate(client_id, f["width"], f["input"], f["output"])
        elif kind == "AddMux":
            adapter.add_gate("add_mux", client_id, f["width"], f["inputs"], f["output"])
        elif kind == "SetNetDrive":
            adapter.set_net_drive(client_id, f["net"], f["bit_plane_0"], f["bit_plane_1"])
        elif kind == "Eval":
            adapter.eval(client_id, f["max_steps"])
            out.append(encode_server_message("Report", sim_state=adapter.sim_state(client_id)))
        elif kind == "QueryReport":
            out.append(encode_server_message("Report", sim_state=adapter.sim_state(client_id)))
        elif kind == "QueryUpdate":
            # sent bare on the unreliable data channel (send_sim_state, server.rs:20-23), not as a ServerMessage
            out.append(msgpack.packb(encode_sim_state(adapter.sim_state(client_id)), use_bin_type=True))
        else:
            raise ServerErrorException(ServerError.Unsupported)
    except ServerErrorException as e:
        out.append(encode_server_message("Error", id=msg_id, error=e.error))
    return out
