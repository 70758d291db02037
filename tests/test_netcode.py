"""The caller side of the boundary: `Adapter` / `process_message` restated from the reference
(crates/digilogic_netcode/src/server.rs:233-470) and the msgpack form of the messages — on CPU
against a fake SimServer, on GPU a whole client session replayed as bytes against B200Server."""
import json
import os

import msgpack
import pytest

from digilogic_b200 import importers
import netcode_harness as netcode
from digilogic_b200.server import ServerError, ServerErrorException
from digilogic_b200.sim_state import SimState

HERE = os.path.dirname(os.path.abspath(__file__))


class FakeServer:
    """A minimal SimServer: engine net ids are 100, 101, ...; only AND and NOT exist (the trait's
    other gate methods default to Err(Unsupported), server.rs:28-41,58-84); nets hold their drive."""

    def __init__(self):
        self.calls = []
        self.nets = {}

    def client_connected(self, cid): self.calls.append(("client_connected", cid))
    def client_disconnected(self, cid): self.calls.append(("client_disconnected", cid))
    def begin_build(self, cid): self.nets.clear(); self.calls.append(("begin_build", cid))
    def end_build(self, cid): self.calls.append(("end_build", cid))

    def add_net(self, cid, width):
        nid = 100 + len(self.nets)
        self.nets[nid] = [width, b"\0" * 32, b"\0" * 32]
        return nid

    def add_and_gate(self, cid, width, inputs, output):
        self.calls.append(("add_and_gate", width, tuple(inputs), output))
        return 7

    def add_not_gate(self, cid, width, input, output):
        self.calls.append(("add_not_gate", width, input, output))
        return 8

    def __getattr__(self, name):
        if name.startswith("add_"):
            def unsupported(*a, **k):
                raise ServerErrorException(ServerError.Unsupported)
            return unsupported
        raise AttributeError(name)

    def set_net_drive(self, cid, net, p0, p1):
        self.nets[net][1] = bytes(p0).ljust(32, b"\xEE")     # stale bytes past the width, like the scratch planes
        self.nets[net][2] = bytes(p1).ljust(32, b"\xEE")

    def eval(self, cid, max_steps):
        if max_steps == 0:
            raise ServerErrorException(ServerError.MaxStepsReached)

    def get_net_state(self, cid, net):
        return tuple(self.nets[net])


def test_message_wire_form_roundtrip():
    raw = netcode.encode_client_message(5, "AddAndGate", width=1, inputs=[0, 1], output=2)
    assert msgpack.unpackb(raw) == [5, {"AddAndGate": [1, [0, 1], 2]}]
    assert netcode.decode_client_message(raw) == (5, "AddAndGate", {"width": 1, "inputs": [0, 1], "output": 2})
    assert msgpack.unpackb(netcode.encode_client_message(6, "BeginBuild")) == [6, "BeginBuild"]
    raw = netcode.encode_client_message(7, "SetNetDrive", net=3, bit_plane_0=b"\1", bit_plane_1=b"\1")
    assert msgpack.unpackb(raw) == [7, {"SetNetDrive": [3, [1], [1]]}]        # plain Vec<u8>: array of ints
    assert msgpack.unpackb(netcode.encode_client_message(8, "Eval", max_steps=10_000)) == [8, {"Eval": [10000]}]
    st = SimState()
    st.push_net(7, b"\x2A", b"\x55")
    raw = netcode.encode_server_message("Report", sim_state=st)
    assert msgpack.unpackb(raw) == {"Report": [0, 7, b"\x2A", b"\x55"]}       # serde_bytes: bin
    kind, f = netcode.decode_server_message(raw)
    assert kind == "Report" and f["sim_state"].get_net(0, 7) == (b"\x2A", b"\x55")
    kind, f = netcode.decode_server_message(netcode.encode_server_message("Error", id=9, error=ServerError.DriverConflict))
    assert (kind, f) == ("Error", {"id": 9, "error": ServerError.DriverConflict})
    assert netcode.decode_server_message(netcode.encode_server_message("BuildingFinished")) == ("BuildingFinished", {})


def test_adapter_translates_ids_and_reports_errors():
    fake = FakeServer()
    ad = netcode.Adapter(fake, bulk_sim_state=False)
    ad.client_connected(1)
    send = lambda i, kind, **f: [netcode.decode_server_message(m) for m in
                                 netcode.process_message(ad, 1, netcode.encode_client_message(i, kind, **f))]
    assert send(1, "BeginBuild") == []
    for i in range(3):
        assert send(2 + i, "AddNet", width=1 if i < 2 else 9) == []
    assert send(5, "AddAndGate", width=1, inputs=[0, 1], output=1) == []
    assert fake.calls[-1] == ("add_and_gate", 1, (100, 101), 101), "client ids 0,1 -> engine ids 100,101"
    assert send(6, "AddNotGate", width=1, input=1, output=0) == []
    assert fake.calls[-1] == ("add_not_gate", 1, 101, 100)
    assert send(7, "AddXorGate", width=1, inputs=[0, 1], output=1) == [("Error", {"id": 7, "error": ServerError.Unsupported})]
    assert send(8, "AddAndGate", width=1, inputs=[0, 5], output=1) == [("Error", {"id": 8, "error": ServerError.InvalidNetId})]
    assert send(9, "EndBuild") == [("BuildingFinished", {})]
    assert send(10, "SetNetDrive", net=2, bit_plane_0=[0xAA, 0x01], bit_plane_1=[0x55, 0x01]) == []
    assert send(11, "SetNetDrive", net=0, bit_plane_0=[1], bit_plane_1=[1]) == []
    (kind, f), = send(12, "Eval", max_steps=10_000)
    st = f["sim_state"]
    assert kind == "Report" and st.bit_len == 11
    assert st.get_net(0, 1) == (b"\1", b"\1") and st.get_net(1, 1) == (b"\0", b"\0")
    assert st.get_net(2, 9) == (b"\xAA\x01", b"\x55\x01"), "stale scratch bytes past the width are masked"
    assert send(13, "Eval", max_steps=0) == [("Error", {"id": 13, "error": ServerError.MaxStepsReached})]
    raw, = netcode.process_message(ad, 1, netcode.encode_client_message(14, "QueryUpdate"))
    assert msgpack.unpackb(raw)[1] == 11, "QueryUpdate sends the bare SimState"
    assert send(15, "BeginBuild") == [] and ad.client_state[1]["net_map"] == []
    ad.client_disconnected(1)
    assert fake.calls[-1] == ("client_connected", 1), "server.rs:296-299 calls client_connected on disconnect"
    assert netcode.decode_server_message(netcode.process_message(ad, 1, netcode.encode_client_message(16, "BeginBuild"))[0]) == \
        ("Error", {"id": 16, "error": ServerError.Other})


@pytest.mark.gpu
@pytest.mark.parametrize("bulk", [True, False])
def test_client_session_as_bytes_against_b200server(bulk):
    """The GUI client's session (client.rs:288-427,178-193,253-269) for the reference's medium.dig
    circuit, as msgpack bytes in and out: build, then for a few input vectors SetNetDrive* + Eval."""
    from digilogic_b200.server import B200Server
    with open(os.path.join(HERE, "golden", "fixtures.json")) as f:
        circ = importers.ImportedCircuit.from_json(json.load(f)["circuits"]["medium.dig"])
    from test_fixtures import TRUTH
    order, expected = TRUTH["medium"]
    ad = netcode.Adapter(B200Server(), bulk_sim_state=bulk)
    cid, mid = 1234, [0]

    def send(kind, **f):
        mid[0] += 1
        return [netcode.decode_server_message(m) for m in netcode.process_message(ad, cid, netcode.encode_client_message(mid[0], kind, **f))]

    ad.client_connected(cid)
    assert send("BeginBuild") == []
    for _ in range(circ.n_nets):
        assert send("AddNet", width=1) == []
    names = {importers.AND: "AddAndGate", importers.OR: "AddOrGate", importers.XOR: "AddXorGate"}
    for kind, ins, out in circ.gates:
        if kind == importers.NOT:
            assert send("AddNotGate", width=1, input=ins[0], output=out) == []
        else:
            assert send(names[kind], width=1, inputs=ins, output=out) == []
    assert send("EndBuild") == [("BuildingFinished", {})]
    for row in (0, 1, 21, 42, 63):
        for k, name in enumerate(order):
            bit = (row >> (len(order) - 1 - k)) & 1
            assert send("SetNetDrive", net=circ.inputs[name][0], bit_plane_0=[bit], bit_plane_1=[1]) == []
        (kind, f), = send("Eval", max_steps=10_000)
        st = f["sim_state"]
        assert kind == "Report" and st.bit_len == circ.n_nets
        for out, table in expected.items():
            assert st.get_net(circ.outputs[out][0], 1) == (bytes([(table >> row) & 1]), b"\1"), (row, out)
    assert send("AddNet", width=1) == [("Error", {"id": mid[0], "error": ServerError.InvalidState})]
