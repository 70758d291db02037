"""SimState packing against the reference's own 18 golden vectors
(crates/digilogic_netcode/src/lib.rs:324-592, extracted by tests/golden/make_sim_state_vectors.py):
the host mirror on CPU, and the GPU bulk read-back (b2_pack_sim_state) on the same vectors."""
import json
import os

import numpy as np
import pytest

from digilogic_b200.sim_state import SimState

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "sim_state_vectors.json")) as f:
    VECTORS = json.load(f)["vectors"]
INSERTS = [v for v in VECTORS if v["kind"] == "insert"]
READS = [v for v in VECTORS if v["kind"] == "read"]


def test_vector_count():
    assert len(VECTORS) == 18 and len(INSERTS) == 9 and len(READS) == 9


@pytest.mark.parametrize("vec", INSERTS, ids=lambda v: v["name"])
def test_push_net_golden(vec):
    st = SimState()
    offsets = [st.push_net(p["width"], bytes(p["plane0"]), bytes(p["plane1"])) for p in vec["pushes"]]
    assert offsets == vec["offsets"]
    assert list(st.bit_plane_0) == vec["plane0"] and list(st.bit_plane_1) == vec["plane1"]


@pytest.mark.parametrize("vec", READS, ids=lambda v: v["name"])
def test_get_net_golden(vec):
    st = SimState(0, vec["bit_len"], bytes(vec["plane0"]), bytes(vec["plane1"]))
    for r in vec["reads"]:
        p0, p1 = st.get_net(r["offset"], r["width"])
        assert list(p0) == r["plane0"] and list(p1) == r["plane1"]


def test_push_masks_stale_bytes():
    """get_net_state hands out 32-byte scratch planes with stale bytes past the width (gsim/lib.rs:232-237);
    push_net must mask them (netcode lib.rs:256-258)."""
    st = SimState()
    st.push_net(3, bytes([0xFF] * 32), bytes([0xFF] * 32))
    st.push_net(1, b"\0", b"\1")
    assert st.bit_len == 4 and st.bit_plane_0 == bytes([0b0111]) and st.bit_plane_1 == bytes([0b1111])


@pytest.mark.gpu
@pytest.mark.parametrize("vec", INSERTS, ids=lambda v: v["name"])
def test_gpu_bulk_pack_golden(vec):
    """Drive wires of the golden widths with the golden planes; the engine's bulk SimState
    read-back and the per-net path through B200Server must both reproduce the golden bytes."""
    from digilogic_b200.server import B200Server
    sv = B200Server()
    sv.client_connected(7)
    sv.begin_build(7)
    nets = [sv.add_net(7, p["width"]) for p in vec["pushes"]]
    sv.end_build(7)
    for net, p in zip(nets, vec["pushes"]):
        sv.set_net_drive(7, net, bytes(p["plane0"]), bytes(p["plane1"]))
    sv.eval(7, 10_000)
    for st in (sv.sim_state(7, nets), sv.sim_state_per_net(7, nets)):
        assert st.bit_len == sum(p["width"] for p in vec["pushes"])
        assert list(st.bit_plane_0) == vec["plane0"] and list(st.bit_plane_1) == vec["plane1"]
