"""The reference's own sample circuits (imported once into tests/golden/fixtures.json), the
Digital / Yosys readers, and BASELINE configs 1 and 2 — on the oracle (CPU) and through the
engine (GPU), against truth tables and arithmetic."""
import json
import os

import numpy as np
import pytest

import oracle
from digilogic_b200 import importers, netgen
from helpers import lane_mask

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "fixtures.json")) as f:
    FIXTURES = {k: importers.ImportedCircuit.from_json(v) for k, v in json.load(f)["circuits"].items()}

# inputs MSB->LSB of the row index; bit r of the constant = output at row r
TRUTH = {
    "small": (["A", "B"], {"Y": 0x4}),
    "medium": (["A", "B", "Cin", "B_inv", "Op_0", "Op_1"], {"Y": 0x3DB5D35B830B38B0, "Cout": 0xFF0FFFF00F00F000}),
    "large": (["A", "B", "Cin", "Bsel_0", "Bsel_1", "Op_0", "Op_1"],
              {"Y": 0xD33D5BB5DD3355BB3883B00B3388BB00, "Cout": 0xFFFFF00FFFFFFF00F00F0000FF000000}),
    "mux": (["S", "A", "B"], {"out0": 0xAC}),
}
FILES = {"small": ["small.dig", "small.yosys"], "medium": ["medium.dig", "medium.yosys"],
         "large": ["large.dig", "large.yosys", "large_opt.yosys"], "mux": ["mux.dig"]}
CASES = [(c, f) for c, fs in FILES.items() for f in fs]


def pack_lanes(bits):
    """bool/0-1 array over lanes -> uint64 lane words."""
    bits = np.asarray(bits, np.uint8)
    pad = (-len(bits)) % 64
    return np.packbits(np.concatenate([bits, np.zeros(pad, np.uint8)]), bitorder="little").view(np.uint64).copy()


def unpack_lanes(words, n):
    return np.unpackbits(np.ascontiguousarray(words).view(np.uint8), bitorder="little")[:n]


def exhaustive_drives(order, n_lanes):
    rows = np.arange(n_lanes)
    n = len(order)
    return {name: pack_lanes((rows >> (n - 1 - k)) & 1) for k, name in enumerate(order)}


def truth_int(words, n_lanes):
    return sum(int(b) << r for r, b in enumerate(unpack_lanes(words, n_lanes)))


@pytest.mark.parametrize("circuit,fname", CASES)
def test_reference_fixtures_on_oracle(circuit, fname):
    order, expected = TRUTH[circuit]
    circ = FIXTURES[fname]
    n_lanes = 1 << len(order)
    lanes = oracle.Lanes(circ.emit(oracle.Builder()), n_lanes)
    ones = lane_mask(n_lanes)
    for name, words in exhaustive_drives(order, n_lanes).items():
        lanes.set_drive(circ.inputs[name][0], words, ones)
    status, _, _ = lanes.run(10_000)
    assert (status == 0).all()
    for out, want in expected.items():
        v, m = lanes.get(circ.outputs[out][0])
        assert (m == ones).all()
        assert truth_int(v, n_lanes) == want, (fname, out)


def test_adder_dig_roundtrip_and_exhaustive_on_oracle():
    """BASELINE config 1: the adder authored as a .dig, read back, 512 vectors one gsim run each."""
    circ = importers.load_dig(importers.adder_dig(4), "adder4")
    assert len(circ.gates) == 20 and circ.n_nets == 9 + 20
    sim = circ.emit(oracle.Builder()).build()
    for vec in range(512):
        a, b, cin = vec & 15, (vec >> 4) & 15, vec >> 8
        for i in range(4):
            sim.set_wire_drive(circ.inputs[f"A{i}"][0], (a >> i) & 1, 1)
            sim.set_wire_drive(circ.inputs[f"B{i}"][0], (b >> i) & 1, 1)
        sim.set_wire_drive(circ.inputs["Cin"][0], cin, 1)
        assert sim.run_sim(10_000) == oracle.RUN_OK
        got = sum(sim.get_wire_state(circ.outputs[f"S{i}"][0])[0] << i for i in range(4))
        got |= sim.get_wire_state(circ.outputs["Cout"][0])[0] << 4
        assert got == a + b + cin


def _multiplier_inputs(circ, n, n_lanes, seed):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 1 << n, n_lanes, dtype=np.uint64)
    b = rng.integers(0, 1 << n, n_lanes, dtype=np.uint64)
    drives = {}
    for j in range(n):
        drives[circ.inputs["A"][j]] = pack_lanes((a >> np.uint64(j)) & np.uint64(1))
        drives[circ.inputs["B"][j]] = pack_lanes((b >> np.uint64(j)) & np.uint64(1))
    return a, b, drives


def _product(get_bit, circ, n, n_lanes):
    p = np.zeros(n_lanes, dtype=object)
    for j in range(2 * n):
        p += unpack_lanes(get_bit(circ.outputs["P"][j]), n_lanes).astype(object) << j
    return p


def test_multiplier_yosys_on_oracle():
    """BASELINE config 2 (CPU sample): 32x32 array multiplier through the Yosys reader, P = A*B."""
    n, n_lanes = 32, 1024
    circ = importers.load_yosys(importers.multiplier_yosys(n))
    assert len(circ.gates) == 5888 and list(circ.inputs) == ["A", "B"] and len(circ.outputs["P"]) == 64
    bs = oracle.BitSliced(circ.emit(oracle.Builder()), n_lanes, circ.n_nets)
    a, b, drives = _multiplier_inputs(circ, n, n_lanes, 0xD1610001)
    ones = lane_mask(n_lanes)
    for w, words in drives.items():
        bs.set_drive(w, words, ones)
    conflict, unsettled = bs.run(10_000)
    assert not conflict.any() and not unsettled.any()
    p = _product(lambda w: bs.get(w)[0], circ, n, n_lanes)
    assert (p == a.astype(object) * b.astype(object)).all()


@pytest.mark.skipif(not os.path.isdir("/root/reference/crates/digilogic/assets/testdata"), reason="reference tree not present")
def test_golden_fixtures_match_reference_files():
    """In the build container: the committed JSON is what the readers produce from the reference's files."""
    src = "/root/reference/crates/digilogic/assets/testdata"
    for fname, circ in FIXTURES.items():
        text = open(os.path.join(src, fname)).read()
        fresh = importers.load_dig(text, fname) if fname.endswith(".dig") else importers.load_yosys(text)
        fresh.name = fname
        assert fresh.to_json() == circ.to_json()


# ------------------------------------------------------------------------------ GPU

@pytest.mark.gpu
@pytest.mark.parametrize("circuit,fname", CASES)
def test_reference_fixtures_on_gpu(circuit, fname):
    from digilogic_b200 import SimulationRunResult, SimulatorBuilder
    order, expected = TRUTH[circuit]
    circ = FIXTURES[fname]
    n_lanes = 1 << len(order)
    sim = circ.emit(SimulatorBuilder()).build()
    sim.set_lanes(n_lanes)
    ones = lane_mask(n_lanes)
    for name, words in exhaustive_drives(order, n_lanes).items():
        sim.set_wire_drive_lanes(circ.inputs[name][0], words, ones)
    res, conflict, unsettled = sim.run_sim_lanes(10_000)
    assert res == SimulationRunResult.Ok
    for out, want in expected.items():
        v, m = sim.get_wire_lanes(circ.outputs[out][0])
        assert (m == ones).all() and truth_int(v, n_lanes) == want, (fname, out)
    # every wire against the scalar oracle as well
    lanes = oracle.Lanes(circ.emit(oracle.Builder()), n_lanes)
    for name, words in exhaustive_drives(order, n_lanes).items():
        lanes.set_drive(circ.inputs[name][0], words, ones)
    lanes.run(10_000)
    rv, rm = lanes.get_all(circ.n_nets)
    gv, gm = sim.gather_wires_host(np.arange(circ.n_nets, dtype=np.uint32))
    assert (((gv ^ rv) | (gm ^ rm)) & ones).max() == 0


@pytest.mark.gpu
def test_fixture_through_server_like_the_gui_client():
    """The GUI flow (client.rs:288-427,178-193): BeginBuild, AddNet*, Add*Gate*, EndBuild, then
    SetNetDrive per input + Eval{max_steps: 10_000} per click, state read back for every net."""
    from digilogic_b200.server import B200Server
    order, expected = TRUTH["medium"]
    circ = FIXTURES["medium.dig"]
    sv = B200Server()
    sv.client_connected(42)
    sv.begin_build(42)
    nets = [sv.add_net(42, 1) for _ in range(circ.n_nets)]
    adders = {importers.AND: sv.add_and_gate, importers.OR: sv.add_or_gate, importers.XOR: sv.add_xor_gate}
    for kind, ins, out in circ.gates:
        if kind == importers.NOT:
            sv.add_not_gate(42, 1, ins[0], out)
        else:
            adders[kind](42, 1, ins, out)
    sv.end_build(42)
    got = {k: 0 for k in expected}
    for row in range(1 << len(order)):
        for k, name in enumerate(order):
            bit = (row >> (len(order) - 1 - k)) & 1
            sv.set_net_drive(42, circ.inputs[name][0], bytes([bit]), bytes([1]))   # LogicState::from_bool
        sv.eval(42, 10_000)
        st = sv.sim_state(42, nets, total_bits_hint=len(nets))
        assert st.bit_len == len(nets)
        for out in expected:
            p0, p1 = st.get_net(circ.outputs[out][0], 1)
            assert p1 == b"\1"
            got[out] |= p0[0] << row
        if row % 16 == 0:
            per_net = sv.sim_state_per_net(42, nets)
            assert per_net.bit_plane_0 == st.bit_plane_0 and per_net.bit_plane_1 == st.bit_plane_1
    assert got == expected


@pytest.mark.gpu
def test_config1_adder_dig_exhaustive_on_gpu():
    from digilogic_b200 import SimulationRunResult, SimulatorBuilder
    circ = importers.load_dig(importers.adder_dig(4), "adder4")
    sim = circ.emit(SimulatorBuilder()).build()
    sim.set_lanes(512)
    vec = np.arange(512)
    ones = lane_mask(512)
    for i in range(4):
        sim.set_wire_drive_lanes(circ.inputs[f"A{i}"][0], pack_lanes((vec >> i) & 1), ones)
        sim.set_wire_drive_lanes(circ.inputs[f"B{i}"][0], pack_lanes((vec >> (4 + i)) & 1), ones)
    sim.set_wire_drive_lanes(circ.inputs["Cin"][0], pack_lanes(vec >> 8), ones)
    res, _, _ = sim.run_sim_lanes(10_000)
    assert res == SimulationRunResult.Ok
    total = np.zeros(512, np.int64)
    for i in range(4):
        total += unpack_lanes(sim.get_wire_lanes(circ.outputs[f"S{i}"][0])[0], 512).astype(np.int64) << i
    total += unpack_lanes(sim.get_wire_lanes(circ.outputs["Cout"][0])[0], 512).astype(np.int64) << 4
    assert (total == (vec & 15) + ((vec >> 4) & 15) + (vec >> 8)).all()


@pytest.mark.gpu
def test_config2_multiplier_yosys_1m_vectors_on_gpu():
    """BASELINE config 2 at full size: 2^20 random vectors, P = A*B on every lane."""
    from digilogic_b200 import SimulationRunResult, SimulatorBuilder
    n, n_lanes = 32, 1 << 20
    circ = importers.load_yosys(importers.multiplier_yosys(n))
    sim = circ.emit(SimulatorBuilder()).build()
    sim.set_lanes(n_lanes)
    a, b, drives = _multiplier_inputs(circ, n, n_lanes, 0xD1610001)
    wires = np.array(list(drives), np.uint32)
    value = np.stack([drives[int(w)] for w in wires])
    sim.set_drives_lanes(wires, value, np.full_like(value, 0xFFFFFFFFFFFFFFFF))
    res, conflict, unsettled = sim.run_sim_lanes(10_000)
    assert res == SimulationRunResult.Ok and not conflict.any() and not unsettled.any()
    pv, pm = sim.gather_wires_host(np.array(circ.outputs["P"], np.uint32))
    assert (pm == np.uint64(0xFFFFFFFFFFFFFFFF)).all()
    lo = np.zeros(n_lanes, np.uint64)
    hi = np.zeros(n_lanes, np.uint64)
    for j in range(32):
        lo |= unpack_lanes(pv[j], n_lanes).astype(np.uint64) << np.uint64(j)
        hi |= unpack_lanes(pv[32 + j], n_lanes).astype(np.uint64) << np.uint64(j)
    with np.errstate(over="ignore"):
        prod = a * b   # both < 2^32: exact in uint64
    assert (lo == (prod & np.uint64(0xFFFFFFFF))).all() and (hi == (prod >> np.uint64(32))).all()
