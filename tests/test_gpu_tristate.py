"""GPU parity for tri-state buffers and nets with several drivers:
the CUDA engine through the C ABI against the CPU oracle — run result, every wire, every lane —
on the resident (shared-memory) kernel and on the HBM unit-delay path with and without the
changed-row frontier."""
import numpy as np
import pytest

import oracle
from digilogic_b200 import AddComponentError, LogicState, SimulationRunResult, SimulatorBuilder
from digilogic_b200.gsim import ComponentError
from helpers import Netlist, bus_enable_drives, lane_mask, random_bus_netlist, random_drive

pytestmark = pytest.mark.gpu

Z, X, L0, L1 = (0, 0), (1, 0), (0, 1), (1, 1)
PATHS = [("resident", 1, 1, 1), ("hbm-device-loop", 0, 1, 1), ("hbm-host-loop", 0, 1, 0), ("hbm-dense", 0, 0, 0)]


def engine(nl, n_lanes, resident, frontier, device_loop=1):
    sim = nl.replay(SimulatorBuilder()).build()
    sim.set_option(sim.OPT_RESIDENT_KERNEL, resident)
    sim.set_option(sim.OPT_FRONTIER, frontier)
    sim.set_option(sim.OPT_DEVICE_LOOP, device_loop)
    sim.set_lanes(n_lanes)
    return sim


@pytest.mark.parametrize("path,resident,frontier,device_loop", PATHS)
def test_buffer_truth_table(path, resident, frontier, device_loop):
    nl = Netlist()
    d, e, o = nl.add_wire(), nl.add_wire(), nl.add_wire()
    nl.add_buffer(d, e, o)
    states = [Z, X, L0, L1]
    pack = lambda f, plane: np.array([sum(states[f(i)][plane] << i for i in range(16))], np.uint64)
    sim = engine(nl, 16, resident, frontier, device_loop)
    bs = oracle.BitSliced(nl.replay(oracle.Builder()), 16, 3)
    for wire, f in ((d, lambda i: i // 4), (e, lambda i: i % 4)):
        sim.set_wire_drive_lanes(wire, pack(f, 0), pack(f, 1))
        bs.set_drive(wire, pack(f, 0), pack(f, 1))
    res, conflict, unsettled = sim.run_sim_lanes(10)
    c, u = bs.run(10)
    assert res == SimulationRunResult.Ok and not conflict.any() and not c.any()
    rv, rm = bs.get_all()
    gv, gm = sim.gather_wires_host(np.arange(3, dtype=np.uint32))
    m = lane_mask(16)
    assert ((gv ^ rv) & m).max() == 0 and ((gm ^ rm) & m).max() == 0
    assert sim.info.cyclic == 1  # tri-state netlists always take the unit-delay path


def test_buffer_validation_and_wide_buffer():
    b = SimulatorBuilder()
    d8, o8, e1, e2, o4 = b.add_wire(8), b.add_wire(8), b.add_wire(1), b.add_wire(2), b.add_wire(4)
    for fn, err in ((lambda: b.add_buffer(d8, e1, 99), AddComponentError.InvalidWireId),
                    (lambda: b.add_buffer(d8, e1, o4), AddComponentError.WireWidthMismatch),
                    (lambda: b.add_buffer(d8, e2, o8), AddComponentError.WireWidthIncompatible)):
        with pytest.raises(ComponentError) as ei:
            fn()
        assert ei.value.error == err
    b.add_buffer(d8, e1, o8)
    sim = b.build()
    sim.set_wire_drive(d8, LogicState(0xA5, 0xF7))  # bit 3 High-Z
    sim.set_wire_drive(e1, LogicState(1, 1))
    assert sim.run_sim(10) == SimulationRunResult.Ok
    assert sim.get_wire_state(o8) == LogicState(0xA5 | 0x08, 0xF7)
    sim.set_wire_drive(e1, LogicState(0, 1))
    assert sim.run_sim(10) == SimulationRunResult.Ok
    assert sim.get_wire_state(o8) == LogicState(0, 0)


@pytest.mark.parametrize("path,resident,frontier,device_loop", PATHS)
def test_shared_bus_single_lane(path, resident, frontier, device_loop):
    """The scalar oracle's bus scenario (tests/test_tristate.py) replayed call for call."""
    def make(b):
        d0, d1, e0, e1, bus, nb = (b.add_wire() for _ in range(6))
        b.add_buffer(d0, e0, bus)
        b.add_buffer(d1, e1, bus)
        b.add_not_gate(bus, nb)
        return d0, d1, e0, e1, bus, nb

    ob = oracle.Builder()
    w = make(ob)
    ref = ob.build()
    gb = SimulatorBuilder()
    make(gb)
    sim = gb.build()
    sim.set_option(sim.OPT_RESIDENT_KERNEL, resident)
    sim.set_option(sim.OPT_FRONTIER, frontier)
    sim.set_option(sim.OPT_DEVICE_LOOP, device_loop)
    d0, d1, e0, e1, bus, nb = w
    script = [
        [(d0, L1), (d1, L0), (e0, L1), (e1, L0)],
        [(e0, L0), (e1, L1)],
        [(e1, L0)],
        [(bus, L1)],
        [(bus, Z)],
        [(bus, L0), (d0, Z)],
        [(e0, L1)],          # buffer turns on against the external drive: conflict
    ]
    results = []
    for drives in script:
        for wire, st in drives:
            ref.set_wire_drive(wire, *st)
            sim.set_wire_drive(wire, LogicState(*st))
        r_ref, r_gpu = ref.run_sim(100), int(sim.run_sim(100))
        assert r_gpu == r_ref, drives
        results.append(r_ref)
        if r_ref == oracle.RUN_OK:
            for wire in w:
                assert sim.get_wire_state(wire) == LogicState(*ref.get_wire_state(wire)), (drives, wire)
    assert results == [0, 0, 0, 0, 0, 0, 2]


@pytest.mark.parametrize("path,resident,frontier,device_loop", PATHS)
def test_transient_conflict_lanes(path, resident, frontier, device_loop):
    nl = Netlist()
    d0, d1, en, nen, bus = (nl.add_wire() for _ in range(5))
    nl.add_not_gate(en, nen)
    nl.add_buffer(d0, en, bus)
    nl.add_buffer(d1, nen, bus)
    n_lanes = 200
    sim = engine(nl, n_lanes, resident, frontier, device_loop)
    bs = oracle.BitSliced(nl.replay(oracle.Builder()), n_lanes, 5)
    words = bs.words
    ones, zeros = np.full(words, (1 << 64) - 1, np.uint64), np.zeros(words, np.uint64)
    for s in (sim.set_wire_drive_lanes, bs.set_drive):
        s(d0, ones, ones); s(d1, zeros, ones); s(en, zeros, ones)
    res, conflict, unsettled = sim.run_sim_lanes(100)
    c, u = bs.run(100)
    assert res == SimulationRunResult.Ok and not c.any()
    rising = np.random.default_rng(5).integers(0, 1 << 64, words, dtype=np.uint64)
    for s in (sim.set_wire_drive_lanes, bs.set_drive):
        s(en, rising, ones)
    res, conflict, unsettled = sim.run_sim_lanes(100)
    c, u = bs.run(100)
    m = lane_mask(n_lanes)
    assert res == SimulationRunResult.Err
    assert ((conflict ^ c) & m).max() == 0 and ((conflict ^ rising) & m).max() == 0
    assert not (unsettled & m).any() and not u.any()


@pytest.mark.parametrize("seed", range(5))
@pytest.mark.parametrize("cyclic", [False, True])
@pytest.mark.parametrize("n_lanes", [100, 64 * 9])
@pytest.mark.parametrize("path,resident,frontier,device_loop", PATHS)
def test_random_bus_netlists_vs_bitsliced(seed, cyclic, n_lanes, path, resident, frontier, device_loop):
    """Random netlists with multi-driver buses: per-lane run result and every wire against the dense CPU
    model (itself checked against the event-driven restatement in tests/test_tristate.py), over warm
    restarts and a max_steps sweep.  Lanes that have conflicted are out of contract from then on."""
    rng = np.random.default_rng(9000 + 10 * seed + cyclic)
    n_in, n_gates, n_bus = 6, 40, 5
    nl, first_in, groups = random_bus_netlist(rng, n_in, n_gates, n_bus, cyclic=cyclic)
    sim = engine(nl, n_lanes, resident, frontier, device_loop)
    bs = oracle.BitSliced(nl.replay(oracle.Builder()), n_lanes, nl.n_wires)
    words = bs.words
    m = lane_mask(n_lanes)
    dirty = np.zeros(words, np.uint64)
    modes, seen = set(), set()
    for rnd, max_steps in enumerate([0, 1, 2, 3, 4, 5, 8, 60, 7, 0, 2, 1000]):
        for i in range(n_in):
            if rng.random() < 0.7:
                v, vm = random_drive(rng, words, four_state=rnd % 3 == 0)
                sim.set_wire_drive_lanes(first_in + i, v, vm)
                bs.set_drive(first_in + i, v, vm)
        for grp in groups:
            if rnd == 0 or rng.random() < 0.5:
                for e, (v, vm) in zip(grp, bus_enable_drives(rng, grp, words)):
                    sim.set_wire_drive_lanes(e, v, vm)
                    bs.set_drive(e, v, vm)
        if rnd == 6:  # an external drive straight onto a bus, High-Z on most lanes
            bus0 = n_in + n_gates
            v, _ = random_drive(rng, words)
            vm = v & rng.integers(0, 1 << 64, words, dtype=np.uint64) & rng.integers(0, 1 << 64, words, dtype=np.uint64)
            sim.set_wire_drive_lanes(bus0, v & vm, vm)
            bs.set_drive(bus0, v & vm, vm)
        res, conflict, unsettled = sim.run_sim_lanes(max_steps)
        c, u = bs.run(max_steps)
        keep = m & ~dirty
        assert ((conflict ^ c) & keep).max() == 0, (rnd, max_steps, "conflict lanes")
        assert ((unsettled ^ u) & keep).max() == 0, (rnd, max_steps, "unsettled lanes")
        dirty |= c
        keep = m & ~dirty
        rv, rm = bs.get_all()
        gv, gm = sim.gather_wires_host(np.arange(nl.n_wires, dtype=np.uint32))
        assert ((gv ^ rv) & keep).max() == 0 and ((gm ^ rm) & keep).max() == 0, (rnd, max_steps)
        modes.add(sim.info.last_mode)
        seen.add("c" if (c & m).any() else "-")
        seen.add("u" if (u & m).any() else "-")
    assert modes == ({3} if resident else {2}), modes
    assert (m & ~dirty).any(), "every lane conflicted: the comparison is vacuous"


@pytest.mark.parametrize("path,resident,frontier,device_loop", PATHS)
def test_conflict_predicate_switch_follows_the_oracle(path, resident, frontier, device_loop):
    """SURVEY.md A.6-2 is one switch in the oracle (`conflict_needs_disagree`) and one option in the engine
    (B2_OPT_CONFLICT_NEEDS_DISAGREE): under the alternative reading two valid drivers that agree do not conflict."""
    nl = Netlist()
    d0, d1, e0, e1, bus, a, c, o = (nl.add_wire() for _ in range(8))
    nl.add_buffer(d0, e0, bus)
    nl.add_buffer(d1, e1, bus)
    nl.add_gate(0, [a, c], o)                      # AND whose output also gets an external drive
    n_lanes = 256
    words = n_lanes // 64
    rng = np.random.default_rng(17)
    ones = np.full(words, (1 << 64) - 1, np.uint64)
    try:
        for lenient in (1, 0):
            oracle.lib().go_set_config(0, lenient, 0, 0)
            sim = engine(nl, n_lanes, resident, frontier, device_loop)
            sim.set_option(sim.OPT_CONFLICT_NEEDS_DISAGREE, lenient)
            lanes = oracle.Lanes(nl.replay(oracle.Builder()), n_lanes)
            for rnd in range(4):
                for w in (d0, d1, e0, e1, a, c, o):
                    v = rng.integers(0, 1 << 64, words, dtype=np.uint64)
                    m = ones if w != o else rng.integers(0, 1 << 64, words, dtype=np.uint64)   # o: High-Z on about half the lanes
                    sim.set_wire_drive_lanes(w, v & m, m)
                    lanes.set_drive(w, v & m, m)
                status, _, _ = lanes.run(100)
                res, conflict, unsettled = sim.run_sim_lanes(100)
                got = oracle.status_from_masks(conflict, unsettled, n_lanes)
                assert (got == status).all(), (lenient, rnd)
                assert (status == oracle.RUN_CONFLICT).any() and (status == oracle.RUN_OK).any()
                break  # one run per simulator: lanes that conflicted are out of contract afterwards
    finally:
        oracle.lib().go_set_config(0, 0, 0, 0)
    # a netlist with two plain gates on one net keeps the strict reading
    b = SimulatorBuilder()
    x, y, z = b.add_wire(), b.add_wire(), b.add_wire()
    b.add_and_gate([x, y], z)
    b.add_or_gate([x, y], z)
    s = b.build()
    with pytest.raises(Exception):
        s.set_option(s.OPT_CONFLICT_NEEDS_DISAGREE, 1)
