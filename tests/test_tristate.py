"""Tri-state buffers and nets with several drivers on the CPU oracle:
known answers, transient conflicts, and the event-driven scalar restatement against the dense
bit-sliced model on random bus netlists."""
import numpy as np
import pytest

import oracle
from helpers import bus_enable_drives, random_bus_netlist, random_drive

Z, X, L0, L1 = (0, 0), (1, 0), (0, 1), (1, 1)


def _buffer_expected(data, enable):
    if enable in (Z, X):
        return X
    if enable == L0:
        return Z
    return X if data == Z else data  # High-Z on the input of an enabled buffer reads X (A.6-6)


def test_buffer_truth_table_scalar_and_bitsliced():
    b = oracle.Builder()
    d, e, o = b.add_wire(), b.add_wire(), b.add_wire()
    b.add_buffer(d, e, o)
    states = [Z, X, L0, L1]
    bs = oracle.BitSliced(b, 16, 3)
    dv = sum(states[i // 4][0] << i for i in range(16)); dm = sum(states[i // 4][1] << i for i in range(16))
    ev = sum(states[i % 4][0] << i for i in range(16)); em = sum(states[i % 4][1] << i for i in range(16))
    bs.set_drive(d, np.array([dv], np.uint64), np.array([dm], np.uint64))
    bs.set_drive(e, np.array([ev], np.uint64), np.array([em], np.uint64))
    c, u = bs.run(10)
    assert not c.any() and not u.any()
    v, m = bs.get(o)
    for i in range(16):
        exp = _buffer_expected(states[i // 4], states[i % 4])
        sim = b.build()
        sim.set_wire_drive(d, *states[i // 4]); sim.set_wire_drive(e, *states[i % 4])
        assert sim.run_sim(10) == oracle.RUN_OK
        assert sim.get_wire_state(o) == exp, (i, exp)
        assert ((int(v[0]) >> i) & 1, (int(m[0]) >> i) & 1) == exp, i


def test_buffer_validation():
    b = oracle.Builder()
    d8, o8, e1, e2, o4 = b.add_wire(8), b.add_wire(8), b.add_wire(1), b.add_wire(2), b.add_wire(4)
    with pytest.raises(oracle.OracleError, match="InvalidWireId"):
        b.add_buffer(d8, e1, 99)
    with pytest.raises(oracle.OracleError, match="WireWidthMismatch"):
        b.add_buffer(d8, e1, o4)
    with pytest.raises(oracle.OracleError, match="WireWidthIncompatible"):
        b.add_buffer(d8, e2, o8)
    b.add_buffer(d8, e1, o8)
    sim = b.build()
    sim.set_wire_drive(d8, 0xA5, 0xF7)  # bit 3 High-Z
    sim.set_wire_drive(e1, 1, 1)
    assert sim.run_sim(10) == oracle.RUN_OK
    assert sim.get_wire_state(o8) == (0xA5 | 0x08, 0xF7)
    sim.set_wire_drive(e1, 0, 1)
    assert sim.run_sim(10) == oracle.RUN_OK
    assert sim.get_wire_state(o8) == (0, 0)


def _bus(builder):
    """Two buffers on one bus; a third wire reads the bus through a NOT."""
    b = builder
    d0, d1, e0, e1, bus, nb = (b.add_wire() for _ in range(6))
    b.add_buffer(d0, e0, bus)
    b.add_buffer(d1, e1, bus)
    b.add_not_gate(bus, nb)
    return d0, d1, e0, e1, bus, nb


def test_shared_bus_scalar():
    b = oracle.Builder()
    d0, d1, e0, e1, bus, nb = _bus(b)
    sim = b.build()
    for w, s in ((d0, L1), (d1, L0), (e0, L1), (e1, L0)):
        sim.set_wire_drive(w, *s)
    assert sim.run_sim(100) == oracle.RUN_OK
    assert sim.get_wire_state(bus) == L1 and sim.get_wire_state(nb) == L0
    sim.set_wire_drive(e0, *L0); sim.set_wire_drive(e1, *L1)  # hand-over in one call: no overlap
    assert sim.run_sim(100) == oracle.RUN_OK
    assert sim.get_wire_state(bus) == L0 and sim.get_wire_state(nb) == L1
    sim.set_wire_drive(e1, *L0)  # nobody drives: the bus floats, its reader sees X
    assert sim.run_sim(100) == oracle.RUN_OK
    assert sim.get_wire_state(bus) == Z and sim.get_wire_state(nb) == X
    sim.set_wire_drive(bus, *L1)  # an external drive on a floating bus is fine
    assert sim.run_sim(100) == oracle.RUN_OK
    assert sim.get_wire_state(bus) == L1 and sim.get_wire_state(nb) == L0
    sim.set_wire_drive(e0, *L1)  # ... until a buffer turns on as well (even with an agreeing value)
    assert sim.run_sim(100) == oracle.RUN_CONFLICT


def test_transient_conflict_through_inverter():
    """Enables `en` and `NOT en`: the fixed point never has both buffers on, but the inverter's unit
    delay overlaps them for one step when `en` rises, so the run reports the conflict — the reason a
    netlist with several drivers on a net is never evaluated in one levelised pass."""
    b = oracle.Builder()
    d0, d1, en, nen, bus = (b.add_wire() for _ in range(5))
    b.add_not_gate(en, nen)
    b.add_buffer(d0, en, bus)
    b.add_buffer(d1, nen, bus)
    sim = b.build()
    for w, s in ((d0, L1), (d1, L0), (en, L0)):
        sim.set_wire_drive(w, *s)
    assert sim.run_sim(100) == oracle.RUN_OK
    assert sim.get_wire_state(bus) == L0
    sim.set_wire_drive(en, *L1)
    assert sim.run_sim(100) == oracle.RUN_CONFLICT
    bs = oracle.BitSliced(b, 64, 5)
    ones, zeros = np.array([(1 << 64) - 1], np.uint64), np.zeros(1, np.uint64)
    bs.set_drive(d0, ones, ones); bs.set_drive(d1, zeros, ones); bs.set_drive(en, zeros, ones)
    c, u = bs.run(100)
    assert not c.any() and not u.any()
    bs.set_drive(en, np.array([0xFF], np.uint64), ones)  # en rises on lanes 0..7 only
    c, u = bs.run(100)
    assert int(c[0]) == 0xFF and not u.any()


@pytest.mark.parametrize("seed", range(8))
@pytest.mark.parametrize("cyclic", [False, True])
def test_scalar_vs_bitsliced_random_buses(seed, cyclic):
    """Status on every lane and state on every wire agree between the two CPU models on netlists with
    multi-driver buses, over warm restarts and a max_steps sweep.  A lane that has reported a conflict
    is dropped from later comparisons: its state is outside the contract."""
    rng = np.random.default_rng(7000 + 10 * seed + cyclic)
    n_in, n_gates, n_bus, n_lanes = 6, 30, 4, 128
    nl, first_in, groups = random_bus_netlist(rng, n_in, n_gates, n_bus, cyclic=cyclic)
    b = nl.replay(oracle.Builder())
    n_w = nl.n_wires
    lanes = oracle.Lanes(b, n_lanes)
    bs = oracle.BitSliced(b, n_lanes, n_w)
    clean = np.ones(n_lanes, bool)
    compared = 0
    for rnd, max_steps in enumerate([0, 1, 2, 3, 4, 5, 8, 50, 7, 0, 2, 1000]):
        for i in range(n_in):
            if rng.random() < 0.7:
                v, m = random_drive(rng, lanes.words, four_state=rnd % 3 == 0)
                lanes.set_drive(first_in + i, v, m)
                bs.set_drive(first_in + i, v, m)
        for grp in groups:
            if rnd == 0 or rng.random() < 0.5:
                for e, (v, m) in zip(grp, bus_enable_drives(rng, grp, lanes.words)):
                    lanes.set_drive(e, v, m)
                    bs.set_drive(e, v, m)
        status, steps, _ = lanes.run(max_steps)
        conflict, unsettled = bs.run(max_steps)
        st2 = oracle.status_from_masks(conflict, unsettled, n_lanes)
        assert (st2[clean] == status[clean]).all(), (rnd, max_steps)
        clean &= status != oracle.RUN_CONFLICT
        sv, sm = lanes.get_all(n_w)
        dv, dm = bs.get_all()
        keep = np.zeros(lanes.words, np.uint64)
        for l in np.nonzero(clean)[0]:
            keep[l // 64] |= np.uint64(1 << (l % 64))
        assert (((sv ^ dv) | (sm ^ dm)) & keep[None, :] == 0).all(), (rnd, max_steps)
        compared += int(clean.sum())
    assert compared > 4 * n_lanes


def test_driver_swap_keeps_the_bus_value():
    """Two buffers with the same data hand the bus over in one step: the component outputs change while
    the wire does not.  gsim stops on "no wire changed" *or* "no component output changed"; sweep
    max_steps across that point — status and state of the dense model follow the event-driven one."""
    ones = np.array([(1 << 64) - 1], np.uint64)
    half = np.array([0xFFFFFFFF], np.uint64)
    for max_steps in range(0, 4):
        b = oracle.Builder()
        data, ea, eb, bus, nb = (b.add_wire() for _ in range(5))
        b.add_buffer(data, ea, bus)
        b.add_buffer(data, eb, bus)
        b.add_not_gate(bus, nb)
        lanes = oracle.Lanes(b, 64)
        bs = oracle.BitSliced(b, 64, 5)
        for sim in (lanes, bs):
            sim.set_drive(data, ones, ones); sim.set_drive(ea, ones, ones); sim.set_drive(eb, 0 * ones, ones)
        st, _, _ = lanes.run(1000)
        c, u = bs.run(1000)
        assert (st == oracle.RUN_OK).all() and not c.any() and not u.any()
        for sim in (lanes, bs):  # lanes 0..31 swap drivers, the others keep A on
            sim.set_drive(ea, ones ^ half, ones); sim.set_drive(eb, half, ones)
        st, steps, _ = lanes.run(max_steps)
        c, u = bs.run(max_steps)
        assert (st == oracle.RUN_OK).all() and (steps[:32] == 1).all() and (steps[32:] == 0).all()
        assert not c.any() and not u.any()
        sv, sm = lanes.get_all(5)
        dv, dm = bs.get_all()
        assert (sv == dv).all() and (sm == dm).all()
        assert int(dv[bus][0]) == (1 << 64) - 1 and int(dm[bus][0]) == (1 << 64) - 1


def test_gate_derived_enables_conflict_on_a_cold_start():
    """Every component output is High-Z until its first evaluation, so an enable that comes from a gate
    is High-Z in the first step, its buffer answers X, and two such buffers on one bus collide — under
    the restated semantics a shared bus needs enables that are inputs."""
    b = oracle.Builder()
    d0, d1, en, e0, e1, bus = (b.add_wire() for _ in range(6))
    b.add_not_gate(en, e0)
    b.add_not_gate(e0, e1)
    b.add_buffer(d0, e0, bus)
    b.add_buffer(d1, e1, bus)
    sim = b.build()
    for w in (d0, d1, en):
        sim.set_wire_drive(w, *L1)
    assert sim.run_sim(100) == oracle.RUN_CONFLICT
    bs = oracle.BitSliced(b, 64, 6)
    ones = np.array([(1 << 64) - 1], np.uint64)
    for w in (d0, d1, en):
        bs.set_drive(w, ones, ones)
    c, u = bs.run(100)
    assert int(c[0]) == (1 << 64) - 1
