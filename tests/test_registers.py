"""Registers (gsim add_register): known answers on the CPU oracle, the
event-driven scalar restatement against the dense bit-sliced model on random netlists with registers, and the
engine against the oracle on the GPU."""
import numpy as np
import pytest

import oracle
from helpers import Netlist, bus_enable_drives, lane_mask, random_drive, random_mixed_netlist, random_register_netlist

Z, X, L0, L1 = (0, 0), (1, 0), (0, 1), (1, 1)
ALL = (1 << 64) - 1


def test_register_known_answers_scalar():
    b = oracle.Builder()
    d, q, en, ck = b.add_wire(4), b.add_wire(4), b.add_wire(), b.add_wire()
    b.add_register(d, q, en, ck)
    sim = b.build()
    wires = {"d": d, "en": en, "ck": ck}

    def step(**kw):
        for k, v in kw.items():
            sim.set_wire_drive(wires[k], *v)
        assert sim.run_sim(10) == oracle.RUN_OK
        return sim.get_wire_state(q)

    assert step(d=(0b1010, 0xF), en=L1, ck=L0) == (0xF, 0)        # Undefined until the first load
    assert step(ck=L1) == (0b1010, 0xF)                           # rising edge loads
    assert step(d=(0b0101, 0xF)) == (0b1010, 0xF)                 # no edge: holds
    assert step(ck=L0) == (0b1010, 0xF) and step(ck=L1) == (0b0101, 0xF)
    assert step(ck=L0, en=L0) == (0b0101, 0xF) and step(ck=L1, d=(0xF, 0xF)) == (0b0101, 0xF)   # enable 0 keeps
    assert step(ck=L0, en=X) == (0b0101, 0xF) and step(ck=L1) == (0xF, 0)                          # enable X: Undefined
    assert step(ck=L0, en=L1, d=(0b0011, 0b0111)) == (0xF, 0) and step(ck=L1) == (0b1011, 0b0111)  # High-Z bit reads X
    assert step(ck=X) == (0b1011, 0b0111) and step(ck=L1, d=(0, 0xF)) == (0b1011, 0b0111)           # X -> 1 is no edge
    assert step(ck=L0) == (0b1011, 0b0111) and step(ck=L1) == (0, 0xF)


def test_register_validation():
    b = oracle.Builder()
    d8, q8, q4, e1, e2, c1 = b.add_wire(8), b.add_wire(8), b.add_wire(4), b.add_wire(1), b.add_wire(2), b.add_wire(1)
    with pytest.raises(oracle.OracleError, match="InvalidWireId"):
        b.add_register(d8, 99, e1, c1)
    with pytest.raises(oracle.OracleError, match="WireWidthMismatch"):
        b.add_register(d8, q4, e1, c1)
    with pytest.raises(oracle.OracleError, match="WireWidthIncompatible"):
        b.add_register(d8, q8, e2, c1)
    with pytest.raises(oracle.OracleError, match="WireWidthIncompatible"):
        b.add_register(d8, q8, e1, e2)
    assert b.add_register(d8, q8, e1, c1) == 0


def lfsr_netlist(bits=8, taps=(7, 5, 4, 3)):
    """Fibonacci LFSR out of 1-bit registers: q[0] <- XOR of the taps, q[i] <- q[i-1]; a load path
    (mux by AND/OR) initialises it: q[i] <- init[i] while `load` is 1."""
    nl = Netlist()
    ck, en, load, nload = nl.add_wire(), nl.add_wire(), nl.add_wire(), nl.add_wire()
    init = [nl.add_wire() for _ in range(bits)]
    q = [nl.add_wire() for _ in range(bits)]
    nl.add_not_gate(load, nload)
    fb = q[taps[0]]
    for t in taps[1:]:
        x = nl.add_wire()
        nl.add_gate(2, [fb, q[t]], x)     # XOR
        fb = x
    for i in range(bits):
        nxt = fb if i == 0 else q[i - 1]
        a, c, dd = nl.add_wire(), nl.add_wire(), nl.add_wire()
        nl.add_gate(0, [nxt, nload], a)   # AND
        nl.add_gate(0, [init[i], load], c)
        nl.add_gate(1, [a, c], dd)        # OR
        nl.add_register(dd, q[i], en, ck)
    return nl, ck, en, load, init, q


def soft_lfsr(state, bits, taps, steps):
    for _ in range(steps):
        fb = 0
        for t in taps:
            fb ^= (state >> t) & 1
        state = ((state << 1) | fb) & ((1 << bits) - 1)
    return state


def test_lfsr_of_registers_on_both_cpu_models():
    bits, taps, n_lanes = 8, (7, 5, 4, 3), 128
    nl, ck, en, load, init, q = lfsr_netlist(bits, taps)
    b = nl.replay(oracle.Builder())
    lanes = oracle.Lanes(b, n_lanes)
    bs = oracle.BitSliced(b, n_lanes, nl.n_wires)
    words = lanes.words
    ones, zeros = np.full(words, ALL, np.uint64), np.zeros(words, np.uint64)
    rng = np.random.default_rng(3)
    pat = [rng.integers(0, 1 << 64, words, dtype=np.uint64) for _ in range(bits)]

    def drive(w, v):
        lanes.set_drive(w, v, ones)
        bs.set_drive(w, v, ones)

    def run():
        st, _, _ = lanes.run(1000)
        c, u = bs.run(1000)
        assert (st == 0).all() and not c.any() and not u.any()
        a, b_ = lanes.get_all(nl.n_wires), bs.get_all()
        assert (a[0] == b_[0]).all() and (a[1] == b_[1]).all()

    drive(en, ones); drive(load, ones); drive(ck, zeros)
    for i in range(bits):
        drive(init[i], pat[i])
    run()
    drive(ck, ones); run()          # load edge
    drive(load, zeros); drive(ck, zeros); run()
    clocks = 25
    for _ in range(clocks):
        drive(ck, ones); run()
        drive(ck, zeros); run()
    got = [lanes.get(q[i])[0] for i in range(bits)]
    for lane in range(n_lanes):
        start = sum(((int(pat[i][lane // 64]) >> (lane % 64)) & 1) << i for i in range(bits))
        have = sum(((int(got[i][lane // 64]) >> (lane % 64)) & 1) << i for i in range(bits))
        assert have == soft_lfsr(start, bits, taps, clocks), lane


@pytest.mark.parametrize("seed", range(8))
def test_scalar_vs_bitsliced_random_register_netlists(seed):
    """Status and every wire on every lane: event-driven restatement (explicit per-register state) against the
    dense model, over warm restarts and a max_steps sweep that ends runs in the middle of clock activity."""
    rng = np.random.default_rng(4200 + seed)
    n_in, n_gates, n_lanes = 6, 40, 128
    nl, first_in = random_register_netlist(rng, n_in, n_gates)
    b = nl.replay(oracle.Builder())
    lanes = oracle.Lanes(b, n_lanes)
    bs = oracle.BitSliced(b, n_lanes, nl.n_wires)
    for rnd, max_steps in enumerate([0, 1, 2, 3, 5, 8, 50, 7, 0, 2, 1000, 4, 1, 30]):
        for i in range(n_in):
            if rng.random() < 0.7:
                v, m = random_drive(rng, lanes.words, four_state=rnd % 4 == 0)
                lanes.set_drive(first_in + i, v, m)
                bs.set_drive(first_in + i, v, m)
        status, _, _ = lanes.run(max_steps)
        conflict, unsettled = bs.run(max_steps)
        assert (oracle.status_from_masks(conflict, unsettled, n_lanes) == status).all(), (rnd, max_steps)
        assert not (status == oracle.RUN_CONFLICT).any()
        sv, sm = lanes.get_all(nl.n_wires)
        dv, dm = bs.get_all()
        assert (sv == dv).all() and (sm == dm).all(), (rnd, max_steps)


def _mixed_rounds(rng, nl, first_in, groups, n_in, words, set_all, run_all):
    """Drives and runs shared by the CPU and GPU mixed tests; yields after every run."""
    for rnd, max_steps in enumerate([0, 1, 2, 3, 5, 8, 50, 7, 0, 2, 1000, 4]):
        for i in range(n_in):
            if rng.random() < 0.7:
                v, m = random_drive(rng, words, four_state=rnd % 4 == 0)
                set_all(first_in + i, v, m)
        for grp in groups:
            if rnd == 0 or rng.random() < 0.5:
                for e, (v, m) in zip(grp, bus_enable_drives(rng, grp, words)):
                    set_all(e, v, m)
        yield rnd, max_steps, run_all(max_steps)


@pytest.mark.parametrize("seed", range(6))
def test_scalar_vs_bitsliced_mixed_netlists(seed):
    """Gates, muxes, registers, buffers and multi-driver buses in one cyclic netlist: event-driven restatement against
    the dense model; lanes that have conflicted are out of contract from then on."""
    rng = np.random.default_rng(6100 + seed)
    n_in, n_gates, n_bus, n_lanes = 6, 40, 3, 128
    nl, first_in, groups = random_mixed_netlist(rng, n_in, n_gates, n_bus)
    b = nl.replay(oracle.Builder())
    lanes = oracle.Lanes(b, n_lanes)
    bs = oracle.BitSliced(b, n_lanes, nl.n_wires)
    clean = np.ones(n_lanes, bool)

    def set_all(w, v, m):
        lanes.set_drive(w, v, m)
        bs.set_drive(w, v, m)

    def run_all(ms):
        st, _, _ = lanes.run(ms)
        c, u = bs.run(ms)
        return st, oracle.status_from_masks(c, u, n_lanes)

    compared = 0
    for rnd, ms, (st, st2) in _mixed_rounds(rng, nl, first_in, groups, n_in, lanes.words, set_all, run_all):
        assert (st2[clean] == st[clean]).all(), (rnd, ms)
        clean &= st != oracle.RUN_CONFLICT
        keep = np.zeros(lanes.words, np.uint64)
        for l in np.nonzero(clean)[0]:
            keep[l // 64] |= np.uint64(1 << (l % 64))
        sv, sm = lanes.get_all(nl.n_wires)
        dv, dm = bs.get_all()
        assert (((sv ^ dv) | (sm ^ dm)) & keep[None, :] == 0).all(), (rnd, ms)
        compared += int(clean.sum())
    assert compared > 2 * n_lanes


# ---------------------------------------------------------------------------------------------- GPU
PATHS = [("resident", 1, 1, 1), ("hbm-device-loop", 0, 1, 1), ("hbm-host-loop", 0, 1, 0), ("hbm-dense", 0, 0, 0)]


def _engine(nl, n_lanes, resident, frontier, device_loop):
    from digilogic_b200 import SimulatorBuilder
    sim = nl.replay(SimulatorBuilder()).build()
    sim.set_option(sim.OPT_RESIDENT_KERNEL, resident)
    sim.set_option(sim.OPT_FRONTIER, frontier)
    sim.set_option(sim.OPT_DEVICE_LOOP, device_loop)
    sim.set_lanes(n_lanes)
    return sim


@pytest.mark.gpu
def test_engine_builder_validation_and_single_lane_register():
    from digilogic_b200 import AddComponentError, LogicState, SimulationRunResult, SimulatorBuilder
    from digilogic_b200.gsim import ComponentError
    b = SimulatorBuilder()
    d8, q8, q4, e1, e2, c1 = b.add_wire(8), b.add_wire(8), b.add_wire(4), b.add_wire(1), b.add_wire(2), b.add_wire(1)
    for fn, err in ((lambda: b.add_register(d8, 99, e1, c1), AddComponentError.InvalidWireId),
                    (lambda: b.add_register(d8, q4, e1, c1), AddComponentError.WireWidthMismatch),
                    (lambda: b.add_register(d8, q8, e2, c1), AddComponentError.WireWidthIncompatible),
                    (lambda: b.add_register(d8, q8, e1, e2), AddComponentError.WireWidthIncompatible)):
        with pytest.raises(ComponentError) as e:
            fn()
        assert e.value.error == err
    # the scalar oracle's known-answer script, call for call
    ob, gb = oracle.Builder(), SimulatorBuilder()
    for bb in (ob, gb):
        d, q, en, ck = bb.add_wire(4), bb.add_wire(4), bb.add_wire(), bb.add_wire()
        bb.add_register(d, q, en, ck)
    ref, sim = ob.build(), gb.build()
    assert sim.info.cyclic == 1
    script = [{d: (0b1010, 0xF), en: L1, ck: L0}, {ck: L1}, {d: (0b0101, 0xF)}, {ck: L0}, {ck: L1}, {ck: L0, en: L0},
              {ck: L1, d: (0xF, 0xF)}, {ck: L0, en: X}, {ck: L1}, {ck: L0, en: L1, d: (0b0011, 0b0111)}, {ck: L1}, {ck: X},
              {ck: L1, d: (0, 0xF)}, {ck: L0}, {ck: L1}, {ck: Z}, {ck: L0}, {ck: L1, en: Z}]
    for i, drives in enumerate(script):
        for w, st in drives.items():
            ref.set_wire_drive(w, *st)
            sim.set_wire_drive(w, LogicState(*st))
        for ms in (10, 0):
            assert int(sim.run_sim(ms)) == ref.run_sim(ms), i
            assert sim.get_wire_state(q) == LogicState(*ref.get_wire_state(q)), (i, drives)


@pytest.mark.gpu
@pytest.mark.parametrize("path,resident,frontier,device_loop", PATHS)
def test_lfsr_of_registers_on_gpu(path, resident, frontier, device_loop):
    from digilogic_b200 import SimulationRunResult
    bits, taps, n_lanes = 8, (7, 5, 4, 3), 64 * 5 + 9
    nl, ck, en, load, init, q = lfsr_netlist(bits, taps)
    sim = _engine(nl, n_lanes, resident, frontier, device_loop)
    words = (n_lanes + 63) // 64
    ones, zeros = np.full(words, ALL, np.uint64), np.zeros(words, np.uint64)
    rng = np.random.default_rng(3)
    pat = [rng.integers(0, 1 << 64, words, dtype=np.uint64) for _ in range(bits)]

    def run():
        res, c, u = sim.run_sim_lanes(1000)
        assert res == SimulationRunResult.Ok

    sim.set_wire_drive_lanes(en, ones, ones); sim.set_wire_drive_lanes(load, ones, ones); sim.set_wire_drive_lanes(ck, zeros, ones)
    for i in range(bits):
        sim.set_wire_drive_lanes(init[i], pat[i], ones)
    run()
    sim.set_wire_drive_lanes(ck, ones, ones); run()
    sim.set_wire_drive_lanes(load, zeros, ones); sim.set_wire_drive_lanes(ck, zeros, ones); run()
    clocks = 30
    for _ in range(clocks):
        sim.set_wire_drive_lanes(ck, ones, ones); run()
        sim.set_wire_drive_lanes(ck, zeros, ones); run()
    got = [sim.get_wire_lanes(q[i])[0] for i in range(bits)]
    for lane in range(n_lanes):
        start = sum(((int(pat[i][lane // 64]) >> (lane % 64)) & 1) << i for i in range(bits))
        have = sum(((int(got[i][lane // 64]) >> (lane % 64)) & 1) << i for i in range(bits))
        assert have == soft_lfsr(start, bits, taps, clocks), lane
    assert sim.info.last_mode == (3 if resident else 2)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(5))
@pytest.mark.parametrize("n_lanes", [100, 64 * 9])
@pytest.mark.parametrize("path,resident,frontier,device_loop", PATHS)
def test_random_register_netlists_on_gpu(seed, n_lanes, path, resident, frontier, device_loop):
    """Per-lane run result and every wire against the dense CPU model (itself checked against the event-driven
    restatement above), warm restarts, max_steps sweep."""
    rng = np.random.default_rng(5200 + seed)
    n_in, n_gates = 6, 48
    nl, first_in = random_register_netlist(rng, n_in, n_gates)
    sim = _engine(nl, n_lanes, resident, frontier, device_loop)
    bs = oracle.BitSliced(nl.replay(oracle.Builder()), n_lanes, nl.n_wires)
    m = lane_mask(n_lanes)
    for rnd, max_steps in enumerate([0, 1, 2, 3, 5, 8, 50, 7, 0, 2, 1000, 4, 1, 30]):
        for i in range(n_in):
            if rng.random() < 0.7:
                v, vm = random_drive(rng, bs.words, four_state=rnd % 4 == 0)
                sim.set_wire_drive_lanes(first_in + i, v, vm)
                bs.set_drive(first_in + i, v, vm)
        res, conflict, unsettled = sim.run_sim_lanes(max_steps)
        c, u = bs.run(max_steps)
        assert ((conflict ^ c) & m).max() == 0 and ((unsettled ^ u) & m).max() == 0, (rnd, max_steps)
        rv, rm = bs.get_all()
        gv, gm = sim.gather_wires_host(np.arange(nl.n_wires, dtype=np.uint32))
        assert ((gv ^ rv) & m).max() == 0 and ((gm ^ rm) & m).max() == 0, (rnd, max_steps)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(4))
@pytest.mark.parametrize("path,resident,frontier,device_loop", PATHS)
def test_mixed_netlists_on_gpu(seed, path, resident, frontier, device_loop):
    """The same mix on the engine against the dense CPU model."""
    rng = np.random.default_rng(6300 + seed)
    n_in, n_gates, n_bus, n_lanes = 6, 48, 3, 64 * 7 + 5
    nl, first_in, groups = random_mixed_netlist(rng, n_in, n_gates, n_bus)
    sim = _engine(nl, n_lanes, resident, frontier, device_loop)
    bs = oracle.BitSliced(nl.replay(oracle.Builder()), n_lanes, nl.n_wires)
    m = lane_mask(n_lanes)
    dirty = np.zeros(bs.words, np.uint64)

    def set_all(w, v, vm):
        sim.set_wire_drive_lanes(w, v, vm)
        bs.set_drive(w, v, vm)

    def run_all(ms):
        res, conflict, unsettled = sim.run_sim_lanes(ms)
        c, u = bs.run(ms)
        return conflict, unsettled, c, u

    for rnd, ms, (conflict, unsettled, c, u) in _mixed_rounds(rng, nl, first_in, groups, n_in, bs.words, set_all, run_all):
        keep = m & ~dirty
        assert ((conflict ^ c) & keep).max() == 0 and ((unsettled ^ u) & keep).max() == 0, (rnd, ms)
        dirty |= c
        keep = m & ~dirty
        rv, rm = bs.get_all()
        gv, gm = sim.gather_wires_host(np.arange(nl.n_wires, dtype=np.uint32))
        assert ((gv ^ rv) & keep).max() == 0 and ((gm ^ rm) & keep).max() == 0, (rnd, ms)
    assert (m & ~dirty).any()
