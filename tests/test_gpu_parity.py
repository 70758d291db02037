"""GPU parity: the CUDA engine, through the C ABI, against the CPU oracle — bit-exact on every
wire, every lane, every run result, across warm restarts and max_steps sweeps."""
import itertools

import numpy as np
import pytest

import oracle
from digilogic_b200 import LogicState, SimulationRunResult, SimulatorBuilder
from helpers import (AND, NAND, NOR, NOT, OR, XNOR, XOR, Netlist, lane_mask, random_drive, random_netlist,
                     stimulus_words)

pytestmark = pytest.mark.gpu

Z, X, L0, L1 = (0, 0), (1, 0), (0, 1), (1, 1)


def both(nl: Netlist, n_lanes: int, resident=True):
    """resident: True = shared-memory kernel, False = HBM path with the device-side loop, "host" = HBM path with the
    host-driven loop."""
    ob = nl.replay(oracle.Builder())
    sim = nl.replay(SimulatorBuilder()).build()
    sim.set_option(sim.OPT_RESIDENT_KERNEL, int(resident is True))
    sim.set_option(sim.OPT_DEVICE_LOOP, int(resident != "host"))
    sim.set_lanes(n_lanes)
    return ob, sim


def gpu_status(conflict, unsettled, n_lanes):
    return oracle.status_from_masks(conflict, unsettled, n_lanes)


def compare_all_wires(sim, ref_value, ref_valid, n_wires, n_lanes, skip_lanes=None):
    gv, gm = sim.gather_wires_host(np.arange(n_wires, dtype=np.uint32))
    mask = lane_mask(n_lanes)
    if skip_lanes is not None:
        mask = mask & ~skip_lanes
    assert ((gv ^ ref_value) & mask).max(initial=0) == 0, "value plane differs"
    assert ((gm ^ ref_valid) & mask).max(initial=0) == 0, "valid plane differs"


@pytest.mark.parametrize("kind", [AND, OR, XOR, NAND, NOR, XNOR])
def test_truth_tables_all_16(kind):
    """All 16 input pairs of every 2-input gate, one pair per lane."""
    nl = Netlist()
    a, b, o = nl.add_wire(), nl.add_wire(), nl.add_wire()
    nl.add_gate(kind, [a, b], o)
    pairs = list(itertools.product([Z, X, L0, L1], repeat=2))
    ob, sim = both(nl, 16)
    pack = lambda idx, plane: np.array([sum(p[idx][plane] << i for i, p in enumerate(pairs))], np.uint64)
    lanes = oracle.Lanes(ob, 16)
    for wire, idx in ((a, 0), (b, 1)):
        sim.set_wire_drive_lanes(wire, pack(idx, 0), pack(idx, 1))
        lanes.set_drive(wire, pack(idx, 0), pack(idx, 1))
    res, conflict, unsettled = sim.run_sim_lanes(100)
    status, _, _ = lanes.run(100)
    assert res == SimulationRunResult.Ok and (status == 0).all()
    rv, rm = lanes.get_all(3)
    compare_all_wires(sim, rv, rm, 3, 16)


def test_single_lane_gsim_surface():
    """The plain gsim calls (one lane): NOT truth table and the A.7 latch trace."""
    for a, exp in [(Z, X), (X, X), (L0, L1), (L1, L0)]:
        b = SimulatorBuilder()
        i, o = b.add_wire(), b.add_wire()
        b.add_not_gate(i, o)
        s = b.build()
        s.set_wire_drive(i, LogicState(*a))
        assert s.run_sim(10) == SimulationRunResult.Ok
        assert s.get_wire_state(o) == LogicState(*exp)

    from test_oracle import A7
    b = SimulatorBuilder()
    s_, r_, q, qn = (b.add_wire() for _ in range(4))
    b.add_nand_gate([s_, qn], q)
    b.add_nand_gate([r_, q], qn)
    sim = b.build()
    assert sim.info.cyclic == 1
    for i, (sd, rd, ms, res, steps, eq, eqn) in enumerate(A7):
        sim.set_wire_drive(s_, LogicState(*sd))
        sim.set_wire_drive(r_, LogicState(*rd))
        assert int(sim.run_sim(ms)) == res, i
        assert sim.get_wire_state(q) == LogicState(*eq), i
        assert sim.get_wire_state(qn) == LogicState(*eqn), i


@pytest.mark.parametrize("seed", range(4))
@pytest.mark.parametrize("cyclic", [False, True])
@pytest.mark.parametrize("n_lanes", [1, 100, 64 * 37])
@pytest.mark.parametrize("resident", [True, False, "host"])
def test_random_netlists_vs_scalar_oracle(seed, cyclic, n_lanes, resident):
    """Random netlists (n-ary gates and muxes included), 4-state stimulus, warm restarts and a
    max_steps sweep that crosses the levelised/unit-delay switch."""
    rng = np.random.default_rng(77 * seed + cyclic)
    n_in, n_gates = 6, 48
    nl, first_in = random_netlist(rng, n_in, n_gates, cyclic=cyclic, wide=True, mux=True)
    ob, sim = both(nl, n_lanes, resident)
    lanes = oracle.Lanes(ob, n_lanes)
    words = lanes.words
    modes = set()
    for rnd, max_steps in enumerate([0, 1, 2, 3, 5, 8, 60, 7, 0, 1000]):
        for i in range(n_in):
            if rng.random() < 0.7:
                v, m = random_drive(rng, words)
                lanes.set_drive(first_in + i, v, m)
                sim.set_wire_drive_lanes(first_in + i, v, m)
        status, _, _ = lanes.run(max_steps)
        res, conflict, unsettled = sim.run_sim_lanes(max_steps)
        assert (gpu_status(conflict, unsettled, n_lanes) == status).all(), (rnd, max_steps)
        assert int(res) == int(status.max()), (rnd, max_steps)
        rv, rm = lanes.get_all(nl.n_wires)
        compare_all_wires(sim, rv, rm, nl.n_wires, n_lanes)
        modes.add(sim.info.last_mode)
    assert (3 in modes) == (resident is True) and (2 in modes) == (resident is not True), modes


@pytest.mark.parametrize("n_lanes,n_gates", [(64 * 64, 3000), (64 * 600, 500)])
def test_random_dag_vs_bitsliced_oracle(n_lanes, n_gates):
    """Bigger DAGs against the bit-sliced CPU model (levelised path, rows wider than one CTA)."""
    rng = np.random.default_rng(n_gates)
    n_in = 32
    nl, first_in = random_netlist(rng, n_in, n_gates, cyclic=False)
    ob, sim = both(nl, n_lanes)
    bs = oracle.BitSliced(ob, n_lanes, nl.n_wires)
    words = bs.words
    ins = np.arange(first_in, first_in + n_in, dtype=np.uint32)
    sim.fill_stimulus(ins, seed=0xD1610003, lane_word_base=5, mode=1)
    value, valid = stimulus_words(0xD1610003, n_in, words, word_base=5, mode=1)
    for i in range(n_in):
        bs.set_drive(first_in + i, value[i], valid[i])
    assert not bs.eval_dag().any()
    res, conflict, unsettled = sim.run_sim_lanes(10_000)
    assert res == SimulationRunResult.Ok and not conflict.any() and not unsettled.any()
    assert sim.info.last_mode == 1
    rv, rm = bs.get_all()
    compare_all_wires(sim, rv, rm, nl.n_wires, n_lanes)
    # the same netlist forced down the unit-delay path (max_steps below the depth) must agree
    depth = sim.info.depth
    sim2 = nl.replay(SimulatorBuilder()).build()
    sim2.set_lanes(n_lanes)
    sim2.fill_stimulus(ins, seed=0xD1610003, lane_word_base=5, mode=1)
    bs2 = oracle.BitSliced(ob, n_lanes, nl.n_wires)
    for i in range(n_in):
        bs2.set_drive(first_in + i, value[i], valid[i])
    for ms in (depth // 2, depth - 1):
        c_ref, u_ref = bs2.run(ms)
        res, conflict, unsettled = sim2.run_sim_lanes(ms)
        assert sim2.info.last_mode in (2, 3)
        assert (conflict == c_ref).all() and ((unsettled ^ u_ref) & lane_mask(n_lanes)).max() == 0
        rv, rm = bs2.get_all()
        compare_all_wires(sim2, rv, rm, nl.n_wires, n_lanes)


def test_conflict_lanes():
    """A driven gate output conflicts exactly in the lanes whose drive is not High-Z; both paths."""
    for cyclic, resident in ((False, True), (True, True), (True, False), (False, False)):
        nl = Netlist()
        a, c, o, q, qn = (nl.add_wire() for _ in range(5))
        nl.add_gate(AND, [a, c], o)
        if cyclic:
            nl.add_gate(NAND, [a, qn], q)
            nl.add_gate(NAND, [c, q], qn)
        ob, sim = both(nl, 64, resident)
        lanes = oracle.Lanes(ob, 64)
        ones = np.array([0xFFFFFFFFFFFFFFFF], np.uint64)
        dv, dm = np.array([0x00FF], np.uint64), np.array([0x0F0F], np.uint64)
        for s in (lanes,):
            s.set_drive(a, ones, ones); s.set_drive(c, ones, ones); s.set_drive(o, dv, dm)
        sim.set_wire_drive_lanes(a, ones, ones); sim.set_wire_drive_lanes(c, ones, ones); sim.set_wire_drive_lanes(o, dv, dm)
        for max_steps in (100, 0):
            status, _, _ = lanes.run(max_steps)
            res, conflict, unsettled = sim.run_sim_lanes(max_steps)
            assert int(conflict[0]) == 0x0FFF
            assert (gpu_status(conflict, unsettled, 64) == status).all()
            assert res == SimulationRunResult.Err
            rv, rm = lanes.get_all(nl.n_wires)
            compare_all_wires(sim, rv, rm, nl.n_wires, 64, skip_lanes=conflict)


def test_multi_driver_always_conflicts():
    b = SimulatorBuilder()
    a, c, o = b.add_wire(), b.add_wire(), b.add_wire()
    b.add_and_gate([a, c], o)
    b.add_or_gate([a, c], o)
    s = b.build()
    assert s.info.multi_driver == 1
    assert s.run_sim(100) == SimulationRunResult.Err


def test_wide_wires_and_mux():
    """Multi-bit wires are bit-blasted; results must equal the scalar oracle on the same calls."""
    def make(b):
        a, c, o = b.add_wire(40), b.add_wire(40), b.add_wire(40)
        b.add_xor_gate([a, c], o)
        sel, m = b.add_wire(2), b.add_wire(40)
        d = [b.add_wire(40) for _ in range(4)]
        b.add_multiplexer(d, sel, m)
        n5 = b.add_wire(40)
        b.add_nand_gate([a, c, d[0], d[1], d[2]], n5)
        return a, c, o, sel, m, d, n5
    ob, gb = oracle.Builder(), SimulatorBuilder()
    a, c, o, sel, m, d, n5 = make(ob)
    make(gb)
    osim, gsim = ob.build(), gb.build()
    full = (1 << 40) - 1
    script = [(a, 0x12345678AB, full), (c, 0x0F0F0F0F0F, 0xFFFFFF00FF)] + \
        [(w, 0x1111111111 * (i + 1), full) for i, w in enumerate(d)]
    for sel_state in [(2, 3), (2, 1), (1, 3), (3, 3), (0, 0)]:
        for w, p0, p1 in script + [(sel, *sel_state)]:
            osim.set_wire_drive(w, p0, p1)
            gsim.set_wire_drive(w, LogicState(p0, p1))
        assert int(gsim.run_sim(10)) == osim.run_sim(10)
        for w in (o, m, n5, a, sel):
            assert gsim.get_wire_state(w) == LogicState(*osim.get_wire_state(w)), (sel_state, w)
        script[5] = (d[1], 0, 0)  # later rounds: a High-Z data input


def test_adder_exhaustive_and_sim_state_pack():
    """4-bit ripple-carry adder, all 512 (A, B, Cin) vectors as lanes; S = A + B + Cin."""
    nl = Netlist()
    A = [nl.add_wire() for _ in range(4)]
    B = [nl.add_wire() for _ in range(4)]
    cin = nl.add_wire()
    S, carry = [], cin
    for i in range(4):
        x1, s, a1, a2, co = (nl.add_wire() for _ in range(5))
        nl.add_gate(XOR, [A[i], B[i]], x1)
        nl.add_gate(XOR, [x1, carry], s)
        nl.add_gate(AND, [A[i], B[i]], a1)
        nl.add_gate(AND, [x1, carry], a2)
        nl.add_gate(OR, [a1, a2], co)
        S.append(s)
        carry = co
    sim = nl.replay(SimulatorBuilder()).build()
    sim.set_lanes(512)
    lanes = np.arange(512, dtype=np.uint64)
    def bits(fn):
        b = np.array([fn(int(l)) for l in lanes], np.uint64)
        return np.array([sum(int(b[w * 64 + i]) << i for i in range(64)) for w in range(8)], np.uint64)
    ones = np.full(8, 0xFFFFFFFFFFFFFFFF, np.uint64)
    for i in range(4):
        sim.set_wire_drive_lanes(A[i], bits(lambda l: (l >> i) & 1), ones)
        sim.set_wire_drive_lanes(B[i], bits(lambda l: (l >> (4 + i)) & 1), ones)
    sim.set_wire_drive_lanes(cin, bits(lambda l: (l >> 8) & 1), ones)
    res, conflict, unsettled = sim.run_sim_lanes(10_000)
    assert res == SimulationRunResult.Ok
    outs, valid = sim.gather_wires_host(np.array(S + [carry], np.uint32))
    assert (valid == ones).all()
    for l in range(512):
        got = sum(((int(outs[i][l // 64]) >> (l % 64)) & 1) << i for i in range(5))
        assert got == (l & 15) + ((l >> 4) & 15) + (l >> 8)
    # bulk SimState pack of lane 0 == per-net get_wire_state pushed LSB-first
    bit_len, p0, p1 = sim.pack_sim_state(np.arange(nl.n_wires, dtype=np.uint32))
    assert bit_len == nl.n_wires and len(p0) == nl.n_wires // 8 + 1
    for w in range(nl.n_wires):
        st = sim.get_wire_state(w)
        assert (p0[w // 8] >> (w % 8)) & 1 == st.plane0 and (p1[w // 8] >> (w % 8)) & 1 == st.plane1


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_valid_plane_elision_is_bit_identical(mode):
    """B2_OPT_VALID_ELISION: rows that are provably valid keep no valid plane in HBM.  Stimulus
    mode 0 = all inputs valid (everything elided), 1 = 4-state on every input (nothing elided),
    2 = only a few inputs carry X/Z (mixed).  Every read-back path must return the same bits as
    the oracle, and a following unit-delay run must start from the same state."""
    rng = np.random.default_rng(31 + mode)
    n_in, n_gates, n_lanes = 24, 1500, 64 * 40
    nl, first_in = random_netlist(rng, n_in, n_gates, cyclic=False, wide=True, mux=True)
    ob, sim = both(nl, n_lanes)
    sim.set_option(sim.OPT_VALID_ELISION, 1)
    bs = oracle.BitSliced(ob, n_lanes, nl.n_wires)
    words = bs.words
    for i in range(n_in):
        four = mode == 1 or (mode == 2 and i % 7 == 0)
        v, m = random_drive(rng, words, four_state=four)
        bs.set_drive(first_in + i, v, m)
        sim.set_wire_drive_lanes(first_in + i, v, m)
    assert not bs.eval_dag().any()
    res, conflict, unsettled = sim.run_sim_lanes(10_000)
    assert res == SimulationRunResult.Ok and sim.info.last_mode == 1
    rv, rm = bs.get_all()
    compare_all_wires(sim, rv, rm, nl.n_wires, n_lanes)             # gather path (reads the elision flags)
    for w in (first_in, first_in + n_in, nl.n_wires - 1, nl.n_wires // 2):
        gv, gm = sim.get_wire_lanes(w)                                # copy path (expands elided planes first)
        assert (gv == rv[w]).all() and (gm == rm[w]).all()
    compare_all_wires(sim, rv, rm, nl.n_wires, n_lanes)
    # second elided run with new drives, then a unit-delay run from that state
    v, m = random_drive(rng, words, four_state=False)
    bs.set_drive(first_in + 1, v, m)
    sim.set_wire_drive_lanes(first_in + 1, v, m)
    bs.run(10_000)
    sim.run_sim_lanes(10_000)
    v, m = random_drive(rng, words, four_state=(mode != 0))
    bs.set_drive(first_in + 2, v, m)
    sim.set_wire_drive_lanes(first_in + 2, v, m)
    c_ref, u_ref = bs.run(2)
    res, conflict, unsettled = sim.run_sim_lanes(2)
    assert sim.info.last_mode in (2, 3)
    assert (conflict == c_ref).all() and ((unsettled ^ u_ref) & lane_mask(n_lanes)).max() == 0
    rv, rm = bs.get_all()
    compare_all_wires(sim, rv, rm, nl.n_wires, n_lanes)
