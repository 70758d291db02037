"""Edge cases through the C ABI on the GPU, each against the scalar oracle: empty and gate-less
netlists, ragged lane counts, self-loops, repeated inputs, step budgets 0 and u64::MAX, 255-bit
wires, re-allocating lanes, and argument errors that must come back as codes, not crashes."""
import ctypes as C

import numpy as np
import pytest

import oracle
from digilogic_b200 import B200Error, LogicState, SimulationRunResult, SimulatorBuilder, _ffi
from helpers import AND, NAND, NOT, OR, XOR, Netlist, lane_mask, random_drive

pytestmark = pytest.mark.gpu
Z, X, L0, L1 = (0, 0), (1, 0), (0, 1), (1, 1)
U64_MAX = (1 << 64) - 1


def test_empty_and_gateless_netlists():
    s = SimulatorBuilder().build()
    assert s.run_sim(10) == SimulationRunResult.Ok
    b = SimulatorBuilder()
    a, c = b.add_wire(), b.add_wire(3)
    s = b.build()
    assert s.get_wire_state(a) == LogicState.HIGH_Z          # never driven: High-Z (A.6-5)
    assert s.run_sim(0) == SimulationRunResult.Ok
    s.set_wire_drive(a, LogicState(1, 1))
    s.set_wire_drive(c, LogicState(0b101, 0b110))
    assert s.get_wire_state(a) == LogicState.HIGH_Z, "a drive is not visible before the next run_sim"
    assert s.run_sim(0) == SimulationRunResult.Ok
    assert s.get_wire_state(a) == LogicState(1, 1) and s.get_wire_state(c) == LogicState(0b101, 0b110)


@pytest.mark.parametrize("n_lanes", [1, 63, 65, 127, 129, 64 * 5 + 1])
@pytest.mark.parametrize("resident", [True, False, "host"])
def test_ragged_lane_counts_and_oscillator(n_lanes, resident):
    """A ring of one inverter oscillates in lanes where the enable is 1 and holds where it is 0;
    lanes beyond n_lanes in the last word must not leak into the result."""
    nl = Netlist()
    en, q, d = nl.add_wire(), nl.add_wire(), nl.add_wire()
    nl.add_gate(NAND, [en, q], d)      # d = ~(en & q)
    nl.add_gate(AND, [d, d], q)        # q = d (repeated input), closes the loop
    ob = nl.replay(oracle.Builder())
    sim = nl.replay(SimulatorBuilder()).build()
    sim.set_option(sim.OPT_RESIDENT_KERNEL, int(resident is True))
    sim.set_option(sim.OPT_DEVICE_LOOP, int(resident != "host"))
    sim.set_lanes(n_lanes)
    lanes = oracle.Lanes(ob, n_lanes)
    rng = np.random.default_rng(n_lanes)
    mask = lane_mask(n_lanes)
    for max_steps in (50, 0, 7, 8):
        v, m = random_drive(rng, lanes.words, four_state=False)
        lanes.set_drive(en, v, m)
        sim.set_wire_drive_lanes(en, v, m)
        status, _, _ = lanes.run(max_steps)
        res, conflict, unsettled = sim.run_sim_lanes(max_steps)
        assert (oracle.status_from_masks(conflict, unsettled, n_lanes) == status).all()
        assert ((conflict | unsettled) & ~mask).max(initial=0) == 0, "padding lanes reported"
        assert int(res) == int(status.max())
        rv, rm = lanes.get_all(3)
        gv, gm = sim.gather_wires_host(np.arange(3, dtype=np.uint32))
        assert (((gv ^ rv) | (gm ^ rm)) & mask).max(initial=0) == 0


def test_step_budget_extremes():
    def mk(builder):
        a, o1, o2, o3 = (builder.add_wire() for _ in range(4))
        builder.add_not_gate(a, o1)
        builder.add_not_gate(o1, o2)
        builder.add_not_gate(o2, o3)
        return builder.build(), a, o3
    osim, a, o3 = mk(oracle.Builder())
    gsim, _, _ = mk(SimulatorBuilder())
    for val, ms in [(1, 0), (0, 1), (1, 2), (0, U64_MAX), (1, U64_MAX - 1), (0, 3)]:
        osim.set_wire_drive(a, val, 1)
        gsim.set_wire_drive(a, LogicState(val, 1))
        assert int(gsim.run_sim(ms)) == osim.run_sim(ms), (val, ms)
        for w in range(4):
            assert gsim.get_wire_state(w) == LogicState(*osim.get_wire_state(w)), (val, ms, w)


def test_255_bit_wires():
    def mk(builder):
        a, b_, o, n = (builder.add_wire(255) for _ in range(4))
        builder.add_xor_gate([a, b_], o)
        builder.add_nor_gate([a, b_, o], n)
        return builder.build(), a, b_, o, n
    osim, a, b_, o, n = mk(oracle.Builder())
    gsim, *_ = mk(SimulatorBuilder())
    rng = np.random.default_rng(255)
    full = (1 << 255) - 1
    for _ in range(3):
        pa, ma, pb, mb = (int.from_bytes(rng.bytes(32), "little") & full for _ in range(4))
        osim.set_wire_drive(a, pa, ma); osim.set_wire_drive(b_, pb, mb)
        gsim.set_wire_drive(a, LogicState(pa, ma)); gsim.set_wire_drive(b_, LogicState(pb, mb))
        assert int(gsim.run_sim(10)) == osim.run_sim(10)
        for w in (a, b_, o, n):
            assert gsim.get_wire_state(w) == LogicState(*osim.get_wire_state(w))
    bit_len, p0, p1 = gsim.pack_sim_state([o, n])
    assert bit_len == 510 and len(p0) == 64
    so, sn = gsim.get_wire_state(o), gsim.get_wire_state(n)
    assert int.from_bytes(p0, "little") == so.plane0 | (sn.plane0 << 255)
    assert int.from_bytes(p1, "little") == so.plane1 | (sn.plane1 << 255)


def test_set_lanes_resets_to_cold_state():
    b = SimulatorBuilder()
    s_, r_, q, qn = (b.add_wire() for _ in range(4))
    b.add_nand_gate([s_, qn], q)
    b.add_nand_gate([r_, q], qn)
    sim = b.build()
    sim.set_wire_drive(s_, LogicState(0, 1)); sim.set_wire_drive(r_, LogicState(1, 1))
    assert sim.run_sim(100) == SimulationRunResult.Ok and sim.get_wire_state(q) == LogicState(1, 1)
    sim.set_lanes(128)
    assert sim.get_wire_state(q) == LogicState.HIGH_Z and sim.get_wire_state(s_) == LogicState.HIGH_Z
    ones = np.full(2, U64_MAX, np.uint64)
    sim.set_wire_drive_lanes(s_, ones, ones); sim.set_wire_drive_lanes(r_, ones, ones)
    res, conflict, unsettled = sim.run_sim_lanes(100)
    assert res == SimulationRunResult.Ok
    v, m = sim.get_wire_lanes(q)
    assert (v == ones).all() and (m == 0).all(), "cold latch with S'=R'=1 settles to X (A.7 row 1)"


def test_argument_errors_are_codes():
    L = _ffi.lib()
    b = SimulatorBuilder()
    a, o = b.add_wire(), b.add_wire()
    b.add_not_gate(a, o)
    sim = b.build()
    sim.set_lanes(64)
    one = np.array([1], np.uint64)
    assert L.b2_set_wire_drive_lanes(sim._h, 99, 0, 0, 1, one.ctypes.data, one.ctypes.data) == 2      # InvalidWireId
    assert L.b2_set_wire_drive_lanes(sim._h, a, 1, 0, 1, one.ctypes.data, one.ctypes.data) == 2       # bit >= width
    assert L.b2_set_wire_drive_lanes(sim._h, a, 0, 1, 1, one.ctypes.data, one.ctypes.data) == 21      # lane-word range
    assert L.b2_set_wire_drive_lanes(sim._h, a, 0, 0, 1, None, one.ctypes.data) == 16                 # NULL
    assert L.b2_get_wire_lanes(sim._h, o, 0, 0, 2, one.ctypes.data, one.ctypes.data) == 21
    assert L.b2_sim_set_lanes(sim._h, 0) == 17
    assert L.b2_sim_set_option(sim._h, 999, 1) == 17
    assert L.b2_run_sim(None, 1) == -16
    p = (C.c_uint32 * 1)()
    assert L.b2_get_wire_state(sim._h, 5, p, p, 1) == 2
    assert L.b2_get_wire_state(sim._h, o, p, p, 0) == 21
    with pytest.raises(B200Error):
        sim.fill_stimulus([a], 1, 0, mode=7)
    # the simulator still works afterwards
    sim.set_wire_drive(a, LogicState(1, 1))
    assert sim.run_sim(10) == SimulationRunResult.Ok and sim.get_wire_state(o) == LogicState(0, 1)


def test_many_clients_independent():
    from digilogic_b200.server import B200Server
    sv = B200Server()
    for cid in range(1, 9):
        sv.client_connected(cid)
        sv.begin_build(cid)
        a, b_, o = sv.add_net(cid, 1), sv.add_net(cid, 1), sv.add_net(cid, 1)
        (sv.add_and_gate if cid % 2 else sv.add_or_gate)(cid, 1, [a, b_], o)
        sv.end_build(cid)
    for cid in range(1, 9):
        sv.set_net_drive(cid, 0, b"\1", b"\1")
        sv.set_net_drive(cid, 1, b"\0", b"\1")
        sv.eval(cid, 10_000)
    for cid in range(1, 9):
        w, p0, p1 = sv.get_net_state(cid, 2)
        assert (w, p0[0] & 1, p1[0] & 1) == (1, 0 if cid % 2 else 1, 1)


@pytest.mark.parametrize("resident", [True, False])
def test_step_budget_switch_moves_with_the_oracle(resident):
    """SURVEY.md A.6-1 is one switch on both sides: with `steps >= max_steps` the oscillating latch of
    A.7 row 6 stops one step earlier (Q = Q' = 1 instead of 0) — engine and oracle must agree either way."""
    for inclusive in (0, 1):
        oracle.lib().go_set_config(inclusive, 0, 0, 0)
        try:
            def mk(builder):
                s_, r_, q, qn = (builder.add_wire() for _ in range(4))
                builder.add_nand_gate([s_, qn], q)
                builder.add_nand_gate([r_, q], qn)
                return builder.build(), s_, r_, q, qn
            osim, s_, r_, q, qn = mk(oracle.Builder())
            gsim, *_ = mk(SimulatorBuilder())
            gsim.set_option(gsim.OPT_RESIDENT_KERNEL, int(resident))
            gsim.set_option(gsim.OPT_STEP_BUDGET_INCLUSIVE, inclusive)
            for sd, rd, ms in [(L0, L0, 100), (L1, L1, 10), (L1, L1, 11), (L0, L1, 0), (L1, L1, 0), (L1, L0, 1), (L1, L1, 3)]:
                osim.set_wire_drive(s_, *sd); osim.set_wire_drive(r_, *rd)
                gsim.set_wire_drive(s_, LogicState(*sd)); gsim.set_wire_drive(r_, LogicState(*rd))
                assert int(gsim.run_sim(ms)) == osim.run_sim(ms), (inclusive, sd, rd, ms)
                for w in (q, qn):
                    assert gsim.get_wire_state(w) == LogicState(*osim.get_wire_state(w)), (inclusive, sd, rd, ms, w)
            if inclusive == 0:
                pass
        finally:
            oracle.lib().go_set_config(0, 0, 0, 0)


def test_drive_queue_order_and_state_snapshot():
    """set_wire_drive is queued on the host and get_wire_state is served from a lane-0 snapshot: the
    call order across the different drive setters still decides the value, a drive is invisible until
    the next run_sim (gsim: drives and states are separate), and the snapshot follows every run."""
    b = SimulatorBuilder()
    a, c, o, w8, n8 = b.add_wire(), b.add_wire(), b.add_wire(), b.add_wire(8), b.add_wire(8)
    b.add_xor_gate([a, c], o)
    b.add_not_gate(w8, n8)
    sim = b.build()
    sim.set_lanes(128)
    one, zero = LogicState(1, 1), LogicState(0, 1)
    ones = np.full(2, (1 << 64) - 1, np.uint64)
    sim.set_wire_drive(a, one)
    sim.set_wire_drive(c, zero)
    sim.set_wire_drive(a, zero)                      # last call wins inside the queue
    assert sim.get_wire_state(o) == LogicState(0, 0)  # nothing ran yet: still High-Z
    assert sim.run_sim(10) == SimulationRunResult.Ok
    assert sim.get_wire_state(o) == zero and sim.get_wire_state(a) == zero
    sim.set_wire_drive(a, one)
    assert sim.get_wire_state(a) == zero              # drive queued, state unchanged until the run
    sim.set_wire_drive_lanes(a, np.zeros(2, np.uint64), ones)   # a later lane-wise drive overrides the queued one
    assert sim.run_sim(10) == SimulationRunResult.Ok
    assert sim.get_wire_state(o) == zero
    sim.set_wire_drive_lanes(a, np.zeros(2, np.uint64), ones)
    sim.set_wire_drive(a, one)                       # ... and a later queued drive overrides the lane-wise one
    sim.set_wire_drive(w8, LogicState(0x5A, 0xFF))
    assert sim.run_sim(10) == SimulationRunResult.Ok
    assert sim.get_wire_state(o) == one
    assert sim.get_wire_state(n8) == LogicState(0xA5, 0xFF)
    v, m = sim.get_wire_lanes(o)
    assert (v == ones).all() and (m == ones).all()   # the queued drive went to every lane
    for k in range(40):                               # many evals: the snapshot is refreshed after each run
        sim.set_wire_drive(c, LogicState(k & 1, 1))
        assert sim.run_sim(10) == SimulationRunResult.Ok
        assert sim.get_wire_state(o) == LogicState(1 ^ (k & 1), 1)
    sim.set_wire_drive(a, one)
    sim.set_lanes(64)                                 # cold again: queued drives are dropped with the store
    assert sim.run_sim(10) == SimulationRunResult.Ok
    assert sim.get_wire_state(a) == LogicState(0, 0)


def test_device_loop_is_deterministic_under_oscillation():
    """Regression: inside the device-side loop a CTA re-reads rows that other CTAs rewrote two
    applications earlier; with cached (weak) loads a stale row came back now and then on oscillating
    lanes near the step budget.  The same run, repeated, must match the oracle every time."""
    from helpers import random_netlist
    for trial in range(6):
        rng = np.random.default_rng(77 * 1 + 1)
        nl, first_in = random_netlist(rng, 6, 48, cyclic=True, wide=True, mux=True)
        n_lanes = 64 * 37
        sim = nl.replay(SimulatorBuilder()).build()
        sim.set_option(sim.OPT_RESIDENT_KERNEL, 0)
        sim.set_lanes(n_lanes)
        lanes = oracle.Lanes(nl.replay(oracle.Builder()), n_lanes)
        mask = lane_mask(n_lanes)
        for max_steps in [0, 1, 2, 3, 5, 8, 60, 7, 0, 1000]:
            for i in range(6):
                if rng.random() < 0.7:
                    v, m = random_drive(rng, lanes.words)
                    lanes.set_drive(first_in + i, v, m)
                    sim.set_wire_drive_lanes(first_in + i, v, m)
            status, _, _ = lanes.run(max_steps)
            res, conflict, unsettled = sim.run_sim_lanes(max_steps)
            assert (oracle.status_from_masks(conflict, unsettled, n_lanes) == status).all(), (trial, max_steps)
            rv, rm = lanes.get_all(nl.n_wires)
            gv, gm = sim.gather_wires_host(np.arange(nl.n_wires, dtype=np.uint32))
            assert (((gv ^ rv) | (gm ^ rm)) & mask).max(initial=0) == 0, (trial, max_steps)
