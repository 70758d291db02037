"""B2_OPT_PERIOD2_SKIP: the unit-delay loops jump over an exact period-2 oscillation.  The jump must be invisible: per-lane
results, every wire and the state left behind (which depends on the parity of the step count) equal
the scalar oracle's, which walks every step — on the resident kernel, the device-side loop and, as the control, the host loop;
oscillators with a longer period (odd rings) must fall through untouched."""
import numpy as np
import pytest

import oracle
from digilogic_b200 import SimulationRunResult, SimulatorBuilder
from helpers import ALL1, NAND, NOR, NOT, XOR, Netlist, lane_mask, random_drive, random_netlist

pytestmark = pytest.mark.gpu


def make(nl, n_lanes, path, skip=1):
    ob = nl.replay(oracle.Builder())
    sim = nl.replay(SimulatorBuilder()).build()
    sim.set_option(sim.OPT_RESIDENT_KERNEL, int(path == "resident"))
    sim.set_option(sim.OPT_DEVICE_LOOP, int(path != "host"))
    sim.set_option(sim.OPT_PERIOD2_SKIP, skip)
    sim.set_lanes(n_lanes)
    return oracle.Lanes(ob, n_lanes), sim


def check(sim, lanes, n_wires, n_lanes, max_steps, tag):
    status, _, _ = lanes.run(max_steps)
    res, conflict, unsettled = sim.run_sim_lanes(max_steps)
    assert (oracle.status_from_masks(conflict, unsettled, n_lanes) == status).all(), tag
    assert int(res) == int(status.max()), tag
    rv, rm = lanes.get_all(n_wires)
    gv, gm = sim.gather_wires_host(np.arange(n_wires, dtype=np.uint32))
    ok = status != 2
    words = (n_lanes + 63) // 64
    okw = np.packbits(np.pad(ok, (0, words * 64 - n_lanes)).reshape(words, 64), axis=1, bitorder="little").view(np.uint64).reshape(words)
    mask = lane_mask(n_lanes) & okw
    assert (((gv ^ rv) | (gm ^ rm)) & mask).max(initial=0) == 0, tag
    return status


def oscillators():
    """Latches forced into unit-delay oscillation (period 2), a NOT ring of one (period 2), rings of three and five (periods 6
    and 10), an XOR feedback whose period depends on the lane's drive, all in one netlist, plus settled logic beside them."""
    nl = Netlist()
    s, r, en, d = (nl.add_wire() for _ in range(4))
    q, qn = nl.add_wire(), nl.add_wire()
    nl.add_gate(NAND, [s, qn], q)
    nl.add_gate(NAND, [r, q], qn)
    n1 = nl.add_wire()
    nl.add_gate(NAND, [en, n1], n1)                  # en = 1: NOT ring of one, en = 0: settles at 1
    ring3 = [nl.add_wire() for _ in range(3)]
    nl.add_gate(NAND, [en, ring3[2]], ring3[0])
    nl.add_not_gate(ring3[0], ring3[1])
    nl.add_not_gate(ring3[1], ring3[2])
    x = nl.add_wire()
    nl.add_gate(XOR, [d, x], x)                      # d = 1: toggles every step; d = 0: holds
    a, b = nl.add_wire(), nl.add_wire()
    nl.add_gate(NOR, [s, b], a)
    nl.add_gate(NOR, [r, a], b)
    return nl, (s, r, en, d)


@pytest.mark.parametrize("path", ["resident", "device", "host"])
@pytest.mark.parametrize("n_lanes", [1, 64 * 3 + 5, 64 * 40])
def test_oscillators_step_parity_and_results(path, n_lanes):
    nl, ins = oscillators()
    lanes, sim = make(nl, n_lanes, path)
    words = lanes.words
    rng = np.random.default_rng(3 + n_lanes)
    ones, zeros = np.full(words, ALL1, np.uint64), np.zeros(words, np.uint64)

    def drive(wire, v, m):
        lanes.set_drive(wire, v, m)
        sim.set_wire_drive_lanes(wire, v, m)

    s, r, en, d = ins
    # per lane: which oscillators are switched on; the latch is first set up with S' = R' = 0, then both rise together
    drive(s, zeros, ones); drive(r, zeros, ones)
    drive(en, rng.integers(0, 1 << 64, words, dtype=np.uint64) if n_lanes > 1 else zeros, ones)
    drive(d, rng.integers(0, 1 << 64, words, dtype=np.uint64) if n_lanes > 1 else zeros, ones)
    check(sim, lanes, nl.n_wires, n_lanes, 50, "setup")
    drive(s, ones, ones); drive(r, ones, ones)
    seen = set()
    for max_steps in [10, 11, 10_000, 9_999, 10_001, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 12, 13, 37, 38, 1000, 1001]:
        st = check(sim, lanes, nl.n_wires, n_lanes, max_steps, (path, max_steps))
        seen |= set(int(x) for x in st)
    assert 1 in seen, "the latch must oscillate"
    if n_lanes == 1:
        # the latch alone (en = d = 0 in lane 0): period 2, so a 10 000-step run must not walk 10 003 applications
        res, _, _ = sim.run_sim_lanes(10_000)
        assert res == SimulationRunResult.MaxStepsReached
        if path != "host":
            assert sim.info.last_applications < 40, sim.info.last_applications


@pytest.mark.parametrize("path", ["resident", "device"])
def test_skip_is_invisible_on_random_cyclic_netlists(path):
    """Random feedback netlists with 4-state drives over warm restarts and step budgets that end inside oscillations: skip on
    equals skip off equals the oracle."""
    for seed in range(6):
        rng = np.random.default_rng(1000 + seed)
        nl, first_in = random_netlist(rng, 5, 40, cyclic=True, wide=True, mux=True)
        n_lanes = 64 * 5
        lanes, on = make(nl, n_lanes, path, 1)
        _, off = make(nl, n_lanes, path, 0)
        for max_steps in [23, 24, 100, 101, 3, 10_000, 10_001]:
            for i in range(5):
                v, m = random_drive(rng, lanes.words)
                lanes.set_drive(first_in + i, v, m)
                on.set_wire_drive_lanes(first_in + i, v, m)
                off.set_wire_drive_lanes(first_in + i, v, m)
            check(on, lanes, nl.n_wires, n_lanes, max_steps, (seed, max_steps))
            r2, c2, u2 = off.run_sim_lanes(max_steps)
            a = on.gather_wires_host(np.arange(nl.n_wires, dtype=np.uint32))
            b = off.gather_wires_host(np.arange(nl.n_wires, dtype=np.uint32))
            clean = ~c2                                   # (the state after a conflict is outside the contract)
            assert (((a[0] ^ b[0]) | (a[1] ^ b[1])) & clean).max(initial=0) == 0, (seed, max_steps)
