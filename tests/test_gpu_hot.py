"""Shared-memory staging of hot fan-in rows in the unit-delay work-list kernel (B2_OPT_HOT_STAGING, k_eval_list: cp.async.bulk +
mbarrier): a clock-like input read by hundreds of gates of a netlist with feedback, against the scalar oracle on every wire and
lane, with the staging on and off, at strides that make one and several column blocks per row."""
import numpy as np
import pytest

import digilogic_b200 as d
import oracle

from helpers import Netlist

pytestmark = pytest.mark.gpu


def clocked_netlist(rng, n_cells=160):
    """`n_cells` gated SR-latch-like cells: every cell reads the shared wires `clk` and `rst` (fan-out = hundreds) and one
    private data input; two- and three-input NAND / NOR gates with feedback inside each cell."""
    nl = Netlist()
    clk, rst = nl.add_wires(1), nl.add_wires(1)
    data = nl.add_wires(n_cells)
    for c in range(n_cells):
        s, r, q, qn, t = (nl.add_wires(1) for _ in range(5))
        nl.add_gate(3, [data + c, clk], s)                 # NAND(d, clk)
        nl.add_gate(3, [s, clk, rst], r)                   # NAND(s, clk, rst): a three-input gate on both hot rows
        nl.add_gate(3, [s, qn], q)                         # cross-coupled pair
        nl.add_gate(3, [r, q, rst], qn)
        nl.add_gate(int(rng.integers(0, 6)), [q, int(data + rng.integers(0, n_cells))], t)
    return nl, clk, rst, data, n_cells


@pytest.mark.parametrize("n_lanes", [64 * 40, 64 * 600])     # 40 words: one column block; 600 words: two blocks of 256 pairs, a short one
def test_clocked_cells_match_the_oracle_and_the_unstaged_kernel(n_lanes):
    rng = np.random.default_rng(n_lanes)
    nl, clk, rst, data, n_cells = clocked_netlist(rng)
    sims = []
    for hot in (1, 0):
        sim = nl.replay(d.SimulatorBuilder()).build()
        sim.set_option(sim.OPT_RESIDENT_KERNEL, 0)
        sim.set_option(sim.OPT_DEVICE_LOOP, 0)               # host loop: the standalone work-list kernels
        sim.set_option(sim.OPT_HOT_STAGING, hot)
        sim.set_lanes(n_lanes)
        staged = sim.hot_rows
        assert (len(staged) == 2) == bool(hot) and (not hot or set(staged) == {clk, rst})   # sources come first: row = wire here
        sims.append(sim)
    words = n_lanes // 64
    n_oracle = min(n_lanes, 256)
    ow = (n_oracle + 63) // 64
    lanes = oracle.Lanes(nl.replay(oracle.Builder()), n_oracle)
    ones = np.full(words, np.uint64(0xFFFFFFFFFFFFFFFF))
    all_wires = np.arange(nl.n_wires, dtype=np.uint32)

    def drive(wire, value, valid):
        for sim in sims:
            sim.set_wire_drive_lanes(wire, value, valid)
        lanes.set_drive(wire, value[:ow], valid[:ow])

    drive(rst, ones, ones)
    for step in range(10):
        for c in range(n_cells):
            if step == 0 or rng.random() < 0.3:
                drive(data + c, rng.integers(0, 1 << 64, words, dtype=np.uint64), ones if step else rng.integers(0, 1 << 64, words, dtype=np.uint64) | ones)
        for level in (ones, np.zeros(words, np.uint64)):
            drive(clk, level if step != 4 else rng.integers(0, 1 << 64, words, dtype=np.uint64), ones)
            status, _, _ = lanes.run(500)
            planes = []
            for sim in sims:
                res, conflict, unsettled = sim.run_sim_lanes(500)
                assert sim.info.last_mode == 2
                planes.append((int(res), conflict, unsettled) + sim.gather_wires_host(all_wires))
            (res, conflict, unsettled, gv, gm), (res0, conflict0, unsettled0, gv0, gm0) = planes
            # every lane of every wire: staged = not staged; the oracle's lanes: = the oracle
            assert res == res0 and (conflict == conflict0).all() and (unsettled == unsettled0).all()
            assert (gv == gv0).all() and (gm == gm0).all(), (step,)
            rv, rm = lanes.get_all(nl.n_wires)
            assert (gv[:, :ow] == rv).all() and (gm[:, :ow] == rm).all(), (step,)
            if words <= ow:
                assert res == int(status.max())
