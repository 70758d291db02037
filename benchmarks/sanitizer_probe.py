#!/usr/bin/env python
"""A small tour of every kernel for `compute-sanitizer --tool memcheck` (one tool per gpurun call):
levelised, unit-delay in HBM, resident, valid-plane elision, wide gates, muxes, multi-bit wires,
drives on gate outputs, all read-back paths."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from digilogic_b200 import LogicState, SimulatorBuilder, netgen  # noqa: E402
from helpers import random_drive, random_netlist  # noqa: E402

rng = np.random.default_rng(3)
for cyclic in (False, True):
    for resident in (True, False):
        for n_lanes in (1, 100, 64 * 19):
            nl, first_in = random_netlist(rng, 6, 60, cyclic=cyclic, wide=True, mux=True)
            sim = nl.replay(SimulatorBuilder()).build()
            sim.set_option(sim.OPT_RESIDENT_KERNEL, int(resident))
            sim.set_option(sim.OPT_VALID_ELISION, 1)
            sim.set_lanes(n_lanes)
            words = (n_lanes + 63) // 64
            for ms in (0, 3, 1000):
                for i in range(6):
                    v, m = random_drive(rng, words, four_state=(i % 2 == 0))
                    sim.set_wire_drive_lanes(first_in + i, v, m)
                sim.run_sim_lanes(ms)
                sim.gather_wires_host(np.arange(nl.n_wires, dtype=np.uint32))
                sim.get_wire_lanes(nl.n_wires - 1)
                sim.pack_sim_state(np.arange(nl.n_wires, dtype=np.uint32))
            # a drive on a gate output (conflict path)
            v, m = random_drive(rng, words)
            sim.set_wire_drive_lanes(nl.n_wires - 1, v, m)
            sim.run_sim_lanes(50)
# multi-bit wires + stimulus generator + bulk drives
b = SimulatorBuilder()
a, c, o = b.add_wire(40), b.add_wire(40), b.add_wire(40)
b.add_xor_gate([a, c], o)
s = b.build()
s.set_wire_drive(a, LogicState(0x123456789A, (1 << 40) - 1))
s.set_wire_drive(c, LogicState(0x0F0F0F0F0F, (1 << 40) - 1))
s.run_sim(10)
assert s.get_wire_state(o) == LogicState(0x123456789A ^ 0x0F0F0F0F0F, (1 << 40) - 1)
gen = netgen.random_dag(n_gates=3000, depth=30, n_inputs=64, n_outputs=16)
sim = gen.emit(SimulatorBuilder()).build()
sim.set_lanes(64 * 130)
sim.fill_stimulus(gen.inputs, 7, 11, 1)
sim.run_sim_lanes(10_000)
v, m = sim.gather_wires_host(gen.inputs)
sim.set_drives_lanes(gen.inputs, v, m)
sim.run_sim_lanes(5)
seq = netgen.sequential(n_latches=8, n_chains=2, bits=8, n_hazard=1)
q = seq.emit(SimulatorBuilder()).build()
q.set_lanes(200)
q.run_sim_lanes(30)
# tri-state buses, registers, wide lists: resident kernel, device-side loop, host loop, dense
from helpers import bus_enable_drives, random_bus_netlist, random_register_netlist  # noqa: E402
for resident, frontier, device_loop in ((1, 1, 1), (0, 1, 1), (0, 1, 0), (0, 0, 0)):
    for kind in ("bus", "reg"):
        if kind == "bus":
            nl, first_in, groups = random_bus_netlist(rng, 6, 40, 4)
        else:
            nl, first_in = random_register_netlist(rng, 6, 40)
            groups = []
        sim = nl.replay(SimulatorBuilder()).build()
        sim.set_option(sim.OPT_RESIDENT_KERNEL, resident)
        sim.set_option(sim.OPT_FRONTIER, frontier)
        sim.set_option(sim.OPT_DEVICE_LOOP, device_loop)
        n_lanes = 64 * 5 + 3
        sim.set_lanes(n_lanes)
        words = (n_lanes + 63) // 64
        for ms in (0, 2, 40, 1000):
            for i in range(6):
                v, m = random_drive(rng, words)
                sim.set_wire_drive_lanes(first_in + i, v, m)
            for grp in groups:
                for e, (v, m) in zip(grp, bus_enable_drives(rng, grp, words)):
                    sim.set_wire_drive_lanes(e, v, m)
            sim.run_sim_lanes(ms)
            sim.gather_wires_host(np.arange(nl.n_wires, dtype=np.uint32))
# the single-lane server path: queued drives, snapshot reads, slices / merges / adders / reductions
b = SimulatorBuilder()
x, y, ssum, lo, red = b.add_wire(12), b.add_wire(12), b.add_wire(12), b.add_wire(4), b.add_wire(1)
b.add_add(x, y, ssum)
b.add_slice(ssum, 8, lo)
b.add_horizontal_gate(2, lo, red)
s = b.build()
for k in range(20):
    s.set_wire_drive(x, LogicState(k * 37, 0xFFF))
    s.set_wire_drive(y, LogicState(k * 91, 0xFFF))
    s.run_sim(100)
    assert s.get_wire_state(ssum) == LogicState((k * 37 + k * 91) & 0xFFF, 0xFFF)
    s.get_wire_state(red)
print("sanitizer probe done")
