#!/usr/bin/env python
"""Latency of the reference's real use — one lane, `Eval{max_steps: 10_000}` per click
(crates/digilogic_netcode/src/client.rs:223,266) — through B200Server, with the on-chip resident
kernel and with the HBM unit-delay path.  Prints one JSON line per case."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from digilogic_b200 import LogicState, SimulationRunResult, SimulatorBuilder, netgen  # noqa: E402


def timed(fn, reps):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


def latch(resident):
    b = SimulatorBuilder()
    s_, r_, q, qn = (b.add_wire() for _ in range(4))
    b.add_nand_gate([s_, qn], q)
    b.add_nand_gate([r_, q], qn)
    sim = b.build()
    sim.set_option(sim.OPT_RESIDENT_KERNEL, int(resident))
    one, zero = LogicState(1, 1), LogicState(0, 1)

    def oscillate():
        sim.set_wire_drive(s_, zero); sim.set_wire_drive(r_, zero)
        assert sim.run_sim(10_000) == SimulationRunResult.Ok
        sim.set_wire_drive(s_, one); sim.set_wire_drive(r_, one)
        assert sim.run_sim(10_000) == SimulationRunResult.MaxStepsReached

    return timed(oscillate, 3 if not resident else 20), sim.info.last_applications


def multiplier(resident):
    gen = netgen.array_multiplier(32)
    sim = gen.emit(SimulatorBuilder()).build()
    sim.set_option(sim.OPT_RESIDENT_KERNEL, int(resident))
    state = [0]

    def click():
        state[0] ^= 1
        sim.set_wire_drive(int(gen.inputs[0]), LogicState(state[0], 1))
        assert sim.run_sim(10_000) == SimulationRunResult.Ok
        sim.get_wire_state(int(gen.outputs[-1]))  # the read-back waits for the run (the levelised path returns early)

    for w in gen.inputs:
        sim.set_wire_drive(int(w), LogicState(1, 1))
    return timed(click, 20), sim.info.last_mode


if __name__ == "__main__":
    for resident in (True, False):
        t, apps = latch(resident)
        print(json.dumps({"case": "NAND latch forced to oscillate, run_sim(10_000) -> MaxStepsReached (2 evals)",
                          "resident_kernel": resident, "seconds_per_pair": t, "applications": apps}))
        t, mode = multiplier(resident)
        print(json.dumps({"case": "32x32 array multiplier (5888 gates, depth 184), one input toggled + run_sim(10_000)",
                          "resident_kernel": resident, "seconds_per_eval": t, "mode": mode}))
