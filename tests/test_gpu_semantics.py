"""The reconstructed points of gsim's behaviour as builder switches of the engine, each tested in lockstep
with the oracle's matching switch on every device path: whichever way a run of the real crate (oracle/gsim_diff) decides, the
engine follows by flipping B2_BOPT_*, no kernel edit.

    A.6-4  B2_BOPT_MUX_Z_PASSTHROUGH     a multiplexer passes a High-Z selected input on as High-Z (default: X)
    A.6-5  B2_BOPT_INITIAL_UNDEFINED     wires and component outputs start Undefined (default: High-Z)
    A.6-6  B2_BOPT_BUFFER_Z_PASSTHROUGH  an enabled buffer passes a High-Z input bit on as High-Z (default: X)
    A.6-7  B2_BOPT_COPY_Z_PASSTHROUGH    slices and merges pass a High-Z input bit on as High-Z (default: X)
"""
import itertools

import numpy as np
import pytest

import oracle
from digilogic_b200 import LogicState, SimulationRunResult, SimulatorBuilder
from helpers import ALL1, lane_mask, random_drive

pytestmark = pytest.mark.gpu

FLAGS = ["mux_z", "init_x", "buf_z", "copy_z"]
COMBOS = [(), ("mux_z",), ("init_x",), ("buf_z",), ("copy_z",), ("mux_z", "init_x", "buf_z", "copy_z")]
PATHS = ["resident", "device", "host", "dense"]


class oracle_config:
    def __init__(self, on):
        self.on = set(on)

    def __enter__(self):
        L = oracle.lib()
        L.go_set_config(0, 0, int("mux_z" in self.on), int("init_x" in self.on))
        L.go_set_buffer_z_passthrough(int("buf_z" in self.on))
        L.go_set_copy_z_passthrough(int("copy_z" in self.on))

    def __exit__(self, *exc):
        L = oracle.lib()
        L.go_set_config(0, 0, 0, 0)
        L.go_set_buffer_z_passthrough(0)
        L.go_set_copy_z_passthrough(0)


def engine_builder(on):
    b = SimulatorBuilder()
    b.set_option(b.OPT_MUX_Z_PASSTHROUGH, int("mux_z" in on))
    b.set_option(b.OPT_INITIAL_UNDEFINED, int("init_x" in on))
    b.set_option(b.OPT_BUFFER_Z_PASSTHROUGH, int("buf_z" in on))
    b.set_option(b.OPT_COPY_Z_PASSTHROUGH, int("copy_z" in on))
    return b


def circuit(b, rng, cyclic, multi=True):
    """Multiplexers (1- and 2-bit select, 4-bit data), buffers onto a shared bus and onto plain wires, slices and merges of
    4-bit wires, reduction gates and plain gates reading all of them; a multiplexer and a buffer share one net (legal only when
    multiplexers can float); with `cyclic`, gates feed back.  multi = False leaves every net with one driver (with outputs
    that start Undefined, two drivers on a net conflict in the first wire update of every lane).  Returns (input wires, all
    wires with widths)."""
    ins1 = [b.add_wire(1) for _ in range(8)]
    ins4 = [b.add_wire(4) for _ in range(4)]
    wires = [(w, 1) for w in ins1] + [(w, 4) for w in ins4]

    def new(width):
        w = b.add_wire(width)
        wires.append((w, width))
        return w

    m1 = new(4); b.add_multiplexer([ins4[0], ins4[1]], ins1[0], m1)
    sel2 = new(2); b.add_merge([ins1[1], ins1[2]], sel2)
    m2 = new(4); b.add_multiplexer([ins4[0], ins4[1], ins4[2], m1], sel2, m2)
    bus = new(4)
    b.add_buffer(ins4[2], ins1[3], bus)
    if multi:
        b.add_buffer(m2, ins1[4], bus)
    lone = new(4); b.add_buffer(ins4[3], ins1[5], lone)
    sl = new(2); b.add_slice(bus, 1, sl)
    mg = new(4); b.add_merge([sl, ins1[6], ins1[7]], mg)
    shared = new(1)                                        # a multiplexer and a buffer on one net
    b.add_multiplexer([ins1[6], ins1[7]], ins1[0], shared)
    if multi:
        b.add_buffer(ins1[1], ins1[2], shared)
    g1 = new(4); b.add_xor_gate([mg, lone], g1)
    g2 = new(4); b.add_nand_gate([m1, bus, g1], g2)
    r1 = new(1); b.add_horizontal_or_gate(g2, r1)
    s1 = new(1); b.add_slice(lone, 3, s1)
    fb = new(1); b.add_nor_gate([r1, s1], fb)
    if cyclic:
        loop = new(1); b.add_nand_gate([fb, loop], loop)   # a gated ring oscillator: fb = 1 lets it toggle
        q = new(1); qn = new(1)
        b.add_nand_gate([shared, qn], q)                   # a latch set and reset by floating-capable nets
        b.add_nand_gate([r1, q], qn)
    return ins1, ins4, wires


def drive_all(rng, targets, ins1, ins4, words, p_z=0.35):
    """Random 4-state drives, High-Z (0, 0) with probability p_z per lane and bit."""
    last = {}
    for w, width in [(w, 1) for w in ins1] + [(w, 4) for w in ins4]:
        for bit in range(width):
            v, m = random_drive(rng, words)
            z = rng.integers(0, 1 << 64, words, dtype=np.uint64) & rng.integers(0, 1 << 64, words, dtype=np.uint64)
            if p_z > 0.3:
                z |= rng.integers(0, 1 << 64, words, dtype=np.uint64) & rng.integers(0, 1 << 64, words, dtype=np.uint64) \
                    & rng.integers(0, 1 << 64, words, dtype=np.uint64)
            v, m = v & ~z, m & ~z
            if w == ins1[3]:
                m = np.full(words, ALL1, np.uint64)        # a valid enable (an X enable makes its buffer drive X)
            if w in (ins1[2], ins1[4]):
                # enables of the second driver of a shared net: on in few lanes only, so that conflicts (which end a lane's
                # comparable life, A.6-3) stay rare — ins1[4] is mostly the complement of ins1[3], ins1[2] mostly 0
                rare = rng.integers(0, 1 << 64, words, dtype=np.uint64)
                for _ in range(4):
                    rare &= rng.integers(0, 1 << 64, words, dtype=np.uint64)
                v = ((~last.get(ins1[3], v)) if w == ins1[4] else np.zeros(words, np.uint64)) | rare
                m = np.full(words, ALL1, np.uint64)
            last[w] = v
            for t in targets:
                if isinstance(t, oracle.Lanes):
                    t.set_drive(w, v, m, bit)
                else:
                    t.set_wire_drive_lanes(w, v, m, bit=bit)


def make_sim(on, rng_seed, cyclic, n_lanes, path):
    b = engine_builder(on)
    ins1, ins4, wires = circuit(b, np.random.default_rng(rng_seed), cyclic, "init_x" not in on)
    sim = b.build()
    sim.set_option(sim.OPT_RESIDENT_KERNEL, int(path == "resident"))
    sim.set_option(sim.OPT_DEVICE_LOOP, int(path in ("resident", "device")))
    sim.set_option(sim.OPT_FRONTIER, int(path != "dense"))
    sim.set_lanes(n_lanes)
    return sim, ins1, ins4, wires


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("on", COMBOS, ids=lambda c: "+".join(c) or "default")
@pytest.mark.parametrize("cyclic", [False, True])
def test_switches_in_lockstep_with_the_oracle(on, cyclic, path):
    n_lanes = 64 * 3 + 17
    with oracle_config(on):
        ob = oracle.Builder()
        circuit(ob, np.random.default_rng(5), cyclic, "init_x" not in on)
        lanes = oracle.Lanes(ob, n_lanes)
        sim, ins1, ins4, wires = make_sim(on, 5, cyclic, n_lanes, path)
        rng = np.random.default_rng(42 + len(on))
        mask = lane_mask(n_lanes)
        alive = np.ones(n_lanes, bool)                    # lanes that have never conflicted: what happens after a conflict is
        seen = set()                                      # outside the contract (A.6-3), in this run and in every later one
        for rnd, max_steps in enumerate([0, 1, 2, 40, 3, 1000, 17, 10_000]):
            drive_all(rng, [lanes, sim], ins1, ins4, lanes.words)
            status, _, _ = lanes.run(max_steps)
            res, conflict, unsettled = sim.run_sim_lanes(max_steps)
            got = oracle.status_from_masks(conflict, unsettled, n_lanes)
            assert (got[alive] == status[alive]).all(), (rnd, max_steps)
            seen |= set(int(x) for x in status[alive])
            alive &= status != 2
            assert alive.sum() > n_lanes // 4, "too few lanes left to compare"
            alive_w = np.packbits(np.pad(alive, (0, lanes.words * 64 - n_lanes)).reshape(lanes.words, 64), axis=1,
                                  bitorder="little").view(np.uint64).reshape(lanes.words)
            clean = mask & alive_w
            for w, width in wires:
                for bit in range(width):
                    rv, rm = lanes.get(w, bit)
                    gv, gm = sim.get_wire_lanes(w, bit=bit)
                    assert (((gv ^ rv) | (gm ^ rm)) & clean).max(initial=0) == 0, (rnd, max_steps, w, bit)
        assert (2 in seen) == ("init_x" not in on), "lanes must conflict exactly where a net has two drivers"


@pytest.mark.parametrize("on", COMBOS, ids=lambda c: "+".join(c) or "default")
def test_single_lane_known_answers(on):
    """The gsim-shaped single-lane calls: what each switch changes, spelled out."""
    on = set(on)
    b = engine_builder(on)
    d0, d1, sel, en, out_m, out_b = (b.add_wire() for _ in range(6))
    wide, sl = b.add_wire(2), b.add_wire(1)
    b.add_multiplexer([d0, d1], sel, out_m)
    b.add_buffer(d0, en, out_b)
    b.add_merge([d0, d1], wide)
    b.add_slice(wide, 0, sl)
    sim = b.build()
    one, zero, z, x = LogicState(1, 1), LogicState(0, 1), LogicState.HIGH_Z, LogicState.UNDEFINED
    # before any run: the initial wire state
    assert sim.get_wire_state(out_m) == (x if "init_x" in on else z)
    sim.set_wire_drive(d0, z); sim.set_wire_drive(d1, one); sim.set_wire_drive(sel, zero); sim.set_wire_drive(en, one)
    assert sim.run_sim(100) == SimulationRunResult.Ok
    assert sim.get_wire_state(out_m) == (z if "mux_z" in on else x)          # High-Z selected input
    assert sim.get_wire_state(out_b) == (z if "buf_z" in on else x)          # enabled buffer, High-Z input
    assert sim.get_wire_state(wide) == (LogicState(0b10, 0b10) if "copy_z" in on else LogicState(0b11, 0b10))
    assert sim.get_wire_state(sl) == (z if "copy_z" in on else x)
    sim.set_wire_drive(sel, one)
    assert sim.run_sim(100) == SimulationRunResult.Ok and sim.get_wire_state(out_m) == one
    sim.set_wire_drive(sel, x)
    assert sim.run_sim(100) == SimulationRunResult.Ok and sim.get_wire_state(out_m) == x   # invalid select: X either way


def test_initial_undefined_conflicts_at_the_cold_start():
    """With outputs starting Undefined a drive on a gate-driven wire, or two buffers on one net, conflict in the very first
    wire update — with High-Z outputs they do not (the drive conflicts one step later, the idle bus never)."""
    for init_x, path in itertools.product([0, 1], ["resident", "device"]):
        with oracle_config(["init_x"] if init_x else []):
            def build(b):
                a, en0, en1, bus, g = (b.add_wire() for _ in range(5))
                b.add_buffer(a, en0, bus)
                b.add_buffer(a, en1, bus)
                b.add_not_gate(a, g)
                return a, en0, en1, bus, g
            ob = oracle.Builder(); build(ob)
            osim = ob.build()
            eb = engine_builder(["init_x"] if init_x else [])
            a, en0, en1, bus, g = build(eb)
            sim = eb.build()
            sim.set_option(sim.OPT_RESIDENT_KERNEL, int(path == "resident"))
            for w in (a, en0, en1):
                osim.set_wire_drive(w, 0, 1)
                sim.set_wire_drive(w, LogicState(0, 1))
            want = osim.run_sim(100)
            got = sim.run_sim(100)
            assert int(got) == int(want), (init_x, path)
            assert int(want) == (2 if init_x else 0)
