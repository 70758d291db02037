"""Level-pair fusion of the levelised pass (B2_OPT_PAIR_FUSION, csrc/pair.cuh): two dependent gates per thread, rows read only
inside their item never stored.  The batch runner's designated outputs must equal the dense CPU model on every lane, with the
fusion on and off, on netlists that exercise every item shape."""
import numpy as np
import pytest

import digilogic_b200 as d
import oracle
from digilogic_b200 import netgen
from digilogic_b200.batch import Batch
from digilogic_b200.gsim import SimulationRunResult

from helpers import Netlist, random_netlist, stimulus_words

pytestmark = pytest.mark.gpu


def _oracle_rows(nl, n_wires, inputs, words, seed, word_begin, mode):
    ob = nl.replay(oracle.Builder()) if hasattr(nl, "replay") else nl.emit(oracle.Builder())
    bs = oracle.BitSliced(ob, 64 * words, n_wires)
    v, m = stimulus_words(seed, len(inputs), words, word_begin, mode)
    for i, w in enumerate(inputs):
        bs.set_drive(int(w), v[i], m[i])
    assert not bs.eval_dag().any()
    return bs.get_all()


def _run(sim, inputs, outputs, tile, slots, words, seed, mode, pair):
    sim.set_option(sim.OPT_PAIR_FUSION, pair)
    b = Batch(sim, inputs, outputs, tile, slots)
    b.set_option(Batch.OPT_SWEEP, 0)
    ov = np.zeros((len(outputs), words), np.uint64)
    om = np.zeros_like(ov)
    for s in (seed ^ 0x55, seed):                  # twice: the slots' rows are reused, the level graph is replayed
        b.run_generated(s, mode, 3, words, 10_000, ov, om)
        assert b.wait() == SimulationRunResult.Ok
    return ov, om, b.info.pair_launches


@pytest.mark.parametrize("tile,slots,n_gates,keep_every", [(16, 2, 300, 7), (32, 3, 2000, 50), (128, 2, 1500, 1), (64, 1, 800, 3)])
def test_random_two_input_netlists_match_the_oracle(tile, slots, n_gates, keep_every):
    """Random fan-ins from anywhere below (items with 0, 1 and 2 children, NOT gates, children pushed to the next launch),
    4-state stimulus, a batch that is not a whole number of slot groups; a sparse or a full set of designated outputs."""
    rng = np.random.default_rng(tile * 7 + n_gates)
    nl, first_in = random_netlist(rng, 20, n_gates, cyclic=False, wide=False, mux=False)
    inputs = np.arange(first_in, first_in + 20, dtype=np.uint32)
    outputs = np.arange(20 + (n_gates % keep_every), nl.n_wires, keep_every, dtype=np.uint32)
    sim = nl.replay(d.SimulatorBuilder()).build()
    words = tile * (2 * slots + 1)       # the replicas take whole tiles
    rv, rm = _oracle_rows(nl, nl.n_wires, inputs, words, 0xFACE, 3, 1)
    ov, om, launches = _run(sim, inputs, outputs, tile, slots, words, 0xFACE, 1, pair=1)
    assert 0 < launches <= sim.info.depth
    assert (ov == rv[outputs]).all() and (om == rm[outputs]).all()
    ov0, om0, launches0 = _run(sim, inputs, outputs, tile, slots, words, 0xFACE, 1, pair=0)
    assert launches0 == 0
    assert (ov0 == ov).all() and (om0 == om).all()


def test_deep_dag_halves_the_launches():
    """The shape of BASELINE config 3 (first fan-in from the level below, second from anywhere), 4-state stimulus."""
    gen = netgen.random_dag(n_gates=40_000, depth=40, n_inputs=256, n_outputs=128, seed=11)
    sim = gen.emit(d.SimulatorBuilder()).build()
    words = 128 * 5
    rv, rm = _oracle_rows(gen, gen.n_wires, gen.inputs, words, 0xD1610003, 3, 1)
    ov, om, launches = _run(sim, gen.inputs, gen.outputs, 128, 3, words, 0xD1610003, 1, pair=1)
    assert 20 <= launches <= 26
    assert (ov == rv[gen.outputs]).all() and (om == rm[gen.outputs]).all()
    assert (om != np.uint64(0xFFFFFFFFFFFFFFFF)).any()


def test_item_shapes_by_hand():
    """Chains of NOT gates (a child that is itself a parent's only reader), a gate with both inputs on one wire, a gate whose
    two fan-ins are written by the same launch (pushed back), a parent with four readers (two children, two pushed back),
    and outputs that would otherwise never be stored."""
    nl = Netlist()
    a, b_ = nl.add_wires(2), None
    b_ = a + 1
    n1 = nl.add_wires(1); nl.add_not_gate(a, n1)              # parent
    n2 = nl.add_wires(1); nl.add_not_gate(n1, n2)             # child of n1 (only reader: n1 is never stored)
    n3 = nl.add_wires(1); nl.add_not_gate(n2, n3)             # parent of the next launch
    x1 = nl.add_wires(1); nl.add_gate(2, [n3, n3], x1)        # XOR of a wire with itself: child, SELF
    g1 = nl.add_wires(1); nl.add_gate(0, [a, b_], g1)         # AND: parent with four readers
    r = [nl.add_wires(1) for _ in range(4)]
    nl.add_gate(1, [g1, a], r[0]); nl.add_gate(3, [b_, g1], r[1]); nl.add_gate(4, [g1, b_], r[2]); nl.add_gate(5, [a, g1], r[3])
    t = nl.add_wires(1); nl.add_gate(2, [r[0], r[1]], t)      # both fan-ins written by one launch
    u = nl.add_wires(1); nl.add_gate(0, [t, x1], u)
    inputs = np.array([a, b_], np.uint32)
    sim = nl.replay(d.SimulatorBuilder()).build()
    words = 48
    rv, rm = _oracle_rows(nl, nl.n_wires, inputs, words, 5, 3, 1)
    for outputs in (np.array([u], np.uint32), np.array([n1, n2, x1, r[3], t, u], np.uint32), np.arange(nl.n_wires, dtype=np.uint32)):
        ov, om, launches = _run(sim, inputs, outputs, 16, 2, words, 5, 1, pair=1)
        assert launches > 0
        assert (ov == rv[outputs]).all() and (om == rm[outputs]).all()


def test_netlists_with_other_gate_classes_keep_the_per_level_kernels():
    rng = np.random.default_rng(3)
    nl, first_in = random_netlist(rng, 16, 400, cyclic=False, wide=True, mux=True)
    inputs = np.arange(first_in, first_in + 16, dtype=np.uint32)
    outputs = np.arange(16, nl.n_wires, 5, dtype=np.uint32)
    sim = nl.replay(d.SimulatorBuilder()).build()
    words = 64
    rv, rm = _oracle_rows(nl, nl.n_wires, inputs, words, 77, 3, 1)
    ov, om, launches = _run(sim, inputs, outputs, 32, 2, words, 77, 1, pair=1)
    assert launches == 0
    assert (ov == rv[outputs]).all() and (om == rm[outputs]).all()


@pytest.mark.parametrize("pair", [2, 1, 0])
def test_plain_simulator_runs_keep_every_row(pair):
    """b2_run_sim_lanes must leave every wire readable: the launches are fused (half as many), but every parent row is stored
    and nothing is dropped."""
    gen = netgen.random_dag(n_gates=4000, depth=10, n_inputs=32, n_outputs=16, seed=3)
    sim = gen.emit(d.SimulatorBuilder()).build()
    sim.set_option(sim.OPT_PAIR_FUSION, pair)
    sim.set_lanes(64 * 32)
    rv, rm = _oracle_rows(gen, gen.n_wires, gen.inputs, 32, 1, 0, 1)
    for _ in range(2):                                   # the second run replays the captured graph
        sim.fill_stimulus(gen.inputs, 1, 0, 1)
        assert sim.run_sim_lanes(100)[0] == SimulationRunResult.Ok and sim.info.last_mode == 1
        launches, children, unstored = sim.pair_info
        assert unstored == 0 and (launches > 0) == (pair != 0) and (children > 0) == (pair != 0)
        gv, gm = sim.gather_wires_host(np.arange(gen.n_wires, dtype=np.uint32))
        assert (gv == rv).all() and (gm == rm).all()


@pytest.mark.parametrize("pair", [1, 0])
@pytest.mark.parametrize("pdl", [1, 2])
def test_programmatic_dependent_launch_keeps_the_row_loads_behind_the_wait(pdl, pair):
    """Tiny tiles, tiny launches: the next launch is resident long before the previous one has written its rows.  A row load
    hoisted above griddepcontrol.wait (ptxas does that to .nc loads) shows up here as wrong outputs; with early release
    (B2_OPT_PDL = 1) on every run."""
    rng = np.random.default_rng(16 * 7 + 300)
    nl, first_in = random_netlist(rng, 20, 300, cyclic=False, wide=False, mux=False)
    inputs = np.arange(first_in, first_in + 20, dtype=np.uint32)
    outputs = np.arange(20, nl.n_wires, 3, dtype=np.uint32)
    words = 16 * 5
    rv, rm = _oracle_rows(nl, nl.n_wires, inputs, words, 0xFACE, 3, 1)
    for rep in range(3):
        sim = nl.replay(d.SimulatorBuilder()).build()
        sim.set_option(sim.OPT_PDL, pdl)
        sim.set_option(sim.OPT_PAIR_FUSION, pair)
        b = Batch(sim, inputs, outputs, 16, 2)
        ov = np.zeros((len(outputs), words), np.uint64)
        om = np.zeros_like(ov)
        for _ in range(3):
            b.run_generated(0xFACE, 1, 3, words, 10_000, ov, om)
            assert b.wait() == SimulationRunResult.Ok
            assert (ov == rv[outputs]).all() and (om == rm[outputs]).all()


def test_config3_shape_fused_equals_per_level_kernels():
    """200 000 gates, depth 200 (the shape of BASELINE config 3 at a fifth of its width), the bench's batch shape: the fused
    plan and the per-level kernels hand out the same output planes, and the plan is the one b2_sim_plan_pairs reports."""
    gen = netgen.random_dag(n_gates=200_000, depth=200, n_inputs=4096, n_outputs=1024, seed=0xD1610003)
    sim = gen.emit(d.SimulatorBuilder()).build()
    words = 128 * 6
    ov1, om1, launches = _run(sim, gen.inputs, gen.outputs, 128, 3, words, 0xD1610003, 0, pair=1)
    plan = sim.plan_pairs(gen.outputs)
    assert launches == plan["launches"] and 100 <= launches <= 125
    assert plan["items"] + plan["children"] == 200_000 and plan["children"] > 80_000 and plan["unstored"] > 20_000
    ov0, om0, launches0 = _run(sim, gen.inputs, gen.outputs, 128, 3, words, 0xD1610003, 0, pair=0)
    assert launches0 == 0
    assert (ov1 == ov0).all() and (om1 == om0).all()
    assert (om1 == np.uint64(0xFFFFFFFFFFFFFFFF)).all()          # two-valued stimulus: every output lane is a valid 0 / 1


def test_automatic_choice_follows_the_level_width():
    """B2_OPT_PAIR_FUSION = 2 (the default): narrow levels are fused, wide ones keep the per-level kernels."""
    narrow = netgen.random_dag(n_gates=6000, depth=60, n_inputs=64, n_outputs=32, seed=2)        # 100 gates per level
    wide = netgen.random_dag(n_gates=100_000, depth=5, n_inputs=64, n_outputs=32, seed=2)        # 20 000 gates per level
    for gen, fused in ((narrow, True), (wide, False)):
        sim = gen.emit(d.SimulatorBuilder()).build()
        words = 64
        b = Batch(sim, gen.inputs, gen.outputs, 32, 2)
        ov = np.zeros((len(gen.outputs), words), np.uint64)
        om = np.zeros_like(ov)
        b.run_generated(11, 1, 0, words, 10_000, ov, om)
        assert b.wait() == SimulationRunResult.Ok
        assert (b.info.pair_launches > 0) == fused
        rv, rm = _oracle_rows(gen, gen.n_wires, gen.inputs, words, 11, 0, 1)
        assert (ov == rv[gen.outputs]).all() and (om == rm[gen.outputs]).all()
