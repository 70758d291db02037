"""GPU parity on the BASELINE configurations that are not covered by tests/test_fixtures.py:
config 3 at full size (sampled against the bit-sliced oracle + size-independent properties) and
config 4 (sequential: flip-flop LFSR chains and SR latches, feedback settle, oscillating lanes)."""
import numpy as np
import pytest

import oracle
from digilogic_b200 import SimulationRunResult, SimulatorBuilder, netgen
from helpers import lane_mask

pytestmark = pytest.mark.gpu

ALL1 = np.uint64(0xFFFFFFFFFFFFFFFF)


def test_config3_full_size_sampled_against_oracle():
    """1M gates, depth 200, 4096 inputs.  The engine settles a 512-word tile whose lane-words are
    words [word_base, word_base+512) of the 2^24-lane batch; 3 of those lane-words are re-computed
    by the CPU bit-sliced model over ALL 1 004 096 wires (two-valued and 4-state stimulus)."""
    gen = netgen.random_dag()
    sim = gen.emit(SimulatorBuilder()).build()
    info = sim.info
    assert (info.n_gates2, info.depth, info.n_sources) == (1_000_000, 200, 4096)
    tile, word_base = 512, 262144 - 512      # the last tile of the batch
    sim.set_lanes(tile * 64)
    ob = gen.emit(oracle.Builder())
    sample = [0, 257, 511]
    for mode in (0, 1):
        sim.fill_stimulus(gen.inputs, seed=0xD1610003, lane_word_base=word_base, mode=mode)
        res, conflict, unsettled = sim.run_sim_lanes(10_000)
        assert res == SimulationRunResult.Ok and not conflict.any() and not unsettled.any()
        assert sim.info.last_mode == 1
        bs = oracle.BitSliced(ob, 64 * len(sample), gen.n_wires)
        v, m = netgen.stimulus_words(0xD1610003, len(gen.inputs), tile, word_base, mode)
        for i, w in enumerate(gen.inputs):
            bs.set_drive(int(w), v[i, sample], m[i, sample])
        assert not bs.eval_dag().any()
        rv, rm = bs.get_all()
        # every 61st wire of the netlist (all levels), all sampled lane-words
        some = np.arange(0, gen.n_wires, 61, dtype=np.uint32)
        gv, gm = sim.gather_wires_host(some)
        assert (gv[:, sample] == rv[some]).all() and (gm[:, sample] == rm[some]).all(), mode
        # all designated outputs and the whole last level, all sampled words, in one gather
        last_level = np.arange(gen.n_wires - 5000, gen.n_wires, dtype=np.uint32)
        gv, gm = sim.gather_wires_host(last_level)
        assert (gv[:, sample] == rv[last_level]).all() and (gm[:, sample] == rm[last_level]).all()
        if mode == 0:
            assert (gm == ALL1).all(), "two-valued stimulus must give valid outputs on every lane"
        else:
            assert (gm != ALL1).any(), "4-state stimulus must leave some X outputs"


def test_config3_tiling_invariance():
    """Size-independent property: the output planes of a batch do not depend on the tile size
    (lane-words are independent simulators and the stimulus is a function of the global index)."""
    import torch
    from digilogic_b200.batch import BatchRunner
    gen = netgen.random_dag(n_gates=100_000, depth=50, n_inputs=512, n_outputs=256)
    words = 2048
    outs = []
    for tile in (2048, 512, 128):
        sim = gen.emit(SimulatorBuilder()).build()
        sim.set_stream(torch.cuda.current_stream().cuda_stream)
        runner = BatchRunner(sim, gen.inputs, gen.outputs, tile)
        ov = torch.zeros((len(gen.outputs), words), dtype=torch.int64, device="cuda")
        om = torch.zeros_like(ov)
        worst = runner.run_generated(1000, words, 0xD1610003, 1, ov, om)
        assert worst == SimulationRunResult.Ok
        torch.cuda.synchronize()
        outs.append((ov.cpu(), om.cpu()))
    for ov, om in outs[1:]:
        assert torch.equal(ov, outs[0][0]) and torch.equal(om, outs[0][1])


def _drive_all(sims, wire, value, valid):
    for s in sims:
        if isinstance(s, oracle.Lanes):
            s.set_drive(wire, value, valid)
        else:
            s.set_wire_drive_lanes(wire, value, valid)


@pytest.mark.parametrize("n_hazard,max_steps", [(0, 10_000), (3, 60)])
@pytest.mark.parametrize("resident", [True, False, "host"])
def test_config4_sequential_vs_oracle(n_hazard, max_steps, resident):
    """Flip-flop LFSR chains + SR latches at reduced size: per-lane init pattern, then clocked
    steps (clk up, clk down = two run_sim calls).  After every call: per-lane result
    (Ok / MaxStepsReached) and every wire of every lane equal the scalar oracle.  With hazard
    latches some lanes oscillate (S' = R' rise together) and must end MaxStepsReached with the
    exact state gsim would leave after max_steps + 1 steps."""
    n_lanes = 192
    seq = netgen.sequential(n_latches=24, n_chains=4, bits=8, n_hazard=n_hazard)
    ob = seq.emit(oracle.Builder())
    sim = seq.emit(SimulatorBuilder()).build()
    assert sim.info.cyclic == 1 and sim.info.n_gates3 > 0
    sim.set_option(sim.OPT_RESIDENT_KERNEL, int(resident is True))
    sim.set_option(sim.OPT_DEVICE_LOOP, int(resident != "host"))
    sim.set_lanes(n_lanes)
    lanes = oracle.Lanes(ob, n_lanes)
    both = (lanes, sim)
    words = lanes.words
    rng = np.random.default_rng(4 + n_hazard)
    ones = np.full(words, ALL1, np.uint64)
    zeros = np.zeros(words, np.uint64)
    mask = lane_mask(n_lanes)
    seen_unsettled = False

    def run_and_compare(tag):
        nonlocal seen_unsettled
        status, _, _ = lanes.run(max_steps)
        res, conflict, unsettled = sim.run_sim_lanes(max_steps)
        got = oracle.status_from_masks(conflict, unsettled, n_lanes)
        assert (got == status).all(), tag
        assert int(res) == int(status.max()), tag
        seen_unsettled |= bool((status == 1).any())
        rv, rm = lanes.get_all(seq.n_wires)
        gv, gm = sim.gather_wires_host(np.arange(seq.n_wires, dtype=np.uint32))
        assert (((gv ^ rv) | (gm ^ rm)) & mask).max() == 0, tag

    _drive_all(both, seq.clk, zeros, ones)
    for b in range(8):
        pattern = rng.integers(0, 1 << 64, words, dtype=np.uint64)
        _drive_all(both, seq.preset_n[b], ~pattern, ones)   # bit set   -> preset' = 0
        _drive_all(both, seq.clear_n[b], pattern, ones)     # bit clear -> clear'  = 0
    run_and_compare("init")
    for b in range(8):
        _drive_all(both, seq.preset_n[b], ones, ones)
        _drive_all(both, seq.clear_n[b], ones, ones)
    run_and_compare("release")
    for step in range(24):
        _drive_all(both, seq.clk, ones, ones)
        run_and_compare(("up", step))
        _drive_all(both, seq.clk, zeros, ones)
        run_and_compare(("down", step))
    assert seen_unsettled == (n_hazard > 0)
    assert sim.info.last_mode == (3 if resident is True else 2)


@pytest.mark.parametrize("device_loop", [1, 0])
def test_config4_full_size_with_hazard_latches_vs_oracle(device_loop):
    """BASELINE config 4 as SURVEY.md 8d specifies it, at full netlist size: 64 chains x 64 NAND flip-flops + 4096 SR latches,
    8 of them hazard latches (S' and R' can rise together: unit-delay oscillation until the step budget), 65 536 lanes with a
    per-lane preset pattern, run_sim(10_000) per clock edge.  After every call: the per-lane result (Ok / MaxStepsReached) and
    EVERY wire of 128 sampled lanes (two lane-words) equal the scalar oracle, which walks all 10 001 steps of an oscillating
    lane; over all lanes some must oscillate.  device_loop = 0 is the control that walks every step on the GPU too."""
    n_lanes, bits, max_steps = 65_536, 64, 10_000
    sample = [0, 517]                                   # lane-words compared wire for wire
    seq = netgen.sequential(n_latches=4096, n_chains=64, bits=bits, n_hazard=8)
    sim = seq.emit(SimulatorBuilder()).build()
    sim.set_option(sim.OPT_DEVICE_LOOP, device_loop)
    sim.set_lanes(n_lanes)
    lanes = oracle.Lanes(seq.emit(oracle.Builder()), 64 * len(sample))
    words = n_lanes // 64
    rng = np.random.default_rng(44)
    ones, zeros = np.full(words, ALL1, np.uint64), np.zeros(words, np.uint64)
    all_wires = np.arange(seq.n_wires, dtype=np.uint32)
    unsettled_lanes = 0

    def drive(wire, value, valid):
        sim.set_wire_drive_lanes(wire, value, valid)
        lanes.set_drive(wire, value[sample], valid[sample])

    def run_and_compare(tag):
        nonlocal unsettled_lanes
        status, _, _ = lanes.run(max_steps, threads=8)
        res, conflict, unsettled = sim.run_sim_lanes(max_steps)
        assert not conflict.any(), tag
        got = oracle.status_from_masks(conflict[sample], unsettled[sample], 64 * len(sample))
        assert (got == status).all(), tag
        unsettled_lanes += int(sum(bin(int(x)).count("1") for x in unsettled))
        assert int(res) == (1 if unsettled.any() else 0), tag
        rv, rm = lanes.get_all(seq.n_wires)
        gv, gm = sim.gather_wires_host(all_wires)
        assert (gv[:, sample] == rv).all() and (gm[:, sample] == rm).all(), tag

    drive(seq.clk, zeros, ones)
    for b in range(bits):
        pattern = rng.integers(0, 1 << 64, words, dtype=np.uint64)
        drive(seq.preset_n[b], ~pattern, ones)
        drive(seq.clear_n[b], pattern, ones)
    run_and_compare("init")
    for b in range(bits):
        drive(seq.preset_n[b], ones, ones)
        drive(seq.clear_n[b], ones, ones)
    run_and_compare("release")
    for step in range(3):                               # 6 run_sim(10_000) calls
        drive(seq.clk, ones, ones)
        run_and_compare(("up", step))
        drive(seq.clk, zeros, ones)
        run_and_compare(("down", step))
    assert unsettled_lanes > 0, "no lane oscillated: the hazard latches were not exercised"


def test_config5_netlist_sampled_against_oracle():
    """10M gates, depth 200, 16 384 inputs (BASELINE config 5): one 32-word tile from the end of the
    2^26-lane batch, two of its lane-words re-computed by the CPU bit-sliced model — every 997th wire
    of the netlist and the whole block of designated outputs."""
    gen = netgen.random_dag(n_gates=10_000_000, depth=200, n_inputs=16384, n_outputs=1024, seed=0xD1610005)
    sim = gen.emit(SimulatorBuilder()).build()
    info = sim.info
    assert (info.n_gates2, info.depth, info.n_sources) == (10_000_000, 200, 16384)
    tile, word_base = 32, (1 << 20) - 32
    sim.set_lanes(tile * 64)
    sim.fill_stimulus(gen.inputs, seed=0xD1610005, lane_word_base=word_base, mode=1)
    res, conflict, unsettled = sim.run_sim_lanes(10_000)
    assert res == SimulationRunResult.Ok and not conflict.any() and not unsettled.any()
    sample = [3, 31]
    bs = oracle.BitSliced(gen.emit(oracle.Builder()), 64 * len(sample), gen.n_wires)
    v, m = netgen.stimulus_words(0xD1610005, len(gen.inputs), tile, word_base, 1)
    for i, w in enumerate(gen.inputs):
        bs.set_drive(int(w), v[i, sample], m[i, sample])
    assert not bs.eval_dag().any()
    some = np.concatenate([np.arange(0, gen.n_wires, 997, dtype=np.uint32), gen.outputs])
    gv, gm = sim.gather_wires_host(some)
    ref_v = np.stack([bs.get(int(w))[0] for w in some])
    ref_m = np.stack([bs.get(int(w))[1] for w in some])
    assert (gv[:, sample] == ref_v).all() and (gm[:, sample] == ref_m).all()


def test_config4_full_netlist_lfsr_property():
    """BASELINE config 4 at its full netlist size (64 chains x 64 NAND flip-flops + 4096 latches) and lane count: after a
    few clocked steps every sampled lane's LFSR state equals a software LFSR of its start pattern (size-independent
    property; the wire-by-wire oracle comparison runs at reduced size above)."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "benchmarks"))
    import configs
    bits = chains = 64
    n_lanes, clocks = 65536, 12
    seq = netgen.sequential(n_latches=4096, n_chains=chains, bits=bits)
    taps = [bits - 1, bits - 2, bits - 4, bits - 5]
    sim = seq.emit(SimulatorBuilder()).build()
    sim.set_lanes(n_lanes)
    words = n_lanes // 64
    ones, zeros = np.full(words, ALL1, np.uint64), np.zeros(words, np.uint64)
    rng = np.random.default_rng(4)
    patterns = [rng.integers(0, 1 << 64, words, dtype=np.uint64) for _ in range(bits)]

    def run():
        res, _, _ = sim.run_sim_lanes(10_000, masks=False)
        assert res == SimulationRunResult.Ok

    sim.set_wire_drive_lanes(seq.clk, zeros, ones)
    for b in range(bits):
        sim.set_wire_drive_lanes(seq.preset_n[b], ~patterns[b], ones)
        sim.set_wire_drive_lanes(seq.clear_n[b], patterns[b], ones)
    run()
    for b in range(bits):
        sim.set_wire_drive_lanes(seq.preset_n[b], ones, ones)
        sim.set_wire_drive_lanes(seq.clear_n[b], ones, ones)
    run()
    for _ in range(clocks):
        sim.set_wire_drive_lanes(seq.clk, ones, ones); run()
        sim.set_wire_drive_lanes(seq.clk, zeros, ones); run()
    assert sim.info.last_mode == 2
    sample = np.concatenate([np.arange(128), np.arange(n_lanes - 128, n_lanes)])
    pat_bits = [configs.unpack(p, n_lanes) for p in patterns]
    for c in range(0, chains, 7):
        start = np.zeros(len(sample), np.uint64)
        got = np.zeros(len(sample), np.uint64)
        for b in range(bits):
            start |= pat_bits[(b + c) % bits][sample].astype(np.uint64) << np.uint64(b)
            got |= configs.unpack(sim.get_wire_lanes(seq.q[c][b])[0], n_lanes)[sample].astype(np.uint64) << np.uint64(b)
        assert (configs.software_lfsr(start, bits, taps, clocks) == got).all(), c


@pytest.mark.parametrize("kind", ["gates", "registers", "buses"])
def test_unit_delay_paths_agree_at_scale(kind):
    """The three HBM unit-delay paths (device-side loop, host loop with work lists, dense) on one netlist of a few
    thousand components and 19 200 lanes — beyond what the CPU models check quickly — must leave every wire of every lane
    and both status masks identical, run after run."""
    from helpers import bus_enable_drives, random_bus_netlist, random_drive, random_netlist, random_register_netlist
    rng = np.random.default_rng({"gates": 1, "registers": 2, "buses": 3}[kind])
    n_in, n_gates, n_lanes = 16, 3000, 64 * 300
    groups = []
    if kind == "gates":
        nl, first_in = random_netlist(rng, n_in, n_gates, cyclic=True, wide=True, mux=True)
    elif kind == "registers":
        nl, first_in = random_register_netlist(rng, n_in, n_gates)
    else:
        nl, first_in, groups = random_bus_netlist(rng, n_in, n_gates, 40)
    sims = []
    for frontier, device_loop in ((1, 1), (1, 0), (0, 0)):
        s = nl.replay(SimulatorBuilder()).build()
        s.set_option(s.OPT_RESIDENT_KERNEL, 0)
        s.set_option(s.OPT_FRONTIER, frontier)
        s.set_option(s.OPT_DEVICE_LOOP, device_loop)
        s.set_lanes(n_lanes)
        sims.append(s)
    words = n_lanes // 64
    wires = np.arange(nl.n_wires, dtype=np.uint32)
    dirty = np.zeros(words, np.uint64)
    for rnd, max_steps in enumerate([0, 3, 40, 7, 1, 500, 12]):
        for i in range(n_in):
            if rng.random() < 0.6:
                v, m = random_drive(rng, words, four_state=rnd % 3 == 0)
                for s in sims:
                    s.set_wire_drive_lanes(first_in + i, v, m)
        for grp in groups:
            for e, (v, m) in zip(grp, bus_enable_drives(rng, grp, words, p_bad=0.002)):
                for s in sims:
                    s.set_wire_drive_lanes(e, v, m)
        outs = []
        for s in sims:
            res, conflict, unsettled = s.run_sim_lanes(max_steps)
            gv, gm = s.gather_wires_host(wires)
            outs.append((int(res), conflict, unsettled, gv, gm))
        keep = ~(dirty | outs[0][1])        # lanes that conflicted are outside the contract from then on
        for o in outs[1:]:
            assert ((o[1] ^ outs[0][1]) & ~dirty).max() == 0 and ((o[2] ^ outs[0][2]) & keep).max() == 0, (rnd, max_steps)
            assert ((o[3] ^ outs[0][3]) & keep).max() == 0 and ((o[4] ^ outs[0][4]) & keep).max() == 0, (rnd, max_steps)
        dirty |= outs[0][1]
    assert (~dirty).any()
