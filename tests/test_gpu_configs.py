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
