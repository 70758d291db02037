"""Slices and merges (gsim add_slice / add_merge): validation with the last of the
reference's error variants (OffsetOutOfRange, crates/digilogic_gsim/src/lib.rs:55-66), known answers on the CPU
oracle, and the engine against the oracle on the GPU."""
import numpy as np
import pytest

import oracle

Z, X, L0, L1 = (0, 0), (1, 0), (0, 1), (1, 1)


def build(b):
    """A 40-bit bus cut into three pieces and put together again in another order, plus a slice that crosses
    an atom boundary and a chain slice -> gate -> merge."""
    w = {}
    w["bus"] = b.add_wire(40)
    w["lo"], w["mid"], w["hi"] = b.add_wire(8), b.add_wire(27), b.add_wire(5)
    b.add_slice(w["bus"], 0, w["lo"])
    b.add_slice(w["bus"], 8, w["mid"])      # bits 8..34: crosses the 32-bit atom boundary
    b.add_slice(w["bus"], 35, w["hi"])
    w["swapped"] = b.add_wire(40)
    b.add_merge([w["hi"], w["mid"], w["lo"]], w["swapped"])
    w["nlo"] = b.add_wire(8)
    b.add_not_gate(w["lo"], w["nlo"])
    w["pair"] = b.add_wire(16)
    b.add_merge([w["lo"], w["nlo"]], w["pair"])
    w["bit"] = b.add_wire(1)
    b.add_slice(w["pair"], 15, w["bit"])
    return w


def expected(value, valid):
    """Python model: High-Z / X bits of the bus read X downstream."""
    def cut(v, m, off, n):
        mask = (1 << n) - 1
        mm = (m >> off) & mask
        return ((v >> off) & mask) | (~mm & mask), mm
    lo, mid, hi = cut(value, valid, 0, 8), cut(value, valid, 8, 27), cut(value, valid, 35, 5)
    swapped = (hi[0] | (mid[0] << 5) | (lo[0] << 32), hi[1] | (mid[1] << 5) | (lo[1] << 32))
    nlo = ((~lo[0] & 0xFF) | (~lo[1] & 0xFF), lo[1])
    pair = (lo[0] | (nlo[0] << 8), lo[1] | (nlo[1] << 8))
    bit = ((pair[0] >> 15) & 1 | (~(pair[1] >> 15) & 1), (pair[1] >> 15) & 1)
    return {"lo": lo, "mid": mid, "hi": hi, "swapped": swapped, "nlo": nlo, "pair": pair, "bit": bit}


CASES = [(0x12345678AB, (1 << 40) - 1), (0xFFFFFFFFFF, 0xFF00FF00FF), (0, 0), (0xA5A5A5A5A5, 0x0F0F0F0F0F)]


def test_slice_merge_validation():
    for B in (oracle.Builder,):
        b = B()
        w8, w4, w12, w1 = b.add_wire(8), b.add_wire(4), b.add_wire(12), b.add_wire(1)
        with pytest.raises(oracle.OracleError, match="OffsetOutOfRange"):
            b.add_slice(w8, 5, w4)            # bits 5..8 of an 8-bit wire
        with pytest.raises(oracle.OracleError, match="OffsetOutOfRange"):
            b.add_slice(w4, 0, w8)            # output wider than the input
        with pytest.raises(oracle.OracleError, match="InvalidWireId"):
            b.add_slice(99, 0, w4)
        with pytest.raises(oracle.OracleError, match="WireWidthMismatch"):
            b.add_merge([w8, w1], w12)
        with pytest.raises(oracle.OracleError, match="TooFewInputs"):
            b.add_merge([], w12)
        assert b.add_slice(w8, 4, w4) == 0 and b.add_merge([w8, w4], w12) == 1


def test_slice_merge_known_answers_on_oracle():
    b = oracle.Builder()
    w = build(b)
    sim = b.build()
    for value, valid in CASES:
        sim.set_wire_drive(w["bus"], value, valid)
        assert sim.run_sim(10) == oracle.RUN_OK
        for name, exp in expected(value & (valid | ~valid), valid).items():
            assert sim.get_wire_state(w[name]) == exp, (hex(value), hex(valid), name)


def test_engine_builder_rejects_the_same_calls():
    from digilogic_b200 import AddComponentError, SimulatorBuilder
    from digilogic_b200.gsim import ComponentError
    b = SimulatorBuilder()
    w8, w4, w12, w1 = b.add_wire(8), b.add_wire(4), b.add_wire(12), b.add_wire(1)
    for fn, err in ((lambda: b.add_slice(w8, 5, w4), AddComponentError.OffsetOutOfRange),
                    (lambda: b.add_slice(w4, 0, w8), AddComponentError.OffsetOutOfRange),
                    (lambda: b.add_slice(99, 0, w4), AddComponentError.InvalidWireId),
                    (lambda: b.add_merge([w8, w1], w12), AddComponentError.WireWidthMismatch),
                    (lambda: b.add_merge([], w12), AddComponentError.TooFewInputs)):
        with pytest.raises(ComponentError) as e:
            fn()
        assert e.value.error == err
    assert b.add_slice(w8, 4, w4) == 0 and b.add_merge([w8, w4], w12) == 1


@pytest.mark.gpu
@pytest.mark.parametrize("resident", [1, 0])
def test_slice_merge_on_gpu_against_oracle(resident):
    from digilogic_b200 import LogicState, SimulationRunResult, SimulatorBuilder
    ob, gb = oracle.Builder(), SimulatorBuilder()
    w = build(ob)
    build(gb)
    ref, sim = ob.build(), gb.build()
    sim.set_option(sim.OPT_RESIDENT_KERNEL, resident)
    for max_steps in (10, 1, 0, 2, 10):          # small budgets: the pieces arrive one step apart
        for value, valid in CASES:
            ref.set_wire_drive(w["bus"], value, valid)
            sim.set_wire_drive(w["bus"], LogicState(value, valid))
            assert int(sim.run_sim(max_steps)) == ref.run_sim(max_steps)
            for name, wire in w.items():
                assert sim.get_wire_state(wire) == LogicState(*ref.get_wire_state(wire)), (max_steps, hex(value), name)


# ------------------------------------------------------------------------------- reduction ("horizontal") gates
from oracle import AND, NAND, NOR, OR, XNOR, XOR  # noqa: E402

H_WIDTHS = [1, 2, 3, 8, 40]
H_KINDS = [AND, OR, XOR, NAND, NOR, XNOR]


def reduce_model(kind, value, valid, width):
    """Kleene reduction over the bits (High-Z / X = unknown)."""
    bits = [(((value >> i) & 1) if (valid >> i) & 1 else None) for i in range(width)]
    base = {AND: AND, NAND: AND, OR: OR, NOR: OR, XOR: XOR, XNOR: XOR}[kind]
    if base == AND:
        r = 0 if 0 in bits else (None if None in bits else 1)
    elif base == OR:
        r = 1 if 1 in bits else (None if None in bits else 0)
    else:
        r = None if None in bits else sum(bits) & 1
    if r is not None and kind in (NAND, NOR, XNOR):
        r ^= 1
    return (1, 0) if r is None else (r, 1)


def build_reductions(b):
    ins, outs = [], {}
    for w in H_WIDTHS:
        i = b.add_wire(w)
        ins.append(i)
        for k in H_KINDS:
            o = b.add_wire(1)
            b.add_horizontal_gate(k, i, o)
            outs[(w, k)] = o
    return ins, outs


def reduction_stimuli(rng, width, n=12):
    full = (1 << width) - 1
    cases = [(0, full), (full, full), (0, 0), (full, 0)]
    for _ in range(n):
        v = int(rng.integers(0, 1 << 62)) & full
        m = full if rng.random() < 0.5 else (int(rng.integers(0, 1 << 62)) | int(rng.integers(0, 1 << 62))) & full
        cases.append((v, m))
    return cases


def test_horizontal_gates_on_oracle():
    b = oracle.Builder()
    w8, w2 = b.add_wire(8), b.add_wire(2)
    with pytest.raises(oracle.OracleError, match="WireWidthIncompatible"):
        b.add_horizontal_gate(AND, w8, w2)
    with pytest.raises(oracle.OracleError, match="InvalidWireId"):
        b.add_horizontal_gate(AND, 99, w2)
    b = oracle.Builder()
    ins, outs = build_reductions(b)
    sim = b.build()
    rng = np.random.default_rng(11)
    for w, i in zip(H_WIDTHS, ins):
        for v, m in reduction_stimuli(rng, w):
            sim.set_wire_drive(i, v, m)
            assert sim.run_sim(10) == oracle.RUN_OK
            for k in H_KINDS:
                assert sim.get_wire_state(outs[(w, k)]) == reduce_model(k, v, m, w), (w, k, hex(v), hex(m))


@pytest.mark.gpu
@pytest.mark.parametrize("resident", [1, 0])
def test_horizontal_gates_on_gpu_against_oracle(resident):
    from digilogic_b200 import LogicState, SimulatorBuilder
    ob, gb = oracle.Builder(), SimulatorBuilder()
    ins, outs = build_reductions(ob)
    build_reductions(gb)
    ref, sim = ob.build(), gb.build()
    sim.set_option(sim.OPT_RESIDENT_KERNEL, resident)
    rng = np.random.default_rng(12)
    for rnd in range(10):
        for w, i in zip(H_WIDTHS, ins):
            v, m = reduction_stimuli(rng, w, n=1)[-1]
            ref.set_wire_drive(i, v, m)
            sim.set_wire_drive(i, LogicState(v, m))
        ms = (10, 0, 1)[rnd % 3]
        assert int(sim.run_sim(ms)) == ref.run_sim(ms)
        for key, o in outs.items():
            assert sim.get_wire_state(o) == LogicState(*ref.get_wire_state(o)), (rnd, key)
