"""Adders and subtractors (gsim add_add / add_sub): known answers against Python integers on the
CPU oracle, the engine against the oracle on the GPU, and a counter built from a register and an adder."""
import numpy as np
import pytest

import oracle

WIDTHS = [1, 4, 33, 64, 255]


def build(b):
    w = {}
    for n in WIDTHS:
        a, c, s, d = b.add_wire(n), b.add_wire(n), b.add_wire(n), b.add_wire(n)
        b.add_add(a, c, s)
        b.add_sub(a, c, d)
        w[n] = (a, c, s, d)
    return w


def stimuli(rng, n, count=6):
    full = (1 << n) - 1
    out = [(0, 0, full, full), (full, 1 & full, full, full), (full, full, full, full)]
    for _ in range(count):
        x = int.from_bytes(rng.bytes(32), "little") & full
        y = int.from_bytes(rng.bytes(32), "little") & full
        out.append((x, y, full, full))
    out.append((5 & full, 3 & full, full ^ 1, full))          # one High-Z / X bit in a: everything Undefined
    out.append((5 & full, 3 & full, full, full ^ (1 << (n - 1))))
    return out


def expect(x, y, ma, mb, n):
    full = (1 << n) - 1
    if ma != full or mb != full:
        return (full, 0), (full, 0)
    return ((x + y) & full, full), ((x - y) & full, full)


def test_add_sub_known_answers_and_validation():
    b = oracle.Builder()
    w8, w4 = b.add_wire(8), b.add_wire(4)
    with pytest.raises(oracle.OracleError, match="WireWidthMismatch"):
        b.add_add(w8, w4, w8)
    with pytest.raises(oracle.OracleError, match="WireWidthMismatch"):
        b.add_sub(w8, w8, w4)
    with pytest.raises(oracle.OracleError, match="InvalidWireId"):
        b.add_add(w8, 99, w8)
    b = oracle.Builder()
    w = build(b)
    sim = b.build()
    rng = np.random.default_rng(21)
    for n in WIDTHS:
        a, c, s, d = w[n]
        for x, y, ma, mb in stimuli(rng, n):
            sim.set_wire_drive(a, x, ma)
            sim.set_wire_drive(c, y, mb)
            assert sim.run_sim(10) == oracle.RUN_OK
            es, ed = expect(x, y, ma, mb, n)
            assert sim.get_wire_state(s) == es and sim.get_wire_state(d) == ed, (n, hex(x), hex(y))


@pytest.mark.gpu
@pytest.mark.parametrize("resident", [1, 0])
def test_add_sub_on_gpu_against_oracle(resident):
    from digilogic_b200 import AddComponentError, LogicState, SimulatorBuilder
    from digilogic_b200.gsim import ComponentError
    vb = SimulatorBuilder()
    w8, w4 = vb.add_wire(8), vb.add_wire(4)
    with pytest.raises(ComponentError) as e:
        vb.add_add(w8, w4, w8)
    assert e.value.error == AddComponentError.WireWidthMismatch
    ob, gb = oracle.Builder(), SimulatorBuilder()
    w = build(ob)
    build(gb)
    ref, sim = ob.build(), gb.build()
    sim.set_option(sim.OPT_RESIDENT_KERNEL, resident)
    rng = np.random.default_rng(22)
    for rnd in range(8):
        for n in WIDTHS:
            a, c, s, d = w[n]
            x, y, ma, mb = stimuli(rng, n, count=1)[3 if rnd < 6 else 4 + rnd % 2]
            ref.set_wire_drive(a, x, ma); ref.set_wire_drive(c, y, mb)
            sim.set_wire_drive(a, LogicState(x, ma)); sim.set_wire_drive(c, LogicState(y, mb))
        ms = (10, 0, 1, 10)[rnd % 4]
        assert int(sim.run_sim(ms)) == ref.run_sim(ms)
        for n in WIDTHS:
            for wire in w[n][2:]:
                assert sim.get_wire_state(wire) == LogicState(*ref.get_wire_state(wire)), (rnd, n)


def counter(b, bits=8):
    """q <= q + 1 on every rising clock edge while enable; `clr` forces the next value to 0 (AND with NOT clr)."""
    ck, en, clr, nclr = b.add_wire(), b.add_wire(), b.add_wire(), b.add_wire()
    one, q, inc, nclr_w, d = (b.add_wire(bits) for _ in range(5))
    b.add_not_gate(clr, nclr)
    b.add_merge([nclr] * bits, nclr_w)
    b.add_add(q, one, inc)
    b.add_and_gate([inc, nclr_w], d)
    b.add_register(d, q, en, ck)
    return ck, en, clr, one, q


def test_counter_of_register_and_adder_on_oracle():
    b = oracle.Builder()
    ck, en, clr, one, q = counter(b)
    sim = b.build()
    sim.set_wire_drive(one, 1, 0xFF); sim.set_wire_drive(en, 1, 1); sim.set_wire_drive(clr, 1, 1); sim.set_wire_drive(ck, 0, 1)
    assert sim.run_sim(100) == oracle.RUN_OK
    sim.set_wire_drive(ck, 1, 1); assert sim.run_sim(100) == oracle.RUN_OK     # loads 0
    sim.set_wire_drive(clr, 0, 1); sim.set_wire_drive(ck, 0, 1); assert sim.run_sim(100) == oracle.RUN_OK
    assert sim.get_wire_state(q) == (0, 0xFF)
    for n in range(1, 300):
        sim.set_wire_drive(ck, 1, 1); assert sim.run_sim(100) == oracle.RUN_OK
        sim.set_wire_drive(ck, 0, 1); assert sim.run_sim(100) == oracle.RUN_OK
        assert sim.get_wire_state(q) == (n & 0xFF, 0xFF), n


@pytest.mark.gpu
@pytest.mark.parametrize("resident", [1, 0])
def test_counter_on_gpu(resident):
    from digilogic_b200 import LogicState, SimulationRunResult, SimulatorBuilder
    b = SimulatorBuilder()
    ck, en, clr, one, q = counter(b)
    sim = b.build()
    sim.set_option(sim.OPT_RESIDENT_KERNEL, resident)
    hi, lo = LogicState(1, 1), LogicState(0, 1)
    sim.set_wire_drive(one, LogicState(1, 0xFF)); sim.set_wire_drive(en, hi); sim.set_wire_drive(clr, hi); sim.set_wire_drive(ck, lo)
    assert sim.run_sim(100) == SimulationRunResult.Ok
    sim.set_wire_drive(ck, hi); assert sim.run_sim(100) == SimulationRunResult.Ok
    sim.set_wire_drive(clr, lo); sim.set_wire_drive(ck, lo); assert sim.run_sim(100) == SimulationRunResult.Ok
    assert sim.get_wire_state(q) == LogicState(0, 0xFF)
    for n in range(1, 300):
        sim.set_wire_drive(ck, hi); assert sim.run_sim(100) == SimulationRunResult.Ok
        sim.set_wire_drive(ck, lo); assert sim.run_sim(100) == SimulationRunResult.Ok
        assert sim.get_wire_state(q) == LogicState(n & 0xFF, 0xFF), n


# ------------------------------------------------------------------------------------------------------------------
# multiplier, comparators, shifters (gsim add_mul / add_compare_* / add_*_shift)

WIDTHS2 = [1, 3, 8, 33, 64, 100]


def build2(b):
    w = {}
    for n in WIDTHS2:
        a, c, p = b.add_wire(n), b.add_wire(n), b.add_wire(n)
        amt = b.add_wire(min(8, max(1, n.bit_length())))
        cmps = [b.add_wire(1) for _ in range(10)]
        shl, lshr, ashr = b.add_wire(n), b.add_wire(n), b.add_wire(n)
        b.add_mul(a, c, p)
        for op in range(10):
            b.add_compare(op, a, c, cmps[op])
        b.add_left_shift(a, amt, shl)
        b.add_logical_right_shift(a, amt, lshr)
        b.add_arithmetic_right_shift(a, amt, ashr)
        w[n] = (a, c, amt, p, cmps, shl, lshr, ashr)
    return w


def _signed(x, n):
    return x - (1 << n) if x >> (n - 1) else x


def expect2(x, y, sh, n):
    full = (1 << n) - 1
    sx, sy = _signed(x, n), _signed(y, n)
    cmps = [x == y, x != y, x < y, x > y, x <= y, x >= y, sx < sy, sx > sy, sx <= sy, sx >= sy]
    return (x * y) & full, [int(v) for v in cmps], (x << sh) & full, x >> sh, (sx >> sh) & full


def test_mul_compare_shift_known_answers_and_validation():
    b = oracle.Builder()
    w8, w4, w1, w9 = b.add_wire(8), b.add_wire(4), b.add_wire(1), b.add_wire(9)
    with pytest.raises(oracle.OracleError, match="WireWidthMismatch"):
        b.add_mul(w8, w4, w8)
    with pytest.raises(oracle.OracleError, match="WireWidthMismatch"):
        b.add_compare(oracle.CMP_LTU, w8, w4, w1)
    with pytest.raises(oracle.OracleError, match="WireWidthIncompatible"):
        b.add_compare(oracle.CMP_EQ, w8, w8, w4)
    with pytest.raises(oracle.OracleError, match="WireWidthIncompatible"):
        b.add_left_shift(w8, w9, w8)
    with pytest.raises(oracle.OracleError, match="WireWidthMismatch"):
        b.add_logical_right_shift(w8, w4, w4)
    b = oracle.Builder()
    w = build2(b)
    sim = b.build()
    rng = np.random.default_rng(5)
    for n in WIDTHS2:
        a, c, amt, p, cmps, shl, lshr, ashr = w[n]
        full, aw = (1 << n) - 1, sim.get_wire_width(amt)
        for k in range(30):
            x = int.from_bytes(rng.bytes(16), "little") & full
            y = x if k % 5 == 0 else int.from_bytes(rng.bytes(16), "little") & full
            sh = int(rng.integers(0, 1 << aw))
            sim.set_wire_drive(a, x, full); sim.set_wire_drive(c, y, full); sim.set_wire_drive(amt, sh, (1 << aw) - 1)
            assert sim.run_sim(10) == oracle.RUN_OK
            ep, ec, el, er, ea = expect2(x, y, sh, n)
            assert sim.get_wire_state(p) == (ep, full), (n, x, y)
            assert [sim.get_wire_state(o) for o in cmps] == [(v, 1) for v in ec], (n, x, y)
            assert sim.get_wire_state(shl) == (el, full) and sim.get_wire_state(lshr) == (er, full) and sim.get_wire_state(ashr) == (ea, full)
        sim.set_wire_drive(amt, 0, (1 << aw) - 2 if aw > 1 else 0)          # an invalid amount bit: the shifts are Undefined
        sim.set_wire_drive(c, 1, full ^ 1)                                  # an invalid operand bit: product and comparisons
        assert sim.run_sim(10) == oracle.RUN_OK
        assert sim.get_wire_state(p) == (full, 0) and sim.get_wire_state(cmps[4]) == (1, 0) and sim.get_wire_state(ashr) == (full, 0)


@pytest.mark.gpu
@pytest.mark.parametrize("path", ["resident", "device", "host", "dense"])
def test_mul_compare_shift_on_gpu_against_oracle(path):
    """Every lane its own operands (4-state: some lanes carry an X or Z bit), all four device paths, against the scalar oracle."""
    from digilogic_b200 import AddComponentError, SimulatorBuilder
    from digilogic_b200.gsim import ComponentError
    from helpers import lane_mask
    vb = SimulatorBuilder()
    w8, w4, w1, w9 = vb.add_wire(8), vb.add_wire(4), vb.add_wire(1), vb.add_wire(9)
    for fn, err in [(lambda: vb.add_mul(w8, w4, w8), AddComponentError.WireWidthMismatch),
                    (lambda: vb.add_compare(2, w8, w8, w4), AddComponentError.WireWidthIncompatible),
                    (lambda: vb.add_left_shift(w8, w9, w8), AddComponentError.WireWidthIncompatible),
                    (lambda: vb.add_arithmetic_right_shift(w8, w4, w4), AddComponentError.WireWidthMismatch)]:
        with pytest.raises(ComponentError) as e:
            fn()
        assert e.value.error == err
    n_lanes = 64 + 37
    ob, gb = oracle.Builder(), SimulatorBuilder()
    w = build2(ob)
    build2(gb)
    lanes, sim = oracle.Lanes(ob, n_lanes), gb.build()
    sim.set_option(sim.OPT_RESIDENT_KERNEL, int(path == "resident"))
    sim.set_option(sim.OPT_DEVICE_LOOP, int(path in ("resident", "device")))
    sim.set_option(sim.OPT_FRONTIER, int(path != "dense"))
    sim.set_lanes(n_lanes)
    rng = np.random.default_rng(8)
    mask = lane_mask(n_lanes)
    words = lanes.words
    for rnd in range(3):
        for n in WIDTHS2:
            a, c, amt = w[n][:3]
            for wire, width in ((a, n), (c, n), (amt, min(8, max(1, n.bit_length())))):
                for bit in range(width):
                    v = rng.integers(0, 1 << 64, words, dtype=np.uint64)
                    bad = rng.integers(0, 1 << 64, words, dtype=np.uint64)
                    for _ in range(6 if width > 8 else 3):                   # rare invalid bits: most lanes stay fully valid
                        bad &= rng.integers(0, 1 << 64, words, dtype=np.uint64)
                    m = ~bad
                    lanes.set_drive(wire, v, m, bit)
                    sim.set_wire_drive_lanes(wire, v, m, bit=bit)
        status, _, _ = lanes.run(20)
        res, conflict, unsettled = sim.run_sim_lanes(20)
        assert (status == 0).all() and int(res) == 0
        valid_products = 0
        for n in WIDTHS2:
            a, c, amt, p, cmps, shl, lshr, ashr = w[n]
            for wire, width in [(p, n), (shl, n), (lshr, n), (ashr, n)] + [(o, 1) for o in cmps]:
                for bit in range(width):
                    rv, rm = lanes.get(wire, bit)
                    gv, gm = sim.get_wire_lanes(wire, bit=bit)
                    assert (((gv ^ rv) | (gm ^ rm)) & mask).max(initial=0) == 0, (rnd, n, wire, bit)
                    valid_products += int(rm[0] != 0) if wire == p else 0
        assert valid_products > 0
