"""The CPU oracle against known answers (SURVEY.md A.3 truth tables, A.7 latch trace,
arithmetic circuits) and the two CPU models against each other."""
import itertools

import numpy as np
import pytest

import oracle
from oracle import AND, OR, XOR, NAND, NOR, XNOR, NOT

# (plane0=value, plane1=valid): reference crates/digilogic/src/ui/palette.rs:326-341
Z, X, L0, L1 = (0, 0), (1, 0), (0, 1), (1, 1)
NAMES = {Z: "Z", X: "X", L0: "0", L1: "1"}


def kleene(kind, a, b):
    """Independent truth-table definition of the 4-state gates (Z reads as X)."""
    def tri(s):
        return {L0: 0, L1: 1}.get(s, None)
    x, y = tri(a), tri(b)
    base = {AND: 0, NAND: 0, OR: 1, NOR: 1, XOR: 2, XNOR: 2}[kind]
    if base == 0:
        r = 0 if (x == 0 or y == 0) else (1 if (x == 1 and y == 1) else None)
    elif base == 1:
        r = 1 if (x == 1 or y == 1) else (0 if (x == 0 and y == 0) else None)
    else:
        r = None if (x is None or y is None) else x ^ y
    if kind in (NAND, NOR, XNOR) and r is not None:
        r ^= 1
    return X if r is None else (L1 if r else L0)


@pytest.mark.parametrize("kind", [AND, OR, XOR, NAND, NOR, XNOR])
def test_truth_tables_scalar_and_bitsliced(kind):
    import ctypes as C
    L = oracle.lib()
    for a, b in itertools.product([Z, X, L0, L1], repeat=2):
        rs, rv = C.c_uint32(), C.c_uint32()
        L.go_logic_op(kind, a[0], a[1], b[0], b[1], C.byref(rs), C.byref(rv))
        got = (rs.value & 1, rv.value & 1)
        assert got == kleene(kind, a, b), (kind, NAMES[a], NAMES[b])
        bv, bm = C.c_uint64(), C.c_uint64()
        L.bs_logic_op(kind, a[0], a[1], b[0], b[1], C.byref(bv), C.byref(bm))
        assert (bv.value & 1, bm.value & 1) == got


def test_not_truth_table():
    for a, exp in [(Z, X), (X, X), (L0, L1), (L1, L0)]:
        b = oracle.Builder()
        i, o = b.add_wire(), b.add_wire()
        b.add_not_gate(i, o)
        s = b.build()
        s.set_wire_drive(i, *a)
        assert s.run_sim(10) == oracle.RUN_OK
        assert s.get_wire_state(o) == exp


def test_builder_validation():
    b = oracle.Builder()
    w1, w1b, w4 = b.add_wire(1), b.add_wire(1), b.add_wire(4)
    with pytest.raises(oracle.OracleError, match="InvalidWireId"):
        b.add_and_gate([w1, 99], w1b)
    with pytest.raises(oracle.OracleError, match="TooFewInputs"):
        b.add_and_gate([w1], w1b)
    with pytest.raises(oracle.OracleError, match="WireWidthMismatch"):
        b.add_and_gate([w1, w4], w1b)
    with pytest.raises(oracle.OracleError, match="InvalidInputCount"):
        b.add_multiplexer([w1, w1b, w1], w1, w1b)
    with pytest.raises(oracle.OracleError, match="WireWidthIncompatible"):
        b.add_multiplexer([w1, w1b], w4, w1b)
    assert b.get_wire_width(w4) == 4


def _latch():
    b = oracle.Builder()
    s_, r_, q, qn = (b.add_wire() for _ in range(4))
    b.add_nand_gate([s_, qn], q)
    b.add_nand_gate([r_, q], qn)
    return b, (s_, r_, q, qn)


# (S', R', max_steps, result, steps, Q, Q')
A7 = [
    (L1, L1, 10000, oracle.RUN_OK, 1, X, X),
    (L0, L1, 10000, oracle.RUN_OK, 2, L1, L0),
    (L1, L1, 10000, oracle.RUN_OK, 0, L1, L0),
    (L1, L0, 10000, oracle.RUN_OK, 2, L0, L1),
    (L0, L0, 10000, oracle.RUN_OK, 1, L1, L1),
    (L1, L1, 10, oracle.RUN_MAX_STEPS, 11, L0, L0),
    (L1, L1, 11, oracle.RUN_MAX_STEPS, 12, L1, L1),
    (L0, L1, 10000, oracle.RUN_OK, 2, L1, L0),
    (X, L1, 10000, oracle.RUN_OK, 0, L1, L0),
    (Z, L0, 10000, oracle.RUN_OK, 2, X, L1),
]


def test_latch_trace_scalar():
    b, (s_, r_, q, qn) = _latch()
    sim = b.build()
    for i, (sd, rd, ms, res, steps, eq, eqn) in enumerate(A7):
        sim.set_wire_drive(s_, *sd)
        sim.set_wire_drive(r_, *rd)
        assert sim.run_sim(ms) == res, i
        assert sim.last_steps == steps, i
        assert sim.get_wire_state(q) == eq and sim.get_wire_state(qn) == eqn, i


def test_latch_trace_bitsliced():
    b, (s_, r_, q, qn) = _latch()
    bs = oracle.BitSliced(b, 64, 4)
    full = lambda s: (np.array([(1 << 64) - 1 if s[0] else 0], np.uint64), np.array([(1 << 64) - 1 if s[1] else 0], np.uint64))
    for i, (sd, rd, ms, res, steps, eq, eqn) in enumerate(A7):
        bs.set_drive(s_, *full(sd))
        bs.set_drive(r_, *full(rd))
        conflict, unsettled = bs.run(ms)
        st = oracle.status_from_masks(conflict, unsettled, 64)
        assert (st == res).all(), i
        for wire, exp in ((q, eq), (qn, eqn)):
            v, m = bs.get(wire)
            assert (int(v[0]) & 1, int(m[0]) & 1) == exp, i


def test_driver_conflict():
    b = oracle.Builder()
    a, c, o = b.add_wire(), b.add_wire(), b.add_wire()
    b.add_and_gate([a, c], o)
    sim = b.build()
    sim.set_wire_drive(a, *L1); sim.set_wire_drive(c, *L1)
    assert sim.run_sim(100) == oracle.RUN_OK
    sim.set_wire_drive(o, *L1)  # agreeing value still conflicts (A.6-2)
    assert sim.run_sim(100) == oracle.RUN_CONFLICT
    b2 = oracle.Builder()
    a, c, o = b2.add_wire(), b2.add_wire(), b2.add_wire()
    b2.add_and_gate([a, c], o); b2.add_or_gate([a, c], o)
    assert b2.build().run_sim(100) == oracle.RUN_CONFLICT


def test_wide_wires_and_mux():
    b = oracle.Builder()
    a, c, o = b.add_wire(40), b.add_wire(40), b.add_wire(40)
    b.add_xor_gate([a, c], o)
    sel, m = b.add_wire(2), b.add_wire(40)
    d = [b.add_wire(40) for _ in range(4)]
    b.add_multiplexer(d, sel, m)
    sim = b.build()
    full = (1 << 40) - 1
    sim.set_wire_drive(a, 0x12345678AB, full); sim.set_wire_drive(c, 0x0F0F0F0F0F, full)
    for i, w in enumerate(d):
        sim.set_wire_drive(w, 0x1111111111 * (i + 1), full)
    sim.set_wire_drive(sel, 2, 3)
    assert sim.run_sim(10) == oracle.RUN_OK
    assert sim.get_wire_state(o) == (0x12345678AB ^ 0x0F0F0F0F0F, full)
    assert sim.get_wire_state(m) == (0x3333333333, full)
    sim.set_wire_drive(sel, 2, 1)  # one select bit invalid -> all X
    assert sim.run_sim(10) == oracle.RUN_OK
    assert sim.get_wire_state(m) == (full, 0)
    sim.set_wire_drive(sel, 1, 3)
    sim.set_wire_drive(d[1], 0, 0)  # selected input High-Z reads X (A.6-4)
    assert sim.run_sim(10) == oracle.RUN_OK
    assert sim.get_wire_state(m) == (full, 0)


def test_nary_gate_fold():
    for kind, fn in ((AND, lambda xs: all(xs)), (NOR, lambda xs: not any(xs)), (XNOR, lambda xs: (sum(xs) & 1) == 0)):
        b = oracle.Builder()
        ins = [b.add_wire() for _ in range(5)]
        o = b.add_wire()
        b.add_gate(kind, ins, o)
        sim = b.build()
        for bits in itertools.product([0, 1], repeat=5):
            for w, v in zip(ins, bits):
                sim.set_wire_drive(w, v, 1)
            assert sim.run_sim(10) == oracle.RUN_OK
            assert sim.get_wire_state(o) == (int(fn(bits)), 1)


def _random_netlist(rng, n_in, n_gates, cyclic, wide=False):
    """Random netlist; `cyclic` lets gate inputs come from any wire (feedback)."""
    b = oracle.Builder()
    first_in = b.add_wires(n_in)
    first_g = b.add_wires(n_gates)
    n_w = n_in + n_gates
    for g in range(n_gates):
        kind = int(rng.integers(0, 7))
        hi = n_w if cyclic else n_in + g
        if kind == NOT:
            b.add_not_gate(int(rng.integers(0, hi)), first_g + g)
        else:
            n = int(rng.integers(2, 5)) if (wide and rng.random() < 0.2) else 2
            b.add_gate(kind, [int(x) for x in rng.integers(0, hi, n)], first_g + g)
    return b, first_in, n_w


def _random_drive(rng, words, four_state):
    value = rng.integers(0, 1 << 64, words, dtype=np.uint64)
    valid = (rng.integers(0, 1 << 64, words, dtype=np.uint64) | rng.integers(0, 1 << 64, words, dtype=np.uint64)) if four_state \
        else np.full(words, (1 << 64) - 1, np.uint64)
    return value, valid


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("cyclic", [False, True])
def test_scalar_vs_bitsliced_random(seed, cyclic):
    """Event-driven scalar restatement == dense bit-sliced model: status, every wire, every
    lane, across warm restarts and a sweep of max_steps (the step-count contract)."""
    rng = np.random.default_rng(1000 * seed + cyclic)
    n_in, n_gates, n_lanes = 6, 40, 128
    b, first_in, n_w = _random_netlist(rng, n_in, n_gates, cyclic, wide=True)
    lanes = oracle.Lanes(b, n_lanes)
    bs = oracle.BitSliced(b, n_lanes, n_w)
    for rnd, max_steps in enumerate([0, 1, 2, 3, 5, 8, 50, 7, 0, 1000]):
        for i in range(n_in):
            if rng.random() < 0.7:
                v, m = _random_drive(rng, lanes.words, four_state=True)
                lanes.set_drive(first_in + i, v, m)
                bs.set_drive(first_in + i, v, m)
        status, steps, _ = lanes.run(max_steps)
        conflict, unsettled = bs.run(max_steps)
        assert (oracle.status_from_masks(conflict, unsettled, n_lanes) == status).all(), (rnd, max_steps)
        assert not (status == oracle.RUN_CONFLICT).any()
        sv, sm = lanes.get_all(n_w)
        dv, dm = bs.get_all()
        assert (sv == dv).all() and (sm == dm).all(), (rnd, max_steps)


def test_bitsliced_dag_shortcut():
    rng = np.random.default_rng(7)
    b, first_in, n_w = _random_netlist(rng, 8, 300, cyclic=False)
    bs1 = oracle.BitSliced(b, 256, n_w)
    bs2 = oracle.BitSliced(b, 256, n_w)
    assert bs1.topological
    for i in range(8):
        v, m = _random_drive(rng, bs1.words, four_state=True)
        bs1.set_drive(first_in + i, v, m); bs2.set_drive(first_in + i, v, m)
    c, u = bs1.run(10_000)
    assert not c.any() and not u.any()
    assert not bs2.eval_dag().any()
    a, b_ = bs1.get_all(), bs2.get_all()
    assert (a[0] == b_[0]).all() and (a[1] == b_[1]).all()


def test_conflict_lanes_agree():
    """A driven gate output conflicts exactly in the lanes where the drive is not High-Z."""
    b = oracle.Builder()
    a, c, o = b.add_wire(), b.add_wire(), b.add_wire()
    b.add_and_gate([a, c], o)
    lanes = oracle.Lanes(b, 64)
    bs = oracle.BitSliced(b, 64, 3)
    ones = np.array([(1 << 64) - 1], np.uint64)
    for sim in (lanes, bs):
        sim.set_drive(a, ones, ones); sim.set_drive(c, ones, ones)
        sim.set_drive(o, np.array([0x00FF], np.uint64), np.array([0x0F0F], np.uint64))
    status, _, _ = lanes.run(100)
    conflict, unsettled = bs.run(100)
    assert int(conflict[0]) == 0x0FFF and not unsettled.any()
    assert (oracle.status_from_masks(conflict, unsettled, 64) == status).all()
