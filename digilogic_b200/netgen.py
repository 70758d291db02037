"""Generators for the benchmark netlists of BASELINE.json (none exist in the reference:
SURVEY.md Appendix C).  Every generator emits calls on any object with the gsim
SimulatorBuilder method names (the engine's builder or the test oracle's)."""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

AND, OR, XOR, NAND, NOR, XNOR, NOT, MUX = range(8)

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x):
    """SplitMix64 finaliser on uint64 arrays (same function as the device stimulus generator)."""
    with np.errstate(over="ignore"):
        x = np.asarray(x, np.uint64) + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def stimulus_words(seed, n_inputs, words, word_base=0, mode=0):
    """Host restatement of b2_fill_stimulus: value/valid words [n_inputs, words]."""
    with np.errstate(over="ignore"):
        i = (np.arange(n_inputs, dtype=np.uint64)[:, None] + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        w = (np.arange(words, dtype=np.uint64)[None, :] + np.uint64(word_base)) * np.uint64(0xD1B54A32D192ED03)
        value = splitmix64(np.uint64(seed) ^ i ^ w)
        if mode == 0:
            valid = np.full_like(value, _M64)
        else:
            valid = splitmix64(value ^ np.uint64(1)) | splitmix64(value ^ np.uint64(2))
    return value, valid


@dataclass
class GeneratedNetlist:
    name: str
    n_wires: int
    n_gates: int
    inputs: np.ndarray            # wire ids of the primary inputs
    outputs: np.ndarray           # designated output wire ids
    depth: int = 0
    kinds: np.ndarray | None = None   # for 2-input netlists: SoA description
    in0: np.ndarray | None = None
    in1: np.ndarray | None = None
    out: np.ndarray | None = None
    extra: dict = field(default_factory=dict)

    def emit(self, builder):
        """Adds the wires and gates to `builder` (bulk path)."""
        first = builder.add_wires(self.n_wires)
        assert first == 0, "generators expect an empty builder"
        builder.add_gates2(self.kinds, self.in0, self.in1, self.out)
        return builder


def random_dag(n_gates=1_000_000, depth=200, n_inputs=4096, n_outputs=1024, seed=0xD1610003) -> GeneratedNetlist:
    """BASELINE config 3/5: `depth` levels of n_gates/depth two-input gates;
    input 0 uniform from the previous level (level 1: the primary inputs), which forces the
    depth; input 1 uniform from all earlier levels and the inputs; kind uniform over
    And/Or/Xor/Nand/Nor/Xnor.  Outputs: the first n_outputs wires of the last level.
    Wire ids: inputs 0..n_inputs-1, gate g drives wire n_inputs+g.  Randomness: SplitMix64 of
    (seed, 3*g + {0,1,2})."""
    assert n_gates % depth == 0
    per = n_gates // depth
    g = np.arange(n_gates, dtype=np.uint64)
    level = (g // np.uint64(per)).astype(np.int64)            # 0-based
    with np.errstate(over="ignore"):
        base = np.uint64(seed) * np.uint64(0x2545F4914F6CDD1D)
        r0 = splitmix64(base + g * np.uint64(3))
        r1 = splitmix64(base + g * np.uint64(3) + np.uint64(1))
        r2 = splitmix64(base + g * np.uint64(3) + np.uint64(2))
    prev_base = np.where(level == 0, 0, n_inputs + (level - 1) * per).astype(np.uint64)
    prev_cnt = np.where(level == 0, n_inputs, per).astype(np.uint64)
    in0 = (prev_base + r0 % prev_cnt).astype(np.uint32)
    earlier = (n_inputs + level * per).astype(np.uint64)
    in1 = (r1 % earlier).astype(np.uint32)
    kinds = (r2 % np.uint64(6)).astype(np.uint8)
    out = (np.uint64(n_inputs) + g).astype(np.uint32)
    outputs = np.arange(n_inputs + (depth - 1) * per, n_inputs + (depth - 1) * per + min(n_outputs, per), dtype=np.uint32)
    return GeneratedNetlist(
        name=f"random_dag_g{n_gates}_d{depth}_i{n_inputs}", n_wires=n_inputs + n_gates, n_gates=n_gates,
        inputs=np.arange(n_inputs, dtype=np.uint32), outputs=outputs, depth=depth,
        kinds=kinds, in0=in0, in1=in1, out=out, extra={"seed": seed, "per_level": per})


def ripple_adder(nbits=4) -> GeneratedNetlist:
    """BASELINE config 1 as a gate list: nbits full adders of 2 Xor, 2 And, 1 Or.
    inputs = A[0..n), B[0..n), Cin; outputs = S[0..n), Cout."""
    n_w = 2 * nbits + 1
    A = list(range(nbits))
    B = list(range(nbits, 2 * nbits))
    cin = 2 * nbits
    kinds, in0, in1, out, S = [], [], [], [], []
    carry = cin
    for i in range(nbits):
        x1, s, a1, a2, co = range(n_w, n_w + 5)
        n_w += 5
        for k, a, b, o in ((XOR, A[i], B[i], x1), (XOR, x1, carry, s), (AND, A[i], B[i], a1), (AND, x1, carry, a2), (OR, a1, a2, co)):
            kinds.append(k); in0.append(a); in1.append(b); out.append(o)
        S.append(s)
        carry = co
    return GeneratedNetlist(
        name=f"ripple_adder_{nbits}", n_wires=n_w, n_gates=len(kinds), inputs=np.array(A + B + [cin], np.uint32),
        outputs=np.array(S + [carry], np.uint32), kinds=np.array(kinds, np.uint8), in0=np.array(in0, np.uint32),
        in1=np.array(in1, np.uint32), out=np.array(out, np.uint32))


def array_multiplier(n=32) -> GeneratedNetlist:
    """BASELINE config 2 as a gate list: n x n unsigned array multiplier from And/Or/Xor cells
    (n^2 partial-product Ands, rows of half/full adders).  inputs = A[0..n), B[0..n);
    outputs = P[0..2n)."""
    wires = 2 * n
    kinds, in0, in1, out = [], [], [], []

    def gate(k, a, b):
        nonlocal wires
        o = wires
        wires += 1
        kinds.append(k); in0.append(a); in1.append(b); out.append(o)
        return o

    def half(a, b):
        return gate(XOR, a, b), gate(AND, a, b)

    def full(a, b, c):
        x = gate(XOR, a, b)
        return gate(XOR, x, c), gate(OR, gate(AND, a, b), gate(AND, x, c))

    A = list(range(n))
    B = list(range(n, 2 * n))
    pp = [[gate(AND, A[j], B[i]) for j in range(n)] for i in range(n)]   # pp[i][j] has weight i + j
    P = [pp[0][0]]
    acc = pp[0][1:]            # weights 1..n-1 of the running sum (n-1 bits), carry-out appended below
    acc_top = None             # weight n + i - 1 bit carried from the previous row
    for i in range(1, n):
        row = pp[i]            # weights i .. i+n-1
        new_acc = []
        carry = None
        for j in range(n):
            a = row[j]
            b = acc[j] if j < len(acc) else acc_top
            if b is None:
                s, c = (a, None) if carry is None else half(a, carry)
            elif carry is None:
                s, c = half(a, b)
            else:
                s, c = full(a, b, carry)
            carry = c
            if j == 0:
                P.append(s)
            else:
                new_acc.append(s)
        acc, acc_top = new_acc, carry
    P.extend(acc)
    P.append(acc_top)
    assert len(P) == 2 * n
    return GeneratedNetlist(
        name=f"array_multiplier_{n}x{n}", n_wires=wires, n_gates=len(kinds), inputs=np.array(A + B, np.uint32),
        outputs=np.array(P, np.uint32), kinds=np.array(kinds, np.uint8), in0=np.array(in0, np.uint32),
        in1=np.array(in1, np.uint32), out=np.array(out, np.uint32))


@dataclass
class SequentialNetlist:
    """BASELINE config 4: NAND-only edge-triggered D flip-flops (async preset/clear) forming
    Fibonacci LFSR chains, plus cross-coupled NAND SR latches fed by the flip-flop outputs."""
    n_wires: int
    clk: int
    preset_n: list            # per bit position: active-low preset input wire
    clear_n: list             # per bit position: active-low clear input wire
    q: list                   # q[chain][bit] wire ids
    latch_q: list             # [(Q, Qn)] wire ids
    hazard_latches: list      # indices into latch_q of latches whose S'/R' can rise together (oscillate)
    ops: list                 # (kind, [inputs], output)
    n_gates: int = 0

    def emit(self, builder):
        first = builder.add_wires(self.n_wires)
        assert first == 0
        for kind, ins, out in self.ops:
            if kind == NOT:
                builder.add_not_gate(ins[0], out)
            else:
                builder.add_gate(kind, ins, out)
        return builder


def sequential(n_latches=4096, n_chains=64, bits=64, n_hazard=0, taps=None) -> SequentialNetlist:
    """Flip-flop (7474 structure, 3-input NANDs):
        g1 = NAND(pre', g4, g2)   g2 = NAND(g1, clr', clk)   g3 = NAND(g2, clk, g4)   g4 = NAND(g3, clr', d)
        q  = NAND(pre', g2, qn)   qn = NAND(q, g3, clr')
    Chain c, bit b uses preset/clear inputs of bit position (b + c) % bits, so one per-lane init
    pattern gives every chain a different start state.  d_0 = XOR of the tap bits, d_b = q_{b-1}.
    Latch j: S' = q, R' = qn of one flip-flop (never 0,0 -> never oscillates); the first
    `n_hazard` latches instead take S' and R' from the q of two different chains at the same bit,
    so S' = R' = 0 -> 1 on one clock edge makes the lane oscillate until max_steps."""
    if taps is None:
        taps = [bits - 1, bits - 2, bits - 4, bits - 5] if bits >= 8 else [bits - 1, bits - 2]
    n_w = 0

    def wire():
        nonlocal n_w
        n_w += 1
        return n_w - 1

    ops = []
    clk = wire()
    preset_n = [wire() for _ in range(bits)]
    clear_n = [wire() for _ in range(bits)]
    q = [[wire() for _ in range(bits)] for _ in range(n_chains)]
    qn = [[wire() for _ in range(bits)] for _ in range(n_chains)]
    for c in range(n_chains):
        # feedback
        fb = q[c][taps[0]]
        for t in taps[1:]:
            x = wire()
            ops.append((XOR, [fb, q[c][t]], x))
            fb = x
        for b in range(bits):
            d = fb if b == 0 else q[c][b - 1]
            pre, clr = preset_n[(b + c) % bits], clear_n[(b + c) % bits]
            g1, g2, g3, g4 = wire(), wire(), wire(), wire()
            ops.append((NAND, [pre, g4, g2], g1))
            ops.append((NAND, [g1, clr, clk], g2))
            ops.append((NAND, [g2, clk, g4], g3))
            ops.append((NAND, [g3, clr, d], g4))
            ops.append((NAND, [pre, g2, qn[c][b]], q[c][b]))
            ops.append((NAND, [q[c][b], g3, clr], qn[c][b]))
    latch_q, hazard = [], []
    for j in range(n_latches):
        c, b = j % n_chains, (j // n_chains) % bits
        if j < n_hazard and n_chains > 1:
            s_n, r_n = q[c][b], q[(c + 1) % n_chains][b]
            hazard.append(j)
        else:
            s_n, r_n = q[c][b], qn[c][b]
        lq, lqn = wire(), wire()
        ops.append((NAND, [s_n, lqn], lq))
        ops.append((NAND, [r_n, lq], lqn))
        latch_q.append((lq, lqn))
    return SequentialNetlist(n_wires=n_w, clk=clk, preset_n=preset_n, clear_n=clear_n, q=q, latch_q=latch_q,
                             hazard_latches=hazard, ops=ops, n_gates=len(ops))


@dataclass
class RegisterLfsrNetlist:
    """BASELINE config 4 rebuilt with gsim's own sequential component: Fibonacci LFSR chains out of one-bit registers
    (`add_register`) with a synchronous load path, plus NAND SR latches fed by q / NOT q of the flip-flops."""
    n_wires: int
    clk: int
    enable: int
    load: int
    init: list                # per bit position: load value input wire
    q: list                   # q[chain][bit]
    latch_q: list
    ops: list                 # (kind, [inputs], output); kind "reg": inputs = [d, enable, clock]
    n_gates: int = 0

    def emit(self, builder):
        first = builder.add_wires(self.n_wires)
        assert first == 0
        for kind, ins, out in self.ops:
            if kind == "reg":
                builder.add_register(ins[0], out, ins[1], ins[2])
            elif kind == NOT:
                builder.add_not_gate(ins[0], out)
            else:
                builder.add_gate(kind, ins, out)
        return builder


def sequential_registers(n_latches=4096, n_chains=64, bits=64, taps=None) -> RegisterLfsrNetlist:
    """Chain c, bit b loads init[(b + c) % bits] while `load` is 1 and shifts otherwise: d = (next AND NOT load) OR
    (init AND load), q <= d on the rising clock edge while `enable`."""
    if taps is None:
        taps = [bits - 1, bits - 2, bits - 4, bits - 5] if bits >= 8 else [bits - 1, bits - 2]
    n_w = 0

    def wire():
        nonlocal n_w
        n_w += 1
        return n_w - 1

    ops = []
    clk, enable, load, nload = wire(), wire(), wire(), wire()
    ops.append((NOT, [load], nload))
    init = [wire() for _ in range(bits)]
    q = [[wire() for _ in range(bits)] for _ in range(n_chains)]
    nq = [[wire() for _ in range(bits)] for _ in range(n_chains)]
    for c in range(n_chains):
        fb = q[c][taps[0]]
        for t in taps[1:]:
            x = wire()
            ops.append((XOR, [fb, q[c][t]], x))
            fb = x
        for b in range(bits):
            nxt = fb if b == 0 else q[c][b - 1]
            a, l, d = wire(), wire(), wire()
            ops.append((AND, [nxt, nload], a))
            ops.append((AND, [init[(b + c) % bits], load], l))
            ops.append((OR, [a, l], d))
            ops.append(("reg", [d, enable, clk], q[c][b]))
            ops.append((NOT, [q[c][b]], nq[c][b]))
    latch_q = []
    for j in range(n_latches):
        c, b = j % n_chains, (j // n_chains) % bits
        lq, lqn = wire(), wire()
        ops.append((NAND, [q[c][b], lqn], lq))
        ops.append((NAND, [nq[c][b], lq], lqn))
        latch_q.append((lq, lqn))
    return RegisterLfsrNetlist(n_wires=n_w, clk=clk, enable=enable, load=load, init=init, q=q, latch_q=latch_q, ops=ops,
                               n_gates=len(ops))
