"""Shared test helpers: netlists recorded once and replayed on the oracle builder and on the
engine's builder (both expose the gsim SimulatorBuilder method names)."""
import numpy as np

AND, OR, XOR, NAND, NOR, XNOR, NOT, MUX, BUF, SLICE, MERGE, REG = range(12)
ALL1 = np.uint64(0xFFFFFFFFFFFFFFFF)


class Netlist:
    """A recorded sequence of builder calls."""

    def __init__(self):
        self.widths = []
        self.ops = []  # (kind, inputs, select, output)

    def add_wire(self, width=1):
        self.widths.append(width)
        return len(self.widths) - 1

    def add_wires(self, n, width=1):
        first = len(self.widths)
        self.widths.extend([width] * n)
        return first

    def add_gate(self, kind, inputs, output):
        self.ops.append((kind, list(inputs), None, output))

    def add_not_gate(self, i, o):
        self.ops.append((NOT, [i], None, o))

    def add_multiplexer(self, inputs, select, output):
        self.ops.append((MUX, list(inputs), select, output))

    def add_buffer(self, input, enable, output):
        self.ops.append((BUF, [input, enable], None, output))

    def add_register(self, data_in, data_out, enable, clock):
        self.ops.append((REG, [data_in, enable, clock], None, data_out))

    @property
    def n_wires(self):
        return len(self.widths)

    def replay(self, builder):
        for w in self.widths:
            builder.add_wire(w)
        simple = all(k not in (MUX, BUF, REG) and len(i) <= 2 for k, i, _, _ in self.ops)
        if simple and len(self.ops) > 64:
            kinds = np.array([k for k, _, _, _ in self.ops], np.uint8)
            in0 = np.array([i[0] for _, i, _, _ in self.ops], np.uint32)
            in1 = np.array([i[-1] for _, i, _, _ in self.ops], np.uint32)
            out = np.array([o for _, _, _, o in self.ops], np.uint32)
            builder.add_gates2(kinds, in0, in1, out)
            return builder
        for kind, inputs, select, output in self.ops:
            if kind == MUX:
                builder.add_multiplexer(inputs, select, output)
            elif kind == NOT:
                builder.add_not_gate(inputs[0], output)
            elif kind == BUF:
                builder.add_buffer(inputs[0], inputs[1], output)
            elif kind == REG:
                builder.add_register(inputs[0], output, inputs[1], inputs[2])
            else:
                builder.add_gate(kind, inputs, output)
        return builder


def random_netlist(rng, n_in, n_gates, cyclic=False, wide=False, mux=False):
    nl = Netlist()
    first_in = nl.add_wires(n_in)
    first_g = nl.add_wires(n_gates)
    n_w = n_in + n_gates
    for g in range(n_gates):
        hi = n_w if cyclic else n_in + g
        r = rng.random()
        if mux and r < 0.1:
            nl.add_multiplexer([int(x) for x in rng.integers(0, hi, 2)], int(rng.integers(0, hi)), first_g + g)
            continue
        kind = int(rng.integers(0, 7))
        if kind == NOT:
            nl.add_not_gate(int(rng.integers(0, hi)), first_g + g)
        else:
            n = int(rng.integers(3, 6)) if (wide and r < 0.3) else 2
            nl.add_gate(kind, [int(x) for x in rng.integers(0, hi, n)], first_g + g)
    return nl, first_in


def random_bus_netlist(rng, n_in, n_gates, n_bus, cyclic=True):
    """Random netlist with tri-state buffers: `n_bus` wires are each driven by two or three buffers
    (real multi-driver nets) whose enables are dedicated input wires (see `bus_enable_drives`), a few
    more buffers drive ordinary wires, the rest are plain gates; gates may read the buses when `cyclic`.
    Returns (netlist, first_in, enable_groups)."""
    nl = Netlist()
    first_in = nl.add_wires(n_in)
    first_g = nl.add_wires(n_gates)
    first_bus = nl.add_wires(n_bus)
    n_w = n_in + n_gates + n_bus

    def pick(g):
        return int(rng.integers(0, n_w)) if cyclic else int(rng.integers(0, n_in + g))

    for g in range(n_gates):
        r = rng.random()
        if r < 0.15:
            nl.add_buffer(pick(g), pick(g), first_g + g)
            continue
        kind = int(rng.integers(0, 7))
        if kind == NOT:
            nl.add_not_gate(pick(g), first_g + g)
        else:
            nl.add_gate(kind, [pick(g), pick(g)], first_g + g)
    groups = []
    for b in range(n_bus):
        k = int(rng.integers(2, 4))
        en = [nl.add_wire() for _ in range(k)]
        for e in en:
            nl.add_buffer(pick(n_gates), e, first_bus + b)
        groups.append(en)
    return nl, first_in, groups


def bus_enable_drives(rng, group, words, p_bad=0.03):
    """(value, valid) planes for the enables of one bus: per lane one driver on or none, except a
    fraction `p_bad` of lanes where two are on (conflict) or an enable is X."""
    k = len(group)
    lanes = words * 64
    choice = rng.integers(0, k + 1, lanes)  # k = nobody drives
    val = np.zeros((k, lanes), np.uint8)
    vld = np.ones((k, lanes), np.uint8)
    for j in range(k):
        val[j] = choice == j
    bad = rng.random(lanes) < p_bad
    both = bad & (rng.random(lanes) < 0.5)
    val[0][both] = 1
    val[1][both] = 1
    unk = bad & ~both
    vld[int(rng.integers(0, k))][unk] = 0

    def pack(bits):
        return np.packbits(bits.reshape(words, 64), axis=1, bitorder="little").view(np.uint64).reshape(words)

    return [(pack(val[j]), pack(vld[j])) for j in range(k)]


def random_register_netlist(rng, n_in, n_gates, p_reg=0.25):
    """Random cyclic netlist in which about `p_reg` of the components are registers whose data, enable and clock
    come from arbitrary wires (inputs, gates, other registers).  Returns (netlist, first_in)."""
    nl = Netlist()
    first_in = nl.add_wires(n_in)
    first_g = nl.add_wires(n_gates)
    n_w = n_in + n_gates
    for g in range(n_gates):
        if rng.random() < p_reg:
            d, en, ck = (int(x) for x in rng.integers(0, n_w, 3))
            if rng.random() < 0.5:
                ck = int(rng.integers(0, n_in))       # half of the clocks straight from an input
            nl.add_register(d, first_g + g, en, ck)
            continue
        kind = int(rng.integers(0, 7))
        if kind == NOT:
            nl.add_not_gate(int(rng.integers(0, n_w)), first_g + g)
        else:
            nl.add_gate(kind, [int(x) for x in rng.integers(0, n_w, 2)], first_g + g)
    return nl, first_in


def random_mixed_netlist(rng, n_in, n_gates, n_bus):
    """Everything at once: plain gates, muxes, registers, single buffers, and `n_bus` multi-driver buses with
    dedicated enable inputs (see bus_enable_drives); cyclic.  Returns (netlist, first_in, enable_groups)."""
    nl = Netlist()
    first_in = nl.add_wires(n_in)
    first_g = nl.add_wires(n_gates)
    first_bus = nl.add_wires(n_bus)
    n_w = n_in + n_gates + n_bus

    def pick():
        return int(rng.integers(0, n_w))

    for g in range(n_gates):
        r = rng.random()
        out = first_g + g
        if r < 0.15:
            ck = int(rng.integers(0, n_in)) if rng.random() < 0.5 else pick()
            nl.add_register(pick(), out, pick(), ck)
        elif r < 0.25:
            nl.add_buffer(pick(), pick(), out)
        elif r < 0.35:
            nl.add_multiplexer([pick(), pick()], pick(), out)
        else:
            kind = int(rng.integers(0, 7))
            if kind == NOT:
                nl.add_not_gate(pick(), out)
            else:
                n = int(rng.integers(3, 6)) if rng.random() < 0.2 else 2
                nl.add_gate(kind, [pick() for _ in range(n)], out)
    groups = []
    for b in range(n_bus):
        en = [nl.add_wire() for _ in range(int(rng.integers(2, 4)))]
        for e in en:
            nl.add_buffer(pick(), e, first_bus + b)
        groups.append(en)
    return nl, first_in, groups


def random_drive(rng, words, four_state=True):
    value = rng.integers(0, 1 << 64, words, dtype=np.uint64)
    if four_state:
        valid = rng.integers(0, 1 << 64, words, dtype=np.uint64) | rng.integers(0, 1 << 64, words, dtype=np.uint64)
    else:
        valid = np.full(words, ALL1, np.uint64)
    return value, valid


def lane_mask(n_lanes):
    words = (n_lanes + 63) // 64
    m = np.full(words, ALL1, np.uint64)
    if n_lanes % 64:
        m[-1] = np.uint64((1 << (n_lanes % 64)) - 1)
    return m


def splitmix64(x):
    x = (np.asarray(x, np.uint64) + np.uint64(0x9E3779B97F4A7C15))
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def stimulus_words(seed, n_inputs, words, word_base=0, mode=0):
    """Host restatement of b2_fill_stimulus (include/digilogic_b200.h)."""
    with np.errstate(over="ignore"):
        i = (np.arange(n_inputs, dtype=np.uint64)[:, None] + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        w = (np.arange(words, dtype=np.uint64)[None, :] + np.uint64(word_base)) * np.uint64(0xD1B54A32D192ED03)
        value = splitmix64(np.uint64(seed) ^ i ^ w)
        if mode == 0:
            valid = np.full_like(value, ALL1)
        else:
            valid = splitmix64(value ^ np.uint64(1)) | splitmix64(value ^ np.uint64(2))
    return value, valid
