"""Shared test helpers: netlists recorded once and replayed on the oracle builder and on the
engine's builder (both expose the gsim SimulatorBuilder method names)."""
import numpy as np

AND, OR, XOR, NAND, NOR, XNOR, NOT, MUX = range(8)
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

    @property
    def n_wires(self):
        return len(self.widths)

    def replay(self, builder):
        for w in self.widths:
            builder.add_wire(w)
        simple = all(k != MUX and len(i) <= 2 for k, i, _, _ in self.ops)
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
