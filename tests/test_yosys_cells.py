"""The Yosys cell types the reference importer names but todo!()s (crates/digilogic_serde/src/yosys.rs:160-211), read by the
harness importer onto the wider component set (rows f1 x f4): a 4-bit enabled counter from $dffe, half-adder logic, a $mux,
a $tribuf output stage, reductions and a constant — simulated on the CPU oracle against a Python model, and on the GPU
against the oracle."""
import json

import numpy as np
import pytest

import oracle
from digilogic_b200 import importers

Z, X, L0, L1 = (0, 0), (1, 0), (0, 1), (1, 1)


def counter_json():
    """bits: clk=2 en=3 sel=4 oe=5; q=10..13; carries 20..; d=30..33; nq=40..43; m=50..53 (mux); y=60..63 (tribuf);
    all=70 (reduce_and), par=71 (reduce_xor), zero=72 (logic_not), same=73 ($xnor of q0,q1), any=74 (reduce_bool)."""
    cells = {}

    def cell(name, typ, conn, params=None):
        cells[name] = {"hide_name": 1, "type": typ, "parameters": params or {}, "attributes": {},
                       "port_directions": {}, "connections": conn}

    q, d, nq, m, y = [10, 11, 12, 13], [30, 31, 32, 33], [40, 41, 42, 43], [50, 51, 52, 53], [60, 61, 62, 63]
    carry = "1"                                   # q + 1: the carry into bit 0 is the constant 1
    for i in range(4):
        cell(f"x{i}", "$xor", {"A": [q[i]], "B": [carry], "Y": [d[i]]})
        if i < 3:
            cell(f"a{i}", "$and", {"A": [q[i]], "B": [carry], "Y": [20 + i]})
            carry = 20 + i
    cell("ff", "$dffe", {"CLK": [2], "EN": [3], "D": d, "Q": q}, {"CLK_POLARITY": "1", "EN_POLARITY": "1", "WIDTH": "100"})
    cell("inv", "$not", {"A": q, "Y": nq})
    cell("mx", "$mux", {"A": q, "B": nq, "S": [4], "Y": m})
    cell("tb", "$tribuf", {"A": m, "EN": [5], "Y": y})
    cell("r_all", "$reduce_and", {"A": q, "Y": [70]})
    cell("r_par", "$reduce_xor", {"A": q, "Y": [71]})
    cell("r_zero", "$logic_not", {"A": q, "Y": [72]})
    cell("same", "$xnor", {"A": [q[0]], "B": [q[1]], "Y": [73]})
    cell("r_any", "$reduce_bool", {"A": q, "Y": [74]})
    ports = {"clk": {"direction": "input", "bits": [2]}, "en": {"direction": "input", "bits": [3]},
             "sel": {"direction": "input", "bits": [4]}, "oe": {"direction": "input", "bits": [5]},
             "q": {"direction": "output", "bits": q}, "y": {"direction": "output", "bits": y},
             "all": {"direction": "output", "bits": [70]}, "par": {"direction": "output", "bits": [71]},
             "zero": {"direction": "output", "bits": [72]}, "same": {"direction": "output", "bits": [73]},
             "any": {"direction": "output", "bits": [74]}}
    return json.dumps({"creator": "test", "modules": {"counter": {"attributes": {"top": "1".zfill(32)}, "ports": ports, "cells": cells,
                                                                   "netnames": {}}}})


def run_script(make_sim, set_drive, run, get):
    circ = importers.load_yosys(counter_json())
    assert list(circ.const_nets.values()) == [1]
    sim = make_sim(circ)
    for n, v in circ.const_nets.items():
        set_drive(sim, n, (v, 1))
    i = {k: v[0] for k, v in circ.inputs.items()}
    o = circ.outputs
    for name, st in (("clk", L0), ("en", L1), ("sel", L0), ("oe", L1)):
        set_drive(sim, i[name], st)
    run(sim)
    # the registers hold Undefined until loaded: force a known state through the enable-less path is impossible in this
    # design, so check X first, then note that q + 1 of X stays X — the counter needs a defined start.  Use the fact that
    # AND with a known 0 is 0: not available either; therefore only X-propagation is checked before the first load.
    assert all(get(sim, n) == X for n in o["q"])
    return circ, sim, i, o


def test_yosys_cells_import_and_x_propagation_on_oracle():
    def make(circ):
        return circ.emit(oracle.Builder()).build()
    circ, sim, i, o = run_script(make, lambda s, n, st: s.set_wire_drive(n, *st), lambda s: s.run_sim(1000), lambda s, n: s.get_wire_state(n))
    kinds = [g[0] for g in circ.gates]
    assert kinds.count(importers.REG) == 4 and kinds.count(importers.BUF) == 4 and kinds.count(importers.MUX) == 4
    # a rising edge with everything X loads X again; reductions of X are X, except where a known bit decides
    sim.set_wire_drive(i["clk"], *L1)
    assert sim.run_sim(1000) == oracle.RUN_OK
    assert all(sim.get_wire_state(n) == X for n in o["q"] + o["y"] + o["all"] + o["par"])
    sim.set_wire_drive(i["oe"], *L0)                  # output stage off: High-Z
    assert sim.run_sim(1000) == oracle.RUN_OK
    assert all(sim.get_wire_state(n) == Z for n in o["y"])


def shift_json(n=6):
    """A shift register that can be initialised: q[0] <= din, q[i] <= q[i-1] ($dff, no enable -> constant-1 enable)."""
    cells, q = {}, list(range(10, 10 + n))
    for k in range(n):
        cells[f"ff{k}"] = {"hide_name": 1, "type": "$dff", "parameters": {"CLK_POLARITY": "1", "WIDTH": "1"}, "attributes": {},
                           "port_directions": {}, "connections": {"CLK": [2], "D": [3 if k == 0 else q[k - 1]], "Q": [q[k]]}}
    cells["par"] = {"hide_name": 1, "type": "$reduce_xor", "parameters": {}, "attributes": {}, "port_directions": {},
                    "connections": {"A": q, "Y": [40]}}
    cells["zero"] = {"hide_name": 1, "type": "$logic_not", "parameters": {}, "attributes": {}, "port_directions": {},
                     "connections": {"A": q, "Y": [41]}}
    ports = {"clk": {"direction": "input", "bits": [2]}, "din": {"direction": "input", "bits": [3]},
             "q": {"direction": "output", "bits": q}, "par": {"direction": "output", "bits": [40]},
             "zero": {"direction": "output", "bits": [41]}}
    return json.dumps({"modules": {"shift": {"ports": ports, "cells": cells, "netnames": {}}}})


def shift_check(make, set_drive, run, get, n=6, clocks=40):
    circ = importers.load_yosys(shift_json(n))
    sim = make(circ)
    for net, v in circ.const_nets.items():
        set_drive(sim, net, (v, 1))
    clk, din = circ.inputs["clk"][0], circ.inputs["din"][0]
    rng = np.random.default_rng(9)
    state = [None] * n
    set_drive(sim, clk, L0)
    run(sim)
    for t in range(clocks):
        bit = int(rng.integers(0, 2))
        set_drive(sim, din, (bit, 1))
        set_drive(sim, clk, L1)
        run(sim)
        set_drive(sim, clk, L0)
        run(sim)
        state = [bit] + state[:-1]
        for k in range(n):
            assert get(sim, circ.outputs["q"][k]) == (X if state[k] is None else (state[k], 1)), (t, k)
        if None not in state:
            assert get(sim, circ.outputs["par"][0]) == (sum(state) & 1, 1)
            assert get(sim, circ.outputs["zero"][0]) == (int(not any(state)), 1)


def test_shift_register_from_dff_cells_on_oracle():
    shift_check(lambda c: c.emit(oracle.Builder()).build(), lambda s, n, st: s.set_wire_drive(n, *st),
                lambda s: s.run_sim(1000), lambda s, n: s.get_wire_state(n))


@pytest.mark.gpu
def test_shift_register_from_dff_cells_on_gpu():
    from digilogic_b200 import LogicState, SimulatorBuilder

    def get(s, n):
        st = s.get_wire_state(n)
        return (st.plane0, st.plane1)

    shift_check(lambda c: c.emit(SimulatorBuilder()).build(), lambda s, n, st: s.set_wire_drive(n, LogicState(*st)),
                lambda s: s.run_sim(1000), get)


# ------------------------------------------------------------------------------------------------------------------
# word-level cells ($add $sub $neg $mul, comparisons, shifts), $pmux, $logic_and/or, $sdff*, $dlatch

def alu_json():
    """a = bits 10..17 (8), b = 20..25 (6, signed in the signed cells), sh = 30..32, s = 40..42 (pmux selects), clk 2, rst 3, en 4."""
    cells = {}

    def cell(name, typ, conn, **params):
        cells[name] = {"hide_name": 1, "type": typ, "parameters": {k: (v if isinstance(v, str) else format(v, "b")) for k, v in params.items()},
                       "attributes": {}, "port_directions": {}, "connections": conn}

    a, b, sh, s = list(range(10, 18)), list(range(20, 26)), [30, 31, 32], [40, 41, 42]
    out = {}

    def y(name, width):
        base = 100 + 20 * len(out)
        out[name] = list(range(base, base + width))
        return out[name]

    cell("add", "$add", {"A": a, "B": b, "Y": y("sum", 9)}, A_SIGNED=0, B_SIGNED=0)
    cell("sub", "$sub", {"A": a, "B": b, "Y": y("diff", 8)}, A_SIGNED=1, B_SIGNED=1)
    cell("mul", "$mul", {"A": a, "B": b, "Y": y("prod", 14)}, A_SIGNED=0, B_SIGNED=0)
    cell("neg", "$neg", {"A": b, "Y": y("negb", 7)}, A_SIGNED=1)
    cell("lt", "$lt", {"A": a, "B": b, "Y": y("lt_s", 1)}, A_SIGNED=1, B_SIGNED=1)
    cell("ge", "$ge", {"A": a, "B": b, "Y": y("ge_u", 1)}, A_SIGNED=0, B_SIGNED=0)
    cell("eq", "$eq", {"A": a, "B": b, "Y": y("eq", 1)}, A_SIGNED=0, B_SIGNED=0)
    cell("shl", "$shl", {"A": a, "B": sh, "Y": y("shl", 8)}, A_SIGNED=0, B_SIGNED=0)
    cell("sshr", "$sshr", {"A": a, "B": sh, "Y": y("sshr", 8)}, A_SIGNED=1, B_SIGNED=0)
    cell("shr", "$shr", {"A": a, "B": sh, "Y": y("shr", 8)}, A_SIGNED=0, B_SIGNED=0)
    cell("pm", "$pmux", {"A": a[:4], "B": a[4:8] + b[:4] + [b[4], b[5], "1", "0"], "S": s, "Y": y("pm", 4)}, WIDTH=4, S_WIDTH=3)
    cell("land", "$logic_and", {"A": a, "B": sh, "Y": y("land", 1)})
    cell("lor", "$logic_or", {"A": sh, "B": s, "Y": y("lor", 1)})
    cell("ff1", "$sdff", {"CLK": [2], "SRST": [3], "D": a[:4], "Q": y("q_sdff", 4)}, CLK_POLARITY=1, SRST_POLARITY=1, SRST_VALUE="1010", WIDTH=4)
    cell("ff2", "$sdffe", {"CLK": [2], "SRST": [3], "EN": [4], "D": a[:4], "Q": y("q_sdffe", 4)}, CLK_POLARITY=1, SRST_POLARITY=1,
         EN_POLARITY=1, SRST_VALUE="0110", WIDTH=4)
    cell("ff3", "$sdffce", {"CLK": [2], "SRST": [3], "EN": [4], "D": a[:4], "Q": y("q_sdffce", 4)}, CLK_POLARITY=1, SRST_POLARITY=0,
         EN_POLARITY=1, SRST_VALUE="0001", WIDTH=4)
    cell("lat", "$dlatch", {"EN": [4], "D": b[:3], "Q": y("q_lat", 3)}, EN_POLARITY=1, WIDTH=3)
    ports = {"a": {"direction": "input", "bits": a}, "b": {"direction": "input", "bits": b}, "sh": {"direction": "input", "bits": sh},
             "s": {"direction": "input", "bits": s}, "clk": {"direction": "input", "bits": [2]}, "rst": {"direction": "input", "bits": [3]},
             "en": {"direction": "input", "bits": [4]}}
    ports.update({k: {"direction": "output", "bits": v} for k, v in out.items()})
    return json.dumps({"creator": "test", "modules": {"alu": {"attributes": {"top": "1".zfill(32)}, "ports": ports, "cells": cells, "netnames": {}}}})


def _s(x, n):
    return x - (1 << n) if x >> (n - 1) else x


def alu_model(a, b, sh, s):
    """Yosys semantics of the combinational cells (operands extended to the result width, signed where both are)."""
    sa, sb = _s(a, 8), _s(b, 6)
    pm = a & 15
    for i, choice in enumerate([(a >> 4) & 15, b & 15, ((b >> 4) & 3) | 4]):      # last case: {0, 1, b5, b4}
        if (s >> i) & 1:
            pm = choice
    return {"sum": (a + b) & 0x1FF, "diff": (sa - sb) & 0xFF, "prod": (a * b) & 0x3FFF, "negb": (-sb) & 0x7F, "lt_s": int(sa < sb),
            "ge_u": int(a >= b), "eq": int(a == b), "shl": (a << sh) & 0xFF, "sshr": (sa >> sh) & 0xFF, "shr": a >> sh, "pm": pm,
            "land": int(a != 0 and sh != 0), "lor": int(sh != 0 or s != 0)}


def test_word_level_cells_on_the_oracle():
    circ = importers.load_yosys(alu_json())
    assert len(circ.wide) == 10 and circ.to_json()["wide"] and importers.ImportedCircuit.from_json(circ.to_json()).wide == circ.wide
    sim = circ.emit(oracle.Builder()).build()
    for n, v in circ.const_nets.items():
        sim.set_wire_drive(n, v, 1)
    rng = np.random.default_rng(12)

    def drive(name, value):
        for j, n in enumerate(circ.inputs[name]):
            sim.set_wire_drive(n, (value >> j) & 1, 1)

    def read(name):
        bits = [sim.get_wire_state(n) for n in circ.outputs[name]]
        assert all(m == 1 for _, m in bits), (name, bits)
        return sum(v << j for j, (v, _) in enumerate(bits))

    drive("clk", 0); drive("rst", 0); drive("en", 0)
    for k in range(60):
        a, b, sh = (int(x) for x in (rng.integers(0, 256), rng.integers(0, 64), rng.integers(0, 8)))
        s = [0, 1, 2, 4, 0][k % 5]
        drive("a", a); drive("b", b); drive("sh", sh); drive("s", s)
        assert sim.run_sim(1000) == oracle.RUN_OK
        want = alu_model(a, b, sh, s)
        got = {k_: read(k_) for k_ in want}
        assert got == want, (a, b, sh, s)
    # sequential cells: reset value with and without enable, load, hold; the latch follows b while en and holds after
    def clock():
        assert sim.run_sim(1000) == oracle.RUN_OK                                 # the new inputs settle before the edge
        drive("clk", 1); assert sim.run_sim(1000) == oracle.RUN_OK
        drive("clk", 0); assert sim.run_sim(1000) == oracle.RUN_OK

    drive("a", 0b0111); drive("b", 0b101); drive("rst", 1); drive("en", 0); clock()
    assert read("q_sdff") == 0b1010 and read("q_sdffe") == 0b0110                 # $sdffe resets without its enable
    drive("rst", 0); drive("en", 1); clock()                                      # rst = 0 is ACTIVE for ff3 (SRST_POLARITY 0)
    assert read("q_sdff") == 0b0111 and read("q_sdffe") == 0b0111 and read("q_sdffce") == 0b0001 and read("q_lat") == 0b101
    drive("rst", 1); drive("en", 0); drive("a", 0b1100); drive("b", 0b010); clock()
    assert read("q_sdff") == 0b1010 and read("q_sdffe") == 0b0110 and read("q_sdffce") == 0b0001 and read("q_lat") == 0b101
    drive("rst", 0); drive("en", 0); clock()
    assert read("q_sdff") == 0b1100 and read("q_sdffe") == 0b0110 and read("q_sdffce") == 0b0001   # not enabled: no reset, no load
    drive("rst", 1); drive("en", 1); clock()                                      # ff3: reset inactive, enabled: loads
    assert read("q_sdffce") == 0b1100 and read("q_lat") == 0b010


@pytest.mark.gpu
@pytest.mark.parametrize("resident", [1, 0])
def test_word_level_cells_on_gpu_against_oracle(resident):
    from digilogic_b200 import LogicState, SimulatorBuilder
    circ = importers.load_yosys(alu_json())
    ref = circ.emit(oracle.Builder()).build()
    sim = circ.emit(SimulatorBuilder()).build()
    sim.set_option(sim.OPT_RESIDENT_KERNEL, resident)
    n_wires = sim.info.n_wires

    def drive(net, v, m=1):
        ref.set_wire_drive(net, v, m)
        sim.set_wire_drive(net, LogicState(v, m))

    for n, v in circ.const_nets.items():
        drive(n, v)
    rng = np.random.default_rng(13)
    all_inputs = [n for nets in circ.inputs.values() for n in nets]
    for rnd in range(40):
        for n in all_inputs:
            v, m = int(rng.integers(0, 2)), int(rng.random() > 0.03)             # now and then an X / Z input bit
            drive(n, v, m)
        ms = (1000, 1000, 3, 0)[rnd % 4]
        assert int(sim.run_sim(ms)) == ref.run_sim(ms), rnd
        for w in range(n_wires):
            assert sim.get_wire_state(w) == LogicState(*ref.get_wire_state(w)), (rnd, w)


# ------------------------------------------------------------------------------------------------------------------
# $mem_v2: lowered onto registers, address decoders and multiplexer trees

MEM = dict(size=12, offset=2, abits=4, width=4)


def mem_json():
    """clk 2; write port: wa 10..13, wd 20..23, we 30..33 (per bit); read port 0 (asynchronous): ra 40..43 -> rd 50..53;
    read port 1 (clocked on clk, enable 3): rb 60..63 -> rq 70..73."""
    p = MEM
    cells = {"mem": {"hide_name": 0, "type": "$mem_v2", "attributes": {}, "port_directions": {},
                     "parameters": {"MEMID": "\\\\m", "SIZE": format(p["size"], "b"), "OFFSET": format(p["offset"], "b"), "ABITS": format(p["abits"], "b"),
                                    "WIDTH": format(p["width"], "b"), "INIT": "x", "RD_PORTS": "10", "WR_PORTS": "1",
                                    "RD_CLK_ENABLE": "10", "RD_CLK_POLARITY": "11", "RD_TRANSPARENCY_MASK": "00", "RD_COLLISION_X_MASK": "00",
                                    "WR_CLK_ENABLE": "1", "WR_CLK_POLARITY": "1"},
                     "connections": {"RD_CLK": ["x", 2], "RD_EN": ["1", 3], "RD_ADDR": [40, 41, 42, 43, 60, 61, 62, 63],
                                     "RD_DATA": [50, 51, 52, 53, 70, 71, 72, 73], "RD_ARST": ["0", "0"], "RD_SRST": ["0", "0"],
                                     "WR_CLK": [2], "WR_EN": [30, 31, 32, 33], "WR_ADDR": [10, 11, 12, 13], "WR_DATA": [20, 21, 22, 23]}}}
    ports = {"clk": {"direction": "input", "bits": [2]}, "ren": {"direction": "input", "bits": [3]},
             "wa": {"direction": "input", "bits": [10, 11, 12, 13]}, "wd": {"direction": "input", "bits": [20, 21, 22, 23]},
             "we": {"direction": "input", "bits": [30, 31, 32, 33]}, "ra": {"direction": "input", "bits": [40, 41, 42, 43]},
             "rb": {"direction": "input", "bits": [60, 61, 62, 63]},
             "rd": {"direction": "output", "bits": [50, 51, 52, 53]}, "rq": {"direction": "output", "bits": [70, 71, 72, 73]}}
    return json.dumps({"creator": "test", "modules": {"ram": {"attributes": {"top": "1".zfill(32)}, "ports": ports, "cells": cells, "netnames": {}}}})


def mem_check(make, set_drive, run, get, clocks=120):
    """Random writes with per-bit enables, reads through both ports, against a per-bit Python model (None = Undefined)."""
    p = MEM
    circ = importers.load_yosys(mem_json())
    kinds = [g[0] for g in circ.gates]
    assert kinds.count(importers.REG) == p["size"] * p["width"] + p["width"]       # stored bits + the clocked read port
    sim = make(circ)
    for n, v in circ.const_nets.items():
        set_drive(sim, n, (v, 1))
    rng = np.random.default_rng(21)
    mem = {a: [None] * p["width"] for a in range(p["offset"], p["offset"] + p["size"])}
    rq = [None] * p["width"]

    def drive(name, value):
        for j, n in enumerate(circ.inputs[name]):
            set_drive(sim, n, ((value >> j) & 1, 1))

    def word(name):
        return [get(sim, n) for n in circ.outputs[name]]

    def want(bits):
        return [X if b is None else (b, 1) for b in bits]

    drive("clk", 0); drive("ren", 0)
    for t in range(clocks):
        wa, wd, we, ra, rb, ren = (int(x) for x in (rng.integers(0, 16), rng.integers(0, 16), rng.integers(0, 16), rng.integers(0, 16),
                                                        rng.integers(0, 16), rng.integers(0, 2)))
        if t % 3 == 0:
            ra = wa                                        # read what is being written: old value before the edge, new after
        if t % 5 == 0:
            rb = wa                                        # the clocked port reads the word before the write lands
        drive("wa", wa); drive("wd", wd); drive("we", we); drive("ra", ra); drive("rb", rb); drive("ren", ren)
        run(sim)
        assert word("rd") == want(mem.get(ra, [None] * p["width"])), t
        drive("clk", 1); run(sim)
        if ren:
            rq = list(mem.get(rb, [None] * p["width"]))    # sampled from the state before the edge
        if wa in mem:
            mem[wa] = [((wd >> j) & 1) if (we >> j) & 1 else mem[wa][j] for j in range(p["width"])]
        assert word("rd") == want(mem.get(ra, [None] * p["width"])), t
        assert word("rq") == want(rq), t
        drive("clk", 0); run(sim)
        assert word("rq") == want(rq), t
    assert any(b is not None for w in mem.values() for b in w) and any(b is None for w in mem.values() for b in w)


def test_mem_v2_on_the_oracle():
    mem_check(lambda c: c.emit(oracle.Builder()).build(), lambda s, n, st: s.set_wire_drive(n, *st),
              lambda s: s.run_sim(1000), lambda s, n: s.get_wire_state(n))


def test_mem_v2_refuses_what_it_cannot_lower():
    doc = json.loads(mem_json())
    cell = doc["modules"]["ram"]["cells"]["mem"]
    for key, value in (("WR_PORTS", "10"), ("INIT", "0101"), ("WR_CLK_ENABLE", "0"), ("RD_TRANSPARENCY_MASK", "01"), ("SIZE", format(300, "b"))):
        bad = json.loads(json.dumps(doc))
        bad["modules"]["ram"]["cells"]["mem"]["parameters"][key] = value
        with pytest.raises(ValueError):
            importers.load_yosys(json.dumps(bad))
    assert cell["parameters"]["WR_PORTS"] == "1"


@pytest.mark.gpu
@pytest.mark.parametrize("resident", [1, 0])
def test_mem_v2_on_gpu(resident):
    from digilogic_b200 import LogicState, SimulatorBuilder

    def make(c):
        s = c.emit(SimulatorBuilder()).build()
        s.set_option(s.OPT_RESIDENT_KERNEL, resident)
        return s

    def get(s, n):
        st = s.get_wire_state(n)
        return (st.plane0, st.plane1)

    mem_check(make, lambda s, n, st: s.set_wire_drive(n, LogicState(*st)), lambda s: s.run_sim(1000), get, clocks=60)


@pytest.mark.gpu
def test_mem_v2_with_undefined_address_bits_gpu_equals_oracle():
    """Every wire, including X / Z on address, enable and data bits, against the oracle."""
    from digilogic_b200 import LogicState, SimulatorBuilder
    circ = importers.load_yosys(mem_json())
    ref = circ.emit(oracle.Builder()).build()
    sim = circ.emit(SimulatorBuilder()).build()
    n_wires = sim.info.n_wires

    def drive(net, v, m=1):
        ref.set_wire_drive(net, v, m)
        sim.set_wire_drive(net, LogicState(v, m))

    for n, v in circ.const_nets.items():
        drive(n, v)
    rng = np.random.default_rng(5)
    clk = circ.inputs["clk"][0]
    others = [n for k, nets in circ.inputs.items() if k != "clk" for n in nets]
    for rnd in range(60):
        for n in others:
            drive(n, int(rng.integers(0, 2)), int(rng.random() > 0.05))
        for level in (1, 0):
            drive(clk, level)
            assert int(sim.run_sim(1000)) == ref.run_sim(1000), rnd
            for w in range(n_wires):
                assert sim.get_wire_state(w) == LogicState(*ref.get_wire_state(w)), (rnd, level, w)


def test_pos_and_case_equality_cells():
    cells = {"p": {"type": "$pos", "parameters": {"A_SIGNED": "1"}, "connections": {"A": [10, 11, 12], "Y": [20, 21, 22, 23, 24]}},
             "e": {"type": "$eqx", "parameters": {"A_SIGNED": "0", "B_SIGNED": "0"}, "connections": {"A": [10, 11, 12], "B": [13, 14, 15], "Y": [30]}},
             "n": {"type": "$nex", "parameters": {"A_SIGNED": "0", "B_SIGNED": "0"}, "connections": {"A": [10, 11, 12], "B": [13, 14, 15], "Y": [31]}}}
    ports = {"a": {"direction": "input", "bits": [10, 11, 12]}, "b": {"direction": "input", "bits": [13, 14, 15]},
             "y": {"direction": "output", "bits": [20, 21, 22, 23, 24]}, "eq": {"direction": "output", "bits": [30]},
             "ne": {"direction": "output", "bits": [31]}}
    circ = importers.load_yosys(json.dumps({"modules": {"m": {"ports": ports, "cells": cells, "netnames": {}}}}))
    sim = circ.emit(oracle.Builder()).build()
    for n, v in circ.const_nets.items():
        sim.set_wire_drive(n, v, 1)
    for a in range(8):
        for b in (a, (a + 3) % 8):
            for j in range(3):
                sim.set_wire_drive(circ.inputs["a"][j], (a >> j) & 1, 1)
                sim.set_wire_drive(circ.inputs["b"][j], (b >> j) & 1, 1)
            assert sim.run_sim(100) == oracle.RUN_OK
            y = sum(sim.get_wire_state(n)[0] << j for j, n in enumerate(circ.outputs["y"]))
            assert y == (a | (0b11000 if a & 4 else 0))                      # sign-extended
            assert sim.get_wire_state(circ.outputs["eq"][0]) == (int(a == b), 1)
            assert sim.get_wire_state(circ.outputs["ne"][0]) == (int(a != b), 1)


def test_memory_port_cells_give_the_same_circuit_as_mem_v2():
    """`$memrd_v2` / `$memwr_v2` (the form before Yosys' memory_collect) are collected into the `$mem_v2` they stand for."""
    doc = json.loads(mem_json())
    mod = doc["modules"]["ram"]
    m = mod["cells"].pop("mem")
    p, c = m["parameters"], m["connections"]
    ab, w = MEM["abits"], MEM["width"]
    mod["memories"] = {"\\m": {"width": w, "start_offset": MEM["offset"], "size": MEM["size"]}}
    for r in range(2):
        mod["cells"][f"rd{r}"] = {"type": "$memrd_v2", "parameters": {"MEMID": "\\m", "ABITS": p["ABITS"], "WIDTH": p["WIDTH"],
                                  "CLK_ENABLE": str(r), "CLK_POLARITY": "1", "TRANSPARENCY_MASK": "0", "COLLISION_X_MASK": "0"},
                                  "connections": {"CLK": [c["RD_CLK"][r]], "EN": [c["RD_EN"][r]], "ARST": ["0"], "SRST": ["0"],
                                                  "ADDR": c["RD_ADDR"][r * ab:(r + 1) * ab], "DATA": c["RD_DATA"][r * w:(r + 1) * w]}}
    mod["cells"]["wr0"] = {"type": "$memwr_v2", "parameters": {"MEMID": "\\m", "ABITS": p["ABITS"], "WIDTH": p["WIDTH"], "CLK_ENABLE": "1",
                           "CLK_POLARITY": "1", "PORTID": "0", "PRIORITY_MASK": "0"},
                           "connections": {"CLK": c["WR_CLK"], "EN": c["WR_EN"], "ADDR": c["WR_ADDR"], "DATA": c["WR_DATA"]}}
    split = importers.load_yosys(json.dumps(doc))
    whole = importers.load_yosys(mem_json())
    assert split.gates == whole.gates and split.inputs == whole.inputs and split.outputs == whole.outputs
    mod["cells"]["init"] = {"type": "$meminit_v2", "parameters": {"MEMID": "\\m"}, "connections": {}}
    with pytest.raises(ValueError):
        importers.load_yosys(json.dumps(doc))


def test_sr_cell_on_the_oracle():
    cells = {"l": {"type": "$sr", "parameters": {"SET_POLARITY": "1", "CLR_POLARITY": "0", "WIDTH": "10"},
                   "connections": {"SET": [10, 11], "CLR": [12, 13], "Q": [20, 21]}}}
    ports = {"s": {"direction": "input", "bits": [10, 11]}, "rn": {"direction": "input", "bits": [12, 13]}, "q": {"direction": "output", "bits": [20, 21]}}
    circ = importers.load_yosys(json.dumps({"modules": {"m": {"ports": ports, "cells": cells, "netnames": {}}}}))
    sim = circ.emit(oracle.Builder()).build()
    for n, v in circ.const_nets.items():
        sim.set_wire_drive(n, v, 1)

    def step(s, rn):
        for j in range(2):
            sim.set_wire_drive(circ.inputs["s"][j], (s >> j) & 1, 1)
            sim.set_wire_drive(circ.inputs["rn"][j], (rn >> j) & 1, 1)
        assert sim.run_sim(100) == oracle.RUN_OK
        return [sim.get_wire_state(n) for n in circ.outputs["q"]]

    assert step(0b00, 0b00) == [L0, L0]            # CLR is active low: both cleared
    assert step(0b01, 0b11) == [L1, L0]            # bit 0 set, bit 1 holds its 0
    assert step(0b00, 0b11) == [L1, L0]            # hold
    assert step(0b11, 0b10) == [L0, L1]            # bit 0: set and clear together -> clear wins; bit 1 set
    assert step(0b00, 0b11) == [L0, L1]
