#!/usr/bin/env python
"""Writes cases.json for the Rust harness: the probes of SURVEY.md A.6/A.7 (latch trace with step
budgets, driver conflicts, mux with Z/X select and data, cold-start read-back), truth tables, wide
wires, n-ary gates, random cyclic netlists with max_steps sweeps."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
KINDS = ["and", "or", "xor", "nand", "nor", "xnor", "not", "mux"]
Z, X, L0, L1 = (0, 0), (1, 0), (0, 1), (1, 1)


def drive(wire, state):
    return [wire, [state[0]], [state[1]]]


def cases():
    out = []
    # A.7 latch trace (also probes A.6-1 and A.6-5)
    a7 = [(L1, L1, 10000), (L0, L1, 10000), (L1, L1, 10000), (L1, L0, 10000), (L0, L0, 10000), (L1, L1, 10), (L1, L1, 11),
          (L0, L1, 10000), (X, L1, 10000), (Z, L0, 10000)]
    out.append({"name": "latch_trace_A7", "wires": [1, 1, 1, 1],
                "comps": [{"kind": "nand", "inputs": [0, 3], "output": 2}, {"kind": "nand", "inputs": [1, 2], "output": 3}],
                "runs": [{"drives": [drive(0, s), drive(1, r)], "max_steps": ms} for s, r, ms in a7]})
    # truth tables: one case per kind, 16 runs
    for k in KINDS[:6]:
        out.append({"name": f"truth_{k}", "wires": [1, 1, 1], "comps": [{"kind": k, "inputs": [0, 1], "output": 2}],
                    "runs": [{"drives": [drive(0, a), drive(1, b)], "max_steps": 10} for a in (Z, X, L0, L1) for b in (Z, X, L0, L1)]})
    # A.6-2/3: conflicts (agreeing and disagreeing drive on a gate output; two gate drivers)
    out.append({"name": "conflict_drive_on_gate_output", "wires": [1, 1, 1], "comps": [{"kind": "and", "inputs": [0, 1], "output": 2}],
                "runs": [{"drives": [drive(0, L1), drive(1, L1)], "max_steps": 10}, {"drives": [drive(2, L1)], "max_steps": 10},
                         {"drives": [drive(2, L0)], "max_steps": 10}, {"drives": [drive(2, X)], "max_steps": 10},
                         {"drives": [drive(2, Z)], "max_steps": 10}]})
    out.append({"name": "conflict_two_drivers", "wires": [1, 1, 1],
                "comps": [{"kind": "and", "inputs": [0, 1], "output": 2}, {"kind": "or", "inputs": [0, 1], "output": 2}],
                "runs": [{"drives": [drive(0, L1), drive(1, L0)], "max_steps": 10}]})
    # A.6-4: mux
    out.append({"name": "mux_z_x", "wires": [1, 1, 1, 1], "comps": [{"kind": "mux", "inputs": [1, 2], "select": 0, "output": 3}],
                "runs": [{"drives": [drive(0, s), drive(1, a), drive(2, b)], "max_steps": 10}
                         for s in (Z, X, L0, L1) for a in (Z, L1) for b in (X, L0)]})
    # tri-state buffers (A.6-6): truth table, a shared bus with hand-over, floating bus,
    # external drive on the bus, a buffer turning on against it, transient overlap through an inverter,
    # gate-derived enables on a cold start
    out.append({"name": "buffer_truth", "wires": [1, 1, 1], "comps": [{"kind": "buffer", "inputs": [0, 1], "output": 2}],
                "runs": [{"drives": [drive(0, d), drive(1, e)], "max_steps": 10} for d in (Z, X, L0, L1) for e in (Z, X, L0, L1)]})
    bus = [{"kind": "buffer", "inputs": [0, 2], "output": 4}, {"kind": "buffer", "inputs": [1, 3], "output": 4},
           {"kind": "not", "inputs": [4], "output": 5}]
    out.append({"name": "buffer_shared_bus", "wires": [1] * 6, "comps": bus,
                "runs": [{"drives": d, "max_steps": 100} for d in (
                    [drive(0, L1), drive(1, L0), drive(2, L1), drive(3, L0)], [drive(2, L0), drive(3, L1)], [drive(3, L0)],
                    [drive(4, L1)], [drive(4, Z)], [drive(4, L0), drive(0, Z)], [drive(2, L1)])]})
    out.append({"name": "buffer_transient_overlap", "wires": [1] * 5,
                "comps": [{"kind": "not", "inputs": [2], "output": 3}, {"kind": "buffer", "inputs": [0, 2], "output": 4},
                          {"kind": "buffer", "inputs": [1, 3], "output": 4}],
                "runs": [{"drives": [drive(0, L1), drive(1, L0), drive(2, L0)], "max_steps": 100},
                         {"drives": [drive(2, L1)], "max_steps": 100}]})
    out.append({"name": "buffer_gate_derived_enables_cold", "wires": [1] * 6,
                "comps": [{"kind": "not", "inputs": [2], "output": 3}, {"kind": "not", "inputs": [3], "output": 4},
                          {"kind": "buffer", "inputs": [0, 3], "output": 5}, {"kind": "buffer", "inputs": [1, 4], "output": 5}],
                "runs": [{"drives": [drive(0, L1), drive(1, L1), drive(2, L1)], "max_steps": 100}]})
    # slices and merges (A.6-7: High-Z bits of the input read X): an 8-bit wire cut in two and swapped
    out.append({"name": "slice_merge_swap", "wires": [8, 3, 5, 8],
                "comps": [{"kind": "slice", "inputs": [0], "select": 0, "output": 1}, {"kind": "slice", "inputs": [0], "select": 3, "output": 2},
                          {"kind": "merge", "inputs": [2, 1], "output": 3}],
                "runs": [{"drives": [[0, [v], [m]]], "max_steps": ms} for v, m, ms in
                         ((0xA5, 0xFF, 10), (0xFF, 0x0F, 10), (0x00, 0x00, 10), (0x3C, 0xFF, 0), (0x3C, 0xFF, 1))]})
    # registers (A.6-8..): Undefined until the first load, rising edge, enable 0 / X, High-Z data bit, X and Z on the clock,
    # and a toggle flip-flop (q -> NOT -> d) whose settle takes more than one step
    reg = [{"kind": "register", "inputs": [0, 2, 3], "output": 1}]
    out.append({"name": "register_script", "wires": [4, 4, 1, 1], "comps": reg,
                "runs": [{"drives": d, "max_steps": ms} for ms in (10,) for d in (
                    [[0, [0b1010], [0xF]], drive(2, L1), drive(3, L0)], [drive(3, L1)], [[0, [0b0101], [0xF]]], [drive(3, L0)], [drive(3, L1)],
                    [drive(3, L0), drive(2, L0)], [drive(3, L1), [0, [0xF], [0xF]]], [drive(3, L0), drive(2, X)], [drive(3, L1)],
                    [drive(3, L0), drive(2, L1), [0, [0b0011], [0b0111]]], [drive(3, L1)], [drive(3, X)], [drive(3, L1), [0, [0], [0xF]]],
                    [drive(3, L0)], [drive(3, L1)], [drive(3, Z)], [drive(3, L0)], [drive(3, L1), drive(2, Z)])]})
    out.append({"name": "register_toggle", "wires": [1, 1, 1, 1],
                "comps": [{"kind": "not", "inputs": [1], "output": 0}, {"kind": "register", "inputs": [0, 2, 3], "output": 1}],
                "runs": [{"drives": d, "max_steps": ms} for d, ms in (
                    ([drive(2, L1), drive(3, L0)], 10), ([drive(3, L1)], 10), ([drive(3, L0)], 0), ([drive(3, L1)], 0), ([drive(3, L0)], 1),
                    ([drive(3, L1)], 1), ([drive(3, L0)], 10), ([drive(3, L1)], 10))]})
    # arithmetic and reduction gates: wrap-around, one invalid bit poisoning the whole sum, reductions with unknown bits
    out.append({"name": "add_sub_8", "wires": [8, 8, 8, 8],
                "comps": [{"kind": "add", "inputs": [0, 1], "output": 2}, {"kind": "sub", "inputs": [0, 1], "output": 3}],
                "runs": [{"drives": [[0, [x], [ma]], [1, [y], [mb]]], "max_steps": 10} for x, y, ma, mb in
                         ((200, 100, 255, 255), (3, 5, 255, 255), (255, 1, 255, 255), (5, 3, 254, 255), (5, 3, 255, 127), (0, 0, 0, 0))]})
    # multiplier, comparators (unsigned and signed), shifters: wrap-around, sign handling, amounts past the width, one invalid bit
    cmps = ("cmp_eq", "cmp_ne", "cmp_ltu", "cmp_gtu", "cmp_leu", "cmp_geu", "cmp_lts", "cmp_gts", "cmp_les", "cmp_ges")
    out.append({"name": "mul_compare_8", "wires": [8, 8, 8] + [1] * 10,
                "comps": [{"kind": "mul", "inputs": [0, 1], "output": 2}] + [{"kind": k, "inputs": [0, 1], "output": 3 + i} for i, k in enumerate(cmps)],
                "runs": [{"drives": [[0, [x], [ma]], [1, [y], [mb]]], "max_steps": 10} for x, y, ma, mb in
                         ((200, 100, 255, 255), (3, 5, 255, 255), (255, 255, 255, 255), (0x80, 0x7F, 255, 255), (5, 5, 255, 255),
                          (5, 3, 254, 255), (5, 3, 255, 127), (0, 0, 0, 0))]})
    out.append({"name": "shifts_8", "wires": [8, 4, 8, 8, 8],
                "comps": [{"kind": "shl", "inputs": [0, 1], "output": 2}, {"kind": "lshr", "inputs": [0, 1], "output": 3},
                          {"kind": "ashr", "inputs": [0, 1], "output": 4}],
                "runs": [{"drives": [[0, [x], [ma]], [1, [n], [mn]]], "max_steps": 10} for x, n, ma, mn in
                         ((0x96, 0, 255, 15), (0x96, 1, 255, 15), (0x96, 7, 255, 15), (0x96, 8, 255, 15), (0x96, 15, 255, 15),
                          (0x16, 3, 255, 15), (0x96, 1, 254, 15), (0x96, 1, 255, 14))]})
    out.append({"name": "horizontal_gates_4", "wires": [4, 1, 1, 1, 1, 1, 1],
                "comps": [{"kind": k, "inputs": [0], "output": i + 1} for i, k in enumerate(("hand", "hor", "hxor", "hnand", "hnor", "hxnor"))],
                "runs": [{"drives": [[0, [v], [m]]], "max_steps": 10} for v, m in
                         ((0xF, 0xF), (0x0, 0xF), (0x5, 0xF), (0xF, 0x7), (0x0, 0xE), (0x8, 0x8), (0x0, 0x0))]})
    # random cyclic netlists, max_steps sweep
    rng = np.random.default_rng(1)
    for n in range(4):
        n_in, n_g = 5, 30
        comps = []
        for g in range(n_g):
            kind = KINDS[int(rng.integers(0, 7))]
            ins = [int(x) for x in rng.integers(0, n_in + n_g, 1 if kind == "not" else int(rng.integers(2, 4)))]
            comps.append({"kind": kind, "inputs": ins, "output": n_in + g})
        runs = []
        for ms in [0, 1, 2, 3, 5, 8, 50, 7, 0, 1000]:
            runs.append({"drives": [drive(int(w), [Z, X, L0, L1][int(rng.integers(0, 4))]) for w in range(n_in) if rng.random() < 0.7],
                         "max_steps": ms})
        out.append({"name": f"random_cyclic_{n}", "wires": [1] * (n_in + n_g), "comps": comps, "runs": runs})
    return out


if __name__ == "__main__":
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "cases.json")
    with open(path, "w") as f:
        json.dump(cases(), f)
    print("wrote", path)
