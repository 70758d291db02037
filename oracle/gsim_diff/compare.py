#!/usr/bin/env python
"""compare.py cases.json gsim_dump.json — replays the cases on the C oracle (oracle/gsim_oracle.c)
and reports every difference from the dump of real gsim 1.1.4 (result per run, planes per wire).
A clean run pins the oracle, and through the GPU parity tests the engine, to the reference."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402

KIND = {"and": 0, "or": 1, "xor": 2, "nand": 3, "nor": 4, "xnor": 5}
HKIND = {"h" + k: v for k, v in KIND.items()}
RESULT = {0: "Ok", 1: "MaxStepsReached", 2: "Err"}


def main(cases_path, dump_path):
    cases = json.load(open(cases_path))
    dump = {c["name"]: c for c in json.load(open(dump_path))}
    bad = 0
    for case in cases:
        ref = dump[case["name"]]
        b = oracle.Builder()
        for w in case["wires"]:
            b.add_wire(w)
        for c in case["comps"]:
            if c["kind"] == "not":
                b.add_not_gate(c["inputs"][0], c["output"])
            elif c["kind"] == "mux":
                b.add_multiplexer(c["inputs"], c["select"], c["output"])
            elif c["kind"] == "buffer":
                b.add_buffer(c["inputs"][0], c["inputs"][1], c["output"])
            elif c["kind"] == "slice":
                b.add_slice(c["inputs"][0], c["select"], c["output"])
            elif c["kind"] == "merge":
                b.add_merge(c["inputs"], c["output"])
            elif c["kind"] == "register":
                b.add_register(c["inputs"][0], c["output"], c["inputs"][1], c["inputs"][2])
            elif c["kind"] in ("add", "sub"):
                (b.add_add if c["kind"] == "add" else b.add_sub)(c["inputs"][0], c["inputs"][1], c["output"])
            elif c["kind"] == "mul":
                b.add_mul(c["inputs"][0], c["inputs"][1], c["output"])
            elif c["kind"].startswith("cmp_"):
                ops = ("cmp_eq", "cmp_ne", "cmp_ltu", "cmp_gtu", "cmp_leu", "cmp_geu", "cmp_lts", "cmp_gts", "cmp_les", "cmp_ges")
                b.add_compare(ops.index(c["kind"]), c["inputs"][0], c["inputs"][1], c["output"])
            elif c["kind"] in ("shl", "lshr", "ashr"):
                b.add_shift(c["kind"], c["inputs"][0], c["inputs"][1], c["output"])
            elif c["kind"] in HKIND:
                b.add_horizontal_gate(HKIND[c["kind"]], c["inputs"][0], c["output"])
            else:
                b.add_gate(KIND[c["kind"]], c["inputs"], c["output"])
        sim = b.build()
        conflicted = False
        for i, (run, rref) in enumerate(zip(case["runs"], ref["runs"])):
            for w, p0, p1 in run["drives"]:
                sim.set_wire_drive(w, sum(x << (32 * k) for k, x in enumerate(p0)), sum(x << (32 * k) for k, x in enumerate(p1)))
            res = RESULT[sim.run_sim(run["max_steps"])]
            if res != rref["result"]:
                print(f"{case['name']} run {i}: result {res} != gsim {rref['result']}")
                bad += 1
            conflicted |= res == "Err"
            if conflicted:
                continue  # state after a conflict is outside the contract; report it separately
            for w, (p0, p1) in enumerate(rref["wires"]):
                width = case["wires"][w]
                mask = (1 << width) - 1
                want = (sum(x << (32 * k) for k, x in enumerate(p0)) & mask, sum(x << (32 * k) for k, x in enumerate(p1)) & mask)
                if sim.get_wire_state(w) != want:
                    print(f"{case['name']} run {i} wire {w}: {sim.get_wire_state(w)} != gsim {want}")
                    bad += 1
    print("differences:", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1], sys.argv[2]))
