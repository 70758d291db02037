"""The cases of the real-gsim differential harness (oracle/gsim_diff): the Python side works
(an oracle-made dump compares clean), and the engine reproduces the same runs on the GPU, so a
dump of real gsim 1.1.4 can be held against both."""
import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "gsim_diff"))
import compare  # noqa: E402
import make_cases  # noqa: E402

import oracle  # noqa: E402


CMPS = ("cmp_eq", "cmp_ne", "cmp_ltu", "cmp_gtu", "cmp_leu", "cmp_geu", "cmp_lts", "cmp_gts", "cmp_les", "cmp_ges")


def build(case, builder):
    for w in case["wires"]:
        builder.add_wire(w)
    for c in case["comps"]:
        if c["kind"] == "not":
            builder.add_not_gate(c["inputs"][0], c["output"])
        elif c["kind"] == "mux":
            builder.add_multiplexer(c["inputs"], c["select"], c["output"])
        elif c["kind"] == "buffer":
            builder.add_buffer(c["inputs"][0], c["inputs"][1], c["output"])
        elif c["kind"] == "slice":
            builder.add_slice(c["inputs"][0], c["select"], c["output"])
        elif c["kind"] == "merge":
            builder.add_merge(c["inputs"], c["output"])
        elif c["kind"] == "register":
            builder.add_register(c["inputs"][0], c["output"], c["inputs"][1], c["inputs"][2])
        elif c["kind"] in ("add", "sub"):
            (builder.add_add if c["kind"] == "add" else builder.add_sub)(c["inputs"][0], c["inputs"][1], c["output"])
        elif c["kind"] == "mul":
            builder.add_mul(c["inputs"][0], c["inputs"][1], c["output"])
        elif c["kind"].startswith("cmp_"):
            builder.add_compare(CMPS.index(c["kind"]), c["inputs"][0], c["inputs"][1], c["output"])
        elif c["kind"] in ("shl", "lshr", "ashr"):
            builder.add_shift(c["kind"], c["inputs"][0], c["inputs"][1], c["output"])
        elif c["kind"] in compare.HKIND:
            builder.add_horizontal_gate(compare.HKIND[c["kind"]], c["inputs"][0], c["output"])
        else:
            builder.add_gate(compare.KIND[c["kind"]], c["inputs"], c["output"])
    return builder.build()


def oracle_dump(cases):
    dump = []
    for case in cases:
        sim = build(case, oracle.Builder())
        runs = []
        for run in case["runs"]:
            for w, p0, p1 in run["drives"]:
                sim.set_wire_drive(w, p0[0], p1[0])
            res = compare.RESULT[sim.run_sim(run["max_steps"])]
            st = [sim.get_wire_state(w) for w in range(len(case["wires"]))]
            runs.append({"result": res, "wires": [[[s[0]], [s[1]]] for s in st]})
        dump.append({"name": case["name"], "build_errors": [], "runs": runs})
    return dump


def test_harness_python_side(tmp_path):
    cases = make_cases.cases()
    assert len(cases) >= 23
    cp, dp = tmp_path / "cases.json", tmp_path / "dump.json"
    cp.write_text(json.dumps(cases))
    dp.write_text(json.dumps(oracle_dump(cases)))
    assert compare.main(str(cp), str(dp)) == 0
    # a corrupted dump is detected
    bad = oracle_dump(cases)
    bad[0]["runs"][5]["result"] = "Ok"
    dp.write_text(json.dumps(bad))
    assert compare.main(str(cp), str(dp)) == 1


@pytest.mark.gpu
def test_engine_reproduces_the_differential_cases():
    from digilogic_b200 import LogicState, SimulatorBuilder
    cases = make_cases.cases()
    for case, ref in zip(cases, oracle_dump(cases)):
        sim = build(case, SimulatorBuilder())
        conflicted = False
        for i, (run, rref) in enumerate(zip(case["runs"], ref["runs"])):
            for w, p0, p1 in run["drives"]:
                sim.set_wire_drive(w, LogicState(p0[0], p1[0]))
            res = compare.RESULT[int(sim.run_sim(run["max_steps"]))]
            assert res == rref["result"], (case["name"], i)
            conflicted |= res == "Err"
            if conflicted:
                continue
            for w, (p0, p1) in enumerate(rref["wires"]):
                assert sim.get_wire_state(w) == LogicState(p0[0], p1[0]), (case["name"], i, w)
