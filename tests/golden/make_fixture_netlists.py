#!/usr/bin/env python
"""Imports the reference's sample circuits (crates/digilogic/assets/testdata/*.dig, *.yosys) with
digilogic_b200.importers and stores the resulting gate lists as tests/golden/fixtures.json, so
that the tests can run where /root/reference does not exist.  Run in the build container."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from digilogic_b200 import importers  # noqa: E402

SRC = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/crates/digilogic/assets/testdata"
FILES = ["small.dig", "medium.dig", "large.dig", "mux.dig", "small.yosys", "medium.yosys", "large.yosys", "large_opt.yosys"]


def main():
    out = {"source": "crates/digilogic/assets/testdata", "circuits": {}}
    for f in FILES:
        text = open(os.path.join(SRC, f)).read()
        circ = importers.load_dig(text, f) if f.endswith(".dig") else importers.load_yosys(text)
        circ.name = f
        out["circuits"][f] = circ.to_json()
        print(f, "nets", circ.n_nets, "gates", len(circ.gates), "in", list(circ.inputs), "out", list(circ.outputs))
    with open(os.path.join(HERE, "fixtures.json"), "w") as fh:
        json.dump(out, fh, separators=(",", ":"))


if __name__ == "__main__":
    main()
