#!/usr/bin/env python
"""Extracts the 18 SimState golden vectors from the reference's own unit tests
(crates/digilogic_netcode/src/lib.rs:324-592) into tests/golden/sim_state_vectors.json.
Run in the build container (needs /root/reference); the JSON is what travels."""
import json
import os
import re
import sys

SRC = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/crates/digilogic_netcode/src/lib.rs"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sim_state_vectors.json")


def ints(s):
    return [int(x, 0) for x in re.findall(r"0[xb][0-9A-Fa-f]+|\d+", s)]


def main():
    text = open(SRC).read()
    text = text[text.index("mod tests"):]
    tests = re.findall(r"#\[test\]\s*fn (\w+)\(\) \{(.*?)\n    \}", text, re.S)
    vectors = []
    for name, body in tests:
        if name.startswith("insert"):
            pushes = [{"width": int(w), "plane0": ints(a), "plane1": ints(b)}
                      for w, a, b in re.findall(r"push_net\(nz!\((\d+)\), &\[(.*?)\], &\[(.*?)\]\)", body)]
            offsets = [int(x) for x in re.findall(r"assert_eq!\(net_index\w*, (\d+)\)", body)]
            p0 = ints(re.search(r"assert_eq!\(sim_state\.bit_plane_0, \[(.*?)\]\)", body).group(1))
            p1 = ints(re.search(r"assert_eq!\(sim_state\.bit_plane_1, \[(.*?)\]\)", body).group(1))
            vectors.append({"name": name, "kind": "insert", "pushes": pushes, "offsets": offsets, "plane0": p0, "plane1": p1})
        else:
            bit_len = int(re.search(r"bit_len: (\d+)", body).group(1))
            p0 = ints(re.search(r"bit_plane_0: vec!\[(.*?)\]", body).group(1))
            p1 = ints(re.search(r"bit_plane_1: vec!\[(.*?)\]", body).group(1))
            reads = []
            for m in re.finditer(r"get_net\((\d+), nz!\((\d+)\),.*?\);\s*assert_eq!\(bit_plane_0, \[(.*?)\]\);\s*assert_eq!\(bit_plane_1, \[(.*?)\]\);", body, re.S):
                reads.append({"offset": int(m.group(1)), "width": int(m.group(2)), "plane0": ints(m.group(3)), "plane1": ints(m.group(4))})
            vectors.append({"name": name, "kind": "read", "bit_len": bit_len, "plane0": p0, "plane1": p1, "reads": reads})
    assert len(vectors) == 18, len(vectors)
    with open(OUT, "w") as f:
        json.dump({"source": "crates/digilogic_netcode/src/lib.rs:324-592", "vectors": vectors}, f, indent=1)
    print(f"wrote {len(vectors)} vectors to {OUT}")


if __name__ == "__main__":
    main()
