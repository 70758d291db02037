#!/usr/bin/env python
"""Latency of the reference server's per-Eval cycle (set every input, eval, read every net back one call
per net — crates/digilogic_netcode/src/server.rs:375-383,447-456) through b2_server_*, from a plain-C
client (benchmarks/server_cycle.c), for the 4-bit adder, the 32x32 multiplier and a 100k-gate DAG.
One JSON line per circuit.  B2_LIB_PATH selects another build of the library for A/B runs."""
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from digilogic_b200 import build, netgen  # noqa: E402


class Recorder:
    """Records builder calls (gsim SimulatorBuilder method names) as the text netlist the C client reads."""

    def __init__(self):
        self.n_wires, self.ops = 0, []

    def add_wire(self, width=1):
        self.n_wires += 1
        return self.n_wires - 1

    def add_wires(self, n, width=1):
        first = self.n_wires
        self.n_wires += n
        return first

    def add_gate(self, kind, inputs, output):
        self.ops.append((int(kind), [int(i) for i in inputs], int(output)))

    def add_not_gate(self, i, o):
        self.ops.append((6, [int(i)], int(o)))

    def add_gates2(self, kinds, in0, in1, out):
        for k, a, b, o in zip(kinds, in0, in1, out):
            self.ops.append((int(k), [int(a)] if int(k) == 6 else [int(a), int(b)], int(o)))

    def __getattr__(self, name):  # add_and_gate ... add_xnor_gate
        kinds = {"add_and_gate": 0, "add_or_gate": 1, "add_xor_gate": 2, "add_nand_gate": 3, "add_nor_gate": 4, "add_xnor_gate": 5}
        if name in kinds:
            return lambda inputs, output: self.add_gate(kinds[name], inputs, output)
        raise AttributeError(name)


def main():
    lib = os.environ.get("B2_LIB_PATH") or build.build()
    lib_dir = os.path.dirname(lib)
    with tempfile.TemporaryDirectory() as tmp:
        exe = os.path.join(tmp, "server_cycle")
        subprocess.run(["gcc", "-O2", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "benchmarks", "server_cycle.c"),
                        "-o", exe, "-L", lib_dir, "-l:" + os.path.basename(lib), "-Wl,-rpath," + lib_dir], check=True)
        cases = [("4-bit ripple-carry adder", netgen.ripple_adder(4), 200),
                 ("32x32 array multiplier", netgen.array_multiplier(32), 50),
                 ("random DAG, 100k gates, depth 100", netgen.random_dag(100_000, 100, 256, 64, seed=0xD1610003), 10)]
        for name, gen, cycles in cases:
            rec = gen.emit(Recorder())
            path = os.path.join(tmp, "netlist.txt")
            with open(path, "w") as f:
                f.write(f"{rec.n_wires} {len(rec.ops)} {len(gen.inputs)}\n")
                f.write(" ".join(str(int(w)) for w in gen.inputs) + "\n")
                for kind, ins, out in rec.ops:
                    f.write(f"{kind} {len(ins)} {' '.join(map(str, ins))} {out}\n")
            res = subprocess.run([exe, path, str(cycles)], capture_output=True, text=True)
            if res.returncode != 0:
                print(json.dumps({"circuit": name, "error": res.returncode, "stderr": res.stderr.strip()[-300:]}))
                continue
            line = json.loads(res.stdout)
            line["circuit"] = name
            print(json.dumps(line))


if __name__ == "__main__":
    main()
