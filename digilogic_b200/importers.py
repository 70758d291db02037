"""Harness-side readers (and writers) for the two circuit file formats behind BASELINE configs 1
and 2 — SURVEY.md §8 row f1.

* Digital `.dig` (H. Neemann's Digital, XML): restates what the reference's importer does —
  symbols are placed with the port geometry of crates/digilogic_core/src/symbol.rs:32-253, wire
  segments join nets only at identical end points, and nets are found by a flood fill over
  positions (crates/digilogic_serde/src/digital.rs:111-215).  Element names:
  crates/digilogic_serde/src/digital/circuitfile.rs:55-79.
* Yosys JSON: 1-bit `$and/$or/$xor/$not` cells joined by bit ids
  (crates/digilogic_serde/src/yosys.rs:115-243, schema yosys/netlist.rs:159-196).  Beyond what the reference translates
  today, the cell types its importer already names but `todo!()`s (yosys.rs:160-211) map onto the engine's wider component
  set: `$xnor`, `$mux`, `$pmux`, `$reduce_and/or/xor/xnor/bool`, `$logic_not/and/or`, `$tribuf`, `$dff`,
  `$dffe`, `$sdff`, `$sdffe`, `$sdffce`, `$dlatch`, and — through gsim's word-level components, on wide wires that `emit()`
  merges from and slices back into the 1-bit nets — `$add`, `$sub`, `$neg`, `$mul`, `$lt/$le/$eq/$ne/$ge/$gt`,
  `$shl/$sshl/$shr/$sshr`, `$pos`, `$eqx/$nex`, `$sr`; `$mem_v2` — or the `$memrd_v2` / `$memwr_v2` port cells of a memory,
  collected first — is lowered onto registers, address decoders and multiplexer trees (`_lower_mem_v2`); constant bits "0"/"1" become two extra input nets (`const_nets`) the caller drives.

Both produce an `ImportedCircuit`, whose `emit()` issues the same calls the GUI client sends
(crates/digilogic_netcode/src/client.rs:288-427): one 1-bit net per net, then one gate per
symbol with inputs in port order (mux: S, A, B so that inputs[0] is the select).
"""
from __future__ import annotations

import json
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field

AND, OR, XOR, NAND, NOR, XNOR, NOT, MUX = range(8)
BUF, REG = "buf", "reg"   # gates-list kinds beyond the reference's: tri-state buffer [A, EN], register [D, EN, CLK]

# port name -> (dx, dy, is_input), in the reference's port order
_GATE2 = [("A", 0, 0, True), ("B", 0, 40, True), ("Y", 80, 20, False)]
_GATE1 = [("A", 0, 0, True), ("Y", 40, 0, False)]
DIG_SYMBOLS = {
    "And": (AND, _GATE2), "Or": (OR, _GATE2), "XOr": (XOR, _GATE2), "Not": (NOT, _GATE1),
    "In": (None, [("Y", 0, 0, False)]), "Out": (None, [("A", 0, 0, True)]),
    "Multiplexer": (MUX, [("S", 20, 40, True), ("A", 0, 0, True), ("B", 0, 40, True), ("Y", 40, 20, False)]),
}
YOSYS_CELLS = {"$and": AND, "$or": OR, "$xor": XOR, "$not": NOT, "$xnor": XNOR}
# cell -> (unsigned, signed) comparator op of the engine (b2_compare_op)
YOSYS_COMPARE = {"$eq": (0, 0), "$ne": (1, 1), "$lt": (2, 6), "$gt": (3, 7), "$le": (4, 8), "$ge": (5, 9),
                 "$eqx": (0, 0), "$nex": (1, 1)}     # case equality: the same comparator on defined inputs
YOSYS_REDUCE = {"$reduce_and": AND, "$reduce_or": OR, "$reduce_bool": OR, "$reduce_xor": XOR, "$reduce_xnor": XNOR, "$logic_not": NOR}


@dataclass
class ImportedCircuit:
    name: str
    n_nets: int
    gates: list = field(default_factory=list)      # (kind, [input nets], output net); mux inputs = [S, A, B]
    inputs: dict = field(default_factory=dict)     # label -> [nets] (LSB first)
    outputs: dict = field(default_factory=dict)
    const_nets: dict = field(default_factory=dict)  # net -> 0 / 1: constants of the source file, to be driven by the caller
    # word-level components: (op, [operand bit-net lists], [output bit nets]); op = "add" "sub" "mul" ("cmp", CMP_*) "shl" "lshr" "ashr"
    wide: list = field(default_factory=list)

    def emit(self, builder):
        """Same sequence as the client's build: AddNet{width:1} per net, then the gates."""
        for _ in range(self.n_nets):
            builder.add_wire(1)
        for kind, ins, out in self.gates:
            if kind == MUX:
                builder.add_multiplexer(ins[1:], ins[0], out)
            elif kind == BUF:
                builder.add_buffer(ins[0], ins[1], out)
            elif kind == REG:
                builder.add_register(ins[0], out, ins[1], ins[2])
            elif kind == NOT:
                builder.add_not_gate(ins[0], out)
            else:
                builder.add_gate(kind, ins, out)
        for op, operands, outs in self.wide:
            # operand bits -> one wide wire each (add_merge), the component, result bits back onto the 1-bit nets (add_slice)
            ws = []
            for bits in operands:
                w = builder.add_wire(len(bits))
                builder.add_merge(list(bits), w)
                ws.append(w)
            cmp_op = None
            if isinstance(op, tuple):
                op, cmp_op = op
            y = builder.add_wire(1 if op == "cmp" else len(operands[0]))
            if op == "add":
                builder.add_add(ws[0], ws[1], y)
            elif op == "sub":
                builder.add_sub(ws[0], ws[1], y)
            elif op == "mul":
                builder.add_mul(ws[0], ws[1], y)
            elif op == "cmp":
                builder.add_compare(cmp_op, ws[0], ws[1], y)
            else:
                builder.add_shift(op, ws[0], ws[1], y)
            for j, out in enumerate(outs):
                builder.add_slice(y, j, out)
        return builder

    def to_json(self) -> dict:
        d = {"name": self.name, "n_nets": self.n_nets, "gates": [[k, list(i), o] for k, i, o in self.gates],
             "inputs": self.inputs, "outputs": self.outputs}
        if self.wide:
            d["wide"] = [[list(op) if isinstance(op, tuple) else op, [list(x) for x in ops], list(outs)] for op, ops, outs in self.wide]
        return d

    @classmethod
    def from_json(cls, d: dict) -> "ImportedCircuit":
        c = cls(d["name"], d["n_nets"], [(k, list(i), o) for k, i, o in d["gates"]], dict(d["inputs"]), dict(d["outputs"]))
        c.wide = [(tuple(op) if isinstance(op, list) else op, [list(x) for x in ops], list(outs)) for op, ops, outs in d.get("wide", [])]
        return c


# ------------------------------------------------------------------------------ Digital .dig

def load_dig(text: str, name: str = "circuit") -> ImportedCircuit:
    root = ET.fromstring(text)
    symbols = []   # (element name, label, [(port name, (x, y), is_input)])
    for ve in root.find("visualElements").findall("visualElement"):
        elem = ve.findtext("elementName")
        if elem not in DIG_SYMBOLS:
            raise ValueError(f"unsupported Digital element {elem!r}")
        pos = ve.find("pos")
        x, y = int(pos.get("x")), int(pos.get("y"))
        label = None
        attrs = ve.find("elementAttributes")
        if attrs is not None:
            for entry in attrs.findall("entry"):
                kids = list(entry)
                if len(kids) == 2 and kids[0].text == "Label":
                    label = kids[1].text
        ports = [(pn, (x + dx, y + dy), is_in) for pn, dx, dy, is_in in DIG_SYMBOLS[elem][1]]
        symbols.append((elem, label, ports))

    # union-find over positions: a port position and the two ends of every wire segment
    parent = {}

    def find(p):
        parent.setdefault(p, p)
        while parent[p] != p:
            parent[p] = parent[parent[p]]
            p = parent[p]
        return p

    for _, _, ports in symbols:
        for _, p, _ in ports:
            find(p)
    wires = root.find("wires")
    for w in (wires.findall("wire") if wires is not None else []):
        p1 = (int(w.find("p1").get("x")), int(w.find("p1").get("y")))
        p2 = (int(w.find("p2").get("x")), int(w.find("p2").get("y")))
        parent[find(p1)] = find(p2)

    # nets in first-seen order of the symbols' ports (any order is a valid client numbering)
    net_of = {}

    def net(p):
        r = find(p)
        if r not in net_of:
            net_of[r] = len(net_of)
        return net_of[r]

    circ = ImportedCircuit(name, 0)
    for elem, label, ports in symbols:
        kind = DIG_SYMBOLS[elem][0]
        if elem == "In":
            circ.inputs.setdefault(label or f"in{len(circ.inputs)}", []).append(net(ports[0][1]))
        elif elem == "Out":
            circ.outputs.setdefault(label or f"out{len(circ.outputs)}", []).append(net(ports[0][1]))
        else:
            ins = [net(p) for _, p, is_in in ports if is_in]
            outs = [net(p) for _, p, is_in in ports if not is_in]
            circ.gates.append((kind, ins, outs[0]))
    circ.n_nets = len(net_of)
    return circ


def write_dig(symbols, wires) -> str:
    """symbols: [(element name, label or None, x, y)], wires: [((x1,y1),(x2,y2))] -> Digital XML v2."""
    out = ['<?xml version="1.0" encoding="utf-8"?>', "<circuit>", "  <version>2</version>", "  <attributes/>", "  <visualElements>"]
    for elem, label, x, y in symbols:
        out.append("    <visualElement>")
        out.append(f"      <elementName>{elem}</elementName>")
        if label is not None:
            out += ["      <elementAttributes>", "        <entry>", "          <string>Label</string>",
                    f"          <string>{label}</string>", "        </entry>", "      </elementAttributes>"]
        elif elem in ("And", "Or", "XOr"):
            out += ["      <elementAttributes>", "        <entry>", "          <string>wideShape</string>",
                    "          <boolean>true</boolean>", "        </entry>", "      </elementAttributes>"]
        else:
            out.append("      <elementAttributes/>")
        out.append(f'      <pos x="{x}" y="{y}"/>')
        out.append("    </visualElement>")
    out += ["  </visualElements>", "  <wires>"]
    for (x1, y1), (x2, y2) in wires:
        out += ["    <wire>", f'      <p1 x="{x1}" y="{y1}"/>', f'      <p2 x="{x2}" y="{y2}"/>', "    </wire>"]
    out += ["  </wires>", "  <measurementOrdering/>", "</circuit>", ""]
    return "\n".join(out)


def adder_dig(nbits: int = 4) -> str:
    """BASELINE config 1: an nbits ripple-carry adder authored as a Digital file from In/Out/And/Or/XOr
    symbols at the reference's port geometry; every connection is one point-to-point wire segment
    (a net with fan-out is a star of segments from the driver's port)."""
    symbols, wires = [], []
    port = {"gate2": {"A": (0, 0), "B": (0, 40), "Y": (80, 20)}}

    def place(elem, label, x, y):
        symbols.append((elem, label, x, y))
        return (x, y)

    def pin(pos, name):
        dx, dy = port["gate2"][name]
        return (pos[0] + dx, pos[1] + dy)

    def connect(src, dst):
        wires.append((src, dst))

    carry = place("In", "Cin", 0, -200)          # In: port Y at (0,0)
    for i in range(nbits):
        y0 = i * 400
        a = place("In", f"A{i}", 0, y0)
        b = place("In", f"B{i}", 0, y0 + 100)
        x1 = place("XOr", None, 200, y0)
        s = place("XOr", None, 500, y0)
        a1 = place("And", None, 200, y0 + 160)
        a2 = place("And", None, 500, y0 + 160)
        co = place("Or", None, 800, y0 + 160)
        so = place("Out", f"S{i}", 1000, y0 + 20)
        for g in (x1, a1):
            connect(a, pin(g, "A"))
            connect(b, pin(g, "B"))
        for g in (s, a2):
            connect(pin(x1, "Y"), pin(g, "A"))
            connect(carry, pin(g, "B"))
        connect(pin(a1, "Y"), pin(co, "A"))
        connect(pin(a2, "Y"), pin(co, "B"))
        connect(pin(s, "Y"), so)
        carry = pin(co, "Y")
    cout = place("Out", "Cout", 1000, nbits * 400)
    connect(carry, cout)
    return write_dig(symbols, wires)


# ------------------------------------------------------------------------------ Yosys JSON

def _collect_memory_ports(mod) -> dict:
    """The cells of a module with every group of `$memrd_v2` / `$memwr_v2` port cells of one memory (what Yosys emits before its
    memory_collect pass; crates/digilogic_serde/src/yosys.rs:204-206) replaced by the `$mem_v2` cell memory_collect would make
    of them; size and offset come from the module's `memories` section.  `$meminit_v2` (initial contents) is refused like a
    defined INIT of `$mem_v2`."""
    cells = mod.get("cells", {})
    if not any(c["type"] in ("$memrd_v2", "$memwr_v2", "$meminit_v2") for c in cells.values()):
        return cells

    def num(v):
        return int(v, 2) if isinstance(v, str) else int(v)

    out, groups = {}, {}
    for name, c in cells.items():
        if c["type"] == "$meminit_v2":
            raise ValueError(f"$meminit_v2 {name}: initial contents are not supported (registers start Undefined)")
        if c["type"] in ("$memrd_v2", "$memwr_v2"):
            memid = c["parameters"]["MEMID"]
            groups.setdefault(memid, {"rd": [], "wr": []})["rd" if c["type"] == "$memrd_v2" else "wr"].append((name, c))
        else:
            out[name] = c
    for memid, g in groups.items():
        info = mod.get("memories", {}).get(memid) or mod.get("memories", {}).get(memid.lstrip("\\"))
        if info is None:
            raise ValueError(f"memory {memid!r}: no entry in the module's `memories` section")
        ports = g["rd"] + g["wr"]
        abits, width = num(ports[0][1]["parameters"]["ABITS"]), int(info["width"])
        if any(num(c["parameters"]["ABITS"]) != abits or num(c["parameters"]["WIDTH"]) != width for _, c in ports):
            raise ValueError(f"memory {memid!r}: ports of differing address or data width (wide ports) are not supported")

        def bits(key, which, default="0"):
            return "".join(str(num(c["parameters"].get(key, default)) & 1) for _, c in reversed(g[which])) or "0"

        params = {"MEMID": memid, "SIZE": format(int(info["size"]), "b"), "OFFSET": format(int(info.get("start_offset", 0)), "b"),
                  "ABITS": format(abits, "b"), "WIDTH": format(width, "b"), "INIT": "x",
                  "RD_PORTS": format(len(g["rd"]), "b"), "WR_PORTS": format(len(g["wr"]), "b"),
                  "RD_CLK_ENABLE": bits("CLK_ENABLE", "rd"), "RD_CLK_POLARITY": bits("CLK_POLARITY", "rd", "1"),
                  "WR_CLK_ENABLE": bits("CLK_ENABLE", "wr"), "WR_CLK_POLARITY": bits("CLK_POLARITY", "wr", "1"),
                  "RD_TRANSPARENCY_MASK": "1" if any(num(c["parameters"].get("TRANSPARENCY_MASK", 0)) for _, c in g["rd"]) else "0",
                  "RD_COLLISION_X_MASK": "1" if any(num(c["parameters"].get("COLLISION_X_MASK", 0)) for _, c in g["rd"]) else "0"}
        conn = {"RD_CLK": [c["connections"]["CLK"][0] for _, c in g["rd"]], "RD_EN": [c["connections"]["EN"][0] for _, c in g["rd"]],
                "RD_ADDR": [b for _, c in g["rd"] for b in c["connections"]["ADDR"]],
                "RD_DATA": [b for _, c in g["rd"] for b in c["connections"]["DATA"]],
                "WR_CLK": [c["connections"]["CLK"][0] for _, c in g["wr"]], "WR_EN": [b for _, c in g["wr"] for b in c["connections"]["EN"]],
                "WR_ADDR": [b for _, c in g["wr"] for b in c["connections"]["ADDR"]],
                "WR_DATA": [b for _, c in g["wr"] for b in c["connections"]["DATA"]]}
        out[f"$mem_v2:{memid}"] = {"type": "$mem_v2", "parameters": params, "connections": conn}
    return out


def _lower_mem_v2(circ, cname, conn, params, net, tmp, flag, param_bits):
    """Yosys `$mem_v2` (crates/digilogic_serde/src/yosys.rs:207, `MemV2 => todo!()`) lowered onto registers, multiplexers and
    gates: one register per stored bit (so a word is Undefined until it has been written, like gsim's register), per write port
    an address decoder (n-ary AND over the address bits or their complements) whose hit, ANDed with the port's per-bit write
    enable, is the register's enable; per read port a binary tree of 2:1 multiplexers over the words, selected by the address
    bits (an address bit that is not a valid 0/1 reads Undefined).  Supported: any number of read ports, asynchronous or clocked
    on a rising edge with an enable (read-before-write: both registers sample the state before the edge); ONE write port clocked
    on a rising edge; SIZE <= 256 words; no reset values, no initial contents other than 'x'."""
    size, abits, width, offset = flag("SIZE"), flag("ABITS"), flag("WIDTH"), flag("OFFSET")
    n_rd, n_wr = flag("RD_PORTS"), flag("WR_PORTS")
    if not (1 <= size <= 256 and 1 <= abits <= 16 and width >= 1):
        raise ValueError(f"$mem_v2 {cname}: SIZE {size} / ABITS {abits} / WIDTH {width} outside the supported range")
    if n_wr > 1:
        raise ValueError(f"$mem_v2 {cname}: {n_wr} write ports (one is supported)")
    init = params.get("INIT", "x")
    if isinstance(init, str) and set(init) - set("xX"):
        raise ValueError(f"$mem_v2 {cname}: initial contents are not supported (registers start Undefined)")
    if n_wr and (flag("WR_CLK_ENABLE") & 1) != 1:
        raise ValueError(f"$mem_v2 {cname}: asynchronous write ports are not supported")
    if n_wr and (flag("WR_CLK_POLARITY") & 1) != 1:
        raise ValueError(f"$mem_v2 {cname}: falling-edge write ports are not supported")
    if flag("RD_TRANSPARENCY_MASK") or flag("RD_COLLISION_X_MASK"):
        raise ValueError(f"$mem_v2 {cname}: transparent / collision-x read ports are not supported")

    def complement(n):
        inv = tmp()
        circ.gates.append((NOT, [n], inv))
        return inv

    def gate(kind, ins):
        out = tmp()
        circ.gates.append((kind, ins if len(ins) > 1 else ins * 2, out))
        return out

    # the stored bits
    q = [[tmp() for _ in range(width)] for _ in range(size)]
    if n_wr:
        wa = [net(b) for b in conn["WR_ADDR"][:abits]]
        wa_n = [complement(a) for a in wa]
        clk = net(conn["WR_CLK"][0])
        wen = [net(b) for b in conn["WR_EN"][:width]]
        wdat = [net(b) for b in conn["WR_DATA"][:width]]
        for e in range(size):
            addr = e + offset
            if addr >> abits:
                raise ValueError(f"$mem_v2 {cname}: word {addr} is not addressable with {abits} address bits")
            hit = gate(AND, [wa[i] if (addr >> i) & 1 else wa_n[i] for i in range(abits)])
            for j in range(width):
                circ.gates.append((REG, [wdat[j], gate(AND, [hit, wen[j]]), clk], q[e][j]))
    floating = None
    for r in range(n_rd):
        ra = [net(b) for b in conn["RD_ADDR"][r * abits:(r + 1) * abits]]
        clocked = (flag("RD_CLK_ENABLE") >> r) & 1
        if clocked and ((flag("RD_CLK_POLARITY") >> r) & 1) != 1:
            raise ValueError(f"$mem_v2 {cname}: falling-edge read ports are not supported")
        for j in range(width):
            # words by address, padded to a power of two with a floating net (reads Undefined through the multiplexer)
            level = []
            for addr in range(1 << abits):
                e = addr - offset
                if 0 <= e < size:
                    level.append(q[e][j])
                else:
                    if floating is None:
                        floating = tmp()
                    level.append(floating)
            out = net(conn["RD_DATA"][r * width + j])
            for i in range(abits):
                last = i == abits - 1 and not clocked
                nxt = []
                for k in range(0, len(level), 2):
                    if level[k] == level[k + 1] == floating and floating is not None:
                        nxt.append(floating)
                        continue
                    y = out if (last and len(level) == 2) else tmp()
                    circ.gates.append((MUX, [ra[i], level[k], level[k + 1]], y))
                    nxt.append(y)
                level = nxt
            if clocked:
                en = net(conn["RD_EN"][r]) if "RD_EN" in conn else net("1")
                circ.gates.append((REG, [level[0], en, net(conn["RD_CLK"][r])], out))
            elif level[0] != out:       # a one-word memory behind floating padding: copy
                circ.gates.append((OR, [level[0], level[0]], out))



def load_yosys(text: str, module: str | None = None) -> ImportedCircuit:
    doc = json.loads(text)
    mods = doc["modules"]
    if module is None:
        tops = [m for m, v in mods.items() if int(v.get("attributes", {}).get("top", "0"), 2) == 1]
        module = tops[0] if tops else next(iter(mods))
    mod = mods[module]
    net_of = {}

    circ = ImportedCircuit(module, 0)

    def net(bit):
        if isinstance(bit, tuple):       # an internal net of a lowered cell
            if bit not in net_of:
                net_of[bit] = len(net_of)
            return net_of[bit]
        if not isinstance(bit, int):
            # "0" / "1": the reference importer todo!()s on constants; here they are shared nets the caller drives
            if bit not in ("0", "1"):
                raise ValueError(f"constant bit {bit!r} is not supported")
            key = ("const", bit)
            if key not in net_of:
                net_of[key] = len(net_of)
                circ.const_nets[net_of[key]] = int(bit)
            return net_of[key]
        if bit not in net_of:
            net_of[bit] = len(net_of)
        return net_of[bit]

    for pname, port in mod.get("ports", {}).items():
        nets = [net(b) for b in port["bits"]]
        if port["direction"] == "input":
            circ.inputs[pname] = nets
        elif port["direction"] == "output":
            circ.outputs[pname] = nets
        else:
            raise ValueError("inout ports are not supported")
    def tmp():
        return net(("tmp", len(net_of)))

    cells = _collect_memory_ports(mod)
    for cname, cell in cells.items():
        ctype = cell["type"]
        conn = cell["connections"]
        params = cell.get("parameters", {})

        def flag(name):
            v = params.get(name, 0)
            return int(v, 2) if isinstance(v, str) else int(v)

        def polarity(name):
            v = params.get(name, 1)
            return int(v, 2) if isinstance(v, str) else int(v)

        def positive(bit, pol_name):
            """the net of `bit`, inverted when the cell's polarity parameter says active-low"""
            if polarity(pol_name) == 1:
                return net(bit)
            inv = tmp()
            circ.gates.append((NOT, [net(bit)], inv))
            return inv

        def param_bits(name, width):
            v = params.get(name, 0)
            text = v if isinstance(v, str) else format(int(v), "b")
            text = text.zfill(width)[-width:]
            return [int(ch) for ch in reversed(text)]          # LSB first

        def extend(bits, signed_param, width):
            """operand bits as nets, extended to `width` with its sign bit (signed) or constant 0, or truncated"""
            nets = [net(b) for b in bits][:width]
            pad = nets[-1] if signed_param and flag(signed_param) and nets else net("0")
            return nets + [pad] * (width - len(nets))
        if ctype in YOSYS_CELLS:
            kind = YOSYS_CELLS[ctype]
            for j in range(len(conn["Y"])):
                ins = [net(conn["A"][j])] if kind == NOT else [net(conn["A"][j]), net(conn["B"][j])]
                circ.gates.append((kind, ins, net(conn["Y"][j])))
        elif ctype in YOSYS_REDUCE:
            # one output bit from all bits of A (further Y bits, if any, are constant 0 in Yosys and not produced here)
            kind, a = YOSYS_REDUCE[ctype], [net(b) for b in conn["A"]]
            if len(a) == 1:
                a = a * 2          # the builder's gates take two inputs or more; x op x = x for AND/OR, handled for XOR below
                kind = {XOR: OR, XNOR: NOR}.get(kind, kind)
            circ.gates.append((kind, a, net(conn["Y"][0])))
        elif ctype == "$mux":      # Y = S ? B : A
            for j in range(len(conn["Y"])):
                circ.gates.append((MUX, [net(conn["S"][0]), net(conn["A"][j]), net(conn["B"][j])], net(conn["Y"][j])))
        elif ctype == "$tribuf":   # Y = EN ? A : z
            for j in range(len(conn["Y"])):
                circ.gates.append((BUF, [net(conn["A"][j]), net(conn["EN"][0])], net(conn["Y"][j])))
        elif ctype in ("$dff", "$dffe"):
            if int(str(params.get("CLK_POLARITY", "1")), 2) != 1 or int(str(params.get("EN_POLARITY", "1")), 2) != 1:
                raise ValueError(f"{ctype} with a negative clock or enable polarity is not supported ({cname})")
            en = net(conn["EN"][0]) if ctype == "$dffe" else net("1")
            for j in range(len(conn["Q"])):
                circ.gates.append((REG, [net(conn["D"][j]), en, net(conn["CLK"][0])], net(conn["Q"][j])))
        elif ctype in ("$sdff", "$sdffe", "$sdffce"):
            # synchronous reset: the value loaded is SRST ? SRST_VALUE : D; $sdffe resets whether enabled or not, $sdffce only
            # when enabled
            if polarity("CLK_POLARITY") != 1:
                raise ValueError(f"{ctype} with a negative clock polarity is not supported ({cname})")
            width = len(conn["Q"])
            rst = positive(conn["SRST"][0], "SRST_POLARITY")
            value = param_bits("SRST_VALUE", width)
            en = positive(conn["EN"][0], "EN_POLARITY") if ctype != "$sdff" else net("1")
            if ctype == "$sdffe":
                both = tmp()
                circ.gates.append((OR, [en, rst], both))
                en = both
            for j in range(width):
                d = tmp()
                circ.gates.append((MUX, [rst, net(conn["D"][j]), net(str(value[j]))], d))
                circ.gates.append((REG, [d, en, net(conn["CLK"][0])], net(conn["Q"][j])))
        elif ctype == "$dlatch":   # transparent while EN: Q = EN ? D : Q
            en = positive(conn["EN"][0], "EN_POLARITY")
            for j in range(len(conn["Q"])):
                q = net(conn["Q"][j])
                circ.gates.append((MUX, [en, q, net(conn["D"][j])], q))
        elif ctype == "$pmux":     # Y = A unless a select bit is set: then that slice of B (selects are one-hot in valid designs)
            width = len(conn["Y"])
            cur = [net(b) for b in conn["A"]]
            for i, sbit in enumerate(conn["S"]):
                last = i == len(conn["S"]) - 1
                nxt = [net(conn["Y"][j]) if last else tmp() for j in range(width)]
                for j in range(width):
                    circ.gates.append((MUX, [net(sbit), cur[j], net(conn["B"][i * width + j])], nxt[j]))
                cur = nxt
        elif ctype in ("$logic_and", "$logic_or"):
            a, b = tmp(), tmp()
            for src, dst in ((conn["A"], a), (conn["B"], b)):
                bits = [net(x) for x in src]
                circ.gates.append((OR, bits if len(bits) > 1 else bits * 2, dst))
            circ.gates.append((AND if ctype == "$logic_and" else OR, [a, b], net(conn["Y"][0])))
        elif ctype in ("$add", "$sub", "$mul", "$neg"):
            width = len(conn["Y"])
            if width > 255:
                raise ValueError(f"{ctype} wider than 255 bits ({cname})")
            if ctype == "$neg":
                ops = [[net("0")] * width, extend(conn["A"], "A_SIGNED", width)]
            else:
                ops = [extend(conn["A"], "A_SIGNED", width), extend(conn["B"], "B_SIGNED", width)]
            circ.wide.append(({"$add": "add", "$sub": "sub", "$mul": "mul", "$neg": "sub"}[ctype], ops, [net(b) for b in conn["Y"]]))
        elif ctype in YOSYS_COMPARE:
            width = max(len(conn["A"]), len(conn["B"]))
            signed = flag("A_SIGNED") and flag("B_SIGNED")
            op = YOSYS_COMPARE[ctype][1 if signed else 0]
            ops = [extend(conn["A"], "A_SIGNED" if signed else None, width), extend(conn["B"], "B_SIGNED" if signed else None, width)]
            circ.wide.append((("cmp", op), ops, [net(conn["Y"][0])]))   # further Y bits are constant 0 in Yosys and not produced
        elif ctype in ("$shl", "$sshl", "$shr", "$sshr"):
            width = len(conn["Y"])
            if len(conn["B"]) > 8:
                raise ValueError(f"{ctype} with a shift amount wider than 8 bits ({cname})")
            op = "shl" if ctype in ("$shl", "$sshl") else ("ashr" if ctype == "$sshr" and flag("A_SIGNED") else "lshr")
            circ.wide.append((op, [extend(conn["A"], "A_SIGNED", width), [net(b) for b in conn["B"]]], [net(b) for b in conn["Y"]]))
        elif ctype == "$sr":       # set / reset latch, reset wins: Q = CLR ? 0 : (SET ? 1 : Q)
            for j in range(len(conn["Q"])):
                q, hold = net(conn["Q"][j]), tmp()
                circ.gates.append((MUX, [positive(conn["SET"][j], "SET_POLARITY"), q, net("1")], hold))
                circ.gates.append((MUX, [positive(conn["CLR"][j], "CLR_POLARITY"), hold, net("0")], q))
        elif ctype == "$pos":      # Y = A, extended to the width of Y
            for j, src in enumerate(extend(conn["A"], "A_SIGNED", len(conn["Y"]))):
                circ.gates.append((OR, [src, src], net(conn["Y"][j])))
        elif ctype == "$mem_v2":
            _lower_mem_v2(circ, cname, conn, params, net, tmp, flag, param_bits)
        else:
            raise ValueError(f"unsupported cell type {ctype!r} ({cname})")
    circ.n_nets = len(net_of)
    return circ


def netlist_to_yosys(gen, name: str, in_ports: dict, out_ports: dict) -> str:
    """Writes a two-input gate list (netgen.GeneratedNetlist with And/Or/Xor/Not only) as Yosys JSON
    with 1-bit cells; wire id w becomes bit id w + 2 (0 and 1 are constants in Yosys)."""
    names = {AND: "$and", OR: "$or", XOR: "$xor", NOT: "$not"}
    cells = {}
    for i in range(gen.n_gates):
        k = int(gen.kinds[i])
        if k not in names:
            raise ValueError("only $and/$or/$xor/$not cells can be written")
        conn = {"A": [int(gen.in0[i]) + 2]}
        dirs = {"A": "input", "Y": "output"}
        if k != NOT:
            conn["B"] = [int(gen.in1[i]) + 2]
            dirs["B"] = "input"
        conn["Y"] = [int(gen.out[i]) + 2]
        cells[f"{names[k]}${i}"] = {"hide_name": 1, "type": names[k], "parameters": {}, "attributes": {},
                                     "port_directions": dirs, "connections": conn}
    ports = {p: {"direction": "input", "bits": [int(w) + 2 for w in ws]} for p, ws in in_ports.items()}
    ports.update({p: {"direction": "output", "bits": [int(w) + 2 for w in ws]} for p, ws in out_ports.items()})
    netnames = {p: {"hide_name": 0, "bits": v["bits"], "attributes": {}} for p, v in ports.items()}
    doc = {"creator": "digilogic_b200.importers", "modules": {name: {"attributes": {"top": "1".zfill(32)}, "ports": ports,
                                                                       "cells": cells, "netnames": netnames}}}
    return json.dumps(doc)


def multiplier_yosys(n: int = 32) -> str:
    """BASELINE config 2: n x n array multiplier as Yosys JSON (ports A, B: n bits, P: 2n bits)."""
    from . import netgen
    gen = netgen.array_multiplier(n)
    return netlist_to_yosys(gen, f"mul{n}x{n}", {"A": gen.inputs[:n], "B": gen.inputs[n:]}, {"P": gen.outputs})
