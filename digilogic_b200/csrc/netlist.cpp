// netlist.cpp — builder validation and the netlist compiler (host only, no CUDA).
#include "netlist.hpp"

#include <algorithm>
#include <numeric>

#include "../../include/digilogic_b200.h"

namespace b2 {

// ---------------------------------------------------------------------------- builder
// Validation order and error variants follow SURVEY.md A.2 / the reference's mapping table
// crates/digilogic_gsim/src/lib.rs:55-66.

int Builder::add_wire(uint8_t w, uint32_t *out) {
    return add_wires(w, 1, out);
}

int Builder::add_wires(uint8_t w, uint32_t n, uint32_t *first) {
    if (w == 0) return B2_ERR_INVALID_ARG;
    if (width.size() + (uint64_t)n > 0xFFFFFFFFull || total_bits + (uint64_t)n * w > kMaxBits) return B2_ERR_OUT_OF_WIRES;
    if (first) *first = (uint32_t)width.size();
    width.insert(width.end(), n, w);
    total_bits += (uint64_t)n * w;
    return B2_OK;
}

int Builder::wire_width(uint32_t wire, uint8_t *out) const {
    if (wire >= width.size()) return B2_ERR_INVALID_WIRE_ID;
    *out = width[wire];
    return B2_OK;
}

static int push(Builder &b, uint8_t kind, const uint32_t *in, size_t n, uint32_t select, uint32_t output, uint32_t *cell) {
    if (b.comps.size() >= 0xFFFFFFFFull || b.inputs.size() + n >= 0xFFFFFFFFull) return B2_ERR_TOO_MANY_COMPONENTS;
    // what compile() makes of this component: 1-bit gates and their inputs (same case analysis as its bit-blast loop)
    const uint64_t wout = b.width[output];
    uint64_t bit_gates = wout, bit_inputs;
    switch (kind) {
        case K_MUX: bit_inputs = wout * (n + b.width[select]); break;
        case K_BUF: bit_inputs = wout * 2; break;
        case K_ADD:
        case K_SUB:
        case K_MUL: bit_inputs = wout * 2 * wout; break;
        case K_CMP: bit_inputs = 2ull * b.width[in[0]]; break;
        case K_SHL:
        case K_LSHR:
        case K_ASHR: bit_inputs = wout * (wout + b.width[in[1]]); break;
        case K_HGATE: bit_inputs = b.width[in[0]] > 2 ? b.width[in[0]] : 2; break;
        case K_REG:
            bit_gates = wout + 1;
            bit_inputs = wout * 4 + 1;
            break;
        case K_SLICE:
        case K_MERGE: bit_inputs = wout; break;
        default: bit_inputs = wout * n; break;
    }
    if (b.total_bit_gates + bit_gates > Builder::kMaxBitGates || b.total_bit_inputs + bit_inputs > Builder::kMaxBitInputs)
        return B2_ERR_TOO_MANY_COMPONENTS;
    b.total_bit_gates += bit_gates;
    b.total_bit_inputs += bit_inputs;
    Component c;
    c.kind = kind;
    c.n_inputs = (uint32_t)n;
    c.in_off = (uint32_t)b.inputs.size();
    c.select = select;
    c.output = output;
    b.inputs.insert(b.inputs.end(), in, in + n);
    if (cell) *cell = (uint32_t)b.comps.size();
    b.comps.push_back(c);
    return B2_OK;
}

int Builder::add_gate(int kind, const uint32_t *in, size_t n, uint32_t output, uint32_t *cell) {
    if (kind < K_AND || kind > K_XNOR) return B2_ERR_INVALID_ARG;
    if (output >= width.size()) return B2_ERR_INVALID_WIRE_ID;
    for (size_t i = 0; i < n; i++)
        if (in[i] >= width.size()) return B2_ERR_INVALID_WIRE_ID;
    if (n < 2) return B2_ERR_TOO_FEW_INPUTS;
    for (size_t i = 0; i < n; i++)
        if (width[in[i]] != width[output]) return B2_ERR_WIRE_WIDTH_MISMATCH;
    return push(*this, (uint8_t)kind, in, n, 0, output, cell);
}

int Builder::add_not(uint32_t in, uint32_t output, uint32_t *cell) {
    if (output >= width.size() || in >= width.size()) return B2_ERR_INVALID_WIRE_ID;
    if (width[in] != width[output]) return B2_ERR_WIRE_WIDTH_MISMATCH;
    return push(*this, K_NOT, &in, 1, 0, output, cell);
}

int Builder::add_mux(const uint32_t *in, size_t n, uint32_t select, uint32_t output, uint32_t *cell) {
    if (output >= width.size() || select >= width.size()) return B2_ERR_INVALID_WIRE_ID;
    for (size_t i = 0; i < n; i++)
        if (in[i] >= width.size()) return B2_ERR_INVALID_WIRE_ID;
    if (n < 2 || (n & (n - 1)) != 0) return B2_ERR_INVALID_INPUT_COUNT;
    uint32_t k = 0;
    while (((size_t)1 << k) < n) k++;
    if (width[select] != k) return B2_ERR_WIRE_WIDTH_INCOMPATIBLE;
    for (size_t i = 0; i < n; i++)
        if (width[in[i]] != width[output]) return B2_ERR_WIRE_WIDTH_MISMATCH;
    return push(*this, K_MUX, in, n, select, output, cell);
}

// gsim SimulatorBuilder::add_buffer (tri-state; no call site in the reference — SURVEY.md 8f row f4)
int Builder::add_buffer(uint32_t in, uint32_t enable, uint32_t output, uint32_t *cell) {
    if (output >= width.size() || in >= width.size() || enable >= width.size()) return B2_ERR_INVALID_WIRE_ID;
    if (width[in] != width[output]) return B2_ERR_WIRE_WIDTH_MISMATCH;
    if (width[enable] != 1) return B2_ERR_WIRE_WIDTH_INCOMPATIBLE;
    const uint32_t both[2] = {in, enable};
    return push(*this, K_BUF, both, 2, 0, output, cell);
}

// gsim SimulatorBuilder::add_slice / add_merge (bit re-wiring; no call site in the reference, but
// `OffsetOutOfRange` of its error map, crates/digilogic_gsim/src/lib.rs:55-66, belongs to add_slice)
int Builder::add_slice(uint32_t in, uint32_t offset, uint32_t output, uint32_t *cell) {
    if (output >= width.size() || in >= width.size()) return B2_ERR_INVALID_WIRE_ID;
    if ((uint64_t)offset + width[output] > width[in]) return B2_ERR_OFFSET_OUT_OF_RANGE;
    return push(*this, K_SLICE, &in, 1, offset, output, cell);
}

int Builder::add_merge(const uint32_t *in, size_t n, uint32_t output, uint32_t *cell) {
    if (output >= width.size()) return B2_ERR_INVALID_WIRE_ID;
    for (size_t i = 0; i < n; i++)
        if (in[i] >= width.size()) return B2_ERR_INVALID_WIRE_ID;
    if (n < 1) return B2_ERR_TOO_FEW_INPUTS;
    size_t total = 0;
    for (size_t i = 0; i < n; i++) total += width[in[i]];
    if (total != width[output]) return B2_ERR_WIRE_WIDTH_MISMATCH;
    return push(*this, K_MERGE, in, n, 0, output, cell);
}

// gsim SimulatorBuilder::add_add / add_sub(input_a, input_b, output): wrapping sum / difference of two wires of the output's
// width, Undefined if any input bit is not a valid 0/1 (the Yosys importer's Add/Sub arms are todo!(), yosys.rs:188-189)
int Builder::add_arith(int kind, uint32_t a, uint32_t b2, uint32_t output, uint32_t *cell) {
    if (kind != K_ADD && kind != K_SUB && kind != K_MUL) return B2_ERR_INVALID_ARG;
    if (output >= width.size() || a >= width.size() || b2 >= width.size()) return B2_ERR_INVALID_WIRE_ID;
    if (width[a] != width[output] || width[b2] != width[output]) return B2_ERR_WIRE_WIDTH_MISMATCH;
    const uint32_t in[2] = {a, b2};
    return push(*this, (uint8_t)kind, in, 2, 0, output, cell);
}

// gsim's comparators add_compare_{equal,not_equal,less_than,greater_than,less_than_or_equal,greater_than_or_equal}{,_signed}
// (the Yosys importer's Lt/Le/Eq/Ne/Ge/Gt arms are todo!(), crates/digilogic_serde/src/yosys.rs:182-187): inputs of one width,
// a one-bit output, Undefined if any input bit is not a valid 0/1
int Builder::add_compare(int op, uint32_t a, uint32_t b2, uint32_t output, uint32_t *cell) {
    if (op < CMP_EQ || op > CMP_GES) return B2_ERR_INVALID_ARG;
    if (output >= width.size() || a >= width.size() || b2 >= width.size()) return B2_ERR_INVALID_WIRE_ID;
    if (width[a] != width[b2]) return B2_ERR_WIRE_WIDTH_MISMATCH;
    if (width[output] != 1) return B2_ERR_WIRE_WIDTH_INCOMPATIBLE;
    const uint32_t in[2] = {a, b2};
    return push(*this, K_CMP, in, 2, (uint32_t)op, output, cell);
}

// gsim's shifters add_left_shift / add_logical_right_shift / add_arithmetic_right_shift(input, shift_amount, output) (Yosys
// Shl/Sshl/Shr/Sshr are todo!(), yosys.rs:174-177): output as wide as the input, the amount an unsigned value of at most 8 bits;
// amounts >= width shift everything out; Undefined if any bit of either input is not a valid 0/1
int Builder::add_shift(int kind, uint32_t a, uint32_t amount, uint32_t output, uint32_t *cell) {
    if (kind != K_SHL && kind != K_LSHR && kind != K_ASHR) return B2_ERR_INVALID_ARG;
    if (output >= width.size() || a >= width.size() || amount >= width.size()) return B2_ERR_INVALID_WIRE_ID;
    if (width[a] != width[output]) return B2_ERR_WIRE_WIDTH_MISMATCH;
    if (width[amount] > 8) return B2_ERR_WIRE_WIDTH_INCOMPATIBLE;
    const uint32_t in[2] = {a, amount};
    return push(*this, (uint8_t)kind, in, 2, 0, output, cell);
}

// gsim SimulatorBuilder::add_horizontal_{and,or,xor,nand,nor,xnor}_gate(input, output): reduction over the bits of one wire
// (the Yosys importer's ReduceAnd/Or/Xor/Xnor/Bool arms are todo!(), crates/digilogic_serde/src/yosys.rs:163-167)
int Builder::add_horizontal_gate(int kind, uint32_t in, uint32_t output, uint32_t *cell) {
    if (kind < K_AND || kind > K_XNOR) return B2_ERR_INVALID_ARG;
    if (output >= width.size() || in >= width.size()) return B2_ERR_INVALID_WIRE_ID;
    if (width[output] != 1) return B2_ERR_WIRE_WIDTH_INCOMPATIBLE;
    return push(*this, K_HGATE, &in, 1, (uint32_t)kind, output, cell);
}

// gsim SimulatorBuilder::add_register(data_in, data_out, enable, clock) — no call site in the reference (the Yosys
// importer's Dff/Dffe arms are todo!(), crates/digilogic_serde/src/yosys.rs:199-200); SURVEY.md 8f row f4.
int Builder::add_register(uint32_t data_in, uint32_t data_out, uint32_t enable, uint32_t clock, uint32_t *cell) {
    if (data_in >= width.size() || data_out >= width.size() || enable >= width.size() || clock >= width.size())
        return B2_ERR_INVALID_WIRE_ID;
    if (width[data_in] != width[data_out]) return B2_ERR_WIRE_WIDTH_MISMATCH;
    if (width[enable] != 1 || width[clock] != 1) return B2_ERR_WIRE_WIDTH_INCOMPATIBLE;
    const uint32_t in[3] = {data_in, enable, clock};
    return push(*this, K_REG, in, 3, 0, data_out, cell);
}

// ---------------------------------------------------------------------------- compiler

namespace {
// A 1-bit gate before ordering. Inputs are bit-nets.
struct BitGate {
    uint8_t kind;
    uint8_t nsel;
    uint8_t aux;
    uint32_t in_off, n_in;  // into bit_inputs (mux: select bits first)
    uint32_t out;           // bit-net, or NO_NET for a hidden gate (it has a row but drives no wire)
    uint32_t level;
};
const uint32_t NO_NET = 0xFFFFFFFFu;
const uint32_t HIDDEN = 0x80000000u;  // a bit_inputs entry with this bit set names the row of gate (entry & ~HIDDEN)
}  // namespace

void compile(const Builder &b, Compiled &c) {
    const uint32_t W = (uint32_t)b.width.size();
    c.width = b.width;
    c.sem_flags = b.sem_flags;
    c.bit_off.resize((size_t)W + 1);
    uint32_t nb = 0;
    for (uint32_t w = 0; w < W; w++) {
        c.bit_off[w] = nb;
        nb += b.width[w];
    }
    c.bit_off[W] = nb;

    // ---- bit-blast
    std::vector<BitGate> gates;
    std::vector<uint32_t> bin;
    size_t est = 0;
    for (const Component &cp : b.comps) est += b.width[cp.output];
    gates.reserve(est);
    c.has_state = false;
    for (const Component &cp : b.comps) {
        const uint32_t *in = b.inputs.data() + cp.in_off;
        const uint32_t wout = b.width[cp.output];
        uint32_t prev_clock_gate = 0;
        if (cp.kind == K_REG) {
            // the register's view of its clock one step ago: a hidden raw copy of the clock wire
            c.has_state = true;
            BitGate h;
            h.kind = K_RAW;
            h.nsel = 0;
            h.aux = 0;
            h.in_off = (uint32_t)bin.size();
            h.n_in = 1;
            h.out = NO_NET;
            h.level = 0;
            bin.push_back(c.bit_off[in[2]]);
            prev_clock_gate = (uint32_t)gates.size();
            gates.push_back(h);
        }
        for (uint32_t j = 0; j < wout; j++) {
            BitGate g;
            g.kind = cp.kind;
            g.nsel = 0;
            g.aux = 0;
            g.in_off = (uint32_t)bin.size();
            g.out = c.bit_off[cp.output] + j;
            g.level = 0;
            if (cp.kind == K_MUX) {
                const uint32_t k = b.width[cp.select];
                g.nsel = (uint8_t)k;
                for (uint32_t s = 0; s < k; s++) bin.push_back(c.bit_off[cp.select] + s);
            }
            if (cp.kind == K_BUF) {
                bin.push_back(c.bit_off[in[0]] + j);  // data bit j
                bin.push_back(c.bit_off[in[1]]);      // the one enable bit
            } else if (cp.kind == K_ADD || cp.kind == K_SUB || cp.kind == K_MUL) {
                // bit j of the result: needs every input bit for the validity and bits 0..j for the carries
                g.nsel = (uint8_t)j;
                for (uint32_t t = 0; t < wout; t++) bin.push_back(c.bit_off[in[0]] + t);
                for (uint32_t t = 0; t < wout; t++) bin.push_back(c.bit_off[in[1]] + t);
            } else if (cp.kind == K_CMP) {
                g.aux = (uint8_t)cp.select;
                for (uint32_t t = 0; t < b.width[in[0]]; t++) bin.push_back(c.bit_off[in[0]] + t);
                for (uint32_t t = 0; t < b.width[in[1]]; t++) bin.push_back(c.bit_off[in[1]] + t);
            } else if (cp.kind == K_SHL || cp.kind == K_LSHR || cp.kind == K_ASHR) {
                g.nsel = (uint8_t)j;
                g.aux = b.width[in[1]];
                for (uint32_t t = 0; t < wout; t++) bin.push_back(c.bit_off[in[0]] + t);
                for (uint32_t t = 0; t < b.width[in[1]]; t++) bin.push_back(c.bit_off[in[1]] + t);
            } else if (cp.kind == K_HGATE) {
                const uint32_t wi = b.width[in[0]];
                if (wi == 1) {  // one bit: the bit itself (High-Z read as X), inverted for NAND / NOR / XNOR
                    if (cp.select >= K_NAND) {
                        g.kind = K_NOT;
                    } else {  // OR(a, a): never High-Z, whatever SEM_COPY_Z says about slices and merges
                        g.kind = K_OR;
                        bin.push_back(c.bit_off[in[0]]);
                    }
                    bin.push_back(c.bit_off[in[0]]);
                } else {
                    g.kind = (uint8_t)cp.select;
                    for (uint32_t t = 0; t < wi; t++) bin.push_back(c.bit_off[in[0]] + t);
                }
            } else if (cp.kind == K_REG) {
                bin.push_back(c.bit_off[in[0]] + j);  // data bit j
                bin.push_back(c.bit_off[in[1]]);      // enable
                bin.push_back(c.bit_off[in[2]]);      // clock
                bin.push_back(HIDDEN | prev_clock_gate);
            } else if (cp.kind == K_SLICE) {
                g.kind = K_COPY;
                bin.push_back(c.bit_off[in[0]] + cp.select + j);
            } else if (cp.kind == K_MERGE) {
                g.kind = K_COPY;
                uint32_t k = 0, lo = 0;  // the input that holds output bit j
                while (lo + b.width[in[k]] <= j) lo += b.width[in[k++]];
                bin.push_back(c.bit_off[in[k]] + (j - lo));
            } else {
                for (uint32_t i = 0; i < cp.n_inputs; i++) bin.push_back(c.bit_off[in[i]] + j);
            }
            g.n_in = (uint32_t)bin.size() - g.in_off;
            gates.push_back(g);
        }
    }
    const uint32_t G = (uint32_t)gates.size();

    // ---- drivers
    const uint32_t NONE = 0xFFFFFFFFu;
    // A net with a buffer among its drivers is a tri-state net: it may have any number of drivers and is
    // resolved per step (k_resolve).  Two drivers that are never High-Z on one net conflict in every run.
    std::vector<uint32_t> driver(nb, NONE), n_firm(nb, 0);
    std::vector<uint8_t> tri(nb, 0);
    c.multi_driver = false;
    c.has_tristate = false;
    for (uint32_t g = 0; g < G; g++) {
        const uint32_t bn = gates[g].out;
        if (bn == NO_NET) continue;
        // components whose output can be High-Z: buffers, and under the A.6-4 / A.6-7 alternatives multiplexers / slices and merges
        const uint8_t gk = gates[g].kind;
        if (gk == K_BUF || (gk == K_MUX && (b.sem_flags & SEM_MUX_Z)) || (gk == K_COPY && (b.sem_flags & SEM_COPY_Z))) {
            tri[bn] = 1;
            c.has_tristate = true;
        } else if (++n_firm[bn] > 1) {
            c.multi_driver = true;
        }
        if (driver[bn] == NONE) driver[bn] = g;
    }
    c.bit_driven.assign(nb, 0);
    for (uint32_t i = 0; i < nb; i++) c.bit_driven[i] = driver[i] != NONE;

    // ---- levelise (Kahn). consumers CSR over bit-nets that are gate driven.
    std::vector<uint32_t> indeg(G, 0), cons_off((size_t)nb + 2, 0);
    for (uint32_t g = 0; g < G; g++)
        for (uint32_t i = 0; i < gates[g].n_in; i++) {
            uint32_t bn = bin[gates[g].in_off + i];
            if (bn & HIDDEN) continue;  // registers: the netlist is evaluated step by step anyway (has_state)
            if (driver[bn] != NONE) {
                indeg[g]++;
                cons_off[bn + 2]++;
            }
        }
    for (uint32_t i = 0; i < nb; i++) cons_off[i + 2] += cons_off[i + 1];
    std::vector<uint32_t> cons(cons_off[nb + 1]);
    for (uint32_t g = 0; g < G; g++)
        for (uint32_t i = 0; i < gates[g].n_in; i++) {
            uint32_t bn = bin[gates[g].in_off + i];
            if (bn & HIDDEN) continue;
            if (driver[bn] != NONE) cons[cons_off[bn + 1]++] = g;
        }
    std::vector<uint32_t> queue;
    queue.reserve(G);
    for (uint32_t g = 0; g < G; g++)
        if (indeg[g] == 0) {
            gates[g].level = 1;
            queue.push_back(g);
        }
    uint32_t depth = 0;
    for (size_t h = 0; h < queue.size(); h++) {
        const uint32_t g = queue[h];
        depth = std::max(depth, gates[g].level);
        // only the registered driver of a net propagates (multi-driver netlists never simulate)
        if (gates[g].out == NO_NET || driver[gates[g].out] != g) continue;
        const uint32_t bn = gates[g].out;
        for (uint32_t k = cons_off[bn]; k < cons_off[bn + 1]; k++) {
            const uint32_t h2 = cons[k];
            gates[h2].level = std::max(gates[h2].level, gates[g].level + 1);
            if (--indeg[h2] == 0) queue.push_back(h2);
        }
    }
    c.cyclic = queue.size() != G || c.has_tristate || c.has_state;
    c.depth = c.cyclic ? 0 : depth;
    if (c.cyclic)
        for (auto &g : gates) g.level = 1;  // one pseudo-level: unit-delay evaluation only

    // ---- order: class (2-input first), then level, then kind, then creation order (stable)
    std::vector<uint32_t> order(G);
    std::iota(order.begin(), order.end(), 0u);
    // class 0: one/two inputs, class 1: three inputs (both on the fast kernels), class 2: CSR kernel
    auto cls = [&](uint32_t g) -> int {
        const uint8_t k = gates[g].kind;
        return k == K_MUX || k == K_REG || (k >= K_ADD && k <= K_ASHR) || gates[g].n_in > 3 ? 2 : (gates[g].n_in == 3 ? 1 : 0);
    };
    auto is_wide = [&](uint32_t g) { return cls(g) == 2; };
    // acyclic: lifetime class of the row of gate g — 0: read two or more levels above, 1: last reader in the level directly
    // above, 2: no reader
    std::vector<uint8_t> short_lived(G, 0);
    if (!c.cyclic)
        for (uint32_t g = 0; g < G; g++) {
            const uint32_t bn = gates[g].out;
            uint32_t last = 0;
            if (bn != NO_NET && driver[bn] == g)
                for (uint32_t k = cons_off[bn]; k < cons_off[bn + 1]; k++) last = std::max(last, gates[cons[k]].level);
            short_lived[g] = last == 0 ? 2 : (last <= gates[g].level + 1 ? 1 : 0);
        }
    std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) {
        const int wx = cls(x), wy = cls(y);
        if (wx != wy) return wx < wy;
        if (gates[x].level != gates[y].level) return gates[x].level < gates[y].level;
        if (short_lived[x] != short_lived[y]) return short_lived[x] < short_lived[y];
        return gates[x].kind < gates[y].kind;
    });

    // ---- rows: sources first (in bit-net order), then gates in `order`
    c.row_of_bit.assign(nb, 0);
    uint32_t row = 0;
    for (uint32_t i = 0; i < nb; i++)
        if (driver[i] == NONE) c.row_of_bit[i] = row++;
    c.n_src = row;
    std::vector<uint32_t> gate_row(G);
    for (uint32_t k = 0; k < G; k++) gate_row[order[k]] = c.n_src + k;
    for (uint32_t i = 0; i < nb; i++)
        if (driver[i] != NONE) c.row_of_bit[i] = gate_row[driver[i]];
    c.n_rows = c.n_src + G;
    // tri-state nets: the wire lives in a resolved row behind the gate rows; its drivers keep their own rows
    c.res_off.clear();
    c.res_in.clear();
    c.res_of_bit.assign(nb, -1);
    if (c.has_tristate) {
        std::vector<uint32_t> cnt((size_t)nb + 1, 0);
        for (uint32_t g = 0; g < G; g++)
            if (gates[g].out != NO_NET && tri[gates[g].out]) cnt[gates[g].out]++;
        c.res_off.push_back(0);
        for (uint32_t i = 0; i < nb; i++)
            if (tri[i]) {
                c.res_of_bit[i] = (int32_t)c.res_off.size() - 1;
                c.row_of_bit[i] = c.n_rows++;
                c.res_off.push_back(c.res_off.back() + cnt[i]);
            }
        c.res_in.resize(c.res_off.back());
        std::vector<uint32_t> fill(c.res_off.begin(), c.res_off.end() - 1);
        for (uint32_t g = 0; g < G; g++)  // creation order, like gsim's driver lists
            if (gates[g].out != NO_NET && tri[gates[g].out]) c.res_in[fill[c.res_of_bit[gates[g].out]]++] = gate_row[g];
    }

    // ---- SoA emission
    auto row_of = [&](uint32_t entry) { return (entry & HIDDEN) ? gate_row[entry & ~HIDDEN] : c.row_of_bit[entry]; };
    uint32_t n2 = 0;
    while (n2 < G && !is_wide(order[n2])) n2++;
    c.fan0.resize(n2);
    c.fan1.resize(n2);
    c.fan2.resize(n2);
    c.kind2.resize(n2);
    c.n_two = 0;
    c.far2.assign(n2, 0);
    // level of the producer of a fan-in entry: 0 for a source
    auto level_of = [&](uint32_t entry) -> uint32_t {
        if (entry & HIDDEN) return gates[entry & ~HIDDEN].level;
        return driver[entry] == NONE ? 0u : gates[driver[entry]].level;
    };
    for (uint32_t k = 0; k < n2; k++) {
        const BitGate &g = gates[order[k]];
        if (!c.cyclic)
            for (uint32_t i = 0; i < g.n_in && i < 3; i++)
                if (level_of(bin[g.in_off + i]) + 1 < g.level) c.far2[k] |= (uint8_t)(1u << i);
        c.kind2[k] = g.kind;
        c.fan0[k] = row_of(bin[g.in_off]);
        c.fan1[k] = g.n_in > 1 ? row_of(bin[g.in_off + 1]) : c.fan0[k];
        c.fan2[k] = g.n_in > 2 ? row_of(bin[g.in_off + 2]) : c.fan0[k];
        if (g.n_in <= 2) c.n_two = k + 1;
    }
    const uint32_t nw = G - n2;
    c.wide_kind.resize(nw);
    c.wide_nsel.resize(nw);
    c.wide_aux.resize(nw);
    c.wide_off.assign((size_t)nw + 1, 0);
    c.wide_in.clear();
    for (uint32_t k = 0; k < nw; k++) {
        const BitGate &g = gates[order[n2 + k]];
        c.wide_kind[k] = g.kind;
        c.wide_nsel[k] = g.nsel;
        c.wide_aux[k] = g.aux;
        c.wide_off[k] = (uint32_t)c.wide_in.size();
        for (uint32_t i = 0; i < g.n_in; i++) c.wide_in.push_back(row_of(bin[g.in_off + i]));
    }
    c.wide_off[nw] = (uint32_t)c.wide_in.size();

    // ---- level table
    c.levels.clear();
    const uint32_t n_levels = c.cyclic ? (G ? 1u : 0u) : depth;
    c.levels.assign(n_levels, Level{0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0});
    {
        // one class: [begin, end) of every level, and the first short-lived gate inside it
        auto ranges = [&](uint32_t first, uint32_t last, uint32_t base, uint32_t Level::*begin, uint32_t Level::*end,
                          uint32_t Level::*shrt, uint32_t Level::*never) {
            uint32_t k = first;
            for (uint32_t l = 1; l <= n_levels; l++) {
                Level &lv = c.levels[l - 1];
                lv.*begin = k - base;
                while (k < last && gates[order[k]].level == l && short_lived[order[k]] == 0) k++;
                lv.*shrt = k - base;
                while (k < last && gates[order[k]].level == l && short_lived[order[k]] == 1) k++;
                lv.*never = k - base;
                while (k < last && gates[order[k]].level == l) k++;
                lv.*end = k - base;
            }
        };
        ranges(0, c.n_two, 0, &Level::g2_begin, &Level::g2_end, &Level::g2_short, &Level::g2_never);
        ranges(c.n_two, n2, 0, &Level::g3_begin, &Level::g3_end, &Level::g3_short, &Level::g3_never);
        ranges(n2, G, n2, &Level::wide_begin, &Level::wide_end, &Level::wide_short, &Level::wide_never);
    }
}

}  // namespace b2
