// netlist.hpp — host-side builder and netlist compiler.
//
// Builder mirrors the validation surface of gsim::SimulatorBuilder as the reference uses it
// (crates/digilogic_gsim/src/lib.rs:76-192; error variants lib.rs:55-66).  compile() is what
// stands where SimulatorBuilder::build() (lib.rs:127) stood: it bit-blasts wires into 1-bit
// rows, analyses drivers, levelises, sorts gates by (level, kind) and emits the SoA arrays the
// kernels stream.
#pragma once
#include <cstdint>
#include <vector>

namespace b2 {

enum Kind : uint8_t {
    K_AND = 0, K_OR = 1, K_XOR = 2, K_NAND = 3, K_NOR = 4, K_XNOR = 5, K_NOT = 6, K_MUX = 7,
    K_BUF = 8,   // tri-state buffer: inputs = {data, enable}; the only component whose output can be High-Z
    K_COPY = 9,  // one bit of a slice / merge: output = input, High-Z read as X
    K_SLICE = 10, K_MERGE = 11,  // builder-level only (bit-blasted into K_COPY)
    K_REG = 12,  // register bit: inputs = {data bit, enable, clock, previous-clock row}; reads its own row as the stored value
    K_RAW = 13,  // hidden one-input gate: raw copy (High-Z stays High-Z) — a register's view of its clock one step ago
    K_HGATE = 14, // builder-level only: reduction gate (kind in `select`), becomes an n-ary gate over the bits of its input
    K_ADD = 15, K_SUB = 16, // one sum / difference bit: inputs = {a_0..a_{w-1}, b_0..b_{w-1}}, bit index in the nsel field
    K_MUL = 17,  // one product bit: inputs as for K_ADD, bit index in nsel
    K_CMP = 18,  // comparator (one bit): inputs = {a_0..a_{w-1}, b_0..b_{w-1}}, CmpOp in the aux field
    K_SHL = 19, K_LSHR = 20, K_ASHR = 21  // one bit of a shifted wire: inputs = {a_0..a_{w-1}, amount_0..amount_{k-1}}, bit index in nsel, k in aux
};

enum CmpOp : uint8_t { CMP_EQ = 0, CMP_NE = 1, CMP_LTU = 2, CMP_GTU = 3, CMP_LEU = 4, CMP_GEU = 5, CMP_LTS = 6, CMP_GTS = 7, CMP_LES = 8, CMP_GES = 9 };

// The points of gsim's behaviour that were reconstructed without its source, one switch each, default
// off = the reading the CPU checker takes by default (its go_config switches; the product never links it).  They are properties of a netlist's
// components, set on the builder (b2_builder_set_option) and compiled in.
enum SemFlags : uint32_t {
    SEM_MUX_Z = 1,   // A.6-4: a multiplexer passes a High-Z selected input on as High-Z (default: as X)
    SEM_INIT_X = 2,  // A.6-5: wires and component outputs start Undefined (default: High-Z)
    SEM_BUF_Z = 4,   // A.6-6: an enabled buffer passes a High-Z input bit on as High-Z (default: as X)
    SEM_COPY_Z = 8   // A.6-7: slices and merges pass a High-Z input bit on as High-Z (default: as X)
};

struct Component {
    uint8_t kind;
    uint32_t n_inputs;  // data inputs
    uint32_t in_off;    // into Builder::inputs
    uint32_t select;    // mux: select wire; slice: bit offset; reduction gate: gate kind
    uint32_t output;
};

struct Builder {
    std::vector<uint8_t> width;
    std::vector<Component> comps;
    std::vector<uint32_t> inputs;
    // Running 64-bit totals of what compile() will bit-blast: the compiled arrays index bit-nets, 1-bit gates and
    // their inputs with 32 bits (bit 31 of a bit-net index is a flag), so the builder refuses to grow past these.
    uint32_t sem_flags = 0;  // SemFlags
    uint64_t total_bits = 0, total_bit_gates = 0, total_bit_inputs = 0;
    static constexpr uint64_t kMaxBits = 1ull << 30, kMaxBitGates = 1ull << 30, kMaxBitInputs = 1ull << 31;

    int add_wire(uint8_t w, uint32_t *out);
    int add_wires(uint8_t w, uint32_t n, uint32_t *first);
    int wire_width(uint32_t wire, uint8_t *out) const;
    int add_gate(int kind, const uint32_t *in, size_t n, uint32_t output, uint32_t *cell);
    int add_not(uint32_t in, uint32_t output, uint32_t *cell);
    int add_mux(const uint32_t *in, size_t n, uint32_t select, uint32_t output, uint32_t *cell);
    int add_buffer(uint32_t in, uint32_t enable, uint32_t output, uint32_t *cell);
    int add_slice(uint32_t in, uint32_t offset, uint32_t output, uint32_t *cell);
    int add_arith(int kind, uint32_t a, uint32_t b, uint32_t output, uint32_t *cell);  // K_ADD / K_SUB / K_MUL
    int add_compare(int op, uint32_t a, uint32_t b, uint32_t output, uint32_t *cell);
    int add_shift(int kind, uint32_t a, uint32_t amount, uint32_t output, uint32_t *cell);  // K_SHL / K_LSHR / K_ASHR
    int add_horizontal_gate(int kind, uint32_t in, uint32_t output, uint32_t *cell);
    int add_register(uint32_t data_in, uint32_t data_out, uint32_t enable, uint32_t clock, uint32_t *cell);
    int add_merge(const uint32_t *in, size_t n, uint32_t output, uint32_t *cell);
};

// One level of the levelised order: one/two-input gates [g2_begin, g2_end), three-input gates
// [g3_begin, g3_end) (both index the fast arrays fan0/fan1/fan2/kind2; all three-input gates come
// after all one/two-input gates) and wide gates [wide_begin, wide_end).  Output rows are implicit:
//   row(fast gate i) = n_src + i,   row(wide gate j) = n_src + n_g2 + n_g3 + j.
struct Level {
    uint32_t g2_begin, g2_end, g3_begin, g3_end, wide_begin, wide_end;
    // Inside a level the gates of a class are ordered by the lifetime of their rows: [begin, short) are read two or more
    // levels above, [short, never) only by the level directly above, [never, end) by no gate at all.  Once their last
    // reader has run, such rows are dead to a level sweep; the batch runner uses this to drop them from L2 instead of letting
    // them be written to HBM (Engine::set_discard).
    uint32_t g2_short, g3_short, wide_short;
    uint32_t g2_never, g3_never, wide_never;
};

struct Compiled {
    // wires as the caller numbered them
    std::vector<uint8_t> width;
    std::vector<uint32_t> bit_off;     // wire -> first bit-net (size n_wires+1)
    std::vector<uint32_t> row_of_bit;  // bit-net -> store row
    std::vector<uint8_t> bit_driven;   // bit-net has a gate driver
    uint32_t n_rows = 0, n_src = 0;
    uint32_t sem_flags = 0;  // SemFlags of the builder
    // fast gates (one to three inputs), in row order: n_g2 one/two-input gates, then n_g3 three-input gates
    std::vector<uint32_t> fan0, fan1, fan2;  // store rows (fan2 = fan0 for gates with fewer inputs)
    std::vector<uint8_t> kind2;
    // acyclic netlists: bit i set = fan-in i was produced more than one level below the gate (or is a source below level 1),
    // i.e. long ago in a level sweep: the kernels read it with an L2 evict-first hint, so that the stream of such one-off
    // reads does not push the rows the previous level has just written out of L2
    std::vector<uint8_t> far2;
    uint32_t n_two = 0;
    // wide gates (n-ary, mux slices), CSR over store rows; mux: first n_sel entries are select bits
    std::vector<uint32_t> wide_off, wide_in;
    std::vector<uint8_t> wide_kind, wide_nsel, wide_aux;  // aux: comparator op / width of a shift amount
    std::vector<Level> levels;         // acyclic: one per level (1..depth); cyclic: a single entry
    uint32_t depth = 0;
    // cyclic: unit-delay evaluation only — feedback, or tri-state nets (their conflicts can be transient,
    // so the steps have to be walked exactly).  multi_driver: some net has two drivers that are never
    // High-Z, so every run conflicts.
    bool cyclic = false, multi_driver = false, has_tristate = false, has_state = false;  // has_state: registers
    // Tri-state nets (every net with a buffer among its drivers): each driver keeps its own output row
    // (the gate's implicit row) and the wire gets a resolved row  n_src + n_gates + r  computed by
    // k_resolve from the driver rows res_in[res_off[r] .. res_off[r+1]) and the net's external drive.
    std::vector<uint32_t> res_off, res_in;
    std::vector<int32_t> res_of_bit;   // bit-net -> index of its resolved row entry, or -1
    uint32_t n_res() const { return res_off.empty() ? 0 : (uint32_t)res_off.size() - 1; }

    uint32_t n_wires() const { return (uint32_t)width.size(); }
    uint32_t n_g2() const { return n_two; }
    uint32_t n_g3() const { return (uint32_t)kind2.size() - n_two; }
    uint32_t n_fast() const { return (uint32_t)kind2.size(); }
    uint32_t n_wide() const { return (uint32_t)wide_kind.size(); }
};

// Never fails on a validated Builder.
void compile(const Builder &b, Compiled &out);

}  // namespace b2
