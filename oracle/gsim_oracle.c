/*
 * gsim_oracle.c — CPU restatement of the gsim 1.1.4 simulation algorithm.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (digilogic_b200/)
 * may link, load or call this file.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs use it, as the checker
 * or as the reported CPU baseline.
 *
 * PARITY UNPINNED: the arithmetic of the reference's hot path lives in the
 * third-party crate `gsim = "1.1.4"` (reference crates/digilogic_gsim/Cargo.toml:15,
 * Cargo.lock:3920-3923), whose source is not vendored under /root/reference
 * and cannot be built here (no cargo/rustc, no network).  The reference holds
 * no test that builds or evaluates a circuit (crates/digilogic_gsim has zero
 * tests).  This file restates gsim's published algorithm as reconstructed in
 * SURVEY.md Appendix A and anchors on the reference's call sites:
 *
 *   SimulatorBuilder::default / add_wire     crates/digilogic_gsim/src/lib.rs:12,135-139
 *   get_wire_width                           crates/digilogic_gsim/src/lib.rs:87-92
 *   add_{and,or,xor,nand,nor,xnor}_gate      crates/digilogic_gsim/src/lib.rs:76-99,141-146
 *   add_not_gate                             crates/digilogic_gsim/src/lib.rs:148-167
 *   add_multiplexer                          crates/digilogic_gsim/src/lib.rs:169-192
 *   build                                    crates/digilogic_gsim/src/lib.rs:123-133
 *   LogicState::from_bit_planes/set_wire_drive  crates/digilogic_gsim/src/lib.rs:194-214
 *   run_sim                                  crates/digilogic_gsim/src/lib.rs:216-219
 *   get_wire_state / to_bit_planes           crates/digilogic_gsim/src/lib.rs:221-238
 *   plane encoding (0/1/Z/X)                 crates/digilogic/src/ui/palette.rs:326-341,
 *                                            crates/digilogic_core/src/components.rs:104-110
 *
 * Structure follows gsim: one simulator = one stimulus vector; per-wire
 * `drive` and `state`, per-component `output`; a wire-update queue and a
 * component-update queue, each sorted + deduplicated per phase; begin_sim /
 * step_sim / run_sim(max_steps).  The five reconstructed points of
 * SURVEY.md A.6 are the switches in go_config below.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#include "gsim_oracle.h"

static go_config_t g_cfg = {0, 0, 0, 0, 0, 0};

void go_set_config(int steps_check_ge, int conflict_needs_disagree, int mux_z_passthrough, int initial_undefined) {
    g_cfg.steps_check_ge = steps_check_ge;
    g_cfg.conflict_needs_disagree = conflict_needs_disagree;
    g_cfg.mux_z_passthrough = mux_z_passthrough;
    g_cfg.initial_undefined = initial_undefined;
}
void go_set_buffer_z_passthrough(int on) { g_cfg.buffer_z_passthrough = on; }
void go_set_copy_z_passthrough(int on) { g_cfg.copy_z_passthrough = on; }

/* ------------------------------------------------------------------ builder */

go_builder *go_builder_new(void) { return (go_builder *)calloc(1, sizeof(go_builder)); }

void go_builder_free(go_builder *b) {
    if (!b) return;
    free(b->width); free(b->comps); free(b->inputs); free(b);
}

/* add_wire(width): returns the new id, or -1 when ids are exhausted (-> OutOfResources, lib.rs:138) */
int64_t go_add_wire(go_builder *b, uint8_t width) {
    if (width == 0) return -1;
    if (b->n_wires == UINT32_MAX) return -1;
    if (b->n_wires == b->cap_wires) {
        b->cap_wires = b->cap_wires ? b->cap_wires * 2 : 1024;
        b->width = (uint8_t *)realloc(b->width, b->cap_wires);
    }
    b->width[b->n_wires] = width;
    return (int64_t)b->n_wires++;
}

int go_add_wires(go_builder *b, uint8_t width, uint32_t n, uint32_t *first) {
    if (first) *first = b->n_wires;
    for (uint32_t i = 0; i < n; i++) if (go_add_wire(b, width) < 0) return GO_ERR_TOO_MANY_COMPONENTS;
    return GO_OK;
}

int go_builder_wire_width(const go_builder *b, uint32_t wire, uint8_t *out) {
    if (wire >= b->n_wires) return GO_ERR_INVALID_WIRE_ID;
    *out = b->width[wire];
    return GO_OK;
}

static int push_comp(go_builder *b, uint8_t kind, const uint32_t *inputs, uint32_t n, uint32_t select, uint32_t output, uint32_t *cell) {
    if (b->n_comps == UINT32_MAX) return GO_ERR_TOO_MANY_COMPONENTS;
    if (b->n_comps == b->cap_comps) {
        b->cap_comps = b->cap_comps ? b->cap_comps * 2 : 1024;
        b->comps = (comp_t *)realloc(b->comps, (size_t)b->cap_comps * sizeof(comp_t));
    }
    while (b->n_inputs + n > b->cap_inputs) {
        b->cap_inputs = b->cap_inputs ? b->cap_inputs * 2 : 4096;
        b->inputs = (uint32_t *)realloc(b->inputs, (size_t)b->cap_inputs * sizeof(uint32_t));
    }
    comp_t *c = &b->comps[b->n_comps];
    c->kind = kind; c->n_inputs = n; c->in_off = b->n_inputs; c->select = select; c->output = output;
    memcpy(b->inputs + b->n_inputs, inputs, (size_t)n * sizeof(uint32_t));
    b->n_inputs += n;
    if (cell) *cell = b->n_comps;
    b->n_comps++;
    return GO_OK;
}

/* add_{and,or,xor,nand,nor,xnor}_gate(inputs, output) — SURVEY.md A.2 validation order */
int go_add_gate(go_builder *b, int kind, const uint32_t *inputs, uint32_t n, uint32_t output, uint32_t *cell) {
    if (kind < GO_AND || kind > GO_XNOR) return GO_ERR_INVALID_INPUT_COUNT;
    if (output >= b->n_wires) return GO_ERR_INVALID_WIRE_ID;
    for (uint32_t i = 0; i < n; i++) if (inputs[i] >= b->n_wires) return GO_ERR_INVALID_WIRE_ID;
    if (n < 2) return GO_ERR_TOO_FEW_INPUTS;
    for (uint32_t i = 0; i < n; i++) if (b->width[inputs[i]] != b->width[output]) return GO_ERR_WIRE_WIDTH_MISMATCH;
    return push_comp(b, (uint8_t)kind, inputs, n, 0, output, cell);
}

int go_add_not(go_builder *b, uint32_t input, uint32_t output, uint32_t *cell) {
    if (output >= b->n_wires || input >= b->n_wires) return GO_ERR_INVALID_WIRE_ID;
    if (b->width[input] != b->width[output]) return GO_ERR_WIRE_WIDTH_MISMATCH;
    return push_comp(b, GO_NOT, &input, 1, 0, output, cell);
}

/* add_multiplexer(inputs, select, output): 2^k data inputs, k-bit select (A.2) */
int go_add_mux(go_builder *b, const uint32_t *inputs, uint32_t n, uint32_t select, uint32_t output, uint32_t *cell) {
    if (output >= b->n_wires || select >= b->n_wires) return GO_ERR_INVALID_WIRE_ID;
    for (uint32_t i = 0; i < n; i++) if (inputs[i] >= b->n_wires) return GO_ERR_INVALID_WIRE_ID;
    if (n < 2 || (n & (n - 1)) != 0) return GO_ERR_INVALID_INPUT_COUNT;
    uint32_t k = 0; while ((1u << k) < n) k++;
    if (b->width[select] != k) return GO_ERR_WIRE_WIDTH_INCOMPATIBLE;
    for (uint32_t i = 0; i < n; i++) if (b->width[inputs[i]] != b->width[output]) return GO_ERR_WIRE_WIDTH_MISMATCH;
    return push_comp(b, GO_MUX, inputs, n, select, output, cell);
}

/* add_buffer(input, enable, output): tri-state buffer (gsim SimulatorBuilder::add_buffer; no call site
   in the reference — SURVEY.md 8f row f4).  Validation as for the other components of A.2: data width
   must equal the output width (WireWidthMismatch), the enable is one bit (WireWidthIncompatible). */
int go_add_buffer(go_builder *b, uint32_t input, uint32_t enable, uint32_t output, uint32_t *cell) {
    if (output >= b->n_wires || input >= b->n_wires || enable >= b->n_wires) return GO_ERR_INVALID_WIRE_ID;
    if (b->width[input] != b->width[output]) return GO_ERR_WIRE_WIDTH_MISMATCH;
    if (b->width[enable] != 1) return GO_ERR_WIRE_WIDTH_INCOMPATIBLE;
    uint32_t in[2] = {input, enable};
    return push_comp(b, GO_BUF, in, 2, 0, output, cell);
}

/* add_slice(input, offset, output) / add_merge(inputs, output): bit re-wiring components of gsim (no call
   site in the reference; `OffsetOutOfRange` in the reference's error map, lib.rs:55-66, is theirs). */
int go_add_slice(go_builder *b, uint32_t input, uint32_t offset, uint32_t output, uint32_t *cell) {
    if (output >= b->n_wires || input >= b->n_wires) return GO_ERR_INVALID_WIRE_ID;
    if ((uint64_t)offset + b->width[output] > b->width[input]) return GO_ERR_OFFSET_OUT_OF_RANGE;
    return push_comp(b, GO_SLICE, &input, 1, offset, output, cell);
}

int go_add_merge(go_builder *b, const uint32_t *inputs, uint32_t n, uint32_t output, uint32_t *cell) {
    if (output >= b->n_wires) return GO_ERR_INVALID_WIRE_ID;
    for (uint32_t i = 0; i < n; i++) if (inputs[i] >= b->n_wires) return GO_ERR_INVALID_WIRE_ID;
    if (n < 1) return GO_ERR_TOO_FEW_INPUTS;
    uint32_t total = 0;
    for (uint32_t i = 0; i < n; i++) total += b->width[inputs[i]];
    if (total != b->width[output]) return GO_ERR_WIRE_WIDTH_MISMATCH;
    return push_comp(b, GO_MERGE, inputs, n, 0, output, cell);
}

/* add_register(data_in, data_out, enable, clock): gsim's clocked register (no call site in the reference;
   the Yosys importer's Dff/Dffe arms are todo!(), crates/digilogic_serde/src/yosys.rs:199-200).  Restated:
   on a rising clock edge (previous update saw a valid 0, this one a valid 1) enable = 1 loads data_in
   (High-Z bits read X), enable = 0 keeps, enable Z/X makes the stored value Undefined; the output is the
   stored value, Undefined until the first load. */
int go_add_register(go_builder *b, uint32_t data_in, uint32_t data_out, uint32_t enable, uint32_t clock, uint32_t *cell) {
    if (data_in >= b->n_wires || data_out >= b->n_wires || enable >= b->n_wires || clock >= b->n_wires) return GO_ERR_INVALID_WIRE_ID;
    if (b->width[data_in] != b->width[data_out]) return GO_ERR_WIRE_WIDTH_MISMATCH;
    if (b->width[enable] != 1 || b->width[clock] != 1) return GO_ERR_WIRE_WIDTH_INCOMPATIBLE;
    uint32_t in[3] = {data_in, enable, clock};
    return push_comp(b, GO_REG, in, 3, 0, data_out, cell);
}

/* add_horizontal_{and,or,xor,nand,nor,xnor}_gate(input, output): the 1-bit output is the gate function folded over
   all bits of the input wire (gsim's reduction gates; the Yosys importer's ReduceAnd/Or/Xor/Xnor/Bool arms are todo!(),
   crates/digilogic_serde/src/yosys.rs:163-167). */
int go_add_horizontal_gate(go_builder *b, int kind, uint32_t input, uint32_t output, uint32_t *cell) {
    if (kind < GO_AND || kind > GO_XNOR) return GO_ERR_INVALID_INPUT_COUNT;
    if (output >= b->n_wires || input >= b->n_wires) return GO_ERR_INVALID_WIRE_ID;
    if (b->width[output] != 1) return GO_ERR_WIRE_WIDTH_INCOMPATIBLE;
    return push_comp(b, GO_HGATE, &input, 1, (uint32_t)kind, output, cell);
}

/* add_add / add_sub(input_a, input_b, output): gsim's adder and subtractor (the Yosys importer's Add/Sub arms are
   todo!(), crates/digilogic_serde/src/yosys.rs:188-189).  All three wires have one width; the result wraps; if any bit of
   either input is not a valid 0/1 the whole output is Undefined. */
int go_add_arith(go_builder *b, int kind, uint32_t a, uint32_t c, uint32_t output, uint32_t *cell) {
    if (kind != GO_ADD && kind != GO_SUB) return GO_ERR_INVALID_INPUT_COUNT;
    if (output >= b->n_wires || a >= b->n_wires || c >= b->n_wires) return GO_ERR_INVALID_WIRE_ID;
    if (b->width[a] != b->width[output] || b->width[c] != b->width[output]) return GO_ERR_WIRE_WIDTH_MISMATCH;
    uint32_t in[2] = {a, c};
    return push_comp(b, (uint8_t)kind, in, 2, 0, output, cell);
}

/* add_mul(input_a, input_b, output): gsim's multiplier (Yosys Mul is todo!() in the reference, crates/digilogic_serde/src/yosys.rs:190).
   Restated like add / sub: one width for all three wires, the product wraps, any invalid input bit makes the output Undefined. */
int go_add_mul(go_builder *b, uint32_t a, uint32_t c, uint32_t output, uint32_t *cell) {
    if (output >= b->n_wires || a >= b->n_wires || c >= b->n_wires) return GO_ERR_INVALID_WIRE_ID;
    if (b->width[a] != b->width[output] || b->width[c] != b->width[output]) return GO_ERR_WIRE_WIDTH_MISMATCH;
    uint32_t in[2] = {a, c};
    return push_comp(b, GO_MUL, in, 2, 0, output, cell);
}

/* add_compare_{equal,not_equal,less_than,greater_than,less_than_or_equal,greater_than_or_equal}{,_signed}(input_a, input_b, output):
   gsim's comparators (Yosys Lt/Le/Eq/Ne/Ge/Gt are todo!() in the reference, yosys.rs:182-187).  Restated: the inputs have one
   width (WireWidthMismatch), the output is one bit (WireWidthIncompatible); any invalid input bit makes it Undefined. */
int go_add_compare(go_builder *b, int op, uint32_t a, uint32_t c, uint32_t output, uint32_t *cell) {
    if (op < GO_CMP_EQ || op > GO_CMP_GES) return GO_ERR_INVALID_INPUT_COUNT;
    if (output >= b->n_wires || a >= b->n_wires || c >= b->n_wires) return GO_ERR_INVALID_WIRE_ID;
    if (b->width[a] != b->width[c]) return GO_ERR_WIRE_WIDTH_MISMATCH;
    if (b->width[output] != 1) return GO_ERR_WIRE_WIDTH_INCOMPATIBLE;
    uint32_t in[2] = {a, c};
    return push_comp(b, GO_CMP, in, 2, (uint32_t)op, output, cell);
}

/* add_left_shift / add_logical_right_shift / add_arithmetic_right_shift(input, shift_amount, output): gsim's shifters (Yosys
   Shl/Sshl/Shr/Sshr are todo!() in the reference, yosys.rs:174-177).  Restated: output as wide as the input (WireWidthMismatch),
   the amount an unsigned value of at most 8 bits (WireWidthIncompatible); amounts >= width shift everything out (zeros, or the
   sign for the arithmetic shift); any invalid bit of either input makes the output Undefined. */
int go_add_shift(go_builder *b, int kind, uint32_t a, uint32_t amount, uint32_t output, uint32_t *cell) {
    if (kind != GO_SHL && kind != GO_LSHR && kind != GO_ASHR) return GO_ERR_INVALID_INPUT_COUNT;
    if (output >= b->n_wires || a >= b->n_wires || amount >= b->n_wires) return GO_ERR_INVALID_WIRE_ID;
    if (b->width[a] != b->width[output]) return GO_ERR_WIRE_WIDTH_MISMATCH;
    if (b->width[amount] > 8) return GO_ERR_WIRE_WIDTH_INCOMPATIBLE;
    uint32_t in[2] = {a, amount};
    return push_comp(b, (uint8_t)kind, in, 2, 0, output, cell);
}

/* Bulk 2-input gate add used by the large generators (same validation per gate). */
int go_add_gates2(go_builder *b, const uint8_t *kinds, const uint32_t *in0, const uint32_t *in1, const uint32_t *out, uint32_t n) {
    for (uint32_t i = 0; i < n; i++) {
        int rc;
        if (kinds[i] == GO_NOT) rc = go_add_not(b, in0[i], out[i], NULL);
        else { uint32_t in[2] = {in0[i], in1[i]}; rc = go_add_gate(b, kinds[i], in, 2, out[i], NULL); }
        if (rc != GO_OK) return rc;
    }
    return GO_OK;
}

/* ------------------------------------------------------------------ build */

static void init_state(go_sim *s) {
    atom_t init = {0, 0};
    if (g_cfg.initial_undefined) { init.state = 0xFFFFFFFFu; init.valid = 0; }
    for (uint32_t i = 0; i < s->n_watoms; i++) { s->drive[i].state = 0; s->drive[i].valid = 0; s->state[i] = init; }
    for (uint32_t i = 0; i < s->n_catoms; i++) s->output[i] = init;
    /* a register holds Undefined until its first load; its clock history starts unknown */
    for (uint32_t c = 0; c < s->n_comps; c++) {
        uint32_t w = s->comps[c].output;
        for (uint32_t i = 0; i < s->natoms[w]; i++) {
            uint32_t top = (i == (uint32_t)s->natoms[w] - 1 && (s->width[w] & 31)) ? (1u << (s->width[w] & 31)) - 1 : 0xFFFFFFFFu;
            s->rdata[s->catom_off[c] + i].state = top; s->rdata[s->catom_off[c] + i].valid = 0;
        }
        s->rprev[c] = -1;
    }
    if (g_cfg.initial_undefined) {
        for (uint32_t w = 0; w < s->n_wires; w++) {
            uint32_t top = s->width[w] & 31;
            if (top) s->state[s->atom_off[w] + s->natoms[w] - 1].state &= (1u << top) - 1;
        }
        for (uint32_t c = 0; c < s->n_comps; c++) {
            uint32_t w = s->comps[c].output, top = s->width[w] & 31;
            if (top) s->output[s->catom_off[c] + s->natoms[w] - 1].state &= (1u << top) - 1;
        }
    }
}

/* build(): freeze the graph into id-indexed arrays (lib.rs:123-133). Does not consume b. */
go_sim *go_build(const go_builder *b) {
    go_sim *s = (go_sim *)calloc(1, sizeof(go_sim));
    uint32_t W = b->n_wires, C = b->n_comps;
    s->n_wires = W; s->n_comps = C;
    s->width = (uint8_t *)malloc(W ? W : 1); memcpy(s->width, b->width, W);
    s->natoms = (uint8_t *)malloc(W ? W : 1);
    s->comps = (comp_t *)malloc((C ? C : 1) * sizeof(comp_t)); memcpy(s->comps, b->comps, (size_t)C * sizeof(comp_t));
    s->inputs = (uint32_t *)malloc((b->n_inputs ? b->n_inputs : 1) * sizeof(uint32_t));
    memcpy(s->inputs, b->inputs, (size_t)b->n_inputs * sizeof(uint32_t));
    s->atom_off = (uint32_t *)malloc((size_t)(W + 1) * sizeof(uint32_t));
    uint32_t off = 0;
    for (uint32_t w = 0; w < W; w++) { s->natoms[w] = (uint8_t)((s->width[w] + 31) / 32); s->atom_off[w] = off; off += s->natoms[w]; }
    s->atom_off[W] = off; s->n_watoms = off;
    s->catom_off = (uint32_t *)malloc((size_t)(C + 1) * sizeof(uint32_t));
    off = 0;
    for (uint32_t c = 0; c < C; c++) { s->catom_off[c] = off; off += s->natoms[s->comps[c].output]; }
    s->catom_off[C] = off; s->n_catoms = off;
    s->drive = (atom_t *)malloc((s->n_watoms ? s->n_watoms : 1) * sizeof(atom_t));
    s->state = (atom_t *)malloc((s->n_watoms ? s->n_watoms : 1) * sizeof(atom_t));
    s->output = (atom_t *)malloc((s->n_catoms ? s->n_catoms : 1) * sizeof(atom_t));
    s->rdata = (atom_t *)malloc((s->n_catoms ? s->n_catoms : 1) * sizeof(atom_t));
    s->rprev = (int8_t *)malloc(C ? C : 1);

    /* drivers CSR */
    s->drivers_off = (uint32_t *)calloc((size_t)W + 2, sizeof(uint32_t));
    for (uint32_t c = 0; c < C; c++) s->drivers_off[s->comps[c].output + 2]++;
    for (uint32_t w = 0; w < W; w++) s->drivers_off[w + 2] += s->drivers_off[w + 1];
    s->drivers = (uint32_t *)malloc((C ? C : 1) * sizeof(uint32_t));
    for (uint32_t c = 0; c < C; c++) s->drivers[s->drivers_off[s->comps[c].output + 1]++] = c;
    /* driving CSR (wire -> components reading it); duplicates removed later by sort+dedup of the queue */
    uint64_t total = b->n_inputs;
    for (uint32_t c = 0; c < C; c++) if (s->comps[c].kind == GO_MUX) total++;
    s->driving_off = (uint32_t *)calloc((size_t)W + 2, sizeof(uint32_t));
    for (uint32_t c = 0; c < C; c++) {
        const comp_t *cp = &s->comps[c];
        for (uint32_t i = 0; i < cp->n_inputs; i++) s->driving_off[s->inputs[cp->in_off + i] + 2]++;
        if (cp->kind == GO_MUX) s->driving_off[cp->select + 2]++;
    }
    for (uint32_t w = 0; w < W; w++) s->driving_off[w + 2] += s->driving_off[w + 1];
    s->driving = (uint32_t *)malloc((total ? total : 1) * sizeof(uint32_t));
    for (uint32_t c = 0; c < C; c++) {
        const comp_t *cp = &s->comps[c];
        for (uint32_t i = 0; i < cp->n_inputs; i++) s->driving[s->driving_off[s->inputs[cp->in_off + i] + 1]++] = c;
        if (cp->kind == GO_MUX) s->driving[s->driving_off[cp->select + 1]++] = c;
    }
    size_t qcap = (size_t)(total > W ? total : W) + C + 16;
    s->wq = (uint32_t *)malloc(qcap * sizeof(uint32_t));
    s->cq = (uint32_t *)malloc(qcap * sizeof(uint32_t));
    s->tmp = (uint32_t *)malloc(qcap * sizeof(uint32_t));
    s->owns_graph = 1;
    init_state(s);
    return s;
}

void go_sim_free(go_sim *s) {
    if (!s) return;
    if (s->owns_graph) {
        free(s->width); free(s->natoms); free(s->comps); free(s->inputs);
        free(s->driving_off); free(s->driving); free(s->drivers_off); free(s->drivers);
        free(s->atom_off); free(s->catom_off);
    }
    free(s->drive); free(s->state); free(s->output); free(s->rdata); free(s->rprev);
    free(s->wq); free(s->cq); free(s->tmp); free(s);
}

/* A simulator with the same graph (shared, immutable; `src` must outlive it) and a
   private copy of the *current* state — one gsim Simulator per stimulus lane. */
go_sim *go_sim_clone(const go_sim *src) {
    go_sim *s = (go_sim *)malloc(sizeof(go_sim));
    *s = *src;
    s->owns_graph = 0;
    uint32_t W = s->n_wires, C = s->n_comps;
#define DUP(field, count, type) do { size_t n_ = (size_t)(count); if (!n_) n_ = 1; s->field = (type *)malloc(n_ * sizeof(type)); memcpy(s->field, src->field, n_ * sizeof(type)); } while (0)
    DUP(drive, s->n_watoms, atom_t); DUP(state, s->n_watoms, atom_t); DUP(output, s->n_catoms, atom_t);
    DUP(rdata, s->n_catoms, atom_t); DUP(rprev, C, int8_t);
#undef DUP
    size_t qcap = (size_t)(src->driving_off[W] > W ? src->driving_off[W] : W) + C + 16;
    s->wq = (uint32_t *)malloc(qcap * sizeof(uint32_t));
    s->cq = (uint32_t *)malloc(qcap * sizeof(uint32_t));
    s->tmp = (uint32_t *)malloc(qcap * sizeof(uint32_t));
    s->n_wq = s->n_cq = 0;
    return s;
}

int go_wire_width(const go_sim *s, uint32_t wire, uint8_t *out) {
    if (wire >= s->n_wires) return GO_ERR_INVALID_WIRE_ID;
    *out = s->width[wire];
    return GO_OK;
}

/* ------------------------------------------------------------------ logic (A.3, gsim-style plane formulas) */

static inline atom_t op_and(atom_t a, atom_t b) {
    atom_t r;
    r.state = (a.state & b.state) | (~a.valid & ~b.valid) | (a.state & ~b.valid) | (b.state & ~a.valid);
    r.valid = (a.valid & b.valid) | (~a.state & a.valid) | (~b.state & b.valid);
    return r;
}
static inline atom_t op_or(atom_t a, atom_t b) {
    atom_t r;
    r.state = a.state | ~a.valid | b.state | ~b.valid;
    r.valid = (a.state & a.valid) | (b.state & b.valid) | (a.valid & b.valid);
    return r;
}
static inline atom_t op_xor(atom_t a, atom_t b) {
    atom_t r;
    r.state = (a.state ^ b.state) | ~a.valid | ~b.valid;
    r.valid = a.valid & b.valid;
    return r;
}
static inline atom_t op_not(atom_t a) {
    atom_t r;
    r.state = ~a.state | ~a.valid;
    r.valid = a.valid;
    return r;
}

static inline uint32_t atom_mask(uint8_t width, uint32_t i) {
    uint32_t rem = (uint32_t)width - i * 32;
    return rem >= 32 ? 0xFFFFFFFFu : ((1u << rem) - 1);
}

/* Exposed for the truth-table tests: op on single atoms, kind as GO_*. */
void go_logic_op(int kind, uint32_t as, uint32_t av, uint32_t bs, uint32_t bv, uint32_t *rs, uint32_t *rv) {
    atom_t a = {as, av}, b = {bs, bv}, r;
    switch (kind) {
    case GO_AND: r = op_and(a, b); break;
    case GO_OR: r = op_or(a, b); break;
    case GO_XOR: r = op_xor(a, b); break;
    case GO_NAND: r = op_not(op_and(a, b)); break;
    case GO_NOR: r = op_not(op_or(a, b)); break;
    case GO_XNOR: r = op_not(op_xor(a, b)); break;
    default: r = op_not(a); break;
    }
    *rs = r.state; *rv = r.valid;
}

/* component update: returns 1 if the output changed */
static int eval_component(go_sim *s, uint32_t c) {
    const comp_t *cp = &s->comps[c];
    const uint32_t *in = s->inputs + cp->in_off;
    uint8_t width = s->width[cp->output];
    uint32_t na = s->natoms[cp->output];
    atom_t *out = s->output + s->catom_off[c];
    atom_t res[GO_MAX_ATOMS];
    if (cp->kind == GO_MUX) {
        uint8_t sw = s->width[cp->select];
        const atom_t *sel = s->state + s->atom_off[cp->select];
        uint32_t smask = atom_mask(sw, 0); /* select width <= 31 bits in practice (2^k inputs) */
        if ((sel[0].valid & smask) == smask) {
            uint32_t idx = sel[0].state & smask;
            const atom_t *d = s->state + s->atom_off[in[idx]];
            for (uint32_t i = 0; i < na; i++) {
                res[i] = d[i];
                if (!g_cfg.mux_z_passthrough) res[i].state |= ~d[i].valid; /* Z -> X */
            }
        } else {
            for (uint32_t i = 0; i < na; i++) { res[i].state = 0xFFFFFFFFu; res[i].valid = 0; }
        }
    } else if (cp->kind == GO_NOT) {
        const atom_t *a = s->state + s->atom_off[in[0]];
        for (uint32_t i = 0; i < na; i++) res[i] = op_not(a[i]);
    } else if (cp->kind == GO_SLICE || cp->kind == GO_MERGE) {
        /* bit j of the output = one bit of one input; High-Z bits read X unless A.6-7 says otherwise */
        for (uint32_t i = 0; i < na; i++) { res[i].state = 0; res[i].valid = 0; }
        uint32_t j = 0;
        for (uint32_t k = 0; k < cp->n_inputs; k++) {
            const atom_t *a = s->state + s->atom_off[in[k]];
            uint32_t lo = cp->kind == GO_SLICE ? cp->select : 0;
            uint32_t cnt = cp->kind == GO_SLICE ? width : s->width[in[k]];
            for (uint32_t t = 0; t < cnt; t++, j++) {
                uint32_t sb = lo + t;
                uint32_t st = (a[sb >> 5].state >> (sb & 31)) & 1, vl = (a[sb >> 5].valid >> (sb & 31)) & 1;
                if (!vl && !g_cfg.copy_z_passthrough) st = 1;
                res[j >> 5].state |= st << (j & 31);
                res[j >> 5].valid |= vl << (j & 31);
            }
        }
    } else if (cp->kind == GO_ADD || cp->kind == GO_SUB) {
        const atom_t *a = s->state + s->atom_off[in[0]], *b2 = s->state + s->atom_off[in[1]];
        int all_valid = 1;
        for (uint32_t i = 0; i < na; i++) {
            uint32_t m = atom_mask(width, i);
            if ((a[i].valid & m) != m || (b2[i].valid & m) != m) all_valid = 0;
        }
        if (!all_valid) {
            for (uint32_t i = 0; i < na; i++) { res[i].state = 0xFFFFFFFFu; res[i].valid = 0; }
        } else {
            uint64_t carry = cp->kind == GO_SUB ? 1 : 0;
            for (uint32_t i = 0; i < na; i++) {
                uint64_t y = cp->kind == GO_SUB ? (uint32_t)~b2[i].state : b2[i].state;
                uint64_t sum = (uint64_t)a[i].state + y + carry;
                res[i].state = (uint32_t)sum; res[i].valid = 0xFFFFFFFFu;
                carry = sum >> 32;
            }
        }
    } else if (cp->kind == GO_MUL || cp->kind == GO_CMP || cp->kind == GO_SHL || cp->kind == GO_LSHR || cp->kind == GO_ASHR) {
        const uint32_t wa = s->width[in[0]], wb = s->width[in[1]], naa = s->natoms[in[0]], nab = s->natoms[in[1]];
        const atom_t *a = s->state + s->atom_off[in[0]], *b2 = s->state + s->atom_off[in[1]];
        int all_valid = 1;
        uint32_t av[GO_MAX_ATOMS] = {0}, bv[GO_MAX_ATOMS] = {0}, rv[GO_MAX_ATOMS] = {0};
        for (uint32_t i = 0; i < naa; i++) { uint32_t m = atom_mask((uint8_t)wa, i); if ((a[i].valid & m) != m) all_valid = 0; av[i] = a[i].state & m; }
        for (uint32_t i = 0; i < nab; i++) { uint32_t m = atom_mask((uint8_t)wb, i); if ((b2[i].valid & m) != m) all_valid = 0; bv[i] = b2[i].state & m; }
        if (!all_valid) {
            for (uint32_t i = 0; i < na; i++) { res[i].state = 0xFFFFFFFFu; res[i].valid = 0; }
        } else {
            #define BIT(x, t) (((x)[(t) >> 5] >> ((t) & 31)) & 1u)
            if (cp->kind == GO_MUL) {
                for (uint32_t i = 0; i < naa; i++) {
                    uint64_t carry = 0;
                    for (uint32_t j = 0; i + j < GO_MAX_ATOMS; j++) {
                        uint64_t t = (uint64_t)av[i] * bv[j] + rv[i + j] + carry;
                        rv[i + j] = (uint32_t)t;
                        carry = t >> 32;
                    }
                }
            } else if (cp->kind == GO_CMP) {
                int lt = 0, eq = 1;                         /* unsigned, from the top bit down */
                for (int t = (int)wa - 1; t >= 0 && eq; t--) {
                    uint32_t x = BIT(av, (uint32_t)t), y = BIT(bv, (uint32_t)t);
                    if (x != y) { eq = 0; lt = x < y; }
                }
                const uint32_t op = cp->select;
                if (op >= GO_CMP_LTS && !eq && BIT(av, wa - 1) != BIT(bv, wa - 1)) lt = BIT(av, wa - 1);   /* signs differ: the negative one is less */
                int r;
                switch (op) {
                case GO_CMP_EQ: r = eq; break;
                case GO_CMP_NE: r = !eq; break;
                case GO_CMP_LTU: case GO_CMP_LTS: r = lt; break;
                case GO_CMP_GTU: case GO_CMP_GTS: r = !lt && !eq; break;
                case GO_CMP_LEU: case GO_CMP_LES: r = lt || eq; break;
                default: r = !lt; break;
                }
                rv[0] = (uint32_t)r;
            } else {
                const uint32_t amt = bv[0];                /* at most 8 bits wide */
                const uint32_t sign = cp->kind == GO_ASHR ? BIT(av, wa - 1) : 0;
                for (uint32_t j = 0; j < wa; j++) {
                    uint32_t bit;
                    if (cp->kind == GO_SHL) bit = j >= amt ? BIT(av, j - amt) : 0;
                    else bit = (uint64_t)j + amt < wa ? BIT(av, j + amt) : sign;
                    rv[j >> 5] |= bit << (j & 31);
                }
            }
            #undef BIT
            for (uint32_t i = 0; i < na; i++) { res[i].state = rv[i]; res[i].valid = 0xFFFFFFFFu; }
        }
    } else if (cp->kind == GO_HGATE) {
        const atom_t *a = s->state + s->atom_off[in[0]];
        const uint32_t wi = s->width[in[0]], k = cp->select;
        atom_t acc = {a[0].state & 1, a[0].valid & 1};
        for (uint32_t t = 1; t < wi; t++) {
            atom_t bit = {(a[t >> 5].state >> (t & 31)) & 1, (a[t >> 5].valid >> (t & 31)) & 1};
            acc = (k == GO_AND || k == GO_NAND) ? op_and(acc, bit) : (k == GO_OR || k == GO_NOR) ? op_or(acc, bit) : op_xor(acc, bit);
            acc.state &= 1; acc.valid &= 1;
        }
        if (wi == 1) { acc.state |= ~acc.valid & 1; }   /* a single bit: the bit itself, High-Z read as X */
        if (k >= GO_NAND) acc = op_not(acc);
        res[0].state = acc.state & 1; res[0].valid = acc.valid & 1;
    } else if (cp->kind == GO_REG) {
        const atom_t *d = s->state + s->atom_off[in[0]];
        const atom_t en = s->state[s->atom_off[in[1]]], ck = s->state[s->atom_off[in[2]]];
        atom_t *data = s->rdata + s->catom_off[c];
        if (s->rprev[c] == 0 && (ck.valid & 1) && (ck.state & 1)) {          /* rising edge */
            for (uint32_t i = 0; i < na; i++) {
                if (!(en.valid & 1)) { data[i].state = 0xFFFFFFFFu; data[i].valid = 0; }
                else if (en.state & 1) { data[i].state = d[i].state | ~d[i].valid; data[i].valid = d[i].valid; }
            }
        }
        s->rprev[c] = (ck.valid & 1) ? (int8_t)(ck.state & 1) : -1;
        for (uint32_t i = 0; i < na; i++) res[i] = data[i];
    } else if (cp->kind == GO_BUF) {
        /* enable = 1: the input (High-Z bits read X); enable = 0: High-Z; enable Z/X: Undefined */
        const atom_t *a = s->state + s->atom_off[in[0]];
        const atom_t en = s->state[s->atom_off[in[1]]];
        for (uint32_t i = 0; i < na; i++) {
            if (!(en.valid & 1)) { res[i].state = 0xFFFFFFFFu; res[i].valid = 0; }
            else if (!(en.state & 1)) { res[i].state = 0; res[i].valid = 0; }
            else { res[i] = a[i]; if (!g_cfg.buffer_z_passthrough) res[i].state |= ~a[i].valid; }
        }
    } else {
        const atom_t *a = s->state + s->atom_off[in[0]];
        for (uint32_t i = 0; i < na; i++) res[i] = a[i];
        for (uint32_t k = 1; k < cp->n_inputs; k++) {
            const atom_t *b = s->state + s->atom_off[in[k]];
            switch (cp->kind) {
            case GO_AND: case GO_NAND: for (uint32_t i = 0; i < na; i++) res[i] = op_and(res[i], b[i]); break;
            case GO_OR: case GO_NOR: for (uint32_t i = 0; i < na; i++) res[i] = op_or(res[i], b[i]); break;
            default: for (uint32_t i = 0; i < na; i++) res[i] = op_xor(res[i], b[i]); break;
            }
        }
        if (cp->kind >= GO_NAND) for (uint32_t i = 0; i < na; i++) res[i] = op_not(res[i]);
    }
    int changed = 0;
    for (uint32_t i = 0; i < na; i++) {
        uint32_t m = atom_mask(width, i);
        res[i].state &= m; res[i].valid &= m;
        if (res[i].state != out[i].state || res[i].valid != out[i].valid) changed = 1;
        out[i] = res[i];
    }
    s->comp_evals++;
    return changed;
}

/* wire update: returns bit0 = changed, bit1 = conflict (A.4) */
static int update_wire(go_sim *s, uint32_t w) {
    uint32_t na = s->natoms[w];
    uint8_t width = s->width[w];
    atom_t *st = s->state + s->atom_off[w];
    const atom_t *dr = s->drive + s->atom_off[w];
    atom_t res[GO_MAX_ATOMS];
    uint32_t conflict = 0;
    for (uint32_t i = 0; i < na; i++) res[i] = dr[i];
    for (uint32_t k = s->drivers_off[w]; k < s->drivers_off[w + 1]; k++) {
        const atom_t *o = s->output + s->catom_off[s->drivers[k]];
        for (uint32_t i = 0; i < na; i++) {
            uint32_t m = atom_mask(width, i);
            uint32_t both = (res[i].state | res[i].valid) & (o[i].state | o[i].valid);
            if (g_cfg.conflict_needs_disagree)
                both &= ~(res[i].valid & o[i].valid & ~(res[i].state ^ o[i].state));
            conflict |= both & m;
            res[i].state |= o[i].state;
            res[i].valid |= o[i].valid;
        }
    }
    int changed = 0;
    for (uint32_t i = 0; i < na; i++) {
        uint32_t m = atom_mask(width, i);
        res[i].state &= m; res[i].valid &= m;
        if (res[i].state != st[i].state || res[i].valid != st[i].valid) changed = 1;
        st[i] = res[i];
    }
    return changed | (conflict ? 2 : 0);
}

/* sort + dedup of a queue (gsim: par_sort_unstable + dedup). LSD radix for big queues. */
static uint32_t sort_dedup(uint32_t *q, uint32_t n, uint32_t *tmp) {
    if (n < 2) return n;
    if (n < 64) {
        for (uint32_t i = 1; i < n; i++) { uint32_t v = q[i]; uint32_t j = i; while (j > 0 && q[j - 1] > v) { q[j] = q[j - 1]; j--; } q[j] = v; }
    } else {
        uint32_t maxv = 0; for (uint32_t i = 0; i < n; i++) if (q[i] > maxv) maxv = q[i];
        uint32_t *src = q, *dst = tmp;
        for (int shift = 0; shift < 32 && (maxv >> shift); shift += 11) {
            uint32_t cnt[2049]; memset(cnt, 0, sizeof(cnt));
            for (uint32_t i = 0; i < n; i++) cnt[((src[i] >> shift) & 2047) + 1]++;
            for (int i = 0; i < 2048; i++) cnt[i + 1] += cnt[i];
            for (uint32_t i = 0; i < n; i++) dst[cnt[(src[i] >> shift) & 2047]++] = src[i];
            uint32_t *t = src; src = dst; dst = t;
        }
        if (src != q) memcpy(q, src, (size_t)n * sizeof(uint32_t));
    }
    uint32_t m = 1;
    for (uint32_t i = 1; i < n; i++) if (q[i] != q[m - 1]) q[m++] = q[i];
    return m;
}

enum { STEP_UNCHANGED = 0, STEP_CHANGED = 1, STEP_ERR = 2 };

static int update_wires(go_sim *s) {
    s->n_cq = 0;
    int conflict = 0;
    for (uint32_t i = 0; i < s->n_wq; i++) {
        uint32_t w = s->wq[i];
        int r = update_wire(s, w);
        if (r & 2) conflict = 1;
        if (r & 1) {
            uint32_t a = s->driving_off[w], e = s->driving_off[w + 1];
            memcpy(s->cq + s->n_cq, s->driving + a, (size_t)(e - a) * sizeof(uint32_t));
            s->n_cq += e - a;
        }
    }
    s->n_wq = 0;
    s->n_cq = sort_dedup(s->cq, s->n_cq, s->tmp);
    if (conflict) return STEP_ERR;
    return s->n_cq ? STEP_CHANGED : STEP_UNCHANGED;
}

static int update_components(go_sim *s) {
    s->n_wq = 0;
    for (uint32_t i = 0; i < s->n_cq; i++) {
        uint32_t c = s->cq[i];
        if (eval_component(s, c)) s->wq[s->n_wq++] = s->comps[c].output;
    }
    s->n_cq = 0;
    s->n_wq = sort_dedup(s->wq, s->n_wq, s->tmp);
    return s->n_wq ? STEP_CHANGED : STEP_UNCHANGED;
}

static int begin_sim(go_sim *s) {
    for (uint32_t w = 0; w < s->n_wires; w++) s->wq[w] = w;
    s->n_wq = s->n_wires;
    if (update_wires(s) == STEP_ERR) return STEP_ERR;
    for (uint32_t c = 0; c < s->n_comps; c++) s->cq[c] = c;
    s->n_cq = s->n_comps;
    return update_components(s);
}

static int step_sim(go_sim *s) {
    int r = update_wires(s);
    if (r == STEP_CHANGED) return update_components(s);
    return r;
}

/* run_sim(max_steps) — A.4 */
int go_run_sim(go_sim *s, uint64_t max_steps) {
    uint64_t steps = 0;
    s->comp_evals = 0;
    int r = begin_sim(s);
    for (;;) {
        if (r == STEP_UNCHANGED) { s->steps_done = steps; return GO_RUN_OK; }
        if (r == STEP_ERR) { s->steps_done = steps; return GO_RUN_CONFLICT; }
        if (g_cfg.steps_check_ge ? (steps >= max_steps) : (steps > max_steps)) { s->steps_done = steps; return GO_RUN_MAX_STEPS; }
        steps++;
        r = step_sim(s);
    }
}

uint64_t go_last_steps(const go_sim *s) { return s->steps_done; }
uint64_t go_last_comp_evals(const go_sim *s) { return s->comp_evals; }

/* LogicState::from_bit_planes(&p0[..words], &p1[..words]) + set_wire_drive (lib.rs:194-214):
   missing words are High-Z. */
int go_set_wire_drive(go_sim *s, uint32_t wire, const uint32_t *p0, const uint32_t *p1, uint32_t words) {
    if (wire >= s->n_wires) return GO_ERR_INVALID_WIRE_ID;
    atom_t *dr = s->drive + s->atom_off[wire];
    for (uint32_t i = 0; i < s->natoms[wire]; i++) {
        uint32_t m = atom_mask(s->width[wire], i);
        dr[i].state = (i < words ? p0[i] : 0) & m;
        dr[i].valid = (i < words ? p1[i] : 0) & m;
    }
    return GO_OK;
}

/* get_wire_state + to_bit_planes(width) (lib.rs:228-230) */
int go_get_wire_state(const go_sim *s, uint32_t wire, uint32_t *p0, uint32_t *p1) {
    if (wire >= s->n_wires) return GO_ERR_INVALID_WIRE_ID;
    const atom_t *st = s->state + s->atom_off[wire];
    for (uint32_t i = 0; i < s->natoms[wire]; i++) { p0[i] = st[i].state; p1[i] = st[i].valid; }
    return GO_OK;
}

/* ------------------------------------------------------------------ lanes
 * L independent simulators of one netlist (one gsim Simulator per stimulus
 * lane), driven/read through 64-lane bit-sliced words so that the harness
 * can compare against the GPU store directly.  Lanes are split over threads
 * (one simulator instance per lane, disjoint lanes per thread) — the same
 * stimulus-parallel sharding the GPU path uses.
 */
typedef struct {
    uint32_t n_lanes;
    go_sim **sims;
} go_lanes;

go_lanes *go_lanes_new(const go_builder *b, uint32_t n_lanes) {
    go_lanes *l = (go_lanes *)calloc(1, sizeof(go_lanes));
    l->n_lanes = n_lanes;
    l->sims = (go_sim **)calloc(n_lanes ? n_lanes : 1, sizeof(go_sim *));
    if (n_lanes) {
        l->sims[0] = go_build(b);
        for (uint32_t i = 1; i < n_lanes; i++) l->sims[i] = go_sim_clone(l->sims[0]);
    }
    return l;
}

void go_lanes_free(go_lanes *l) {
    if (!l) return;
    for (uint32_t i = l->n_lanes; i-- > 0;) go_sim_free(l->sims[i]); /* sims[0] owns the graph: last */
    free(l->sims); free(l);
}

/* Drive bit `bit` of `wire` in every lane from bit-sliced words (lane i = bit i%64 of word i/64). */
int go_lanes_set_drive_bit(go_lanes *l, uint32_t wire, uint32_t bit, const uint64_t *value, const uint64_t *valid) {
    if (!l->n_lanes) return GO_OK;
    if (wire >= l->sims[0]->n_wires || bit >= l->sims[0]->width[wire]) return GO_ERR_INVALID_WIRE_ID;
    for (uint32_t i = 0; i < l->n_lanes; i++) {
        go_sim *s = l->sims[i];
        atom_t *dr = s->drive + s->atom_off[wire] + bit / 32;
        uint32_t m = 1u << (bit & 31);
        uint32_t v = (uint32_t)((value[i / 64] >> (i & 63)) & 1), k = (uint32_t)((valid[i / 64] >> (i & 63)) & 1);
        dr->state = (dr->state & ~m) | (v ? m : 0);
        dr->valid = (dr->valid & ~m) | (k ? m : 0);
    }
    return GO_OK;
}

int go_lanes_get_bit(const go_lanes *l, uint32_t wire, uint32_t bit, uint64_t *value, uint64_t *valid) {
    if (!l->n_lanes) return GO_OK;
    if (wire >= l->sims[0]->n_wires || bit >= l->sims[0]->width[wire]) return GO_ERR_INVALID_WIRE_ID;
    uint32_t words = (l->n_lanes + 63) / 64;
    memset(value, 0, (size_t)words * 8); memset(valid, 0, (size_t)words * 8);
    for (uint32_t i = 0; i < l->n_lanes; i++) {
        const go_sim *s = l->sims[i];
        const atom_t *st = s->state + s->atom_off[wire] + bit / 32;
        value[i / 64] |= (uint64_t)((st->state >> (bit & 31)) & 1) << (i & 63);
        valid[i / 64] |= (uint64_t)((st->valid >> (bit & 31)) & 1) << (i & 63);
    }
    return GO_OK;
}

/* All 1-bit wires at once: planes[w * words + k]; wires wider than 1 bit report bit 0. */
void go_lanes_get_all(const go_lanes *l, uint64_t *value, uint64_t *valid) {
    if (!l->n_lanes) return;
    uint32_t W = l->sims[0]->n_wires, words = (l->n_lanes + 63) / 64;
    memset(value, 0, (size_t)W * words * 8); memset(valid, 0, (size_t)W * words * 8);
    for (uint32_t i = 0; i < l->n_lanes; i++) {
        const go_sim *s = l->sims[i];
        for (uint32_t w = 0; w < W; w++) {
            const atom_t *st = s->state + s->atom_off[w];
            value[(size_t)w * words + i / 64] |= (uint64_t)(st->state & 1) << (i & 63);
            valid[(size_t)w * words + i / 64] |= (uint64_t)(st->valid & 1) << (i & 63);
        }
    }
}

typedef struct { go_lanes *l; uint32_t lo, hi; uint64_t max_steps; uint8_t *status; uint64_t *steps; uint64_t evals; } run_job;

static void *run_worker(void *p) {
    run_job *j = (run_job *)p;
    j->evals = 0;
    for (uint32_t i = j->lo; i < j->hi; i++) {
        int r = go_run_sim(j->l->sims[i], j->max_steps);
        if (j->status) j->status[i] = (uint8_t)r;
        if (j->steps) j->steps[i] = j->l->sims[i]->steps_done;
        j->evals += j->l->sims[i]->comp_evals;
    }
    return NULL;
}

/* run_sim(max_steps) on every lane, `threads` host threads. Returns total component evaluations. */
uint64_t go_lanes_run(go_lanes *l, uint64_t max_steps, uint32_t threads, uint8_t *status, uint64_t *steps) {
    if (threads < 1) threads = 1;
    if (threads > l->n_lanes) threads = l->n_lanes ? l->n_lanes : 1;
    if (threads > 256) threads = 256;
    pthread_t tid[256]; run_job jobs[256];
    for (uint32_t t = 0; t < threads; t++) {
        jobs[t].l = l; jobs[t].max_steps = max_steps; jobs[t].status = status; jobs[t].steps = steps;
        jobs[t].lo = (uint32_t)((uint64_t)l->n_lanes * t / threads);
        jobs[t].hi = (uint32_t)((uint64_t)l->n_lanes * (t + 1) / threads);
        if (threads == 1) run_worker(&jobs[t]); else pthread_create(&tid[t], NULL, run_worker, &jobs[t]);
    }
    uint64_t evals = 0;
    for (uint32_t t = 0; t < threads; t++) { if (threads > 1) pthread_join(tid[t], NULL); evals += jobs[t].evals; }
    return evals;
}
