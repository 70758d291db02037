// capi.cpp — the extern "C" boundary declared in include/digilogic_b200.h, and the
// B200Server state machine that mirrors GsimServer (reference
// crates/digilogic_gsim/src/lib.rs:5-52,101-239).
#include <algorithm>
#include <cstring>
#include <memory>
#include <new>
#include <unordered_map>

#include "../../include/digilogic_b200.h"
#include "engine.hpp"
#include "netlist.hpp"
#include "pair_plan.hpp"

#include "capi_types.hpp"

#define GUARD_BEGIN try {
#define GUARD_END(err)                        \
    }                                         \
    catch (const std::bad_alloc &) {          \
        return err;                           \
    }                                         \
    catch (...) {                             \
        b2::set_last_error("internal error"); \
        return err;                           \
    }

extern "C" {

const char *b2_last_error(void) { return b2::get_last_error(); }
const char *b2_version(void) { return "digilogic_b200 0.1.0 (sm_100a)"; }
uint64_t b2_kernel_launches(void) { return b2::kernel_launches(); }

// ------------------------------------------------------------------ builder
b2_builder *b2_builder_new(void) { return new (std::nothrow) b2_builder(); }
void b2_builder_free(b2_builder *b) { delete b; }

int b2_add_wire(b2_builder *b, uint8_t width, uint32_t *out_wire) {
    if (!b) return B2_ERR_NULL;
    GUARD_BEGIN
    return b->b.add_wire(width, out_wire);
    GUARD_END(B2_ERR_OUT_OF_MEMORY)
}

int b2_add_wires(b2_builder *b, uint8_t width, uint32_t n, uint32_t *out_first) {
    if (!b) return B2_ERR_NULL;
    if (width == 0) return B2_ERR_INVALID_ARG;
    GUARD_BEGIN
    return b->b.add_wires(width, n, out_first);
    GUARD_END(B2_ERR_OUT_OF_MEMORY)
}

int b2_builder_wire_width(const b2_builder *b, uint32_t wire, uint8_t *out_width) {
    if (!b || !out_width) return B2_ERR_NULL;
    return b->b.wire_width(wire, out_width);
}

int b2_add_gate(b2_builder *b, int kind, const uint32_t *inputs, size_t n, uint32_t output, uint32_t *out_cell) {
    if (!b || (n && !inputs)) return B2_ERR_NULL;
    GUARD_BEGIN
    return b->b.add_gate(kind, inputs, n, output, out_cell);
    GUARD_END(B2_ERR_OUT_OF_MEMORY)
}

int b2_add_not(b2_builder *b, uint32_t input, uint32_t output, uint32_t *out_cell) {
    if (!b) return B2_ERR_NULL;
    GUARD_BEGIN
    return b->b.add_not(input, output, out_cell);
    GUARD_END(B2_ERR_OUT_OF_MEMORY)
}

int b2_add_buffer(b2_builder *b, uint32_t input, uint32_t enable, uint32_t output, uint32_t *out_cell) {
    if (!b) return B2_ERR_NULL;
    GUARD_BEGIN
    return b->b.add_buffer(input, enable, output, out_cell);
    GUARD_END(B2_ERR_OUT_OF_MEMORY)
}

int b2_add_add(b2_builder *b, uint32_t input_a, uint32_t input_b, uint32_t output, uint32_t *out_cell) {
    if (!b) return B2_ERR_NULL;
    GUARD_BEGIN
    return b->b.add_arith(b2::K_ADD, input_a, input_b, output, out_cell);
    GUARD_END(B2_ERR_OUT_OF_MEMORY)
}

int b2_add_sub(b2_builder *b, uint32_t input_a, uint32_t input_b, uint32_t output, uint32_t *out_cell) {
    if (!b) return B2_ERR_NULL;
    GUARD_BEGIN
    return b->b.add_arith(b2::K_SUB, input_a, input_b, output, out_cell);
    GUARD_END(B2_ERR_OUT_OF_MEMORY)
}

int b2_add_mul(b2_builder *b, uint32_t input_a, uint32_t input_b, uint32_t output, uint32_t *out_cell) {
    if (!b) return B2_ERR_NULL;
    GUARD_BEGIN
    return b->b.add_arith(b2::K_MUL, input_a, input_b, output, out_cell);
    GUARD_END(B2_ERR_OUT_OF_MEMORY)
}

int b2_add_compare(b2_builder *b, int op, uint32_t input_a, uint32_t input_b, uint32_t output, uint32_t *out_cell) {
    if (!b) return B2_ERR_NULL;
    GUARD_BEGIN
    return b->b.add_compare(op, input_a, input_b, output, out_cell);
    GUARD_END(B2_ERR_OUT_OF_MEMORY)
}

int b2_add_shift(b2_builder *b, int kind, uint32_t input, uint32_t amount, uint32_t output, uint32_t *out_cell) {
    if (!b) return B2_ERR_NULL;
    GUARD_BEGIN
    const int k = kind == B2_SHIFT_LEFT ? b2::K_SHL : kind == B2_SHIFT_LOGICAL_RIGHT ? b2::K_LSHR : kind == B2_SHIFT_ARITHMETIC_RIGHT ? b2::K_ASHR : -1;
    if (k < 0) return B2_ERR_INVALID_ARG;
    return b->b.add_shift(k, input, amount, output, out_cell);
    GUARD_END(B2_ERR_OUT_OF_MEMORY)
}

int b2_add_horizontal_gate(b2_builder *b, int kind, uint32_t input, uint32_t output, uint32_t *out_cell) {
    if (!b) return B2_ERR_NULL;
    GUARD_BEGIN
    return b->b.add_horizontal_gate(kind, input, output, out_cell);
    GUARD_END(B2_ERR_OUT_OF_MEMORY)
}

int b2_add_register(b2_builder *b, uint32_t data_in, uint32_t data_out, uint32_t enable, uint32_t clock, uint32_t *out_cell) {
    if (!b) return B2_ERR_NULL;
    GUARD_BEGIN
    return b->b.add_register(data_in, data_out, enable, clock, out_cell);
    GUARD_END(B2_ERR_OUT_OF_MEMORY)
}

int b2_add_slice(b2_builder *b, uint32_t input, uint32_t offset, uint32_t output, uint32_t *out_cell) {
    if (!b) return B2_ERR_NULL;
    GUARD_BEGIN
    return b->b.add_slice(input, offset, output, out_cell);
    GUARD_END(B2_ERR_OUT_OF_MEMORY)
}

int b2_add_merge(b2_builder *b, const uint32_t *inputs, size_t n, uint32_t output, uint32_t *out_cell) {
    if (!b || (n && !inputs)) return B2_ERR_NULL;
    GUARD_BEGIN
    return b->b.add_merge(inputs, n, output, out_cell);
    GUARD_END(B2_ERR_OUT_OF_MEMORY)
}

int b2_add_mux(b2_builder *b, const uint32_t *inputs, size_t n, uint32_t select, uint32_t output, uint32_t *out_cell) {
    if (!b || (n && !inputs)) return B2_ERR_NULL;
    GUARD_BEGIN
    return b->b.add_mux(inputs, n, select, output, out_cell);
    GUARD_END(B2_ERR_OUT_OF_MEMORY)
}

int b2_add_gates2(b2_builder *b, const uint8_t *kinds, const uint32_t *in0, const uint32_t *in1, const uint32_t *out, size_t n) {
    if (!b || (n && (!kinds || !in0 || !in1 || !out))) return B2_ERR_NULL;
    GUARD_BEGIN
    b->b.comps.reserve(b->b.comps.size() + n);
    b->b.inputs.reserve(b->b.inputs.size() + 2 * n);
    // all or nothing: on the first invalid entry the builder is rolled back to its state before the call
    const size_t comps0 = b->b.comps.size(), inputs0 = b->b.inputs.size();
    const uint64_t gates0 = b->b.total_bit_gates, bin0 = b->b.total_bit_inputs;
    for (size_t i = 0; i < n; i++) {
        int rc;
        if (kinds[i] == B2_NOT) {
            rc = b->b.add_not(in0[i], out[i], nullptr);
        } else {
            const uint32_t in[2] = {in0[i], in1[i]};
            rc = b->b.add_gate(kinds[i], in, 2, out[i], nullptr);
        }
        if (rc) {
            b->b.comps.resize(comps0);
            b->b.inputs.resize(inputs0);
            b->b.total_bit_gates = gates0;
            b->b.total_bit_inputs = bin0;
            return rc;
        }
    }
    return B2_OK;
    GUARD_END(B2_ERR_OUT_OF_MEMORY)
}

b2_sim *b2_build(b2_builder *b) {
    if (!b) return nullptr;
    try {
        b2::Compiled c;
        b2::compile(b->b, c);
        auto *s = new b2_sim();
        s->e.reset(new b2::Engine(std::move(c)));
        const uint32_t sem = b->b.sem_flags;
        b->b = b2::Builder();  // std::mem::take(builder)
        b->b.sem_flags = sem;  // the switches describe the engine's reading of gsim, not one netlist: they stay
        return s;
    } catch (...) {
        return nullptr;
    }
}

void b2_sim_free(b2_sim *s) { delete s; }

b2_sim *b2_sim_clone(const b2_sim *s) {
    if (!s) return nullptr;
    try {
        auto *c = new b2_sim();
        c->e.reset(new b2::Engine(s->e->share()));
        c->e->copy_options_from(*s->e);
        return c;
    } catch (...) {
        return nullptr;
    }
}

int b2_builder_set_option(b2_builder *b, int option, int64_t value) {
    if (!b) return B2_ERR_NULL;
    uint32_t bit;
    switch (option) {
        case B2_BOPT_MUX_Z_PASSTHROUGH: bit = b2::SEM_MUX_Z; break;
        case B2_BOPT_INITIAL_UNDEFINED: bit = b2::SEM_INIT_X; break;
        case B2_BOPT_BUFFER_Z_PASSTHROUGH: bit = b2::SEM_BUF_Z; break;
        case B2_BOPT_COPY_Z_PASSTHROUGH: bit = b2::SEM_COPY_Z; break;
        default: return B2_ERR_INVALID_ARG;
    }
    if (value) b->b.sem_flags |= bit;
    else b->b.sem_flags &= ~bit;
    return B2_OK;
}

// ------------------------------------------------------------------ simulator
int b2_sim_get_info(const b2_sim *s, b2_sim_info *out) {
    if (!s || !out) return B2_ERR_NULL;
    const b2::Compiled &n = s->e->net();
    std::memset(out, 0, sizeof(*out));
    out->n_wires = n.n_wires();
    out->n_rows = n.n_rows;
    out->n_sources = n.n_src;
    out->n_gates2 = n.n_g2();
    out->n_gates3 = n.n_g3();
    out->n_gates_wide = n.n_wide();
    out->depth = n.depth;
    out->cyclic = n.cyclic;
    out->multi_driver = n.multi_driver;
    out->n_lanes = s->e->n_lanes();
    out->lane_words = s->e->lane_words();
    out->row_stride = s->e->row_stride();
    out->store_bytes = s->e->store_bytes();
    out->last_mode = s->e->last_mode();
    out->last_applications = s->e->last_applications();
    return B2_OK;
}

int b2_sim_get_pair_info(const b2_sim *s, uint64_t *launches, uint64_t *children, uint64_t *unstored) {
    if (!s) return B2_ERR_NULL;
    s->e->pair_info(launches, children, unstored);
    return B2_OK;
}

int b2_sim_get_hot_rows(b2_sim *s, uint32_t out_rows[4], uint32_t *out_n) {
    if (!s || !out_rows || !out_n) return B2_ERR_NULL;
    GUARD_BEGIN
    s->e->hot_rows(out_rows, out_n);
    return B2_OK;
    GUARD_END(B2_ERR_OUT_OF_MEMORY)
}

int b2_sim_plan_pairs(const b2_sim *s, const uint32_t *keep_wires, uint32_t n_keep, uint64_t out[6]) {
    if (!s || !out || (n_keep && !keep_wires)) return B2_ERR_NULL;
    GUARD_BEGIN
    const b2::Compiled &n = s->e->net();
    std::vector<uint32_t> keep;
    for (uint32_t i = 0; i < n_keep; i++) {
        if (keep_wires[i] >= n.n_wires()) return B2_ERR_INVALID_WIRE_ID;
        for (uint32_t j = 0; j < n.width[keep_wires[i]]; j++) keep.push_back(n.row_of_bit[n.bit_off[keep_wires[i]] + j]);
    }
    std::sort(keep.begin(), keep.end());
    keep.erase(std::unique(keep.begin(), keep.end()), keep.end());
    b2::PairPlanHost plan;
    if (!b2::plan_pairs(n, keep, plan)) return B2_ERR_INVALID_ARG;
    uint64_t now = 0;
    for (const b2::PairItem &it : plan.items)
        now += ((it.meta & b2::PM_NOW_P) != 0) + ((it.meta & b2::PM_NOW_C1) != 0) + ((it.meta & b2::PM_NOW_C2) != 0);
    out[0] = plan.launches.size();
    out[1] = plan.items.size();
    out[2] = plan.n_children;
    out[3] = plan.n_unstored;
    out[4] = plan.disc_rows.size();
    out[5] = now;
    return B2_OK;
    GUARD_END(B2_ERR_OUT_OF_MEMORY)
}

int b2_wire_width(const b2_sim *s, uint32_t wire, uint8_t *out_width) {
    if (!s || !out_width) return B2_ERR_NULL;
    if (wire >= s->e->net().n_wires()) return B2_ERR_INVALID_WIRE_ID;
    *out_width = s->e->net().width[wire];
    return B2_OK;
}

int b2_sim_set_device(b2_sim *s, int device) { return s ? s->e->set_device(device) : B2_ERR_NULL; }
int b2_sim_set_stream(b2_sim *s, void *stream) { return s ? s->e->set_stream(stream) : B2_ERR_NULL; }
int b2_sim_synchronize(b2_sim *s) { return s ? s->e->synchronize() : B2_ERR_NULL; }

int b2_sim_set_option(b2_sim *s, int option, int64_t value) {
    if (!s) return B2_ERR_NULL;
    switch (option) {
        case B2_OPT_RESIDENT_KERNEL: s->e->set_use_resident(value != 0); return B2_OK;
        case B2_OPT_VALID_ELISION: s->e->set_valid_elision(value != 0); return B2_OK;
        case B2_OPT_LEVEL_GRAPH: s->e->set_use_graph(value != 0); return B2_OK;
        case B2_OPT_FRONTIER: s->e->set_use_frontier(value != 0); return B2_OK;
        case B2_OPT_STEP_BUDGET_INCLUSIVE: s->e->set_budget_inclusive(value != 0); return B2_OK;
        case B2_OPT_EVAL_UNROLL: s->e->set_eval_unroll((uint32_t)value); return B2_OK;
        case B2_OPT_DEVICE_LOOP: s->e->set_device_loop(value != 0); return B2_OK;
        case B2_OPT_L2_HINTS: s->e->set_l2_hints(value != 0); return B2_OK;
        case B2_OPT_PERIOD2_SKIP: s->e->set_period2_skip(value != 0); return B2_OK;
        case B2_OPT_PDL: s->e->set_pdl((uint32_t)value); return B2_OK;
        case B2_OPT_PAIR_FUSION: s->e->set_pair_fusion((uint32_t)value); return B2_OK;
        case B2_OPT_HOT_STAGING: s->e->set_hot_staging(value != 0); return B2_OK;
        case B2_OPT_CONFLICT_NEEDS_DISAGREE: return s->e->set_lenient_conflicts(value != 0) ? B2_OK : B2_ERR_INVALID_ARG;
        default: return B2_ERR_INVALID_ARG;
    }
}

int b2_sim_set_lanes(b2_sim *s, uint64_t n_lanes) {
    if (!s) return B2_ERR_NULL;
    GUARD_BEGIN
    return s->e->set_lanes(n_lanes);
    GUARD_END(B2_ERR_OUT_OF_MEMORY)
}

uint64_t b2_sim_max_lanes(const b2_sim *s, uint64_t bytes) { return s ? s->e->max_lanes(bytes) : 0; }

int b2_set_wire_drive(b2_sim *s, uint32_t wire, const uint32_t *plane0, const uint32_t *plane1, size_t words) {
    if (!s) return B2_ERR_NULL;
    GUARD_BEGIN
    return s->e->set_wire_drive(wire, plane0, plane1, words);
    GUARD_END(B2_ERR_OUT_OF_MEMORY)
}

int b2_run_sim(b2_sim *s, uint64_t max_steps) {
    if (!s) return -B2_ERR_NULL;
    GUARD_BEGIN
    return s->e->run(max_steps, nullptr, nullptr, false);
    GUARD_END(-B2_ERR_OUT_OF_MEMORY)
}

int b2_get_wire_state(b2_sim *s, uint32_t wire, uint32_t *plane0, uint32_t *plane1, size_t words) {
    if (!s) return B2_ERR_NULL;
    GUARD_BEGIN
    return s->e->get_wire_state(wire, plane0, plane1, words);
    GUARD_END(B2_ERR_OUT_OF_MEMORY)
}

int b2_set_wire_drive_lanes(b2_sim *s, uint32_t wire, uint32_t bit, uint32_t lane_word0, uint32_t n_words, const uint64_t *value,
                            const uint64_t *valid) {
    if (!s) return B2_ERR_NULL;
    GUARD_BEGIN
    return s->e->set_wire_drive_lanes(wire, bit, lane_word0, n_words, value, valid);
    GUARD_END(B2_ERR_OUT_OF_MEMORY)
}

int b2_set_drives_lanes(b2_sim *s, const uint32_t *wires, uint32_t n, const uint64_t *value, const uint64_t *valid,
                        size_t src_stride_words) {
    if (!s) return B2_ERR_NULL;
    GUARD_BEGIN
    return s->e->set_drives_lanes(wires, n, value, valid, src_stride_words);
    GUARD_END(B2_ERR_OUT_OF_MEMORY)
}

int b2_fill_stimulus(b2_sim *s, const uint32_t *wires, uint32_t n, uint64_t seed, uint64_t lane_word_base, int mode) {
    if (!s) return B2_ERR_NULL;
    GUARD_BEGIN
    return s->e->fill_stimulus(wires, n, seed, lane_word_base, mode);
    GUARD_END(B2_ERR_OUT_OF_MEMORY)
}

int b2_run_sim_lanes(b2_sim *s, uint64_t max_steps, uint64_t *conflict, uint64_t *unsettled) {
    if (!s) return -B2_ERR_NULL;
    GUARD_BEGIN
    return s->e->run(max_steps, conflict, unsettled, false);
    GUARD_END(-B2_ERR_OUT_OF_MEMORY)
}

int b2_run_sim_lanes_async(b2_sim *s, uint64_t max_steps) {
    if (!s) return -B2_ERR_NULL;
    GUARD_BEGIN
    return s->e->run(max_steps, nullptr, nullptr, true);
    GUARD_END(-B2_ERR_OUT_OF_MEMORY)
}

int b2_get_wire_lanes(b2_sim *s, uint32_t wire, uint32_t bit, uint32_t lane_word0, uint32_t n_words, uint64_t *value,
                      uint64_t *valid) {
    if (!s) return B2_ERR_NULL;
    GUARD_BEGIN
    return s->e->get_wire_lanes(wire, bit, lane_word0, n_words, value, valid);
    GUARD_END(B2_ERR_OUT_OF_MEMORY)
}

int b2_gather_wires_device(b2_sim *s, const uint32_t *wires, uint32_t n, uint64_t *dev_value, uint64_t *dev_valid,
                           size_t dst_stride_words) {
    if (!s) return B2_ERR_NULL;
    GUARD_BEGIN
    return s->e->gather_wires(wires, n, dev_value, dev_valid, dst_stride_words, true);
    GUARD_END(B2_ERR_OUT_OF_MEMORY)
}

int b2_gather_wires_host(b2_sim *s, const uint32_t *wires, uint32_t n, uint64_t *value, uint64_t *valid, size_t dst_stride_words) {
    if (!s) return B2_ERR_NULL;
    GUARD_BEGIN
    return s->e->gather_wires(wires, n, value, valid, dst_stride_words, false);
    GUARD_END(B2_ERR_OUT_OF_MEMORY)
}

int b2_gather_wires_host_async(b2_sim *s, const uint32_t *wires, uint32_t n, uint64_t *value, uint64_t *valid,
                               size_t dst_stride_words) {
    if (!s) return B2_ERR_NULL;
    GUARD_BEGIN
    return s->e->gather_wires(wires, n, value, valid, dst_stride_words, false, true);
    GUARD_END(B2_ERR_OUT_OF_MEMORY)
}

int b2_pack_sim_state(b2_sim *s, const uint32_t *wires, uint32_t n, uint8_t *plane0, uint8_t *plane1, size_t cap_bytes,
                      uint64_t *out_bit_len) {
    if (!s) return B2_ERR_NULL;
    GUARD_BEGIN
    return s->e->pack_sim_state(wires, n, plane0, plane1, cap_bytes, out_bit_len);
    GUARD_END(B2_ERR_OUT_OF_MEMORY)
}

}  // extern "C"

// ------------------------------------------------------------------ server (GsimServer mirror)

namespace {
// enum ClientState { Building(SimulatorBuilder), Simulating(Simulator) }   lib.rs:5-14
struct ClientState {
    b2_builder builder;
    b2_sim *sim = nullptr;  // non-null = Simulating
    ~ClientState() { b2_sim_free(sim); }
};

// fn component_error_to_server_error   lib.rs:55-66
int component_error_to_server_error(int rc) {
    switch (rc) {
        case B2_OK: return B2S_OK;
        case B2_ERR_TOO_MANY_COMPONENTS: return B2S_OUT_OF_RESOURCES;
        case B2_ERR_INVALID_WIRE_ID: return B2S_INVALID_NET_ID;
        case B2_ERR_WIRE_WIDTH_MISMATCH: return B2S_WIDTH_MISMATCH;
        case B2_ERR_WIRE_WIDTH_INCOMPATIBLE: return B2S_WIDTH_INCOMPATIBLE;
        case B2_ERR_OFFSET_OUT_OF_RANGE: return B2S_OUT_OF_RANGE;
        case B2_ERR_TOO_FEW_INPUTS: return B2S_INVALID_INPUT_COUNT;
        case B2_ERR_INVALID_INPUT_COUNT: return B2S_INVALID_INPUT_COUNT;
        case B2_ERR_OUT_OF_MEMORY: return B2S_OUT_OF_RESOURCES;
        default: return B2S_OTHER;
    }
}
}  // namespace

struct b2_server {
    std::unordered_map<uint64_t, std::unique_ptr<ClientState>> clients;
    uint8_t bit_plane_0[32];
    uint8_t bit_plane_1[32];
    b2_server() {
        std::memset(bit_plane_0, 0, 32);
        std::memset(bit_plane_1, 0, 32);
    }
    // The reference panics on an unknown client id (lib.rs:26,30); a panic must not cross the C ABI.
    ClientState *get(uint64_t id) {
        auto it = clients.find(id);
        return it == clients.end() ? nullptr : it->second.get();
    }
};

extern "C" {

b2_server *b2_server_new(void) { return new (std::nothrow) b2_server(); }
void b2_server_free(b2_server *sv) { delete sv; }
uint64_t b2_server_max_clients(b2_server *) { return UINT64_MAX; }

void b2_server_client_connected(b2_server *sv, uint64_t client) {
    if (!sv) return;
    try {
        sv->clients[client].reset(new ClientState());  // insert replaces: a second connect resets the client
    } catch (...) {
    }
}

void b2_server_client_disconnected(b2_server *sv, uint64_t client) {
    if (sv) sv->clients.erase(client);
}

int b2_server_begin_build(b2_server *sv, uint64_t client) {
    if (!sv) return B2S_OTHER;
    if (!sv->get(client)) return B2S_OTHER;
    GUARD_BEGIN
    sv->clients[client].reset(new ClientState());
    return B2S_OK;
    GUARD_END(B2S_OUT_OF_RESOURCES)
}

int b2_server_end_build(b2_server *sv, uint64_t client) {
    ClientState *c = sv ? sv->get(client) : nullptr;
    if (!c) return B2S_OTHER;
    if (c->sim) return B2S_INVALID_STATE;
    c->sim = b2_build(&c->builder);
    return c->sim ? B2S_OK : B2S_OUT_OF_RESOURCES;
}

int b2_server_add_net(b2_server *sv, uint64_t client, uint8_t width, uint32_t *out_net) {
    ClientState *c = sv ? sv->get(client) : nullptr;
    if (!c) return B2S_OTHER;
    if (c->sim) return B2S_INVALID_STATE;
    if (width == 0) return B2S_OTHER;  // NonZeroU8 in the reference: unrepresentable
    return b2_add_wire(&c->builder, width, out_net) == B2_OK ? B2S_OK : B2S_OUT_OF_RESOURCES;
}

// gate_impl! (lib.rs:76-99): width must equal the output wire's width
static int check_output(ClientState *c, uint8_t width, uint32_t output) {
    uint8_t ow = 0;
    if (b2_builder_wire_width(&c->builder, output, &ow) != B2_OK) return B2S_INVALID_NET_ID;
    if (width != ow) return B2S_WIDTH_MISMATCH;
    return B2S_OK;
}

int b2_server_add_gate(b2_server *sv, uint64_t client, int kind, uint8_t width, const uint32_t *inputs, size_t n, uint32_t output,
                       uint32_t *out_cell) {
    ClientState *c = sv ? sv->get(client) : nullptr;
    if (!c) return B2S_OTHER;
    if (c->sim) return B2S_INVALID_STATE;
    if (kind < B2_AND || kind > B2_XNOR) return B2S_UNSUPPORTED;
    if (int rc = check_output(c, width, output)) return rc;
    return component_error_to_server_error(b2_add_gate(&c->builder, kind, inputs, n, output, out_cell));
}

int b2_server_add_not_gate(b2_server *sv, uint64_t client, uint8_t width, uint32_t input, uint32_t output, uint32_t *out_cell) {
    ClientState *c = sv ? sv->get(client) : nullptr;
    if (!c) return B2S_OTHER;
    if (c->sim) return B2S_INVALID_STATE;
    if (int rc = check_output(c, width, output)) return rc;
    return component_error_to_server_error(b2_add_not(&c->builder, input, output, out_cell));
}

int b2_server_add_mux(b2_server *sv, uint64_t client, uint8_t width, const uint32_t *inputs, size_t n, uint32_t output,
                      uint32_t *out_cell) {
    ClientState *c = sv ? sv->get(client) : nullptr;
    if (!c) return B2S_OTHER;
    if (c->sim) return B2S_INVALID_STATE;
    if (int rc = check_output(c, width, output)) return rc;
    if (n == 0 || !inputs) return B2S_INVALID_INPUT_COUNT;  // the reference indexes inputs[0] and would panic
    return component_error_to_server_error(b2_add_mux(&c->builder, inputs + 1, n - 1, inputs[0], output, out_cell));
}

int b2_server_set_net_drive(b2_server *sv, uint64_t client, uint32_t net, const uint8_t *bit_plane_0, size_t len0,
                            const uint8_t *bit_plane_1, size_t len1) {
    ClientState *c = sv ? sv->get(client) : nullptr;
    if (!c) return B2S_OTHER;
    if (!c->sim) return B2S_INVALID_STATE;
    if (len0 > 32 || len1 > 32) return B2S_OUT_OF_RANGE;  // the reference panics on the slice index (lib.rs:205-206)
    if ((len0 && !bit_plane_0) || (len1 && !bit_plane_1)) return B2S_OTHER;
    uint32_t words0[8] = {0}, words1[8] = {0};
    if (len0) std::memcpy(words0, bit_plane_0, len0);
    if (len1) std::memcpy(words1, bit_plane_1, len1);
    const size_t word_count = ((len0 < len1 ? len0 : len1) + 3) / 4;
    const int rc = b2_set_wire_drive(c->sim, net, words0, words1, word_count);
    if (rc == B2_OK) return B2S_OK;
    return rc == B2_ERR_INVALID_WIRE_ID ? B2S_INVALID_NET_ID : B2S_OTHER;
}

// fn simulation_result_to_server_result   lib.rs:68-74
int b2_server_eval(b2_server *sv, uint64_t client, uint64_t max_steps) {
    ClientState *c = sv ? sv->get(client) : nullptr;
    if (!c) return B2S_OTHER;
    if (!c->sim) return B2S_INVALID_STATE;
    switch (b2_run_sim(c->sim, max_steps)) {
        case B2_RUN_OK: return B2S_OK;
        case B2_RUN_MAX_STEPS_REACHED: return B2S_MAX_STEPS_REACHED;
        case B2_RUN_CONFLICT: return B2S_DRIVER_CONFLICT;
        default: return B2S_OTHER;
    }
}

int b2_server_get_net_state(b2_server *sv, uint64_t client, uint32_t net, uint8_t *out_width, const uint8_t **plane0,
                            const uint8_t **plane1) {
    ClientState *c = sv ? sv->get(client) : nullptr;
    if (!c || !out_width || !plane0 || !plane1) return B2S_OTHER;
    if (!c->sim) return B2S_INVALID_STATE;
    uint8_t width = 0;
    if (b2_wire_width(c->sim, net, &width) != B2_OK) return B2S_INVALID_NET_ID;  // the reference unwrap()s (lib.rs:228)
    uint32_t w0[8] = {0}, w1[8] = {0};
    if (b2_get_wire_state(c->sim, net, w0, w1, 8) != B2_OK) return B2S_OTHER;
    const size_t bytes = ((size_t)(width + 31) / 32) * 4;
    std::memcpy(sv->bit_plane_0, w0, bytes);
    std::memcpy(sv->bit_plane_1, w1, bytes);
    *out_width = width;
    *plane0 = sv->bit_plane_0;
    *plane1 = sv->bit_plane_1;
    return B2S_OK;
}

int b2_server_sim_state(b2_server *sv, uint64_t client, const uint32_t *nets, uint32_t n, uint8_t *plane0, uint8_t *plane1,
                        size_t cap_bytes, uint64_t *out_bit_len) {
    ClientState *c = sv ? sv->get(client) : nullptr;
    if (!c) return B2S_OTHER;
    if (!c->sim) return B2S_INVALID_STATE;
    const int rc = b2_pack_sim_state(c->sim, nets, n, plane0, plane1, cap_bytes, out_bit_len);
    if (rc == B2_OK) return B2S_OK;
    if (rc == B2_ERR_INVALID_WIRE_ID) return B2S_INVALID_NET_ID;
    return rc == B2_ERR_RANGE ? B2S_OUT_OF_RANGE : B2S_OTHER;
}

}  // extern "C"
