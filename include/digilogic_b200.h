/*
 * digilogic_b200.h — C ABI of the B200-native gate-level simulation engine.
 *
 * This is the drop-in boundary for the hot path of rj45/digilogic: everything the
 * reference's engine adapter `GsimServer` (crates/digilogic_gsim/src/lib.rs:101-239)
 * calls on the third-party crate gsim 1.1.4 has one entry point here, so that a
 * `B200Server: SimServer` written like that file can stand behind
 * `SimulationEngine::GsimCompute` (crates/digilogic/src/main.rs:312-318,373).
 *
 * Conventions: plain pointers and sizes only; caller-owned buffers; no exceptions or
 * aborts cross this boundary — every failure is a return code.  There is NO CPU
 * fallback: simulation entry points return B2_ERR_NO_DEVICE without a CUDA device.
 *
 * Value encoding (reference crates/digilogic/src/ui/palette.rs:326-341,
 * crates/digilogic_core/src/components.rs:104-110): two planes per bit,
 *   plane0 = value/state, plane1 = valid:  Z=(0,0)  X=(1,0)  0=(0,1)  1=(1,1).
 * A "lane" is one independent stimulus vector (= one gsim `Simulator`); lane l lives in
 * bit l%64 of 64-bit lane-word l/64 of every wire row.
 */
#ifndef DIGILOGIC_B200_H
#define DIGILOGIC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b2_builder b2_builder; /* replaces gsim::SimulatorBuilder (lib.rs:6,12) */
typedef struct b2_sim b2_sim;         /* replaces gsim::Simulator        (lib.rs:7)    */
typedef struct b2_server b2_server;   /* mirrors GsimServer              (lib.rs:16-22) */

/* Gate kinds accepted by b2_add_gate (0..5), plus the ids used in bulk adds / reports. */
enum b2_gate_kind { B2_AND = 0, B2_OR = 1, B2_XOR = 2, B2_NAND = 3, B2_NOR = 4, B2_XNOR = 5, B2_NOT = 6, B2_MUX = 7 };

/* 1..7 are gsim's AddComponentError variants in the order the reference maps them
   (crates/digilogic_gsim/src/lib.rs:55-66); >= 16 are this library's own. */
enum b2_status {
    B2_OK = 0,
    B2_ERR_TOO_MANY_COMPONENTS = 1,
    B2_ERR_INVALID_WIRE_ID = 2,
    B2_ERR_WIRE_WIDTH_MISMATCH = 3,
    B2_ERR_WIRE_WIDTH_INCOMPATIBLE = 4,
    B2_ERR_OFFSET_OUT_OF_RANGE = 5,
    B2_ERR_TOO_FEW_INPUTS = 6,
    B2_ERR_INVALID_INPUT_COUNT = 7,
    B2_ERR_NULL = 16,          /* a required pointer was NULL */
    B2_ERR_INVALID_ARG = 17,
    B2_ERR_NO_DEVICE = 18,     /* no CUDA device / kernels not loadable: there is no CPU path */
    B2_ERR_CUDA = 19,          /* a CUDA call failed; see b2_last_error() */
    B2_ERR_OUT_OF_MEMORY = 20,
    B2_ERR_RANGE = 21,         /* lane-word / bit range outside the store */
    B2_ERR_OUT_OF_WIRES = 22   /* add_wire returned None (-> ServerError::OutOfResources, lib.rs:138) */
};

/* gsim::SimulationRunResult (lib.rs:68-74). Run entry points return one of these, or
   -(b2_status) on failure. */
enum b2_run_result { B2_RUN_OK = 0, B2_RUN_MAX_STEPS_REACHED = 1, B2_RUN_CONFLICT = 2 };

/* Thread-local text of the last B2_ERR_CUDA / B2_ERR_NO_DEVICE on this thread ("" if none). */
const char *b2_last_error(void);
/* Library version string. */
const char *b2_version(void);
/* Number of kernel launches issued by this library in this process (bench.py's gpu_launches). */
uint64_t b2_kernel_launches(void);

/* ------------------------------------------------------------------------------------
 * Builder — replaces gsim::SimulatorBuilder
 * ---------------------------------------------------------------------------------- */

/* SimulatorBuilder::default()  (lib.rs:12) */
b2_builder *b2_builder_new(void);
void b2_builder_free(b2_builder *b);

/* SimulatorBuilder::add_wire(width) -> Option<WireId>  (lib.rs:135-139). width 1..255.
   Capacity: the compiled netlist indexes 1-bit nets and 1-bit gates with 32 bits, so a builder holds at most
   2^30 wire bits (B2_ERR_OUT_OF_WIRES beyond) and 2^30 bit-blasted gates / 2^31 bit-blasted gate inputs
   (B2_ERR_TOO_MANY_COMPONENTS beyond) — counted in 64 bits as components are added, never wrapped. */
int b2_add_wire(b2_builder *b, uint8_t width, uint32_t *out_wire);
/* Batch extension: n wires of one width with consecutive ids starting at *out_first. */
int b2_add_wires(b2_builder *b, uint8_t width, uint32_t n, uint32_t *out_first);
/* SimulatorBuilder::get_wire_width(wire)  (lib.rs:87-92,157-162,178-183) */
int b2_builder_wire_width(const b2_builder *b, uint32_t wire, uint8_t *out_width);

/* SimulatorBuilder::add_{and,or,xor,nand,nor,xnor}_gate(inputs, output) -> Result<ComponentId, AddComponentError>
   (lib.rs:76-99,141-146).  kind = B2_AND..B2_XNOR; n >= 2 inputs, all of the output's width. */
int b2_add_gate(b2_builder *b, int kind, const uint32_t *inputs, size_t n, uint32_t output, uint32_t *out_cell);
/* SimulatorBuilder::add_not_gate(input, output)  (lib.rs:148-167) */
int b2_add_not(b2_builder *b, uint32_t input, uint32_t output, uint32_t *out_cell);
/* SimulatorBuilder::add_multiplexer(inputs, select, output)  (lib.rs:169-192): n = 2^k data
   inputs of the output's width, select of width k. */
int b2_add_mux(b2_builder *b, const uint32_t *inputs, size_t n, uint32_t select, uint32_t output, uint32_t *out_cell);
/* SimulatorBuilder::add_buffer(input, enable, output) — tri-state buffer.  gsim has it; the reference never
   calls it (SimServer has no such method, crates/digilogic_netcode/src/server.rs:43-102; the Yosys
   importer's `TriBuf` arm is todo!(), crates/digilogic_serde/src/yosys.rs:197) — SURVEY.md 8f row f4.
   enable = 1: output = input with High-Z bits read as X; enable = 0: High-Z; enable Z/X: Undefined.
   A net with a buffer among its drivers may have any number of drivers; it is resolved every step
   (plane-wise OR, conflict where two drivers or the external drive are not High-Z) and such netlists
   always run the unit-delay path.  Errors: InvalidWireId, WireWidthMismatch (input vs output),
   WireWidthIncompatible (enable wider than one bit). */
int b2_add_buffer(b2_builder *b, uint32_t input, uint32_t enable, uint32_t output, uint32_t *out_cell);
/* SimulatorBuilder::add_add / add_sub(input_a, input_b, output): output = a + b or a - b modulo 2^width; the whole output is
   Undefined if any bit of either input is not a valid 0/1.  All three wires have one width (WireWidthMismatch).  No call site
   in the reference (Yosys Add/Sub are todo!(), crates/digilogic_serde/src/yosys.rs:188-189).  Each result bit is one gate on
   the CSR kernel reading the input bits below it: correct for every width, quadratic in it. */
int b2_add_add(b2_builder *b, uint32_t input_a, uint32_t input_b, uint32_t output, uint32_t *out_cell);
int b2_add_sub(b2_builder *b, uint32_t input_a, uint32_t input_b, uint32_t output, uint32_t *out_cell);
/* SimulatorBuilder::add_mul(input_a, input_b, output): output = a * b modulo 2^width, Undefined if any input bit is not a valid 0/1;
   one width for all three wires (WireWidthMismatch).  No call site in the reference (Yosys Mul is todo!(), yosys.rs:190).  Each
   result bit is one gate on the CSR kernel with a bit-sliced column adder: correct for every width, cubic in it per wire. */
int b2_add_mul(b2_builder *b, uint32_t input_a, uint32_t input_b, uint32_t output, uint32_t *out_cell);
/* SimulatorBuilder::add_compare_{equal,not_equal,less_than,greater_than,less_than_or_equal,greater_than_or_equal}{,_signed}
   (input_a, input_b, output): inputs of one width (WireWidthMismatch), a one-bit output (WireWidthIncompatible), Undefined if any
   input bit is not a valid 0/1.  No call site in the reference (Yosys Lt/Le/Eq/Ne/Ge/Gt are todo!(), yosys.rs:182-187). */
enum b2_compare_op { B2_CMP_EQ = 0, B2_CMP_NE = 1, B2_CMP_LTU = 2, B2_CMP_GTU = 3, B2_CMP_LEU = 4, B2_CMP_GEU = 5,
                     B2_CMP_LTS = 6, B2_CMP_GTS = 7, B2_CMP_LES = 8, B2_CMP_GES = 9 };
int b2_add_compare(b2_builder *b, int op, uint32_t input_a, uint32_t input_b, uint32_t output, uint32_t *out_cell);
/* SimulatorBuilder::add_left_shift / add_logical_right_shift / add_arithmetic_right_shift(input, shift_amount, output): output as
   wide as the input (WireWidthMismatch), the amount an unsigned value of at most 8 bits (WireWidthIncompatible); amounts >= width
   shift everything out (zeros, or the sign); Undefined if any bit of either input is not a valid 0/1.  No call site in the
   reference (Yosys Shl/Sshl/Shr/Sshr are todo!(), yosys.rs:174-177). */
enum b2_shift_kind { B2_SHIFT_LEFT = 0, B2_SHIFT_LOGICAL_RIGHT = 1, B2_SHIFT_ARITHMETIC_RIGHT = 2 };
int b2_add_shift(b2_builder *b, int kind, uint32_t input, uint32_t amount, uint32_t output, uint32_t *out_cell);
/* SimulatorBuilder::add_horizontal_{and,or,xor,nand,nor,xnor}_gate(input, output): the one-bit output is the gate folded over
   all bits of `input` (kind = B2_AND .. B2_XNOR); compiled into an ordinary n-ary gate over the input's bit rows.  No call site
   in the reference (Yosys ReduceAnd/Or/Xor/Xnor/Bool are todo!(), crates/digilogic_serde/src/yosys.rs:163-167).
   Errors: InvalidWireId; WireWidthIncompatible when the output is wider than one bit. */
int b2_add_horizontal_gate(b2_builder *b, int kind, uint32_t input, uint32_t output, uint32_t *out_cell);
/* SimulatorBuilder::add_register(data_in, data_out, enable, clock) — gsim's clocked register (no call site in the
   reference: SimServer has no such method and the Yosys importer's Dff/Dffe arms are todo!(),
   crates/digilogic_serde/src/yosys.rs:199-200).  Restated (parity unpinned like the rest): on a rising clock edge — the
   register's previous update saw a valid 0, this one sees a valid 1 — enable = 1 loads data_in (High-Z bits read X),
   enable = 0 keeps the value, enable Z/X makes it Undefined; data_out carries the stored value, Undefined until the first
   load.  enable and clock are one bit wide (WireWidthIncompatible), data_in as wide as data_out (WireWidthMismatch).
   Netlists with registers always run the unit-delay path. */
int b2_add_register(b2_builder *b, uint32_t data_in, uint32_t data_out, uint32_t enable, uint32_t clock, uint32_t *out_cell);
/* SimulatorBuilder::add_slice(input, offset, output) and add_merge(inputs, output) — bit re-wiring components of
   gsim, one step of delay like every component, High-Z input bits read as X (no call site in the reference;
   `OffsetOutOfRange` of its error map, lib.rs:55-66, belongs to add_slice).
   slice: output = input[offset .. offset + width(output)); OffsetOutOfRange when that exceeds the input.
   merge: output = inputs concatenated, inputs[0] in the low bits; TooFewInputs for none, WireWidthMismatch
   when the widths do not add up to the output's. */
int b2_add_slice(b2_builder *b, uint32_t input, uint32_t offset, uint32_t output, uint32_t *out_cell);
int b2_add_merge(b2_builder *b, const uint32_t *inputs, size_t n, uint32_t output, uint32_t *out_cell);
/* Batch extension: n one- or two-input gates at once (kinds[i] in B2_AND..B2_NOT; in1 ignored
   for B2_NOT), validated exactly like the single calls.  All or nothing: on the first invalid entry
   its error is returned and the builder is left as it was before the call. */
int b2_add_gates2(b2_builder *b, const uint8_t *kinds, const uint32_t *in0, const uint32_t *in1, const uint32_t *out, size_t n);

/* SimulatorBuilder::build()  (lib.rs:123-133): compiles the netlist (bit-blast, driver
   analysis, levelisation, kind/level-sorted SoA arrays).  Like `std::mem::take(builder).build()`
   it leaves *b an empty default builder.  Host-only: no CUDA call is made until lanes are
   allocated.  Returns NULL only on allocation failure or b == NULL. */
b2_sim *b2_build(b2_builder *b);
void b2_sim_free(b2_sim *s);

/* Builder options: the points of gsim 1.1.4's behaviour that were reconstructed without its source,
   one switch each, default 0 = the reading oracle/gsim_oracle.c takes by default.  They exist so that a verdict from the real
   crate (oracle/gsim_diff) is followed by flipping a switch, not by editing a kernel; each is tested on every device path in
   lockstep with the oracle's matching switch.  (A.6-1 and A.6-2 are simulator options: B2_OPT_STEP_BUDGET_INCLUSIVE,
   B2_OPT_CONFLICT_NEEDS_DISAGREE.)  The switches survive b2_build, which otherwise leaves the builder empty. */
enum b2_builder_option {
    B2_BOPT_MUX_Z_PASSTHROUGH = 1,    /* A.6-4: a multiplexer passes a High-Z bit of the selected input on as High-Z (default: X).
                                         Its output can then float, so a net with a multiplexer among its drivers is resolved
                                         per step like a tri-state net */
    B2_BOPT_INITIAL_UNDEFINED = 2,    /* A.6-5: wires and component outputs start Undefined (default: High-Z) */
    B2_BOPT_BUFFER_Z_PASSTHROUGH = 3, /* A.6-6: an enabled buffer passes a High-Z input bit on as High-Z (default: X) */
    B2_BOPT_COPY_Z_PASSTHROUGH = 4    /* A.6-7: slices and merges pass a High-Z input bit on as High-Z (default: X) */
};
int b2_builder_set_option(b2_builder *b, int option, int64_t value);

/* A second simulator of the same compiled netlist (batch extension): it shares the host arrays and the device copy of the
   netlist with `s` — nothing is recompiled or uploaded again — and has its own lanes, wire-state store, drives and stream.
   Options set on `s` so far are copied.  NULL on allocation failure. */
b2_sim *b2_sim_clone(const b2_sim *s);

/* ------------------------------------------------------------------------------------
 * Simulator — replaces gsim::Simulator
 * ---------------------------------------------------------------------------------- */

typedef struct b2_sim_info {
    uint32_t n_wires;       /* wires as added by the caller */
    uint32_t n_rows;        /* 1-bit rows of the wire-state store (sources first, then gate outputs) */
    uint32_t n_sources;     /* rows without a gate driver (primary inputs and floating nets) */
    uint32_t n_gates2;      /* 1-bit gates with one or two inputs (fast kernels) */
    uint32_t n_gates_wide;  /* 1-bit gates with four or more inputs and multiplexer slices (CSR kernel) */
    uint32_t depth;         /* number of levels; 0 when cyclic */
    uint32_t cyclic;        /* 1 if the netlist has a combinational feedback loop */
    uint32_t multi_driver;  /* 1 if some wire has two or more gate drivers (every run conflicts) */
    uint64_t n_lanes;       /* lanes allocated (0 before b2_sim_set_lanes / first use) */
    uint32_t lane_words;    /* 64-lane words per row in use */
    uint32_t row_stride;    /* allocated words per row */
    uint64_t store_bytes;   /* device bytes held by this simulator */
    uint32_t last_mode;     /* 0 none, 1 levelised, 2 unit-delay (Jacobi) in HBM, 3 unit-delay resident in shared memory */
    uint32_t n_gates3;      /* 1-bit gates with three inputs (fast kernels, third operand) */
    uint64_t last_applications; /* whole-netlist evaluations made by the last run */
} b2_sim_info;

int b2_sim_get_info(const b2_sim *s, b2_sim_info *out);
/* Level-pair fusion of the last levelised run (B2_OPT_PAIR_FUSION): launches of one pass, gates evaluated as the second gate of an
   item, rows never stored; all 0 when the last run used the per-level kernels.  Any pointer may be NULL. */
int b2_sim_get_pair_info(const b2_sim *s, uint64_t *launches, uint64_t *children, uint64_t *unstored);
/* The level-pair plan this netlist would get with `keep_wires` as the designated outputs, computed on the host (no device
   needed): out[0] launches, out[1] items, out[2] gates evaluated as the second gate of an item, out[3] rows never stored,
   out[4] rows dropped from L2 by a later launch, out[5] rows dropped right after their store.  B2_ERR_INVALID_ARG when the
   netlist is not eligible (feedback, tri-state nets, registers, gates with three or more inputs). */
/* Store rows the unit-delay work-list kernel stages in shared memory (B2_OPT_HOT_STAGING): out_rows[0 .. *out_n), at most 4. */
int b2_sim_get_hot_rows(b2_sim *s, uint32_t out_rows[4], uint32_t *out_n);
int b2_sim_plan_pairs(const b2_sim *s, const uint32_t *keep_wires, uint32_t n_keep, uint64_t out[6]);

/* Simulator::get_wire_width(wire)  (lib.rs:228) */
int b2_wire_width(const b2_sim *s, uint32_t wire, uint8_t *out_width);

/* Device / stream selection (before the first call that touches the device). `stream` is a
   cudaStream_t used as given (e.g. torch.cuda.current_stream().cuda_stream; 0 = CUDA's legacy
   default stream).  Without this call the simulator creates a private non-blocking stream. */
int b2_sim_set_device(b2_sim *s, int device);
int b2_sim_set_stream(b2_sim *s, void *stream);
/* Engine options (all default to the values shown). */
enum b2_option {
    B2_OPT_RESIDENT_KERNEL = 1 /* 1: netlists whose store fits in shared memory run the whole run_sim loop in
                                  one on-chip kernel when the path is unit-delay or the lane count is tiny; 0: never */
    ,
    B2_OPT_VALID_ELISION = 2   /* default 0.  1: in levelised runs, rows that are provably valid on every lane of the
                                  store (all fan-ins valid, starting from drives whose valid plane is all-ones) keep no
                                  valid plane in HBM — it is neither read nor written and is re-created on read-back.
                                  Results are bit-identical; rows that may carry X/Z take the full two-plane path. */
    ,
    B2_OPT_LEVEL_GRAPH = 3     /* default 1.  The per-level launches of the levelised path are captured once into a
                                  CUDA graph and replayed per run. */
    ,
    B2_OPT_EVAL_UNROLL = 4,    /* default 0 = automatic.  2 or 4: gates per thread in the per-level kernel (tuning knob). */
    B2_OPT_STEP_BUDGET_INCLUSIVE = 6, /* default 0.  gsim's run_sim is restated as `if steps > max_steps { MaxStepsReached }`
                                  (at most max_steps + 1 steps after begin_sim).  1 switches to `>=`, should the differential
                                  run against the real crate (oracle/gsim_diff) say so — SURVEY.md A.6-1. */
    B2_OPT_FRONTIER = 5,       /* default 1.  Unit-delay runs in HBM re-evaluate a gate only if one of its fan-in rows or
                                  its own output row changed (on any lane) in the previous application. */
    B2_OPT_CONFLICT_NEEDS_DISAGREE = 8, /* default 0.  gsim's driver conflict is restated as "two operands of a wire are not
                                  High-Z" (SURVEY.md A.6-2); 1 selects the alternative reading in which two valid operands that
                                  agree do not conflict — the oracle's `conflict_needs_disagree` switch, for the day a run of real
                                  gsim decides.  Refused (InvalidArgument) for a netlist with two plain gates on one net. */
    B2_OPT_L2_HINTS = 9,       /* default 0 (measured: no gain).  Levelised pass: fan-in rows produced more than one level below the reading gate are
                                  loaded with an L2 evict-first hint (createpolicy + ld.L2::cache_hint), so that these one-off reads
                                  do not displace the rows the previous level has just written and the next kernel re-reads. */
    B2_OPT_PDL = 11,           /* Levelised pass: consecutive level kernels of a simulator are launched with programmatic stream
                                  serialisation (griddepcontrol): the next level's launch and index loads overlap this level's tail.
                                  0 off, 1 dependents released when every CTA has started, 2 (default) when every CTA is about to finish. */
    B2_OPT_PAIR_FUSION = 12,   /* 0 off, 1 on, 2 (default) automatic: on up to 16 384 gates per level on average.  Measured: the 32x32
                                  multiplier of config 2 (32 gates per level) settles 2^20 vectors in 0.54 ms instead of 0.77 ms, config 3
                                  (5000 per level) gains 5 %, config 5 (50 000 per level) loses 2 % and keeps the per-level kernels.  Levelised pass of a simulator that need not keep rows after their last reader (the replicas
                                  of a batch, B2_BATCH_OPT_DISCARD) and whose netlist has only one/two-input gates: two dependent gates
                                  are evaluated by one thread — a launch holds items of one gate whose fan-ins are older than the launch
                                  plus up to two gates reading its output, which they get in registers; a row read only inside its item
                                  is never stored.  Half the launches, 17 % less traffic between the SMs and L2; kept rows are bit-identical. */
    B2_OPT_HOT_STAGING = 13,   /* default 0 (measured on config 4: 7 % fewer L2 reads, 2-3 % slower — the 126 MB L2 serves such rows as
                                  well).  Unit-delay work-list kernel (k_eval_list): the rows with the largest fan-out — at most four,
                                  each read by a hundred gates or more: a clock, a reset — are copied once per CTA into shared memory
                                  (cp.async.bulk + mbarrier) and the listed gates take those operands from there instead of from L2. */
    B2_OPT_PERIOD2_SKIP = 10,  /* default 1.  Unit-delay loops (device-side loop, resident kernel) detect an exact period-2 oscillation
                                  — every row an application writes equals the row two applications earlier — and jump ahead by an
                                  even number of applications: the state, lane results and step accounting are those of walking
                                  every step (the step function depends on the state alone), an oscillating latch under
                                  max_steps = 10 000 costs a dozen applications instead of 10 003. */
    B2_OPT_DEVICE_LOOP = 7     /* default 1.  Unit-delay runs in HBM execute the whole run_sim loop in one cooperative launch
                                  (grid barriers between the phases of an application, stop decision on the device);
                                  0: one set of launches per application, stop decision on the host. */
};
int b2_sim_set_option(b2_sim *s, int option, int64_t value);
/* Block until everything queued by this simulator has finished. */
int b2_sim_synchronize(b2_sim *s);

/* Batch extension: allocate the HBM-resident wire-state store for n_lanes independent
   stimulus lanes (default 1 = gsim behaviour).  Resets all lanes to the cold initial state
   (every wire and drive High-Z). */
int b2_sim_set_lanes(b2_sim *s, uint64_t n_lanes);
/* Largest n_lanes whose store would fit in `bytes` of device memory (0 = use currently free memory). */
uint64_t b2_sim_max_lanes(const b2_sim *s, uint64_t bytes);

/* LogicState::from_bit_planes(&p0[..words], &p1[..words]) + Simulator::set_wire_drive(wire, &state)
   (lib.rs:194-214).  plane words are little-endian u32, missing words are High-Z.  With more
   than one lane the drive is applied to every lane.  Like gsim's, the call is a host-side write: the
   drive is validated and queued, and reaches the device in one upload + one launch before the next
   run (or before any other drive setter, so the call order still decides); the last drive per wire
   wins.  A device error of that deferred upload is reported by the run. */
int b2_set_wire_drive(b2_sim *s, uint32_t wire, const uint32_t *plane0, const uint32_t *plane1, size_t words);

/* Simulator::run_sim(max_steps)  (lib.rs:216-219).  Returns b2_run_result of lane 0 when one
   lane is allocated; with more lanes the worst result over all lanes (conflict > max-steps > ok). */
int b2_run_sim(b2_sim *s, uint64_t max_steps);

/* Simulator::get_wire_state(wire) + LogicState::to_bit_planes(width)  (lib.rs:228-230): lane 0.
   Writes ceil(width/32) words per plane; `words` is the capacity of each buffer.  The reference reads
   every net back after each eval, one call per net (netcode server.rs:375-383): the first call after a
   run takes a bit-packed snapshot of lane 0 of all wires (one launch + one copy), later calls are host
   reads of that snapshot. */
int b2_get_wire_state(b2_sim *s, uint32_t wire, uint32_t *plane0, uint32_t *plane1, size_t words);

/* ---- batch-lane extension (no reference counterpart: gsim has one lane per Simulator) ---- */

/* Drive bit `bit` of `wire` in lanes [64*lane_word0, 64*(lane_word0+n_words)) from host words. */
int b2_set_wire_drive_lanes(b2_sim *s, uint32_t wire, uint32_t bit, uint32_t lane_word0, uint32_t n_words,
                            const uint64_t *value, const uint64_t *valid);
/* Drive bit 0 of n wires for all lane-words from host planes laid out [n][src_stride]: one
   staged copy + one scatter kernel on a private copy stream.  Returns when the host planes
   have been consumed — it does not wait for settles still running on the main stream, so the
   drives of the next tile can be uploaded while the current tile settles. */
int b2_set_drives_lanes(b2_sim *s, const uint32_t *wires, uint32_t n, const uint64_t *value, const uint64_t *valid,
                        size_t src_stride_words);
/* Counter-based stimulus generated on the device: input i (position in `wires`), global
   lane-word w = lane_word_base + local word:
     value = splitmix64(seed ^ (i+1)*0x9E3779B97F4A7C15 ^ w*0xD1B54A32D192ED03)
     mode 0: valid = ~0 (two-valued);  mode 1: valid = splitmix64(value ^ 1) | splitmix64(value ^ 2)
   (about 25 % of the lanes not valid: X where value=1, Z where value=0). */
int b2_fill_stimulus(b2_sim *s, const uint32_t *wires, uint32_t n, uint64_t seed, uint64_t lane_word_base, int mode);

/* run_sim(max_steps) on every lane.  conflict / unsettled (each lane_words long, may be NULL)
   receive the lanes that ended in Err(conflict) / MaxStepsReached.  Returns the worst result. */
int b2_run_sim_lanes(b2_sim *s, uint64_t max_steps, uint64_t *conflict, uint64_t *unsettled);
/* Same, but nothing is copied to the host and the call does not wait for the device when the
   levelised path is taken (the returned result is then B2_RUN_OK if no wire can conflict). */
int b2_run_sim_lanes_async(b2_sim *s, uint64_t max_steps);

/* Read bit `bit` of `wire` for lane-words [lane_word0, lane_word0+n_words) into host buffers. */
int b2_get_wire_lanes(b2_sim *s, uint32_t wire, uint32_t bit, uint32_t lane_word0, uint32_t n_words,
                      uint64_t *value, uint64_t *valid);
/* Gather bit 0 of n wires, all lane-words, into DEVICE planes laid out [n][dst_stride_words]
   (e.g. a torch tensor that is then all-gathered with NCCL). */
int b2_gather_wires_device(b2_sim *s, const uint32_t *wires, uint32_t n, uint64_t *dev_value, uint64_t *dev_valid,
                           size_t dst_stride_words);
/* Same into HOST planes [n][dst_stride_words] (staged through a device buffer, one copy per plane). */
int b2_gather_wires_host(b2_sim *s, const uint32_t *wires, uint32_t n, uint64_t *value, uint64_t *valid,
                         size_t dst_stride_words);

/* Same, but returns once the gather and the copy are queued (the copy runs on a private copy
   stream and overlaps later settles); the host planes (pinned memory) are valid after the next
   b2_sim_synchronize().  At most one such read-back is in flight: a second call waits on the
   device, not on the host, for the first copy to drain the staging buffer. */
int b2_gather_wires_host_async(b2_sim *s, const uint32_t *wires, uint32_t n, uint64_t *value, uint64_t *valid,
                               size_t dst_stride_words);

/* Bulk read-back in the netcode `SimState` layout (replaces the per-net get_net_state +
   SimState::push_net loop of crates/digilogic_netcode/src/server.rs:375-383 and
   src/lib.rs:241-270): lane 0 of `n` wires appended LSB-first with no padding between
   nets.  plane buffers must hold bit_len/8 + 1 bytes (the "at least one unused bit"
   invariant, lib.rs:168).  *out_bit_len receives the total bit count. */
int b2_pack_sim_state(b2_sim *s, const uint32_t *wires, uint32_t n, uint8_t *plane0, uint8_t *plane1, size_t cap_bytes,
                      uint64_t *out_bit_len);

/* ------------------------------------------------------------------------------------
 * Batch runner and multi-GPU gather (no reference counterpart: gsim runs one stimulus vector per Simulator::run_sim,
 * crates/digilogic_gsim/src/lib.rs:216-219; a batch of L lanes is L such runs of one netlist).
 *
 * b2_batch owns what a host would otherwise script around b2_sim: the tiling of the lanes, `n_slots` simulator replicas
 * (sharing ONE device copy of the compiled netlist) that work on alternate tiles on their own CUDA streams, the stimulus
 * source and the destinations of the designated output rows.  Every tile runs with exactly the semantics of
 * b2_run_sim_lanes (levelised pass replayed from a CUDA graph for an acyclic netlist, unit-delay loop otherwise).
 * B2_BATCH_OPT_SWEEP selects an alternative engine for acyclic single-driver netlists whose inputs are undriven nets and a
 * step budget that covers the depth: the persistent level-sweep kernel (csrc/sweep.cuh), one launch per batch, which also
 * reads page-locked host stimulus in place and stores output rows into peer GPUs itself; on BASELINE config 3 it is 8-15 %
 * slower than the replicas (profiles/README.md), so it is not the default.
 * Inputs and outputs are 1-bit wires.  All work is queued on the batch's stream; b2_batch_wait blocks.
 * ---------------------------------------------------------------------------------- */
typedef struct b2_batch b2_batch;
typedef struct b2_comm b2_comm;   /* an NCCL communicator, one rank per GPU */

typedef struct b2_batch_info {
    uint32_t tile_words, n_slots;
    uint32_t sweep;            /* 1: the level-sweep kernel is selected and serves this netlist; 0: simulator replicas */
    uint32_t tasks_per_pass;   /* sweep: tasks of one tile (after the first run) */
    uint32_t grid_ctas, gates_per_task;
    uint32_t n_ranks, rank, peers_mapped;
    uint32_t pair_launches;    /* replicas: launches per tile of the level-pair plan (B2_OPT_PAIR_FUSION) after the first run; 0 = one per level */
    uint64_t store_bytes;
    uint64_t words_local;      /* lane-words per rank of the gathered planes (0 without a communicator) */
    uint64_t last_tiles, last_tickets;
} b2_batch_info;

enum b2_batch_option {
    B2_BATCH_OPT_SWEEP = 1,       /* default 0; 1: the persistent level-sweep kernel where the netlist allows it */
    B2_BATCH_OPT_TASK_ITERS = 2,  /* default 1: 256-thread iterations per gate task of the sweep (tuning) */
    B2_BATCH_OPT_UNROLL = 3,      /* default 4: gates per thread of the one/two-input class in the sweep (2 or 4) */
    B2_BATCH_OPT_CTAS_PER_SM = 4, /* default 0 = as many as fit */
    B2_BATCH_OPT_DISCARD = 6,     /* default 1: a batch run only hands out the designated output rows, so the replicas do not keep a
                                     row after its last reader: rows read by nobody, or only by the level directly above, are
                                     dropped from L2 (discard.global.L2) by a later level's kernel before they are written back to
                                     HBM.  Every gate is still evaluated and its row written; outputs are bit-identical (tested). */
    B2_BATCH_OPT_AUTO_GATHER = 5  /* default 1: with a communicator bound every run ends with b2_batch_gather_outputs, so a
                                     step is one call; 0: the host calls it */
};

/* `sim` names the compiled netlist (and the engine options to copy); it is not modified and may be freed afterwards.
   tile_words: 64-lane words per tile and slot (even; the sweep takes at most 512); n_slots: tiles in flight (1..64). */
int b2_batch_new(const b2_sim *sim, const uint32_t *inputs, uint32_t n_inputs, const uint32_t *outputs, uint32_t n_outputs,
                 uint32_t tile_words, uint32_t n_slots, b2_batch **out);
void b2_batch_free(b2_batch *b);
int b2_batch_set_device(b2_batch *b, int device);
int b2_batch_set_stream(b2_batch *b, void *stream);   /* cudaStream_t, used as given */
int b2_batch_set_option(b2_batch *b, int option, int64_t value);
int b2_batch_get_info(b2_batch *b, b2_batch_info *out);

/* One settle (run_sim(max_steps) on every lane) of lanes [0, 64 * n_words) of the batch, queued on the batch's stream.
   Every lane is a freshly built simulator (gsim: SimulatorBuilder::build() then one run_sim): the slots are reused from tile to
   tile, so no state is carried from one run to the next — a clocked sequence over resident lanes is what b2_sim_set_lanes +
   b2_run_sim_lanes are for.
   Stimulus: generated on the device exactly like b2_fill_stimulus (input i = position in `inputs`, global lane-word
   lane_word_begin + w), or read from planes [n_inputs][in_stride_words] in device memory or in host memory (copied tile by
   tile on a copy stream while earlier tiles settle; the sweep engine reads page-locked memory in place over PCIe and
   page-locks pageable memory for the duration of the run).
   Outputs: planes [n_outputs][out_stride_words], device or host memory, may be NULL when a communicator is bound.
   *_ring_words (0 = none): the planes hold that many words per row and are used as a ring (a multiple of tile_words).
   With a communicator bound, the output rows also go to this rank's share of the gathered planes — on every rank whose
   planes are mapped (b2_batch_import_gather_handles), stored over NVLink by the kernel that reads the rows of a tile,
   while the other slots' tiles settle. */
int b2_batch_run_generated(b2_batch *b, uint64_t seed, int mode, uint64_t lane_word_begin, uint64_t n_words, uint64_t max_steps,
                           uint64_t *out_value, uint64_t *out_valid, size_t out_stride_words, size_t out_ring_words);
int b2_batch_run_planes(b2_batch *b, const uint64_t *in_value, const uint64_t *in_valid, size_t in_stride_words, size_t in_ring_words,
                        uint64_t n_words, uint64_t max_steps, uint64_t *out_value, uint64_t *out_valid, size_t out_stride_words,
                        size_t out_ring_words);
/* Blocks until everything queued has finished (including a pending gather); *out_worst (may be NULL) = worst b2_run_result
   over all lanes (and, with a communicator, all ranks) since the previous wait. */
int b2_batch_wait(b2_batch *b, int *out_worst);

/* Multi-GPU (lanes shard over GPUs, a full netlist replica per GPU, no traffic during the settle, the output
   planes gathered at the end).  The host moves two small blobs between the ranks by whatever channel it has: the 128-byte
   NCCL unique id (rank 0 -> all) and one 128-byte handle of each rank's gathered planes (all -> all). */
int b2_comm_get_unique_id(uint8_t id[128]);                                   /* ncclGetUniqueId */
int b2_comm_init_rank(b2_comm **out, int n_ranks, int rank, const uint8_t id[128], int device);  /* ncclCommInitRank */
void b2_comm_free(b2_comm *c);
/* Allocates this rank's gathered planes [n_ranks][n_outputs][words_local] (value and valid) and binds the communicator. */
int b2_batch_bind_comm(b2_batch *b, b2_comm *comm, uint64_t words_local);
/* CUDA IPC handles of this rank's gathered planes (value: bytes 0-63, valid: 64-127) / the handles of all ranks
   [n_ranks][128] in rank order: from then on every run stores its output rows into all ranks' planes directly. */
int b2_batch_export_gather_handle(b2_batch *b, uint8_t handle[128]);
int b2_batch_import_gather_handles(b2_batch *b, const uint8_t *handles);
/* Completes the gather of the last run (SURVEY.md 8b `b2_gather_outputs(sim, ncclComm...)`): with mapped peers the rows
   are already in place and the ranks only meet (a MAX all-reduce of the run result); otherwise ncclAllGather moves them. */
int b2_batch_gather_outputs(b2_batch *b);
/* Device pointers of this rank's gathered planes, valid after b2_batch_gather_outputs in stream order. */
int b2_batch_gathered(b2_batch *b, uint64_t **dev_value, uint64_t **dev_valid);

/* ------------------------------------------------------------------------------------
 * Server — mirrors `impl SimServer for GsimServer` (crates/digilogic_gsim/src/lib.rs:101-239;
 * trait at crates/digilogic_netcode/src/server.rs:43-102).  Return value: 0 = Ok(..), else
 * 1 + the ordinal of the ServerError variant (crates/digilogic_netcode/src/lib.rs:67-85).
 * ---------------------------------------------------------------------------------- */
enum b2_server_error {
    B2S_OK = 0,
    B2S_UNSUPPORTED = 1, B2S_OTHER = 2, B2S_INVALID_STATE = 3, B2S_OUT_OF_RESOURCES = 4, B2S_INVALID_NET_ID = 5,
    B2S_WIDTH_MISMATCH = 6, B2S_WIDTH_INCOMPATIBLE = 7, B2S_OUT_OF_RANGE = 8, B2S_INVALID_INPUT_COUNT = 9,
    B2S_MAX_STEPS_REACHED = 10, B2S_DRIVER_CONFLICT = 11
};

b2_server *b2_server_new(void);                                   /* GsimServer::default() */
void b2_server_free(b2_server *sv);
uint64_t b2_server_max_clients(b2_server *sv);                    /* lib.rs:105-107 */
void b2_server_client_connected(b2_server *sv, uint64_t client);  /* lib.rs:109-111 */
void b2_server_client_disconnected(b2_server *sv, uint64_t client); /* lib.rs:113-115 */
int b2_server_begin_build(b2_server *sv, uint64_t client);        /* lib.rs:117-121 */
int b2_server_end_build(b2_server *sv, uint64_t client);          /* lib.rs:123-133 */
int b2_server_add_net(b2_server *sv, uint64_t client, uint8_t width, uint32_t *out_net); /* lib.rs:135-139 */
/* gate_impl!(add_*_gate) lib.rs:76-99: kind = B2_AND..B2_XNOR */
int b2_server_add_gate(b2_server *sv, uint64_t client, int kind, uint8_t width, const uint32_t *inputs, size_t n,
                       uint32_t output, uint32_t *out_cell);
int b2_server_add_not_gate(b2_server *sv, uint64_t client, uint8_t width, uint32_t input, uint32_t output,
                           uint32_t *out_cell);           /* lib.rs:148-167 */
/* inputs[0] is the select net, inputs[1..] the data nets (lib.rs:185-187) */
int b2_server_add_mux(b2_server *sv, uint64_t client, uint8_t width, const uint32_t *inputs, size_t n, uint32_t output,
                      uint32_t *out_cell);                /* lib.rs:169-192 */
int b2_server_set_net_drive(b2_server *sv, uint64_t client, uint32_t net, const uint8_t *bit_plane_0, size_t len0,
                            const uint8_t *bit_plane_1, size_t len1); /* lib.rs:194-214 */
int b2_server_eval(b2_server *sv, uint64_t client, uint64_t max_steps); /* lib.rs:216-219 */
/* lib.rs:221-238: the two returned plane pointers point at the server's 32-byte scratch planes, valid until the
   next call on this server; bytes past ceil(width/8) are stale. */
int b2_server_get_net_state(b2_server *sv, uint64_t client, uint32_t net, uint8_t *out_width, const uint8_t **plane0,
                            const uint8_t **plane1);
/* Bulk replacement for Adapter::sim_state (server.rs:375-383): all `n` nets of the client in
   SimState layout, see b2_pack_sim_state. */
int b2_server_sim_state(b2_server *sv, uint64_t client, const uint32_t *nets, uint32_t n, uint8_t *plane0, uint8_t *plane1,
                        size_t cap_bytes, uint64_t *out_bit_len);

#ifdef __cplusplus
}
#endif
#endif /* DIGILOGIC_B200_H */
